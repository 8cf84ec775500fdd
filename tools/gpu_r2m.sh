#!/usr/bin/env bash
# round-2 session M: pre-zeroed accumulators, split bank insert, 1024-thread row kernels
set -u
o=gpurun_out; mkdir -p $o
{
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -8
echo "== memory-bound kernels"; REPS=3 timeout 600 python tools/membound_only.py 2>&1 | sed -n 2,3p
echo "== bench"; timeout 600 python bench.py > $o/bench_n1.json 2> $o/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], json.dumps(d['modes'])[:300])
PY
echo "== bench NR_SPLIT_INSERT=0"; NR_SPLIT_INSERT=0 timeout 600 python bench.py --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "== trace"; timeout 300 python tools/trace_step.py --out $o/r2_trace_graph_step_n1.txt 2>&1 | tail -3
} > $o/r2m.txt 2>&1
tail -c 5000 $o/r2m.txt
