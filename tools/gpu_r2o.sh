#!/usr/bin/env bash
# round-2 session O: redux arg-max in the row-loss kernels, detached weight-gradient branch, high-priority global branches
set -u
o=gpurun_out; mkdir -p $o
{
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
echo "== memory-bound kernels"; REPS=3 timeout 600 python tools/membound_only.py 2>&1 | sed -n 2,3p
echo "== bench"; timeout 600 python bench.py > $o/bench_n1.json 2> $o/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], json.dumps(d['modes'])[:200])
PY
echo "== trace"; timeout 300 python tools/trace_step.py --out $o/r2_trace_graph_step_n1.txt 2>&1 | tail -2
} > $o/r2o.txt 2>&1
tail -c 3000 $o/r2o.txt
