#!/usr/bin/env bash
# round-2 session R: one logits fill, loss-value reductions off the critical path; final N=1 evidence
set -u
o=gpurun_out; mkdir -p $o
{
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -4 | tee $o/r2_gpu_tests_final.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py > $o/r2_bench_n1.json 2> $o/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], json.dumps(d['modes'])[:200], json.dumps(d['eval'])[:300])
PY
echo "== bench activitynet"; timeout 600 python bench.py --shape activitynet --no-extra > $o/r2_bench_activitynet_n1.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_activitynet_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
PY
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
echo "== trace"; timeout 300 python tools/trace_step.py --out $o/r2_trace_graph_step_n1.txt 2>&1 | tail -1
echo "== K1 alone"; timeout 120 python tools/k2_only.py; timeout 120 python tools/k2_only.py 1024 512; timeout 120 python tools/k2_only.py 128 64 1024 64; timeout 120 python tools/k2_only.py rank 1000
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > /dev/null 2>&1; wc -l $o/r2_launches_bench.csv
} > $o/r2r.txt 2>&1
tail -c 4500 $o/r2r.txt
