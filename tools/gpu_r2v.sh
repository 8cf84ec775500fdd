#!/usr/bin/env bash
# round-2 final N=1 evidence: full GPU suite, smoke, bench line, reference arm
set -u
o=gpurun_out; mkdir -p $o
{
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -3 | tee $o/r2_gpu_tests_final.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py > $o/r2_bench_n1.json 2> $o/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], json.dumps(d['modes'])[:200], json.dumps(d['eval'])[:260])
PY
echo "== trace"; timeout 300 python tools/trace_step.py --out $o/r2_trace_graph_step_n1.txt 2>&1 | tail -1
} > $o/r2v.txt 2>&1
tail -c 3000 $o/r2v.txt
