#!/usr/bin/env bash
# round-2 session X: one-launch softmax of both modalities, arg-max of the row direction only where it is saved
set -u
o=gpurun_out; mkdir -p $o
{
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -3 | tee $o/r2_gpu_tests_final.txt
echo "== bench"; timeout 600 python bench.py > $o/r2_bench_n1.json 2> $o/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], json.dumps(d['modes'])[:200], json.dumps(d['eval'])[:260])
PY
echo "== count pass"; timeout 120 python tools/k2_only.py rank 1000; timeout 120 python tools/k2_only.py rank 8192
} > $o/r2x.txt 2>&1
tail -c 2500 $o/r2x.txt
