#!/usr/bin/env bash
# round-2 session K: fused evaluation parity (shared token weights), cached top-k selection in the row-loss / top-k kernels
set -u
o=gpurun_out; mkdir -p $o
{
echo "== fused eval parity"; timeout 600 python -m pytest tests/test_gpu_eval_fused.py -q -m gpu 2>&1 | tail -15
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8
echo "== memory-bound kernels"; REPS=3 timeout 600 python tools/membound_only.py 2>&1 | tail -30
echo "== bench"; timeout 600 python bench.py > $o/bench_n1.json 2> $o/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], json.dumps(d['eval'])[:900])
PY
} > $o/r2k.txt 2>&1
tail -c 7000 $o/r2k.txt
