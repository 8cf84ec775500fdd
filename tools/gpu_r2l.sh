#!/usr/bin/env bash
# round-2 session L: staged-slab input pipeline (one D2D copy, prefetch issued after the replay), 16-byte row staging
set -u
o=gpurun_out; mkdir -p $o
{
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -8
echo "== memory-bound kernels"; REPS=3 timeout 600 python tools/membound_only.py 2>&1 | tail -30
echo "== bench"; timeout 600 python bench.py > $o/bench_n1.json 2> $o/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d['e2e']), json.dumps(d['modes'])[:300])
PY
} > $o/r2l.txt 2>&1
tail -c 7000 $o/r2l.txt
