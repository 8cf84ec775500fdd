#!/usr/bin/env bash
# round-2 session T: elected barrier arrivals in K1 (single and pair), side-branch buffer lifetime fix
set -u
o=gpurun_out; mkdir -p $o
{
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -4 | tee $o/r2_gpu_tests_final.txt
echo "== pair forced parity"; NR_TC2_PAIR=2 timeout 600 python -m pytest tests/test_gpu_tc2.py tests/test_gpu_x3.py -x -q -m gpu 2>&1 | tail -2
for p in 0 1; do
  echo "== NR_TC2_PAIR=$p"
  NR_TC2_PAIR=$p timeout 120 python tools/k2_only.py
  NR_TC2_PAIR=$p timeout 120 python tools/k2_only.py 1024 512
  NR_TC2_PAIR=$p timeout 120 python tools/k2_only.py 128 64 1024 64
  NR_TC2_PAIR=$p timeout 120 python tools/k2_only.py rank 8192
done
for p in 0 1; do
echo "== bench NR_TC2_PAIR=$p"; NR_TC2_PAIR=$p timeout 600 python bench.py --no-cpu-baseline > $o/bench_pair$p.json 2> $o/bench_n1.err; python - $p <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/bench_pair{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['e2e']['value'], json.dumps(d['modes'])[:120])
PY
done
} > $o/r2t.txt 2>&1
tail -c 4500 $o/r2t.txt
