#!/usr/bin/env bash
# A/B on one box: elected barrier releases in K1, single and pair mode
set -u
o=gpurun_out; mkdir -p $o
{
for rep in 1 2; do
for cfg in "NR_TC2_PAIR=0 NR_TC2_ELECT=0" "NR_TC2_PAIR=0 NR_TC2_ELECT=1" "NR_TC2_PAIR=1 NR_TC2_ELECT=0" "NR_TC2_PAIR=1 NR_TC2_ELECT=1"; do
  echo "== $cfg"
  env $cfg timeout 120 python tools/k2_only.py 2>&1 | cut -c1-110
  env $cfg timeout 120 python tools/k2_only.py 1024 512 2>&1 | cut -c1-110
  env $cfg timeout 120 python tools/k2_only.py 128 64 1024 64 2>&1 | cut -c1-110
done; done
for cfg in "NR_TC2_ELECT=0" "NR_TC2_ELECT=1" "NR_TC2_ELECT=0" "NR_TC2_ELECT=1"; do
echo "== bench $cfg"; env $cfg timeout 600 python bench.py --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['e2e']['value'])"
done
} > $o/r2u.txt 2>&1
tail -c 5000 $o/r2u.txt
