#!/usr/bin/env bash
# cta_group::2 variant of nr_maxsim2_fwd: parity (forced on every shape that allows it) and timing against the
# single-CTA kernel.  Every command under its own timeout: a wrong barrier protocol hangs.
o=gpurun_out; mkdir -p $o
{
echo "== parity, pair forced"; NR_TC2_PAIR=2 timeout 600 python -m pytest tests/test_gpu_tc2.py tests/test_gpu_x3.py -x -q -m gpu 2>&1 | tail -3
echo "== parity, single CTA"; NR_TC2_PAIR=0 timeout 600 python -m pytest tests/test_gpu_tc2.py tests/test_gpu_x3.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for p in 0 1; do
  echo "== NR_TC2_PAIR=$p"
  NR_TC2_PAIR=$p timeout 120 python tools/k2_only.py
  NR_TC2_PAIR=$p timeout 120 python tools/k2_only.py 1024 512
  NR_TC2_PAIR=$p timeout 120 python tools/k2_only.py 128 64 1024 64
  NR_TC2_PAIR=$p NR_TC2_TRACE=1 timeout 120 python tools/k2_trace.py 2>&1 | tail -14
done
for p in 0 1; do
  echo "== bench NR_TC2_PAIR=$p"
  NR_TC2_PAIR=$p timeout 300 python bench.py --steps 200 --warmup 20 --no-extra 2>&1 | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['e2e']['value'])"
done
} > $o/pair.txt 2>&1
tail -60 $o/pair.txt
