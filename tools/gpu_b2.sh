#!/usr/bin/env bash
# backward routing kernel after a change: parity tests that reach it, its time alone, a short bench
o=gpurun_out; mkdir -p $o
{
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_tc2.py tests/test_gpu_x3.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
echo "== alone"
timeout 120 python tools/b2_only.py
timeout 120 python tools/b2_step.py
timeout 120 python tools/k2_only.py
echo "== bench"
timeout 300 python bench.py --steps 200 --warmup 20 --no-extra 2>&1 | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['e2e']['value'])"
} > $o/b2.txt 2>&1
tail -40 $o/b2.txt
