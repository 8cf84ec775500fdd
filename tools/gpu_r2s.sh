#!/usr/bin/env bash
set -u
o=gpurun_out; mkdir -p $o
{
timeout 900 python -m pytest tests/test_gpu_install.py tests/test_gpu_parity.py tests/test_gpu_bank_ring.py -q -m gpu -p no:cacheprovider -x 2>&1 | grep -E "^E  |Error|assert|passed|failed|^tests" | cut -c1-400 | head -60
} > $o/r2s.txt 2>&1
tail -c 5000 $o/r2s.txt
