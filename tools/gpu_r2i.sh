#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_zz_fullsize.py tests/test_gpu_parity.py -q -p no:cacheprovider -rP 2>&1 | grep -E "^cfg|passed|failed|^FAILED|^E  " | cut -c1-300 | tail -25
REPS=3 python tools/membound_only.py 2>&1 | head -4
