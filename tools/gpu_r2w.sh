#!/usr/bin/env bash
set -u
o=gpurun_out; mkdir -p $o
{
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_gpu_install.py -q -m gpu -p no:cacheprovider -k "assigning_a_new_bank" 2>&1 | grep -E "AssertionError|passed|failed" | cut -c1-300; done
} > $o/r2w.txt 2>&1
cat $o/r2w.txt
