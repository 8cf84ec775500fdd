#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_x3.py tests/test_gpu_install.py -q -x -p no:cacheprovider 2>&1 | grep -v "Warning\|warnings.warn" | tail -120 > $out/r2c_fail1.log
timeout 600 python -m pytest "tests/test_gpu_install.py" -q -p no:cacheprovider -rP 2>&1 | grep -E "^step|Error|assert|^E " | head -80 > $out/r2c_install.log
timeout 600 python -m pytest "tests/test_gpu_x3.py" -q -p no:cacheprovider -rP 2>&1 | grep -E "^x3|^head|Error|assert|^E " | head -60 > $out/r2c_x3.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -25 > $out/r2c_gpu_tests.log
tail -12 $out/r2c_gpu_tests.log
python tools/trace_step.py --out $out/r2c_trace_n1.txt > /dev/null 2>&1 || echo trace failed
python bench.py --steps 20 > $out/r2c_bench_n1.json 2> $out/r2c_bench_n1.err || tail -20 $out/r2c_bench_n1.err
tail -c 1200 $out/r2c_bench_n1.json
