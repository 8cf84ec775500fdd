#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -rP -p no:cacheprovider 2>&1 | tail -80 > $out/r2b_gemm.log
tail -4 $out/r2b_gemm.log
timeout 600 python -m pytest tests/test_gpu_x3.py tests/test_gpu_install.py "tests/test_zz_fullsize.py::test_cfg2_msrvtt_head_global_batch_1024" -q -rP -p no:cacheprovider 2>&1 | grep -v "Warning\|warnings.warn\|^  " | tail -150 > $out/r2b_new_tests.log
tail -8 $out/r2b_new_tests.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -60 > $out/r2b_gpu_tests.log
tail -5 $out/r2b_gpu_tests.log
python tools/k2_trace.py > $out/r2b_trace.log 2>&1; cat $out/r2b_trace.log
NR_TC2_HALVES=1 python tools/k2_trace.py > $out/r2b_trace_h1.log 2>&1; cat $out/r2b_trace_h1.log
python tools/k2_only.py | tail -1; python tools/k2_only.py 1024 512 | tail -1
python bench.py --check --steps 10 > $out/r2b_bench_n1.json 2> $out/r2b_bench_n1.err || tail -20 $out/r2b_bench_n1.err
tail -c 2500 $out/r2b_bench_n1.json
