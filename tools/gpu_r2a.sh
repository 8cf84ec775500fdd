#!/usr/bin/env bash
# round-2 first GPU pass: new tests first, then the whole suite, kernel A/B timings, a bench line
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_x3.py tests/test_gpu_install.py -q -x -rP -p no:cacheprovider 2>&1 | tail -60 > $out/r2a_new_tests.log
tail -5 $out/r2a_new_tests.log
python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -40 > $out/r2a_gpu_tests.log
tail -5 $out/r2a_gpu_tests.log
for h in 1 2; do
  NR_TC2_HALVES=$h python tools/k2_only.py 2>&1 | tail -1
  NR_TC2_HALVES=$h python tools/k2_only.py 1024 512 2>&1 | tail -1
  NR_TC2_HALVES=$h python tools/k2_only.py 128 64 1024 64 2>&1 | tail -1
done > $out/r2a_k2_ab.log 2>&1
cat $out/r2a_k2_ab.log
python bench.py --check --steps 10 > $out/r2a_bench_n1.json 2> $out/r2a_bench_n1.err || tail -20 $out/r2a_bench_n1.err
tail -c 1500 $out/r2a_bench_n1.json
