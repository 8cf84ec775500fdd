#!/usr/bin/env bash
# round-2 session J: fused evaluation ranks (nr_maxsim2_rank) — parity, full GPU suite, bench line, 100k x 100k on ONE GPU
set -u
o=gpurun_out; mkdir -p $o
{
echo "== fused eval parity"; timeout 600 python -m pytest tests/test_gpu_eval_fused.py -x -q -m gpu 2>&1 | tail -15
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8
echo "== bench"; timeout 600 python bench.py > $o/bench_n1.json 2> $o/bench_n1.err; tail -c 6000 $o/bench_n1.json; tail -3 $o/bench_n1.err
echo "== eval 100k x 100k, one GPU, fused"; timeout 900 python bench.py --workload eval --eval-size 100000 --steps 2 --warmup 1 > $o/eval100k_n1_fused.json 2> $o/eval100k_n1.err; cat $o/eval100k_n1_fused.json; tail -3 $o/eval100k_n1.err
echo "== eval 20k x 20k, one GPU, fused vs materialised"
timeout 600 python bench.py --workload eval --eval-size 20000 --steps 3 --warmup 1 2>/dev/null | tee $o/eval20k_n1_fused.json
timeout 600 python bench.py --workload eval --eval-size 20000 --steps 3 --warmup 1 --eval-ranks materialised 2>/dev/null | tee $o/eval20k_n1_mat.json
} > $o/r2j.txt 2>&1
tail -c 9000 $o/r2j.txt
