#!/usr/bin/env bash
# round-2 session P: band-ordered K1 tiles — parity, large problems (8192 x 8192 count pass, b = 1024 step shape, 100k evaluation)
set -u
o=gpurun_out; mkdir -p $o
{
echo "== GPU suite (tc2, x3, eval, parity, fullsize)"; timeout 1500 python -m pytest tests/test_gpu_tc2.py tests/test_gpu_x3.py tests/test_gpu_eval_fused.py tests/test_gpu_parity.py tests/test_zz_fullsize.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
echo "== K1 shapes"; timeout 120 python tools/k2_only.py; timeout 120 python tools/k2_only.py 1024 512; timeout 120 python tools/k2_only.py 1024 24 8192 12; NR_TC2_BAND=100000 timeout 120 python tools/k2_only.py 1024 24 8192 12
echo "== count pass"; timeout 120 python tools/k2_only.py rank 8192; NR_TC2_BAND=100000 timeout 120 python tools/k2_only.py rank 8192
for b in 16 32 130; do echo "band $b"; NR_TC2_BAND=$b timeout 120 python tools/k2_only.py rank 8192; done
echo "== eval 100k x 100k, one GPU, fused"; timeout 900 python bench.py --workload eval --eval-size 100000 --steps 2 --warmup 1 > $o/r2_eval100k_n1_fused.json 2> $o/eval100k_n1.err; cat $o/r2_eval100k_n1_fused.json | cut -c1-400; tail -2 $o/eval100k_n1.err
echo "== ncu count pass"; timeout 600 ncu --set full --clock-control none -k regex:maxsim2_fwd_tc -s 3 -c 1 -o $o/r2_rank_count -f python tools/k2_only.py rank 8192 > $o/ncu_rank.log 2>&1; tail -1 $o/ncu_rank.log
} > $o/r2p.txt 2>&1
tail -c 4000 $o/r2p.txt
