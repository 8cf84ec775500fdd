#!/usr/bin/env bash
# round-2 session Q: why does the 8192 x 8192 contraction read its operands 400x from HBM?  L2 hints / promotion / band width
set -u
o=gpurun_out; mkdir -p $o
run() { echo "-- $*"; env "$@" timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:maxsim2_fwd_tc -s 3 -c 1 python tools/k2_only.py rank 8192 2>&1 | grep -E "gpu__time_duration|dram__bytes_read|hit_rate"; }
{
run NR_TC2_BAND=100000
run NR_TC2_BAND=65
run NR_TC2_BAND=16
run NR_TC2_BAND=65 NR_TC2_L2HINT=1
run NR_TC2_BAND=16 NR_TC2_L2HINT=1
run NR_TC2_BAND=65 NR_TC2_L2HINT=2
run NR_TC2_BAND=65 NR_TMA_PROMO=0
run NR_TC2_BAND=65 NR_TMA_PROMO=128
run NR_TC2_BAND=8 NR_TC2_L2HINT=1
echo "== no-ncu timing"; for e in "NR_TC2_BAND=65" "NR_TC2_BAND=65 NR_TC2_L2HINT=1" "NR_TC2_BAND=16 NR_TC2_L2HINT=1" "NR_TC2_BAND=65 NR_TMA_PROMO=0"; do echo "-- $e"; env $e timeout 120 python tools/k2_only.py rank 8192; done
} > $o/r2q.txt 2>&1
tail -c 5000 $o/r2q.txt
