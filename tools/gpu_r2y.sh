#!/usr/bin/env bash
# round-2 final profile refresh: ncu --set full of the two contraction kernels (final code), kernels alone, memory-bound kernels
set -u
out=gpurun_out; mkdir -p $out
{
python tools/k2_only.py | tail -1 > $out/r2_kernels_alone.txt; python tools/k2_only.py 1024 512 | tail -1 >> $out/r2_kernels_alone.txt
python tools/k2_only.py 128 64 1024 64 | tail -1 >> $out/r2_kernels_alone.txt; python tools/k2_only.py rank 1000 | tail -1 >> $out/r2_kernels_alone.txt
python tools/k2_only.py rank 8192 | tail -1 >> $out/r2_kernels_alone.txt
python tools/b2_step.py | tail -1 >> $out/r2_kernels_alone.txt; python tools/gemm_only.py | tail -2 >> $out/r2_kernels_alone.txt
NR_TC2_TRACE=1 timeout 120 python tools/k2_trace.py >> $out/r2_kernels_alone.txt 2>&1
cat $out/r2_kernels_alone.txt | cut -c1-160
REPS=5 timeout 300 python tools/membound_only.py > $out/r2_membound_events.txt 2>&1; cat $out/r2_membound_events.txt
for spec in "k2_fwd:maxsim2_fwd:tools/k2_only.py:2" "b2_bwd:maxsim2_bwd_tc:tools/b2_step.py:2"; do
  IFS=: read -r name regex script skip <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 2 -f -o "$out/r2_${name}" python "$script" > "$out/r2_${name}_ncu.log" 2>&1
  tail -1 "$out/r2_${name}_ncu.log"
done
} > $out/r2y.txt 2>&1
tail -c 3500 $out/r2y.txt
