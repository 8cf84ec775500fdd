#!/usr/bin/env bash
# round-2 evidence pass (1 GPU): tests, bench line, launch list, ncu --set full of the tensor-core kernels, DRAM
# counters of the memory-bound kernels.  Summaries are written under profiles/ afterwards (tools/ncu_summary.py).
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -rP 2>&1 | grep -E "^(x3|bf16|bf16x3|head x3|step|cfg|token_weights|mlp_|nr_)|passed|failed|^FAILED" | cut -c1-600 > $out/r2_gpu_tests_final.txt
tail -3 $out/r2_gpu_tests_final.txt
python bench.py > $out/r2_bench_n1.json 2> $out/r2_bench_n1.err || tail -5 $out/r2_bench_n1.err
python bench.py --shape activitynet --no-cpu-baseline --steps 10 > $out/r2_bench_activitynet_n1.json 2> /dev/null
python tools/trace_step.py --out $out/r2_trace_graph_step_n1.txt > /dev/null 2>&1
python tools/k2_only.py | tail -1 > $out/r2_kernels_alone.txt; python tools/k2_only.py 1024 512 | tail -1 >> $out/r2_kernels_alone.txt
python tools/k2_only.py 128 64 1024 64 | tail -1 >> $out/r2_kernels_alone.txt
python tools/b2_step.py | tail -1 >> $out/r2_kernels_alone.txt; python tools/gemm_only.py | tail -2 >> $out/r2_kernels_alone.txt
python tools/k2_trace.py >> $out/r2_kernels_alone.txt 2>&1
cat $out/r2_kernels_alone.txt
REPS=5 python tools/membound_only.py > $out/r2_membound_events.txt 2>&1; cat $out/r2_membound_events.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $out/r2_ncu_bench.log 2>&1
for spec in "k2_fwd:maxsim2_fwd:tools/k2_only.py:2" "b2_bwd:maxsim2_bwd_tc:tools/b2_step.py:2" "gemm:gemm_bf16_tc:tools/gemm_only.py:2"; do
  IFS=: read -r name regex script skip <<< "$spec"
  ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 2 -f -o "$out/r2_${name}" \
      python "$script" > "$out/r2_${name}_ncu.log" 2>&1
  tail -1 "$out/r2_${name}_ncu.log"
done
REPS=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --csv --log-file $out/r2_membound_ncu.csv python tools/membound_only.py > $out/r2_membound_ncu.log 2>&1
tail -2 $out/r2_membound_ncu.log
