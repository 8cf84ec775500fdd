#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -rP 2>&1 | grep -E "^(x3|bf16|bf16x3|head x3|step|cfg|token_weights|mlp_|nr_)|passed|failed|^FAILED" | cut -c1-600 > $out/r2_gpu_tests_final.txt
tail -3 $out/r2_gpu_tests_final.txt
python tools/k2_only.py | tail -1 > $out/r2_kernels_alone.txt; python tools/k2_only.py 1024 512 | tail -1 >> $out/r2_kernels_alone.txt
python tools/k2_only.py 128 64 1024 64 | tail -1 >> $out/r2_kernels_alone.txt
python tools/b2_step.py | tail -1 >> $out/r2_kernels_alone.txt; python tools/gemm_only.py | tail -2 >> $out/r2_kernels_alone.txt
python tools/k2_trace.py >> $out/r2_kernels_alone.txt 2>&1
cat $out/r2_kernels_alone.txt
python bench.py > $out/r2_bench_n1.json 2> $out/r2_bench_n1.err || tail -5 $out/r2_bench_n1.err
REPS=5 python tools/membound_only.py > $out/r2_membound_events.txt 2>&1; tail -4 $out/r2_membound_events.txt
REPS=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --csv --log-file $out/r2_membound_ncu.csv python tools/membound_only.py > $out/r2_membound_ncu.log 2>&1
tail -2 $out/r2_membound_ncu.log
ncu --set full --clock-control none --import-source on -k "regex:maxsim2_fwd" -s 2 -c 2 -f -o "$out/r2_k2_fwd" python tools/k2_only.py > "$out/r2_k2_fwd_ncu.log" 2>&1
python - <<'PY'
import json
d=None
for l in open('gpurun_out/r2_bench_n1.json'):
    if l.startswith('{'): d=json.loads(l)
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'])
print(d['modes']); print(d['eval']['resident_ms'], d['eval']['e2e_ms'], d['eval']['roofline']['frac'])
PY
