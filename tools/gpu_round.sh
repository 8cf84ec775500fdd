#!/usr/bin/env bash
# One gpurun call that produces the evidence of a round (run from the repo root ON the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/gpu_round.sh rN'
# 1. pytest -m gpu            -> gpurun_out/<tag>_gpu_tests.log
# 2. bench.py (N=1)           -> gpurun_out/<tag>_bench_n1.json  (never under a profiler)
# 3. ncu launch list of bench -> gpurun_out/<tag>_launches.csv   (cold-cache, serialised: compare shares)
# 4. ncu --set full of the two contraction kernels and the multi-sentence kernels -> gpurun_out/<tag>_*.ncu-rep
# Afterwards, here (no GPU needed): python tools/ncu_summary.py gpurun_out/<tag>_k2_fwd.ncu-rep "<header>" > profiles/...
set -u
tag=${1:-rX}
out=gpurun_out
mkdir -p "$out"
python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -30 > "$out/${tag}_gpu_tests.log"
tail -3 "$out/${tag}_gpu_tests.log"
python bench.py > "$out/${tag}_bench_n1.json" 2> "$out/${tag}_bench_n1.err" || { echo "bench failed"; tail -5 "$out/${tag}_bench_n1.err"; exit 1; }
tail -c 400 "$out/${tag}_bench_n1.json"; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$out/${tag}_launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$out/${tag}_ncu_bench.log" 2>&1
# name : kernel regex : script : launches to skip (warm-up launches of the script)
for spec in "k2_fwd:maxsim2_fwd:tools/k2_only.py:2" "b2_bwd:maxsim2_bwd_tc:tools/b2_step.py:2" "ms:rank_target|group_max:tools/ms_only.py:0"; do
  IFS=: read -r name regex script skip <<< "$spec"
  python "$script" > "$out/${tag}_${name}_plain.log" 2>&1 || { echo "$script failed without ncu"; continue; }
  ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 2 -f -o "$out/${tag}_${name}" \
      python "$script" > "$out/${tag}_${name}_ncu.log" 2>&1
  tail -1 "$out/${tag}_${name}_ncu.log"
done
