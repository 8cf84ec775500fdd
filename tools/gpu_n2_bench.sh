#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > $out/r2_bench_n2.json 2> $out/r2_bench_n2.err
echo "rc=$?"
grep -v "Warning\|warn\|run_backward\|^\*\*\*\|OMP_NUM" $out/r2_bench_n2.err | tail -40
head -c 600 $out/r2_bench_n2.json
