#!/usr/bin/env bash
# 2-GPU pass: NCCL parity of the eager and the graph-captured sharded step, bench N=2 with --check, timeline
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -p no:cacheprovider 2>&1 | tail -30 > $out/r2_n2_sharded_tests.log
tail -8 $out/r2_n2_sharded_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_graph_check.py --precision bf16 --b 128 --mrows 512 2>&1 | grep -E "^rank|Error|error" | cut -c1-700 > $out/r2_n2_graph_check_b128.log
cat $out/r2_n2_graph_check_b128.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 3 > $out/r2_bench_n2.json 2> $out/r2_bench_n2.err || tail -30 $out/r2_bench_n2.err
python - <<'PY'
import json
try:
    l=json.load(open('gpurun_out/r2_bench_n2.json'))
    print({k:l[k] for k in ('value','ms_per_step','gpu_launches','parity_checked')}); print('e2e',l['e2e']['value']); print(l['modes'])
except Exception as e: print('bench parse failed', e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/trace_step.py --out $out/r2_trace_n2.txt > /dev/null 2> $out/r2_trace_n2.err || tail -5 $out/r2_trace_n2.err
head -3 $out/r2_trace_n2.txt
