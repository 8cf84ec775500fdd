#!/usr/bin/env bash
# 2-GPU pass (round 2, late): NCCL parity of the eager and the graph-captured sharded step after the shared-code changes
# (pre-zeroed accumulators, one-kernel gradient flatten), bench N=2 with --check, NCCL protocol variants, timeline
set -u
out=gpurun_out; mkdir -p $out
{
nvidia-smi -L | head -3
echo "== default band"; timeout 120 python tools/k2_only.py rank 8192
echo "== sharded tests"; timeout 900 python -m pytest tests/test_gpu_sharded.py -q -p no:cacheprovider 2>&1 | tail -6
bench2() { env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 3 --no-extra 2>/dev/null | grep '^{' | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['value'], l['ms_per_step'], l['e2e']['value'], l['parity_checked'])"; }
echo "== bench N=2 (default)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 3 > $out/r2_bench_n2.json 2> $out/r2_bench_n2.err || tail -30 $out/r2_bench_n2.err
python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','gpu_launches','parity_checked')}); print('e2e',l['e2e']['value'])
except Exception as e: print('bench parse failed', e)
PY
echo "== NCCL_PROTO=Simple"; bench2 NCCL_PROTO=Simple
echo "== NCCL_PROTO=LL128"; bench2 NCCL_PROTO=LL128
echo "== NCCL_ALGO=Tree"; bench2 NCCL_ALGO=Tree
echo "== trace"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/trace_step.py --out $out/r2_trace_n2.txt > /dev/null 2> $out/r2_trace_n2.err || tail -5 $out/r2_trace_n2.err
head -2 $out/r2_trace_n2.txt
} > $out/r2_n2b.txt 2>&1
tail -c 4000 $out/r2_n2b.txt
