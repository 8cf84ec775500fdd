#!/usr/bin/env bash
# 8-GPU pass (round 2, final code): bench line with the in-bench parity check; A/B of ordering the contraction behind the
# Sinkhorn chain at global batch 1024 (NR_SINKHORN_SERIAL)
set -u
out=gpurun_out; mkdir -p $out
b8() { env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 20 --warmup 3 --no-extra 2>$out/n8.err | grep '^{' | tail -1; }
{
echo "== default"; b8 NR_X=0 > $out/r2_bench_n8.json; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1]); print(l['value'], l['ms_per_step'], l['e2e']['value'], l['parity_checked'])
PY
echo "== NR_SINKHORN_SERIAL=1024"; b8 NR_SINKHORN_SERIAL=1024 > $out/r2_bench_n8_serial.json; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_n8_serial.json').read().strip().splitlines()[-1]); print(l['value'], l['ms_per_step'], l['e2e']['value'], l['parity_checked'])
PY
tail -3 $out/n8.err
} > $out/r2_n8c.txt 2>&1
tail -c 3000 $out/r2_n8c.txt
