#!/usr/bin/env bash
# 4-GPU pass (round 2, final code): bench line with the in-bench parity check
set -u
out=gpurun_out; mkdir -p $out
{
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 3 --no-extra 2>$out/n4.err | grep '^{' | tail -1 > $out/r2_bench_n4.json
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_n4.json').read().strip().splitlines()[-1]); print(l['value'], l['ms_per_step'], l['e2e']['value'], l['parity_checked'])
PY
tail -2 $out/n4.err
} > $out/r2_n4.txt 2>&1
cat $out/r2_n4.txt | cut -c1-600
