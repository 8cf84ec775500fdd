#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
run2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for rows in 4 8 16 32; do
  NR_SINKHORN_ROWS=$rows run2 2957$((rows % 10)) bench.py --gpus 2 --steps 20 --warmup 3 --no-extra --no-check > $out/r2_sk_n2_$rows.json 2> /dev/null
  python - $rows <<'PY'
import json,sys
d=None
for l in open(f'gpurun_out/r2_sk_n2_{sys.argv[1]}.json'):
    if l.startswith('{'): d=json.loads(l)
print('rows', sys.argv[1], {k:d.get(k) for k in ('value','ms_per_step')} if d else 'no json')
PY
done
# larger global batch on 2 GPUs (B = 1024, as at 8 ranks): rows 8 / 16 / 24
for rows in 8 16 24; do
  NR_SINKHORN_ROWS=$rows run2 2958$((rows % 10)) bench.py --gpus 2 --per-gpu-batch 512 --steps 10 --warmup 3 --no-extra --no-check > $out/r2_sk_b512_$rows.json 2> /dev/null
  python - $rows <<'PY'
import json,sys
d=None
for l in open(f'gpurun_out/r2_sk_b512_{sys.argv[1]}.json'):
    if l.startswith('{'): d=json.loads(l)
print('b512 rows', sys.argv[1], {k:d.get(k) for k in ('value','ms_per_step')} if d else 'no json')
PY
done
