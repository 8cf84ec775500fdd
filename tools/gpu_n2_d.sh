#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
run2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for i in 1 2; do
run2 2955$i bench.py --gpus 2 --per-gpu-batch 1024 --steps 3 --no-extra > $out/r2_b1024_n2.json 2> $out/r2_b1024_n2.err; echo "b1024 n2 rc=$?"
grep "check FAILED" $out/r2_b1024_n2.err | cut -c1-300
done
python - <<'PY'
import json
d=None
for l in open('gpurun_out/r2_b1024_n2.json'):
    if l.startswith('{'): d=json.loads(l)
print({k:d.get(k) for k in ('value','ms_per_step')} if d else 'no json', d.get('parity_checked') if d else '')
PY
