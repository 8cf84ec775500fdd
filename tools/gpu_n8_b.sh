#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 "${@:4}"; }
run 600 8 29561 bench.py --gpus 8 --steps 20 --warmup 3 > $out/r2_bench_n8.json 2> $out/r2_bench_n8.err; echo "bench rc=$?"
run 600 8 29562 tools/trace_step.py --out $out/r2_trace_n8.txt > /dev/null 2> $out/r2_trace_n8.err; echo "trace rc=$?"
run 900 8 29564 bench.py --gpus 8 --per-gpu-batch 1024 --steps 5 --warmup 3 --no-extra > $out/r2_bench_b8192_n8.json 2> $out/r2_bench_b8192_n8.err; echo "b8192 rc=$?"
run 600 4 29566 bench.py --gpus 4 --steps 20 --warmup 3 > $out/r2_bench_n4.json 2> $out/r2_bench_n4.err; echo "bench n4 rc=$?"
for f in r2_bench_n8 r2_bench_b8192_n8 r2_bench_n4; do
  echo "== $f"
  python - "$out/$f.json" <<'PY'
import json,sys
d=None
for l in open(sys.argv[1]):
    if l.startswith('{'): d=json.loads(l)
if d is None: print('no json'); sys.exit()
print({k:d.get(k) for k in ('metric','value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d.get('e2e',{}).get('value'))
print('parity', {k:d['parity_checked'].get(k) for k in ('world','loss_rel','grad_rel_l2','bank_equal')} if d.get('parity_checked') else None)
print('roofline', {k:d['roofline'].get(k) for k in ('achieved','frac','avg_launch_ms')} if d.get('roofline') else None)
print('modes', d.get('modes'))
PY
done
