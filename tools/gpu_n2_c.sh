#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
run2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_tc2.py tests/test_gpu_parity.py -q -p no:cacheprovider 2>&1 | tail -4
run2 29551 bench.py --gpus 2 --steps 20 --warmup 3 > $out/r2_bench_n2.json 2> $out/r2_bench_n2.err; echo "bench n2 rc=$?"
run2 29554 tools/trace_step.py --out $out/r2_trace_n2.txt > /dev/null 2> $out/r2_trace_n2.err
python bench.py --steps 20 > $out/r2g_bench_n1.json 2> $out/r2g_bench_n1.err; echo "bench n1 rc=$?"
for f in r2_bench_n2 r2g_bench_n1; do
  echo "== $f"; grep -v "Warning\|warn\|run_backward\|^\*\*\*\|OMP_NUM" $out/$f.err | tail -5
  python - "$out/$f.json" <<'PY'
import json,sys
d=None
for l in open(sys.argv[1]):
    if l.startswith('{'): d=json.loads(l)
if d is None: print('no json'); sys.exit()
print({k:d.get(k) for k in ('metric','value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'])
print('parity', d.get('parity_checked')); print('roofline', {k:d['roofline'].get(k) for k in ('achieved','frac','avg_launch_ms')} if d.get('roofline') else None)
print('modes', d.get('modes'))
PY
done
