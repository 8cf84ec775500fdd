#!/usr/bin/env bash
# 8-GPU session: the scaling bench at N=8 (with --check), its timeline, ActivityNet shape, global batch 8192, 100k eval
set -u
out=gpurun_out; mkdir -p $out
N=${1:-8}
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
nvidia-smi -L | wc -l
run 600 29561 bench.py --gpus $N --steps 20 --warmup 3 > $out/r2_bench_n$N.json 2> $out/r2_bench_n$N.err; echo "bench rc=$?"
run 600 29562 tools/trace_step.py --out $out/r2_trace_n$N.txt > /dev/null 2> $out/r2_trace_n$N.err; echo "trace rc=$?"
run 900 29563 bench.py --gpus $N --shape activitynet --steps 10 --warmup 3 > $out/r2_bench_act_n$N.json 2> $out/r2_bench_act_n$N.err; echo "act rc=$?"
run 900 29564 bench.py --gpus $N --per-gpu-batch 1024 --steps 5 --warmup 3 --no-extra > $out/r2_bench_b8192_n$N.json 2> $out/r2_bench_b8192_n$N.err; echo "b8192 rc=$?"
run 900 29565 bench.py --gpus $N --workload eval --eval-size 100000 --steps 3 --warmup 1 > $out/r2_eval100k_n$N.json 2> $out/r2_eval100k_n$N.err; echo "eval100k rc=$?"
if [ "$N" = "8" ]; then
  N=4; run 600 29566 bench.py --gpus 4 --steps 20 --warmup 3 > $out/r2_bench_n4.json 2> $out/r2_bench_n4.err; echo "bench n4 rc=$?"; N=8
fi
for f in r2_bench_n$N r2_bench_act_n$N r2_bench_b8192_n$N r2_eval100k_n$N r2_bench_n4; do
  [ -f $out/$f.json ] || continue
  echo "== $f"; grep -v "Warning\|warn\|run_backward\|^\*\*\*\|OMP_NUM\|NCCL" $out/$f.err | grep -i "error\|fail\|Traceback" | head -5
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
