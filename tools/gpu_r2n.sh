#!/usr/bin/env bash
# round-2 session N: raw-only late insert, shared capture workspace; ncu of the B = 8192 row-loss forward and of the count pass
set -u
o=gpurun_out; mkdir -p $o
{
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
echo "== bench"; timeout 600 python bench.py > $o/bench_n1.json 2> $o/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], json.dumps(d['modes'])[:200])
PY
echo "== trace"; timeout 300 python tools/trace_step.py --out $o/r2_trace_graph_step_n1.txt 2>&1 | tail -2
echo "== count pass"; timeout 120 python tools/k2_only.py rank 1000; timeout 120 python tools/k2_only.py rank 8192
echo "== ncu row fwd"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_losses_fwd_kernel -c 1 -o $o/r2_rowfwd -f python tools/membound_only.py > $o/ncu_rowfwd.log 2>&1; tail -2 $o/ncu_rowfwd.log
echo "== ncu count pass"; timeout 600 ncu --set full --clock-control none -k regex:maxsim2_fwd_tc -s 3 -c 1 -o $o/r2_rank_count -f python tools/k2_only.py rank 8192 > $o/ncu_rank.log 2>&1; tail -2 $o/ncu_rank.log
} > $o/r2n.txt 2>&1
tail -c 3000 $o/r2n.txt
