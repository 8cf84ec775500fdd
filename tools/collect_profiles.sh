#!/usr/bin/env bash
# gpurun_out/ (scratch) -> profiles/ (tracked): the round-2 evidence.  Run here after the GPU sessions.
set -u
g=gpurun_out; p=profiles
lastjson() { python - "$1" "$2" <<'PY'
import json,sys
d=None
try:
    for l in open(sys.argv[1]):
        if l.startswith('{'): d=l
except FileNotFoundError:
    sys.exit(0)
if d: open(sys.argv[2],'w').write(d)
PY
}
for f in bench_n1 bench_n2 bench_n4 bench_n8 bench_act_n8 bench_activitynet_n1 bench_b8192_n8 eval100k_n8; do lastjson $g/r2_$f.json $p/r2_$f.json; done
for f in trace_graph_step_n1 trace_n2 trace_n8 kernels_alone membound_events gpu_tests_final; do [ -f $g/r2_$f.txt ] && cp $g/r2_$f.txt $p/r2_$f.txt; done
[ -f $g/r2_k2_fwd.ncu-rep ] && python tools/ncu_summary.py $g/r2_k2_fwd.ncu-rep \
  "ncu --set full --clock-control none -k regex:maxsim2_fwd -s 2 -c 2  python tools/k2_only.py   (round 2: 16 epilogue warps, v+2 keys in the bf16 mode)" \
  "nr_maxsim2_fwd: forward contractions of one MSR-VTT-shaped head step (b=128, M=512; batch pair + two bank pairs) in ONE launch: 1579 tiles, 43.5 GFLOP algorithmic" > $p/r2_k2_fwd_ncu_full.txt
[ -f $g/r2_b2_bwd.ncu-rep ] && python tools/ncu_summary.py $g/r2_b2_bwd.ncu-rep \
  "ncu --set full --clock-control none -k regex:maxsim2_bwd_tc -s 2 -c 2  python tools/b2_step.py   (round 2)" \
  "nr_maxsim2_bwd: the four token-gradient contractions of one MSR-VTT-shaped head step in ONE launch, 48.3 GFLOP algorithmic" > $p/r2_b2_bwd_ncu_full.txt
[ -f $g/r2_gemm.ncu-rep ] && python tools/ncu_summary.py $g/r2_gemm.ncu-rep \
  "ncu --set full --clock-control none -k regex:gemm_bf16_tc -s 2 -c 2  python tools/gemm_only.py   (round 2)" \
  "nr_mlp_fwd_pair: the token-weight MLP first layers of both modalities (15360 + 7680 tokens, 512 -> 1024) with bias + ReLU + second-layer dot in the epilogue, 24.2 GFLOP" > $p/r2_gemm_ncu_full.txt
[ -f $g/r2_membound_ncu.csv ] && python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_membound_ncu.csv')) if len(r)>5]
hdr=None; per=collections.OrderedDict()
for r in rows:
    if r[0]=='ID': hdr=r; continue
    if hdr is None: continue
    i=r[hdr.index('ID')]; name=r[hdr.index('Kernel Name')].split('(')[0][:70]; m=r[hdr.index('Metric Name')]; v=float(r[hdr.index('Metric Value')].replace(',','')); u=r[hdr.index('Metric Unit')]
    per.setdefault((i,name),{})[m]=(v,u)
def to(v,u,kind):
    f={'nsecond':1e-3,'usecond':1,'msecond':1e3,'ns':1e-3,'us':1,'ms':1e3,'byte':1e-6,'Kbyte':1e-3,'Mbyte':1,'Gbyte':1e3,'%':1}.get(u,1)
    return v*f
out=["# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput... --clock-control none  python tools/membound_only.py  (REPS=1)",
     "# memory-bound kernels at the sizes where they leave L2 (global batch 8192, 100k-video gallery); achieved = DRAM bytes / time; peak = MEASURED_PEAKS.json hbm_gbs 6537.6",
     f"# {'kernel':70s} {'us':>10s} {'dram rd MB':>11s} {'dram wr MB':>11s} {'GB/s':>8s} {'of 6537.6':>9s} {'dram %pk':>8s}"]
for (i,name),m in per.items():
    if not name.startswith(('nr::','void nr::')): continue
    t=to(*m['gpu__time_duration.sum'],'t'); rd=to(*m['dram__bytes_read.sum'],'b'); wr=to(*m['dram__bytes_write.sum'],'b')
    if t < 20: continue
    gbs=(rd+wr)/t*1e3
    out.append(f"{name:72s} {t:10.1f} {rd:11.1f} {wr:11.1f} {gbs*1e0:8.0f} {gbs/6537.6:9.2f} {m['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][0]:8.1f}")
open('profiles/r2_membound_ncu.txt','w').write("\n".join(out)+"\n")
print("\n".join(out))
PY
ls $p | grep r2_ | tr '\n' ' '
