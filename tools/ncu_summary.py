"""Summarise an `ncu --set full` report (read here, without a GPU) into the text kept under profiles/:
    python tools/ncu_summary.py gpurun_out/k2_fwd_r1.ncu-rep "header line" > profiles/r1_k2_fwd_ncu_full.txt"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, body = rows[0], rows[1], rows[2:]
for line in sys.argv[2:]:
    print("# " + line)
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:75s} {units[i]:16s} " + " | ".join(r[i] for r in body))
