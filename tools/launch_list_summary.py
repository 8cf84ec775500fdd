"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares:
    python tools/launch_list_summary.py gpurun_out/r2_launches_bench.csv "header line" > profiles/r2_launches_bench.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES of the step, not absolute times."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if r[0] == "ID")
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows:
    if r[0] == "ID" or not r[0].isdigit():
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("at::", "")[:92]
    us = float(r[iv].replace(",", "")) * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3}.get(r[iu], 1e-3)
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + us)
total = sum(t for _, t in tot.values())
for line in sys.argv[2:]:
    print("# " + line)
print(f"# {sum(n for n, _ in tot.values())} launches, {total:.0f} us summed (cold-cache, serialised under ncu: compare SHARES, not absolute times)")
print(f"# {'kernel':92s} {'launches':>8s} {'us':>10s} {'share':>7s}")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:94s} {n:8d} {t:10.1f} {100 * t / total:6.1f}%")
