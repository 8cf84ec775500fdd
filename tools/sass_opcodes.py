"""Per-kernel counts of the SASS opcodes that prove tcgen05 / TMEM / TMA / redux use in the built library:
    cuobjdump -sass neighborretr_b200/libnrhead.so | python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

WANT = ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "SYNCS", "REDUX", "MUFU")
cur, cnt = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cnt[cur]["_total"] += 1
        for w in WANT:
            if op.startswith(w):
                cnt[cur][w] += 1
names = list(cnt)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass neighborretr_b200/libnrhead.so (sm_100a): per-kernel instruction counts")
print("# UTCHMMA = tcgen05.mma (bf16 -> fp32 TMEM), UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA load),")
print("# UTMAPF = tensor-map prefetch, SYNCS = mbarrier ops, REDUX = redux.sync, MUFU = special-function unit (exp / log / rcp)")
rows = []
for k, d in zip(names, dem):
    c = cnt[k]
    if c["UTCHMMA"] + c["LDTM"] + c["UTMALDG"] + c["REDUX"] == 0:
        continue
    d = re.sub(r"\(.*", "", d).replace("void ", "")
    rows.append((d[:70], c))
for name, c in sorted(rows):
    print(f"{name:72s} total {c['_total']:6d} | " + " ".join(f"{w}={c[w]}" for w in WANT if c[w]))
