"""Launch only the fused two-direction max-sim kernel (nr_maxsim2_fwd): the forward contractions of one
MSR-VTT-shaped head step (batch pair + two bank pairs in ONE launch), or a single X x Y problem given as
`rx nx ry ny`.  Target of the `ncu --set full` capture; prints the CUDA-event time per launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import ops, synth  # noqa: E402

d = 512
n = 8


def prep(r, nt, nv, seed):
    h = synth.make_batch(r, nt, nv, d=d, seed=seed).to("cuda")
    T = ops.Prepared(h.text_feat, bf16=True, mask=h.text_mask)
    V = ops.Prepared(h.video_feat, bf16=True, mask=h.video_mask)
    tw = torch.full((r, nt), 1.0 / nt, device="cuda")
    vw = torch.full((r, nv), 1.0 / nv, device="cuda")
    return T, V, tw, vw


if len(sys.argv) >= 3 and sys.argv[1] == "rank":
    # evaluation ranks straight from the accumulator (nr_maxsim2_rank, count pass): q x q MSR-VTT-shaped test set
    q, nt, nv = int(sys.argv[2]), 24, 12
    h = synth.make_batch(q, nt, nv, d=d, seed=7).to("cuda")
    tw = torch.full((q, nt), 1.0 / nt, device="cuda"); vw = torch.full((q, nv), 1.0 / nv, device="cuda")
    r = ops.FusedRanker(h.text_feat, h.video_feat, tw, vw, h.text_mask, h.video_mask, "bf16")
    diag = r.diagonal(q)
    evs = []
    for i in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r._launch(2, diag, r.counts(diag) if i == 0 else cnt); e.record()
        cnt = tuple(torch.zeros_like(c) for c in r.counts(diag)) if i == 0 else cnt
        evs.append((s, e))
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b_) * 1e3 for a, b_ in evs]
    avg = sum(ts[2:]) / len(ts[2:])
    fl = 2.0 * q * nt * q * nv * d
    print(f"nr_maxsim2_rank count pass {q} x {q}: {min(ts):.1f} us best, {avg:.1f} us avg -> {fl / avg / 1e6:.0f} TFLOP/s; "
          f"S ({q * q * 4 / 1e6:.0f} MB) is never written")
    sys.exit(0)
if len(sys.argv) >= 5:
    rx, nx, ry, ny = [int(v) for v in sys.argv[1:5]]
    X, _, wx, _ = prep(rx, nx, ny, 7)
    _, Y, _, wy = prep(ry, nx, ny, 8)
    out = torch.empty(rx, ry, device="cuda")
    probs = [dict(X=X, Y=Y, wx=wx, wy=wy, alpha=0.5, out=out, strides=(ry, 1))]
    fl = 2.0 * rx * nx * ry * ny * d
    name = f"{rx}x{nx} vs {ry}x{ny}"
else:
    b, m, nt, nv = 128, 512, 24, 12
    if len(sys.argv) >= 3:
        b, m = int(sys.argv[1]), int(sys.argv[2])
    T, V, tw, vw = prep(b, nt, nv, 7)
    MT, MV, tw_mb, vw_mb = prep(m, nt, nv, 8)
    S = torch.empty(b, b, device="cuda"); ST = torch.empty(b, b, device="cuda")
    A = torch.empty(b, m, device="cuda"); C = torch.empty(b, m, device="cuda")
    probs = [dict(X=T, Y=MV, wx=tw, wy=vw_mb, alpha=0.5, out=A, strides=(m, 1)),
             dict(X=MT, Y=V, wx=tw_mb, wy=vw, alpha=0.5, out=C, strides=(1, m)),
             dict(X=T, Y=V, wx=tw, wy=vw, alpha=0.5, out=S, strides=(b, 1), out2=ST, strides2=(1, b))]
    fl = 2.0 * nt * nv * d * (b * b + 2 * b * m)
    name = f"msrvtt step b={b} M={m} (3 problems)"
evs = []
for i in range(n):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.maxsim2_fwd(probs)
    e.record()
    evs.append((s, e))
torch.cuda.synchronize()
ts = [a.elapsed_time(b_) * 1e3 for a, b_ in evs]
avg = sum(ts[2:]) / len(ts[2:])
print(f"nr_maxsim2_fwd {name}: {min(ts):.1f} us best, {avg:.1f} us avg -> {fl / avg / 1e6:.0f} TFLOP/s "
      f"(algorithmic: every token pair once, both directions)")
