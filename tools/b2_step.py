"""The four token-gradient contractions of one MSR-VTT-shaped head step (b=128, M=512) as ONE nr_maxsim2_bwd launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import ops, synth  # noqa: E402

d, nt, nv = 512, 24, 12
b, m = 128, 512
if len(sys.argv) >= 3:
    b, m = int(sys.argv[1]), int(sys.argv[2])


def prep(r, seed):
    h = synth.make_batch(r, nt, nv, d=d, seed=seed).to("cuda")
    return (ops.Prepared(h.text_feat, bf16=True, mask=h.text_mask), ops.Prepared(h.video_feat, bf16=True, mask=h.video_mask),
            torch.full((r, nt), 1.0 / nt, device="cuda"), torch.full((r, nv), 1.0 / nv, device="cuda"))


T, V, tw, vw = prep(b, 7)
MT, MV, tw_mb, vw_mb = prep(m, 8)
S = torch.empty(b, b, device="cuda"); A = torch.empty(b, m, device="cuda"); C = torch.empty(b, m, device="cuda")
svA, svC, sv1 = ops.maxsim2_fwd([dict(X=T, Y=MV, wx=tw, wy=vw_mb, alpha=0.5, out=A, strides=(m, 1)),
                                 dict(X=MT, Y=V, wx=tw_mb, wy=vw, alpha=0.5, out=C, strides=(1, m)),
                                 dict(X=T, Y=V, wx=tw, wy=vw, alpha=0.5, out=S, strides=(b, 1))])
dS = torch.randn(b, b, device="cuda"); dc = torch.randn(2, b, device="cuda")
dtn = torch.zeros_like(T.xn); dvn = torch.zeros_like(V.xn)
for P in (T, V, MT, MV):
    P.bwd_source(1)
jobs = [(0, V, tw, vw, sv1[1], sv1[3], dS, b, 1, 0.5, b, b, dtn), (0, MV, tw, vw_mb, svA[1], svA[3], dc[0], 1, 0, 0.5 / m, b, m, dtn),
        (1, T, tw, vw, sv1[1], sv1[3], dS, b, 1, 0.5, b, b, dvn), (1, MT, tw_mb, vw, svC[1], svC[3], dc[1], 0, 1, 0.5 / m, m, b, dvn)]
evs = []
for i in range(8):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.maxsim2_bwd_multi(jobs, nt, nv, d)
    e.record()
    evs.append((s, e))
torch.cuda.synchronize()
ts = [x.elapsed_time(y) * 1e3 for x, y in evs]
fl = 2.0 * nt * nv * d * (2 * b * b + 2 * b * m)
avg = sum(ts[2:]) / len(ts[2:])
print(f"nr_maxsim2_bwd step b={b} M={m} debug={os.environ.get('NR_B2_DEBUG','0')} per={os.environ.get('NR_B2_PER','auto')}: "
      f"{min(ts):.1f} us best, {avg:.1f} us avg -> {fl / avg / 1e6:.0f} TFLOP/s")
