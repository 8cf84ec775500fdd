"""Per-CTA timeline of nr_maxsim2_fwd (NR_TC2_TRACE=1): globaltimer stamps written by the kernel into its workspace.
Prints, relative to the earliest CTA start, the distribution over CTAs of each stamp (ns)."""
import os
import sys

os.environ["NR_TC2_TRACE"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import ops, synth  # noqa: E402

d = 512
b, m, nt, nv = 128, 512, 24, 12
if len(sys.argv) >= 3:
    b, m = int(sys.argv[1]), int(sys.argv[2])


def prep(r, seed):
    h = synth.make_batch(r, nt, nv, d=d, seed=seed).to("cuda")
    return (ops.Prepared(h.text_feat, bf16=True, mask=h.text_mask), ops.Prepared(h.video_feat, bf16=True, mask=h.video_mask),
            torch.full((r, nt), 1.0 / nt, device="cuda"), torch.full((r, nv), 1.0 / nv, device="cuda"))


T, V, tw, vw = prep(b, 7)
MT, MV, tw_mb, vw_mb = prep(m, 8)
S = torch.empty(b, b, device="cuda"); ST = torch.empty(b, b, device="cuda")
A = torch.empty(b, m, device="cuda"); C = torch.empty(b, m, device="cuda")
probs = [dict(X=T, Y=MV, wx=tw, wy=vw_mb, alpha=0.5, out=A, strides=(m, 1)),
         dict(X=MT, Y=V, wx=tw_mb, wy=vw, alpha=0.5, out=C, strides=(1, m)),
         dict(X=T, Y=V, wx=tw, wy=vw, alpha=0.5, out=S, strides=(b, 1), out2=ST, strides2=(1, b))]
for _ in range(3):
    ops.maxsim2_fwd(probs)
torch.cuda.synchronize()
ws = ops._tile_workspace(torch.device("cuda", 0))
ws[4:].zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.maxsim2_fwd(probs); e1.record()
torch.cuda.synchronize()
tr = ws[4:].view(torch.int64).view(-1, 16)[:148].cpu()
t0 = int(tr[:, 0][tr[:, 0] > 0].min())
names = {0: "setup done", 1: "first claim", 2: "claimed end marker", 4: "first operands landed", 5: "first tile issued",
         6: "last tile issued", 8: "set0 first acc ready", 9: "set0 last tile done", 10: "set1 first acc ready",
         11: "set1 last tile done", 12: "CTA done"}
print(f"event time {e0.elapsed_time(e1) * 1e3:.1f} us; tiles per CTA: min {int(tr[:, 3].min())} max {int(tr[:, 3].max())}")
for k, nme in names.items():
    v = (tr[:, k] - t0).float() / 1e3
    v = v[tr[:, k] > 0]
    print(f"{nme:24s} min {v.min():7.2f}  mean {v.mean():7.2f}  max {v.max():7.2f} us")
