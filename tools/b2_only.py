"""Launch only the fused backward contraction (nr_maxsim2_bwd) for one (X, Y) pair: `side rx nx ry ny`
(tools/b2_step.py runs the four contractions of a whole step in one launch)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import ops, synth  # noqa: E402

side, rx, nx, ry, ny = 0, 128, 24, 512, 12
if len(sys.argv) >= 6:
    side, rx, nx, ry, ny = [int(v) for v in sys.argv[1:6]]
d = 512
hx = synth.make_batch(rx, nx, ny, d=d, seed=7).to("cuda")
hy = synth.make_batch(ry, nx, ny, d=d, seed=8).to("cuda")
X = ops.Prepared(hx.text_feat, bf16=True, mask=hx.text_mask)
Y = ops.Prepared(hy.video_feat, bf16=True, mask=hy.video_mask)
wx = torch.full((rx, nx), 1.0 / nx, device="cuda"); wy = torch.full((ry, ny), 1.0 / ny, device="cuda")
out = torch.empty(rx, ry, device="cuda")
p_x, y_s, p_y, x_s = ops.maxsim2_fwd([dict(X=X, Y=Y, wx=wx, wy=wy, alpha=0.5, out=out, strides=(ry, 1))])[0]
g = torch.randn(rx, ry, device="cuda")
dst = torch.zeros_like(X.xn if side == 0 else Y.xn)
src = Y if side == 0 else X
src.bwd_source(1)
evs = []
for i in range(8):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.maxsim2_bwd(side, src, wx, wy, y_s, x_s, g, ry, 1, 0.5, rx, nx, ry, ny, d, dst)
    e.record()
    evs.append((s, e))
torch.cuda.synchronize()
ts = [a.elapsed_time(b) * 1e3 for a, b in evs]
fl = 2.0 * rx * nx * ry * ny * d
avg = sum(ts[2:]) / len(ts[2:])
print(f"nr_maxsim2_bwd side {side} {rx}x{nx} vs {ry}x{ny} debug={os.environ.get('NR_B2_DEBUG','0')} ks={os.environ.get('NR_B2_KS','auto')}: "
      f"{min(ts):.1f} us best, {avg:.1f} us avg -> {fl / avg / 1e6:.0f} TFLOP/s")
