"""Launch only the dominant kernel (nr_maxsim_fwd, text x bank-video block of the MSR-VTT-shaped step) a few times:
the target of the `ncu --set full` capture.  Prints the CUDA-event time per launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import ops, synth  # noqa: E402
from neighborretr_b200._lib import NR_PREC_BF16  # noqa: E402

rx, nx, ry, ny, d = 128, 24, 512, 12, 512
if len(sys.argv) > 1:
    rx, nx, ry, ny = [int(v) for v in sys.argv[1:5]]
hx = synth.make_batch(rx, nx, ny, d=d, seed=7).to("cuda")
hy = synth.make_batch(ry, nx, ny, d=d, seed=8).to("cuda")
X = ops.Prepared(hx.text_feat, bf16=True)
Y = ops.Prepared(hy.video_feat, bf16=True)
wx = torch.full((rx, nx), 1.0 / nx, device="cuda")
out = torch.empty(rx, ry, device="cuda")
n = 6
evs = []
for i in range(n):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops._maxsim_dir_fwd(NR_PREC_BF16, X, Y, wx, hx.text_mask, hy.video_mask, 1.0, out, ry, 1, None, 0, 0, 0, True)
    e.record()
    evs.append((s, e))
torch.cuda.synchronize()
ts = [a.elapsed_time(b) * 1e3 for a, b in evs]
fl = 2.0 * rx * nx * ry * ny * d
print(f"nr_maxsim_fwd {rx}x{nx} vs {ry}x{ny}: {min(ts):.1f} us best, {sum(ts[2:])/len(ts[2:]):.1f} us avg -> "
      f"{fl/ (sum(ts[2:])/len(ts[2:])) / 1e6:.0f} TFLOP/s (one orientation)")
