"""Per-entry-point CUDA-event breakdown of one eager head step (no profiler needed):
    python tools/profile_step.py [--shape msrvtt] [--steps 5] [--b 128]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import ops, synth  # noqa: E402
from neighborretr_b200.modeling import NeighborRetr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="msrvtt")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--b", type=int, default=128)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
nt, nv, mrows = synth.SHAPES[a.shape]
dev = torch.device("cuda")
cfg = synth.default_config()
m = NeighborRetr(cfg, width=512)
for n, sd in synth.make_mlp_params().items():
    getattr(m, n).load_state_dict(sd)
m.clip.logit_scale.data.fill_(4.6052)
m.head_precision = a.precision
m = m.to(dev).train()
bank = synth.make_bank(mrows, nt, nv)
for n in ("mb_ind", "mb_feat_t", "mb_feat_v", "mb_mask_t", "mb_mask_v"):
    setattr(m, n, getattr(bank, n).to(dev))
h = synth.make_batch(a.b, nt, nv).to(dev)


def step():
    t = h.text_feat.clone().requires_grad_(True); v = h.video_feat.clone().requires_grad_(True)
    gt = h.global_text.clone().requires_grad_(True); gv = h.global_video.clone().requires_grad_(True)
    m.zero_grad(set_to_none=True)
    l = m.head_forward(t, v, h.text_mask, h.video_mask, h.idx, global_feats=(gt, gv))
    l[0].backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
ops.KERNEL_TIMER.enable("*")
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(a.steps):
    step()
e.record()
tab = ops.KERNEL_TIMER.table()
tot = s.elapsed_time(e) / a.steps
print(f"eager step {tot*1e3:.0f} us  ({a.shape}, b={a.b}, {a.precision})")
acc = 0.0
for n, (c, t) in sorted(tab.items(), key=lambda kv: -kv[1][1]):
    print(f"  {t/a.steps*1e3:9.1f} us  {c//a.steps:3d} calls  {n}")
    acc += t / a.steps
print(f"  sum of C-ABI entry points {acc*1e3:.0f} us (event-to-event, includes launch gaps)")
