"""Kernel timeline of ONE replay of the whole-step CUDA graph (torch.profiler / CUPTI activity records; nsys is
not in the image): start offset, duration, stream and name of every kernel, plus the busy/critical summary.
    python tools/trace_step.py [--shape msrvtt] [--b 128] [--out gpurun_out/trace.txt]"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import synth  # noqa: E402
from neighborretr_b200.graph import FIELDS, GraphedHeadStep  # noqa: E402
from neighborretr_b200.modeling import NeighborRetr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="msrvtt")
ap.add_argument("--b", type=int, default=128)
ap.add_argument("--out", default=None)
a = ap.parse_args()
nt, nv, mrows = synth.SHAPES[a.shape]
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:      # torchrun --nproc-per-node W tools/trace_step.py: the row-block sharded step with captured NCCL
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
m = NeighborRetr(synth.default_config(world_size=world, local_rank=local, rank=rank), width=512)
for n, sd in synth.make_mlp_params().items():
    getattr(m, n).load_state_dict(sd)
m.clip.logit_scale.data.fill_(4.6052)
m = m.to(dev).train()
bank = synth.make_bank(mrows, nt, nv)
for n in ("mb_ind", "mb_feat_t", "mb_feat_v", "mb_mask_t", "mb_mask_v"):
    setattr(m, n, getattr(bank, n).to(dev))
h = synth.make_batch(a.b, nt, nv, seed=1234, rank=rank).to(dev)
batch = [getattr(h, f) for f in FIELDS]
g = GraphedHeadStep(m, batch)
for _ in range(3):
    g(*batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g(*batch)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
end = max(e.time_range.end for e in evs)
lines = [f"# one graph replay, {a.shape} b={a.b}: {len(evs)} device activities, span {(end - t0):.1f} us"]
busy = 0.0
cur_end = t0
for e in evs:
    s, t = e.time_range.start, e.time_range.end
    if t > cur_end:
        busy += t - max(s, cur_end)
        cur_end = t
    lines.append(f"{s - t0:9.1f} {t - s:8.1f}  {e.name[:110]}")
lines.append(f"# GPU busy (union of kernel intervals) {busy:.1f} us of {end - t0:.1f} us")
txt = "\n".join(lines)
if rank != 0:
    torch.cuda.synchronize()
    os._exit(0)
if a.out:
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    open(a.out, "w").write(txt + "\n")
print(txt)
if world > 1:
    sys.stdout.flush()
    os._exit(0)
