import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator().manual_seed(0)
a = 3 * torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1)
b = 3 * torch.nn.functional.normalize(a + 0.3 * torch.randn(B, 512, generator=g), dim=-1)
G = (a @ b.t()).cuda(); GT = G.t().contiguous()
for _ in range(3):
    ops.sinkhorn_duals(G, GT)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    r = ops.sinkhorn_duals(G, GT)
e.record(); torch.cuda.synchronize()
print(os.environ.get("NR_SINKHORN_VARIANT", "default"), "B", B, f"{s.elapsed_time(e)/10*1e3:.1f} us", float(r[0].sum()), float(r[3].sum()))
