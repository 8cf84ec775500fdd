"""The memory-bound kernels of the head at the sizes where they leave L2 (BASELINE.json configs[4]: global batch 8192,
100 000-video gallery): CUDA-event time per launch, algorithmic bytes and achieved GB/s against the
measured HBM copy bandwidth.  Target of the ncu pass that records dram__bytes_read/write per kernel
(tools/gpu_r2_profiles.sh -> profiles/r2_membound_*.txt)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import ops, synth  # noqa: E402
from neighborretr_b200._lib import NR_LOSS_CENTRALITY, NR_LOSS_KL, NR_LOSS_NEIGHBOR, NR_LOSS_UNIFORM, NR_NSAVE  # noqa: E402
from neighborretr_b200.ops import _call, _p, _stream  # noqa: E402

ALL = NR_LOSS_CENTRALITY | NR_LOSS_NEIGHBOR | NR_LOSS_KL | NR_LOSS_UNIFORM
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
dev = "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
reps = int(os.environ.get("REPS", "5"))


def timed(name, nbytes, fn):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    gbs = nbytes / t / 1e3
    print(f"{name:44s} {t:9.1f} us  {nbytes / 1e6:9.1f} MB algorithmic  {gbs:7.0f} GB/s  {gbs / PEAK:5.2f} of HBM copy peak", flush=True)


g = torch.Generator(device=dev).manual_seed(1)
B, k = 8192, 20
S = torch.randn(B, B, generator=g, device=dev) * 0.1
G = torch.randn(B, B, generator=g, device=dev)
cb = torch.rand(B, generator=g, device=dev); w = torch.rand(B, generator=g, device=dev) + 0.5
ls = torch.tensor([100.0], device=dev)
GT = G.t().contiguous()
duals = torch.empty(4, B, device=dev)
ws, nws = ops.sinkhorn_workspace(B, torch.device(dev))
# Sinkhorn: 2 chains x 2 half-iterations x iters passes over a [B,B] matrix
it = 5
timed(f"nr_sinkhorn B={B} ({it} iterations)", 4.0 * B * B * 4 * it,
      lambda: _call("nr_sinkhorn", _p(G), _p(GT), B, it, _p(duals[0]), _p(duals[1]), _p(duals[2]), _p(duals[3]), _p(ws), nws, _stream()))
row_out = torch.zeros(4, B, device=dev); nbr = torch.empty(B, k, dtype=torch.int32, device=dev)
saved = torch.empty(B, NR_NSAVE, device=dev)
timed(f"nr_row_losses_fwd B={B} (4 losses, S and G)", 2.0 * B * B * 4,
      lambda: _call("nr_row_losses_fwd", _p(S), B, _p(G), B, _p(cb), _p(w), _p(duals[0]), _p(duals[1]), B, B, 0, _p(ls), k, 3.0,
                    3.0, 0.7, ALL, _p(row_out), _p(nbr), _p(saved), _stream()))
gscale = torch.ones(4, device=dev); dS = torch.empty(B, B, device=dev); dG = torch.empty(B, B, device=dev)
dc = torch.zeros(B, device=dev); dw = torch.zeros(B, device=dev); dls = torch.zeros(1, device=dev)
timed(f"nr_row_losses_bwd B={B} (reads S,G; writes dS,dG)", 4.0 * B * B * 4,
      lambda: _call("nr_row_losses_bwd", _p(S), B, _p(G), B, _p(cb), _p(w), _p(duals[0]), _p(duals[1]), B, B, 0, _p(ls), k, 3.0,
                    3.0, 0.7, ALL, _p(nbr), _p(saved), _p(gscale), _p(dS), B, _p(dG), B, _p(dc), _p(dw), _p(dls), _stream()))
out = torch.empty(B, B, device=dev)
timed(f"nr_transpose_add B={B} (dS + dS2^T)", 3.0 * B * B * 4,
      lambda: _call("nr_transpose_add", _p(dS), B, _p(dG), B, _p(out), B, B, B, 1.0, 1.0, _stream()))
del S, G, GT, dS, dG, out
# token preparation of a global batch (text tokens): read fp32, write fp32 + bf16
R, N, D = 8192, 24, 512
x = torch.randn(R, N, D, generator=g, device=dev)
mask = torch.ones(R, N, dtype=torch.int64, device=dev)
P2 = ops.Prepared(x, bf16=True, colsum=True, mask=mask)


def reprep():
    _call("nr_prep_tokens", _p(x), R * N, D, _p(P2.xn), _p(P2.xn_bf16), _p(P2.inv_norm), _p(P2.partials), _p(mask), _stream())


timed(f"nr_prep_tokens {R * N} x {D} (f32 in; f32 + bf16 out)", R * N * D * 10.0, reprep)
dxn = torch.randn(R * N, D, generator=g, device=dev); dx = torch.empty_like(x)
timed(f"nr_prep_tokens_bwd {R * N} x {D}", R * N * D * 12.0, lambda: P2.backward(dxn, out=dx))
P2.alloc_transposed()


def retrans():
    _call("nr_transpose_tokens_bf16", _p(P2.xn_bf16), R * N, D, _p(P2.xnT_bf16), P2.xnT_bf16.shape[1], _stream())


timed(f"nr_transpose_tokens_bf16 {R * N} x {D}", R * N * D * 4.0, retrans)
xb = torch.empty(R * N, D, dtype=torch.bfloat16, device=dev)
timed(f"nr_cast_bf16 {R * N} x {D}", R * N * D * 6.0, lambda: _call("nr_cast_bf16", _p(x), _p(xb), R * N * D, _stream()))
del x, dxn, dx, P2, xb
# evaluation ranking on a 100 000-video gallery
Q, Ng = 8192, 100_000
S = torch.randn(Q, Ng, generator=g, device=dev)
timed(f"nr_rank_count {Q} x {Ng}", Q * Ng * 4.0, lambda: ops.rank_counts(S))
Ssh = S[:, :12500].contiguous()              # one of 8 column shards (the top-k kernel keeps a row in shared memory)
timed(f"nr_topk_rows k=10 {Q} x 12500 (shard)", Q * 12500 * 4.0, lambda: ops.topk_rows(Ssh, 10))
new = torch.randn(1024, 24, 512, generator=g, device=dev); old = torch.randn(1920, 24, 512, generator=g, device=dev)
timed("nr_fifo_update 1024 new + 1920 bank rows (text)", 2.0 * 1920 * 24 * 512 * 4, lambda: ops.fifo_update(new, old, 1920))
