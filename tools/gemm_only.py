"""The token-weight MLP GEMMs of one MSR-VTT-shaped head step on the tcgen05 kernel (csrc/gemm_tc.cu), alone:
forward pair (text 15360 + video 7680 tokens, 512 -> 1024, bias + ReLU + second-layer dot) and backward pair (dW1 and
dx of both).  CUDA-event time per launch and TFLOP/s; target of the ncu capture."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import _lib, ops  # noqa: E402
from neighborretr_b200.ops import _call, _stream  # noqa: E402

D, H = 512, 1024
g = torch.Generator(device="cuda").manual_seed(0)
sides = []
for T, Ta in ((15360, 3072), (7680, 1536)):
    bf = lambda *s: torch.randn(*s, generator=g, device="cuda").to(torch.bfloat16)
    sides.append(dict(T=T, Ta=Ta, x=bf(T, D), w1=bf(H, D) * 0.05, b1=torch.zeros(H, device="cuda"), w2=torch.randn(H, device="cuda") * 0.05,
                      h=torch.empty(T, H, dtype=torch.bfloat16, device="cuda"), logits=torch.zeros(T, device="cuda"), dh=bf(T, H),
                      dw1=torch.zeros(H, D, device="cuda"), dx=torch.zeros(Ta, D, device="cuda")))
arr = (_lib.MlpSide * 2)()
for i, s in enumerate(sides):
    a = arr[i]
    a.x_bf16, a.w1_bf16, a.T, a.T_dx = s["x"].data_ptr(), s["w1"].data_ptr(), s["T"], s["Ta"]
    a.b1, a.w2, a.h_bf16, a.logits = s["b1"].data_ptr(), s["w2"].data_ptr(), s["h"].data_ptr(), s["logits"].data_ptr()
    a.dh_bf16, a.dw1, a.dx = s["dh"].data_ptr(), s["dw1"].data_ptr(), s["dx"].data_ptr()
p = ctypes.cast(arr, ctypes.c_void_p)
fl_f = sum(2.0 * s["T"] * D * H for s in sides)
fl_b = sum(2.0 * s["T"] * D * H + 2.0 * s["Ta"] * D * H for s in sides)
for name, fn, fl in (("nr_mlp_fwd_pair", lambda: _call("nr_mlp_fwd_pair", p, 2, D, H, 0, _stream()), fl_f),
                     ("nr_mlp_bwd_pair", lambda: _call("nr_mlp_bwd_pair", p, 2, D, H, _stream()), fl_b)):
    ts = []
    for i in range(8):
        torch.cuda._sleep(200000)
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(); fn(); b_.record()
        torch.cuda.synchronize()
        ts.append(a_.elapsed_time(b_) * 1e3)
    t = sorted(ts[2:])[len(ts[2:]) // 2]
    print(f"{name}: {t:.1f} us -> {fl / t / 1e6:.0f} TFLOP/s ({fl / 1e9:.1f} GFLOP)")
