import os, sys, torch
from types import SimpleNamespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import cuda_losses, make_head, rel_l2, set_bank
from neighborretr_b200 import ops, synth
D, NT, NV, B, M = 512, 24, 12, 32, 64
cfg = synth.default_config(); params = synth.make_mlp_params(d=D)
h = synth.make_batch(B, NT, NV, d=D, seed=501)
for ragged in (False, True):
    hb = synth.make_batch(M, NT, NV, d=D, seed=500, ragged=ragged)
    bank = SimpleNamespace(mb_ind=hb.idx, mb_feat_t=hb.text_feat, mb_feat_v=hb.video_feat, mb_mask_t=hb.text_mask, mb_mask_v=hb.video_mask)
    res = {}
    for prec in ("fp32", "bf16x3", "bf16"):
        for fused in ((True, False) if prec == "bf16" else (True,)):
            m = make_head(D, cfg, params, prec); set_bank(m, bank)
            ops.USE_FUSED_MAXSIM = fused
            losses, grads = cuda_losses(m, h, cfg)
            ops.USE_FUSED_MAXSIM = True
            res[(prec, fused)] = (losses, grads)
    ref = res[("fp32", True)]
    for k, (l, g) in res.items():
        print("ragged bank" if ragged else "full bank  ", k, "loss rel", float((l / ref[0] - 1).abs().max()),
              {n: round(rel_l2(g[n], ref[1][n]), 5) for n in ("text", "video", "text_weight_fc.0.weight", "video_weight_fc.2.weight")})
