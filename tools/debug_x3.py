import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import ops, synth
b, nt, nv, d = 40, 24, 12, 512
h = synth.make_batch(b, nt, nv, d=d, seed=91).to("cuda")
for prec, roles in (("x3", (ops.ROLE_X, ops.ROLE_Y)), ("bf16", (0, 0))):
    T = ops.Prepared(h.text_feat, bf16=True, mask=h.text_mask, split=roles[0])
    V = ops.Prepared(h.video_feat, bf16=True, mask=h.video_mask, split=roles[1])
    tw = torch.full((b, nt), 1.0 / nt, device="cuda"); vw = torch.full((b, nv), 1.0 / nv, device="cuda")
    S = torch.empty(b, b, device="cuda")
    (pmx, yst, pmy, xst), = ops.maxsim2_fwd([dict(X=T, Y=V, wx=tw, wy=vw, alpha=0.5, out=S, strides=(b, 1))])
    t = torch.nn.functional.normalize(h.text_feat.double(), dim=-1) * h.text_mask[..., None]
    v = torch.nn.functional.normalize(h.video_feat.double(), dim=-1) * h.video_mask[..., None]
    r = torch.einsum("atd,bvd->abtv", t, v)                  # [A,B,Nt,Nv]
    cmax, carg = r.max(dim=2)                                # over t: [A,B,Nv]
    rmax, rarg = r.max(dim=3)                                # over v: [A,B,Nt]
    xs = xst.long(); ys = yst.long()
    got_c = torch.gather(r, 2, xs.unsqueeze(2)).squeeze(2)   # value at the kernel's arg
    got_r = torch.gather(r, 3, ys.unsqueeze(3)).squeeze(3)
    gap_c = (cmax - got_c); gap_r = (rmax - got_r)
    print(prec, "col: mismatched args", int((xs != carg).sum()), "of", xs.numel(), "max gap", float(gap_c.max()),
          "n gap>1e-6", int((gap_c > 1e-6).sum()), "| row: mismatched", int((ys != rarg).sum()), "max gap", float(gap_r.max()),
          "n gap>1e-6", int((gap_r > 1e-6).sum()))
    bad = (gap_c > 1e-6).nonzero()[:5]
    for a_, b_, y_ in bad.tolist():
        col = r[a_, b_, :, y_]
        print("   pair", a_, b_, "y", y_, "kernel x", int(xs[a_, b_, y_]), "true x", int(carg[a_, b_, y_]), "pmax_y", float(pmy[a_, b_, y_]),
              "vals", [round(float(z), 5) for z in col.tolist()], "tmask", h.text_mask[a_].tolist())
