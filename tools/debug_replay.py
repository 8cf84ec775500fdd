import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import ops, selfcheck, synth
from neighborretr_b200.graph import FIELDS, GraphedHeadStep
dev = torch.device("cuda", 0)
nt, nv, M, b = 24, 12, 64, 32
for prec, explicit, own in (("bf16", False, True), ("bf16", True, True), ("bf16", True, False), ("bf16x3", True, True)):
    ops.USE_OWN_GEMM = own
    bank = synth.make_bank(M, nt, nv)
    model = selfcheck.make_model(synth.default_config(), dev, prec)
    selfcheck.set_bank(model, bank, dev)
    h = synth.make_batch(b, nt, nv, seed=5).to(dev)
    batch = [getattr(h, f) for f in FIELDS]
    g = GraphedHeadStep(model, batch, explicit_grads=explicit)
    outs = []
    for i in range(3):
        g.set_bank(bank)
        g(*batch)
        torch.cuda.synchronize()
        gr = g.grad_list if explicit else [g.grads[f] for f in ("text_feat", "video_feat", "global_text", "global_video")] + [p.grad for p in g.params if p.grad is not None]
        outs.append([x.clone() for x in gr if x is not None] + [g.losses.clone()])
    for i in (1, 2):
        d = [float((a - b_).abs().max() / b_.abs().max().clamp_min(1e-30)) for a, b_ in zip(outs[i], outs[0])]
        print(prec, "explicit" if explicit else "backward", "own" if own else "lib", f"replay {i} vs 0: max rel diffs", ["%.1e" % x for x in d])
