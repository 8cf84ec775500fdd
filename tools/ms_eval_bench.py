"""Multi-sentence evaluation kernels (csrc/multisent.cu) against the HBM roofline: nr_rank_count_target and
nr_group_max_t each read the caption x video matrix once (4*T*V bytes).  Matrices larger than L2 (126 MB) so every
launch streams from HBM; MSVD-sized (27763 x 670, 74 MB) reported too.  Usage: python tools/ms_eval_bench.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import ops  # noqa: E402

PEAK = 6537.6
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


for name, V, per in (("msvd", 670, 41), ("large", 4096, 24), ("large_unaligned", 4097, 24)):
    rng = np.random.RandomState(0)
    lens = rng.randint(1, 2 * per, size=V)
    T = int(lens.sum())
    gs = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)).cuda()
    tgt = torch.from_numpy(np.repeat(np.arange(V, dtype=np.int32), lens)).cuda()
    S = torch.randn(T, V, device="cuda")
    gt = torch.zeros(T, dtype=torch.int32, device="cuda")
    eq = torch.zeros(T, dtype=torch.int32, device="cuda")
    t_rank = timed(lambda: ops.rank_counts_target(S, tgt, gt=gt, eq_before=eq, want_valid=False))
    t_gmax = timed(lambda: ops.group_max_t(S, gs))
    nbytes = 4.0 * T * V
    print(json.dumps({"case": name, "captions": T, "videos": V, "matrix_MB": round(nbytes / 1e6, 1),
                      "rank_count_target_us": round(t_rank * 1e3, 1), "rank_GBps": round(nbytes / t_rank / 1e6, 1),
                      "rank_frac_hbm": round(nbytes / t_rank / 1e6 / PEAK, 3),
                      "group_max_t_us": round(t_gmax * 1e3, 1),
                      "gmax_GBps": round((nbytes + 4.0 * V * V) / t_gmax / 1e6, 1),
                      "gmax_frac_hbm": round((nbytes + 4.0 * V * V) / t_gmax / 1e6 / PEAK, 3), "hbm_peak_GBps": PEAK}))
