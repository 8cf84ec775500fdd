"""One launch of each multi-sentence evaluation kernel on a matrix larger than L2 (for `ncu --set full`)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborretr_b200 import ops  # noqa: E402

V, per = 4096, 24
rng = np.random.RandomState(0)
lens = rng.randint(1, 2 * per, size=V)
T = int(lens.sum())
gs = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)).cuda()
tgt = torch.from_numpy(np.repeat(np.arange(V, dtype=np.int32), lens)).cuda()
S = torch.randn(T, V, device="cuda")
gt, eqb, valid = ops.rank_counts_target(S, tgt)
out = ops.group_max_t(S, gs)
torch.cuda.synchronize()
print("captions", T, "videos", V, "bytes", 4 * T * V, int(gt.sum()), float(out[0, 0]))
