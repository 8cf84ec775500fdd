"""Evaluation similarity matrix: _run_on_single_gpu with the reference's signature and return value
(reference NeighborRetr/training/evaluator.py:21-63).

The reference walks 64x64 tiles, runs local_level (including the token-weight MLPs) per tile pair and copies
every tile to the host.  Tiling does not change values, so here the token weights are
evaluated once per modality, the whole [Nq,Ng] matrix comes from one pair of max-sim launches, and there is a
single device->host copy.  ``mini_batch`` is accepted for signature compatibility and bounds the MLP batch.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .modeling import _token_weights


def similarity_matrix(model, t_mask_list, v_mask_list, t_feat_list, v_feat_list, mini_batch=64):
    """CUDA tensor [Nq, Ng] of local_level similarities."""
    with torch.no_grad():
        t_mask = t_mask_list.view(-1, t_mask_list.shape[-1])
        v_mask = v_mask_list.view(-1, v_mask_list.shape[-1])
        chunk = max(int(mini_batch), 1) * 64
        prec = model._head_precision() if hasattr(model, "_head_precision") else "fp32"
        lowp = prec == "bf16"
        tw = torch.cat([_token_weights(model.text_weight_fc, f, m, lowp)
                        for f, m in zip(torch.split(t_feat_list, chunk), torch.split(t_mask, chunk))])
        vw = torch.cat([_token_weights(model.video_weight_fc, f, m, lowp)
                        for f, m in zip(torch.split(v_feat_list, chunk), torch.split(v_mask, chunk))])
        s, _ = ops.maxsim(t_feat_list, v_feat_list, tw, vw, t_mask, v_mask, prec)
    return s


def _run_on_single_gpu(model, t_mask_list, v_mask_list, t_feat_list, v_feat_list, mini_batch=64):
    s = similarity_matrix(model, t_mask_list, v_mask_list, t_feat_list, v_feat_list, mini_batch)
    sim_matrix = s.cpu().numpy()
    return sim_matrix, sim_matrix.T
