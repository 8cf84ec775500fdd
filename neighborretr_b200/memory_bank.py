"""Memory-bank prefill and reset with the reference's manager surface
(reference NeighborRetr/utils/memory_bank.py:22-268; state layout NeighborRetr/models/modeling.py:175-184).

``MemoryBankManager(args)`` keeps the reference's methods and their meaning — ``load_memory_bank(model,
memory_bank_dataloader, device, epoch) -> rows``, ``clear_memory_bank(model) -> model``,
``create_memory_bank_dataloader()`` — and writes the same five plain tensor attributes plus ``mb_batch`` on the
model (``mb_ind, mb_feat_t, mb_feat_v, mb_mask_t, mb_mask_v``), in the same row order: the ``mb_batch`` loader
batches of rank 0, then those of rank 1, ...

What is different is the data movement, which is all this step is besides the (out-of-scope) encoders:
* the reference appends every batch to five Python lists, ``torch.cat``s them, gathers each with the list API
  (one more ``cat``) and empties the allocator cache after every batch; here each batch is written once, at its
  row offset, into a staging buffer sized on the first batch, and every tensor crosses the ranks with ONE
  ``all_gather_into_tensor`` into its final contiguous ``[W*rows, ...]`` buffer (until_module.AllGather);
* the buffers become the FIFO the head updates in place afterwards (``update_memory_bank`` / the step graph).
The encoders run under no_grad and autocast exactly as in the reference (:128-131).
"""
from __future__ import annotations

import torch

from .until_module import AllGather

class _Staging:
    """Five row-major staging buffers filled batch by batch (allocated on the first batch)."""

    def __init__(self, max_batches):
        self.max_batches = max_batches
        self.bufs = None
        self.rows = 0

    def append(self, tensors):
        n = tensors[0].shape[0]
        if self.bufs is None:
            cap = n * self.max_batches
            self.bufs = [torch.empty((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for t in tensors]
        for k, t in enumerate(tensors):
            buf = self.bufs[k]
            if tuple(buf.shape[1:]) != tuple(t.shape[1:]) or t.shape[0] != n:
                raise ValueError(f"memory bank: batch tensor {tuple(t.shape)} does not match the first batch "
                                 f"{tuple(buf.shape)}")
            if self.rows + n > buf.shape[0]:          # a later batch larger than the first one: grow once more
                grown = torch.empty((self.rows + n * self.max_batches,) + tuple(buf.shape[1:]), dtype=buf.dtype,
                                    device=buf.device)
                grown[:self.rows].copy_(buf[:self.rows])
                self.bufs[k] = buf = grown
            buf[self.rows:self.rows + n].copy_(t)
        self.rows += n

    def result(self):
        return [b[:self.rows] for b in self.bufs]


class MemoryBankManager:
    """Drop-in for the reference class (utils/memory_bank.py:22)."""

    def __init__(self, args):
        self.args = args
        self.logger = getattr(args, "logger", None)
        self.mb_batch = getattr(args, "mb_batch", 10)             # reference :45
        self.batch_size = getattr(args, "batch_size", None)
        self.memory_bank_dataloader = None

    # -- logging helpers (the reference logs through args.logger unconditionally) -------------------------------
    def _info(self, msg):
        if self.logger is not None:
            self.logger.info(msg)

    def _error(self, msg):
        if self.logger is not None:
            self.logger.error(msg)

    def create_memory_bank_dataloader(self):
        """Reference :49-79: a second *train* dataloader built by the reference's own dataloader factory.  The input
        pipeline (datasets, video decoding, tokenizer) is outside the retrieval head, so this delegates to the
        reference package when it is importable and fails loudly otherwise."""
        try:
            from argparse import Namespace
            from NeighborRetr.dataloaders.data_dataloaders import DATALOADER_DICT
            from NeighborRetr.models.tokenization_clip import SimpleTokenizer as ClipTokenizer
        except ImportError as e:
            raise RuntimeError("create_memory_bank_dataloader needs the reference's dataloaders "
                               "(NeighborRetr.dataloaders) on sys.path; pass a dataloader to load_memory_bank "
                               f"instead ({e})") from e
        entry = DATALOADER_DICT.get(getattr(self.args, "datatype", None))
        if entry is None or entry["train"] is None:
            self._error(f"Cannot create memory bank dataloader: datatype {getattr(self.args, 'datatype', None)} "
                        "not found")
            return None
        loader, _, _ = entry["train"](Namespace(**vars(self.args)), ClipTokenizer())
        self._info(f"Created memory bank dataloader with batch size {self.batch_size}; up to {self.mb_batch} batches")
        self.memory_bank_dataloader = loader
        return loader

    def load_memory_bank(self, model, memory_bank_dataloader, device, epoch):
        """Reference :80-229.  Runs ``model.get_text_video_feat`` on the first ``min(mb_batch, len(loader))`` batches
        under no_grad (+ autocast on CUDA), gathers across ranks when ``args.distributed``, assigns the bank on the
        model and returns its number of rows (0 if nothing was processed)."""
        target = model.module if hasattr(model, "module") else model
        target = target.to(device)
        target.eval()
        if memory_bank_dataloader is None:
            memory_bank_dataloader = self.memory_bank_dataloader or self.create_memory_bank_dataloader()
            if memory_bank_dataloader is None:
                self._error("Failed to create memory bank dataloader")
                return 0
        n_batches = min(self.mb_batch, len(memory_bank_dataloader))
        self._info(f"Memory bank loading (epoch {epoch}): {n_batches} of {len(memory_bank_dataloader)} batches")
        if n_batches <= 0:
            return 0
        staging = _Staging(n_batches)
        on_cuda = torch.cuda.is_available()
        with torch.no_grad(), torch.autocast("cuda", enabled=on_cuda):
            for batch_idx, batch in enumerate(memory_bank_dataloader):
                if batch_idx >= n_batches:
                    break
                text_ids, text_mask, video, video_mask, indices, _ = (t.to(device=device, non_blocking=True)
                                                                      for t in batch)
                text_feat, video_feat = target.get_text_video_feat(text_ids, text_mask, video, video_mask)
                staging.append((indices, text_feat, text_mask, video_feat, video_mask))
        if staging.rows == 0:
            return 0
        ind, feat_t, mask_t, feat_v, mask_v = staging.result()
        ind = ind.squeeze()                                        # reference :176 ([n,1] loader indices -> [n])
        if (getattr(self.args, "distributed", False) and getattr(self.args, "world_size", 1) > 1
                and torch.distributed.is_available() and torch.distributed.is_initialized()):
            ind, feat_t, mask_t, feat_v, mask_v = (AllGather.apply(t.contiguous(), self.args)
                                                   for t in (ind, feat_t, mask_t, feat_v, mask_v))
            ind = ind.squeeze()
        target.mb_ind, target.mb_feat_t, target.mb_mask_t = ind, feat_t, mask_t
        target.mb_feat_v, target.mb_mask_v = feat_v, mask_v
        target.mb_batch = feat_t.size(0)
        nbytes = sum(t.numel() * t.element_size() for t in (ind, feat_t, mask_t, feat_v, mask_v))
        self._info(f"Memory bank size: {target.mb_batch} samples, {nbytes / 2**30:.3f} GB; text "
                   f"{tuple(feat_t.shape)}, video {tuple(feat_v.shape)}")
        return target.mb_batch

    def clear_memory_bank(self, model):
        """Reference :231-268: back to the empty state of modeling.py:175-184 (same shapes and dtypes)."""
        target = model.module if hasattr(model, "module") else model
        if getattr(target, "mb_batch", 0) > 0:
            self._info(f"Clearing memory bank with {target.mb_batch} samples")
        device = next(target.parameters()).device
        target.mb_ind = torch.tensor([], dtype=torch.long, device=device)
        target.mb_feat_t = torch.empty((0, 0, 0), dtype=torch.float, device=device)
        target.mb_feat_v = torch.empty((0, 0, 0), dtype=torch.float, device=device)
        target.mb_mask_t = torch.empty((0, 0), dtype=torch.float, device=device)
        target.mb_mask_v = torch.empty((0, 0), dtype=torch.float, device=device)
        target.mb_batch = 0
        self._info("Memory bank cleared")
        return model
