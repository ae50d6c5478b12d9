"""Batched offline teacher sampling.

The reference produces its rectified-flow training set one utterance at a time
(`/root/reference/component/binarizer/svs.py:193-223`: condition -> ``x_T = randn`` -> ``x_0 = teacher.diffusion(...)`` ->
item dict) and stores items with ``IndexedDatasetBuilder`` (`/root/reference/utils/indexed_datasets.py:51-76`).  Here the
same items are produced from length-bucketed, padded batches: every batch carries its per-utterance lengths down to the
kernels (``pd_plan_set_lengths``), so each row is computed exactly as that utterance alone -- batching changes
throughput, not results.  Buckets are dealt round-robin to the ranks of a ``torch.distributed`` job (no collective).

The on-disk format is the reference's: ``<prefix>_<segment>.data`` = concatenated ``pickle.dumps(item)``,
``<prefix>_<segment>.idx`` = ``np.save({'offsets': [...]})``, ``segment_size`` items per segment.
"""
from __future__ import annotations

import os
import pickle
from copy import deepcopy
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


# ---------------------------------------------------------------------------------------------- on-disk format
class IndexedDatasetBuilder:
    """Writer compatible with utils/indexed_datasets.py:51-76 (same file names, same bytes)."""

    def __init__(self, path, prefix, segment_size=1024):
        self.path, self.prefix, self.segment_size = path, prefix, segment_size
        self.segment_idx, self.segment_item_count = 0, 0
        self.byte_offsets = [0]
        self.out_file = open(self._name("data"), "wb")

    def _name(self, ext):
        return os.path.join(self.path, f"{self.prefix}_{self.segment_idx}.{ext}")

    def add_item(self, item):
        n = self.out_file.write(pickle.dumps(item))
        self.byte_offsets.append(self.byte_offsets[-1] + n)
        self.segment_item_count += 1
        if self.segment_item_count >= self.segment_size:     # segment full: close it and open the next (:64-69)
            self.finalize()
            self.segment_idx += 1
            self.segment_item_count = 0
            self.out_file = open(self._name("data"), "wb")
            self.byte_offsets = [0]

    def finalize(self):
        self.out_file.close()
        with open(self._name("idx"), "wb") as f:
            np.save(f, {"offsets": self.byte_offsets})


class IndexedDataset:
    """Reader compatible with utils/indexed_datasets.py:7-49.  ``len()`` is the number of items (the reference's
    ``__len__`` returns the segment count minus one, :48-49, which its own loaders never rely on)."""

    def __init__(self, path, prefix, num_cache=1, segment_size=1024):
        self.path, self.segment_size, self.num_cache = path, segment_size, num_cache
        n_seg = len([f for f in os.listdir(path) if f.startswith(prefix + "_") and f.endswith(".idx")])
        self.data_offsets = [np.load(os.path.join(path, f"{prefix}_{i}.idx"), allow_pickle=True).item()["offsets"]
                             for i in range(n_seg)]
        self.data_files = [open(os.path.join(path, f"{prefix}_{i}.data"), "rb") for i in range(n_seg)]
        self.total_size = sum(len(o) - 1 for o in self.data_offsets)
        self.cache: List[Tuple[int, dict]] = []

    def __getitem__(self, i):
        if i < 0 or i >= self.total_size:
            raise IndexError("index out of range")
        for j, item in self.cache:
            if j == i:
                return item
        seg, idx = divmod(i, self.segment_size)
        off = self.data_offsets[seg]
        self.data_files[seg].seek(off[idx])
        item = pickle.loads(self.data_files[seg].read(off[idx + 1] - off[idx]))
        if self.num_cache > 0:
            self.cache = ([(i, deepcopy(item))] + self.cache)[:self.num_cache]
        return item

    def __len__(self):
        return self.total_size

    def close(self):
        for f in self.data_files:
            f.close()
        self.data_files = []

    def __del__(self):
        self.close()


# ---------------------------------------------------------------------------------------------- bucketing
def bucket_by_length(lengths: Sequence[int], frames_per_batch: int = 65536, max_rows: int = 256) -> List[List[int]]:
    """Group item indices into batches of similar length: longest first, a batch is closed when adding the next item would
    push rows x (longest length, rounded up to the 128-frame tile) past ``frames_per_batch`` or rows past ``max_rows``.
    Every index appears exactly once; inside a batch indices are in decreasing length order."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    batches, cur, cur_T = [], [], 0
    for i in order:
        n = int(lengths[i])
        if n <= 0:
            raise ValueError(f"item {i} has length {n}")
        T = cur_T if cur else -(-n // 128) * 128
        if cur and ((len(cur) + 1) * T > frames_per_batch or len(cur) + 1 > max_rows):
            batches.append(cur)
            cur, T = [], -(-n // 128) * 128
        cur.append(i)
        cur_T = T
    if cur:
        batches.append(cur)
    return batches


def my_batches(n_batches: int, rank: Optional[int] = None, world: Optional[int] = None) -> List[int]:
    """Round-robin share of the batch list for this rank (defaults from torch.distributed when initialised)."""
    import torch.distributed as dist
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
        else:
            rank, world = 0, 1
    return list(range(rank, n_batches, world))


# ---------------------------------------------------------------------------------------------- the sampling loop
@torch.no_grad()
def sample_items(diffusion, conditions: Sequence[torch.Tensor], infer_step: Optional[int] = None, *, use_x_T: bool = False,
                 generator: Optional[torch.Generator] = None, frames_per_batch: int = 65536, rank: Optional[int] = None,
                 world: Optional[int] = None) -> Dict[int, Dict[str, np.ndarray]]:
    """The sampler part of ``SVSRectifiedDiffusionBinarizer.process_item`` (svs.py:214-222) for many utterances at once.

    ``conditions[i]``: [T_i, H] fp32 (what ``forward_condition`` returned for item i, batch dim squeezed), any device.
    ``diffusion``: a ``prodiff_b200.GaussianDiffusion`` or ``RectifiedFlow`` on a CUDA device.
    Returns {item index: {"condition" [T,H], "x_T" [T,M], "x_0" [T,M]}} as numpy arrays -- the three entries the
    reference adds to the item dict -- for the items of this rank's batches.

    ``use_x_T=False`` reproduces the reference as written: it draws ``x_T`` and passes it on, but both samplers ignore the
    argument and start from their own random draw (prodiff.py:136,147; reflow.py:32,88), so the stored ``x_T`` is NOT the
    start of the stored ``x_0``.  ``use_x_T=True`` makes the stored ``x_T`` the actual start.
    """
    from .diffusion import GaussianDiffusion
    device = next(diffusion.parameters()).device
    M = diffusion.mel_bins if isinstance(diffusion, GaussianDiffusion) else diffusion.out_dims
    lengths = [int(c.shape[0]) for c in conditions]
    batches = bucket_by_length(lengths, frames_per_batch)
    out: Dict[int, Dict[str, np.ndarray]] = {}
    for bi in my_batches(len(batches), rank, world):
        idx = batches[bi]
        lens = [lengths[i] for i in idx]
        T = max(lens)
        H = conditions[idx[0]].shape[1]
        cond = torch.zeros(len(idx), T, H, device=device)
        for r, i in enumerate(idx):
            cond[r, :lens[r]] = conditions[i].to(device=device, dtype=torch.float32)
        # svs.py:216: one randn per item, shape [1, 1, M, T_i]; drawn per row so a row's x_T does not depend on its batch
        x_T = torch.zeros(len(idx), 1, M, T, device=device)
        for r in range(len(idx)):
            x_T[r, :, :, :lens[r]] = torch.randn(1, M, lens[r], device=device, generator=generator)
        kw = dict(lengths=lens)
        if infer_step is not None:
            kw["infer_step"] = infer_step
        if use_x_T:
            kw["x_T" if isinstance(diffusion, GaussianDiffusion) else "x_start"] = x_T
        x_0 = diffusion(cond, infer=True, **kw)                                   # [b, T, M]
        for r, i in enumerate(idx):
            n = lens[r]
            out[i] = {"condition": cond[r, :n].cpu().numpy(),                    # [T, Hidden]   (svs.py:219)
                      "x_T": x_T[r, 0, :, :n].transpose(0, 1).cpu().numpy(),      # [T, M]        (svs.py:218,220)
                      "x_0": x_0[r, :n].cpu().numpy()}                            # [T, M]        (svs.py:221)
    return out


__all__ = ["IndexedDatasetBuilder", "IndexedDataset", "bucket_by_length", "my_batches", "sample_items"]
