"""Bit-level run-length index codec (``'index': 'rle'``).

Parity with reference pytorch/deepreduce.py:805-846: the d-bit occupancy bitmap
is coded as alternating run lengths starting with a run of zeros (a leading 1
gives a first run of length 0), run lengths are bit-packed with
``bitpack.pack``; values are reordered to ascending index order (:816-817).
The reference loops over all d bits in Python on the CPU; here runs are derived
directly from the sorted indices (O(K)), with a warp-cooperative CUDA path.
"""
from __future__ import annotations

import torch

from . import bitpack
from .base import SparseCompressor, register, use_cuda


def runs_from_sorted_oracle(idxs: torch.Tensor, d: int) -> torch.Tensor:
    """ascending unique int64[K] -> int64 runs [z0, o0, z1, o1, ..., (tail zeros)]."""
    K = idxs.numel()
    if K == 0:
        return torch.tensor([d], dtype=torch.int64, device=idxs.device)
    brk = torch.ones(K, dtype=torch.bool, device=idxs.device)
    brk[1:] = idxs[1:] != idxs[:-1] + 1
    starts = idxs[brk]
    end_mask = torch.ones(K, dtype=torch.bool, device=idxs.device)
    end_mask[:-1] = brk[1:]
    ends = idxs[end_mask]
    ones = ends - starts + 1
    prev_end = torch.cat([idxs.new_tensor([-1]), ends[:-1]])
    zeros = starts - prev_end - 1
    runs = torch.stack([zeros, ones], dim=1).flatten()
    tail = d - 1 - int(ends[-1].item())
    if tail > 0:
        runs = torch.cat([runs, runs.new_tensor([tail])])
    return runs


def indices_from_runs_oracle(runs: torch.Tensor) -> torch.Tensor:
    runs = runs.long()
    if runs.numel() < 2:
        return torch.empty(0, dtype=torch.int64, device=runs.device)
    n_pairs = runs.numel() // 2
    z = runs[: 2 * n_pairs: 2]
    o = runs[1: 2 * n_pairs: 2]
    csum = torch.cumsum(z + o, dim=0)
    starts = csum - o
    total = int(o.sum().item())
    if total == 0:
        return torch.empty(0, dtype=torch.int64, device=runs.device)
    seg = torch.repeat_interleave(torch.arange(n_pairs, device=runs.device), o)
    first = torch.cumsum(o, dim=0) - o
    within = torch.arange(total, device=runs.device) - first[seg]
    return starts[seg] + within


@register("rle")
class RunLength(SparseCompressor):
    order_preserving = False
    kind = "index"

    @staticmethod
    def compress(sparse_tensor, params):
        vals, idxs, shape = sparse_tensor
        idxs, mapping = idxs.long().sort(descending=False)
        vals = vals[mapping]
        d = shape.numel()
        if use_cuda(idxs):
            from .. import ops
            runs = ops.rle_runs(idxs, d)
        else:
            runs = runs_from_sorted_oracle(idxs, d)
        return vals, bitpack.pack(runs), shape

    @staticmethod
    def decompress(rle_sparse_tensor, params):
        vals, enc, shape = rle_sparse_tensor
        runs = bitpack.unpack(enc)
        if use_cuda(runs):
            from .. import ops
            idxs = ops.rle_indices(runs, int(vals.numel()))
        else:
            idxs = indices_from_runs_oracle(runs)
        return vals, idxs, shape
