"""Communicators (GRACE ``communicator`` key): allgather / allreduce.

``Allgather`` reproduces the reference data path exactly (SURVEY Appendix A,
C1): one ``all_gather`` per wire component, a size exchange + pad-to-max when
``tensors_size_are_same`` is False, W separate decodes, aggregate, /W.  It is
the *compatibility* path (works on gloo/CPU and NCCL); the product path is the
fused bucket engine in ``parallel/engine.py`` which replaces all of this with
one kernel and in-kernel P2P stores.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist

from .base import Communicator


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class Allgather(Communicator):
    def __init__(self, compressor, memory, world_size: int | None = None, group=None):
        super().__init__(compressor, memory)
        self.group = group
        self.world_size = world_size if world_size is not None else _world(group)

    def _gather_same(self, x: torch.Tensor):
        out = [torch.empty_like(x) for _ in range(self.world_size)]
        dist.all_gather(out, x.contiguous(), group=self.group)
        return out

    def send_receive(self, tensors: Sequence[torch.Tensor], name, ctx):
        W = self.world_size
        tensors = [t if torch.is_tensor(t) else torch.as_tensor(t) for t in tensors]
        self.bytes_sent += sum(t.numel() * t.element_size() for t in tensors)
        if W == 1:
            per_rank = [list(tensors)]
        elif self.compressor.tensors_size_are_same:
            gathered = [self._gather_same(t) for t in tensors]
            per_rank = [[g[r] for g in gathered] for r in range(W)]
        else:
            dev = tensors[0].device
            sizes = torch.tensor([t.numel() for t in tensors], dtype=torch.int64, device=dev)
            all_sizes = self._gather_same(sizes)
            all_sizes = torch.stack(all_sizes).cpu()
            max_sizes = all_sizes.max(dim=0).values.tolist()
            gathered = []
            for c, t in enumerate(tensors):
                flat = t.flatten()
                pad = max_sizes[c] - flat.numel()
                if pad:
                    flat = torch.cat([flat, flat.new_zeros(pad)])
                gathered.append(self._gather_same(flat))
            per_rank = []
            for r in range(W):
                per_rank.append([gathered[c][r][: int(all_sizes[r, c])] for c in range(len(tensors))])
        dense = [self.compressor.decompress(parts, ctx) for parts in per_rank]
        out = self.compressor.aggregate(dense)
        return out / W if self.compressor.average else out


class Allreduce(Communicator):
    """Dense baseline: ``'communicator': 'allreduce'`` (reference run_deepreduce.sh:51)."""

    def __init__(self, compressor, memory, world_size: int | None = None, group=None):
        super().__init__(compressor, memory)
        self.group = group
        self.world_size = world_size if world_size is not None else _world(group)

    def send_receive(self, tensors, name, ctx):
        out = []
        for t in tensors:
            self.bytes_sent += t.numel() * t.element_size()
            # only VALUE components are summed: an index component (randomk ships int64 coordinates that are
            # identical on every rank by construction — same (step, name) seed) must reach decompress() unchanged,
            # summing it would scatter to W*idx
            if self.world_size > 1 and torch.is_tensor(t) and t.is_floating_point():
                dist.all_reduce(t, group=self.group)
            out.append(t)
        dense = self.compressor.decompress(out, ctx)
        return dense / self.world_size if self.compressor.average else dense
