"""GRACE-compatible abstract interfaces.

The reference is a plug-in for GRACE (``grace_dl.dist``; reference
pytorch/deepreduce.py:7-8,29) which is not vendored.  These classes reproduce
the contract the reference's call sites rely on:
``Compressor.compress(tensor, name) -> (tensors, ctx)``,
``decompress(tensors, ctx)``, ``aggregate``, attrs ``average`` and
``tensors_size_are_same``; ``Memory.compensate/update``;
``Communicator.step(tensor, name)``.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Sequence, Tuple

import torch


class Memory(ABC):
    @abstractmethod
    def compensate(self, tensor: torch.Tensor, name: str) -> torch.Tensor:
        """Return the tensor with the stored residual folded in."""

    def update(self, tensor, name, compressor, tensor_compressed, ctx) -> None:
        """Store what compression lost."""

    def state_dict(self) -> dict:
        return {}

    def load_state_dict(self, state: dict, device=None) -> None:
        pass


class Compressor(ABC):
    """Interface for compressing and decompressing a given tensor."""

    def __init__(self, average: bool = True, tensors_size_are_same: bool = True):
        self.average = average
        self.tensors_size_are_same = tensors_size_are_same

    @abstractmethod
    def compress(self, tensor: torch.Tensor, name: str) -> Tuple[Sequence[torch.Tensor], Any]:
        """Return (list_of_wire_tensors, ctx)."""

    @abstractmethod
    def decompress(self, tensors: Sequence[torch.Tensor], ctx: Any) -> torch.Tensor:
        """Rebuild the dense tensor."""

    def aggregate(self, tensors: Sequence[torch.Tensor]) -> torch.Tensor:
        return sum(tensors)


class Communicator(ABC):
    """``step`` is the per-tensor entry point a trainer calls after backward
    (SURVEY §3.2)."""

    def __init__(self, compressor: Compressor, memory: Memory):
        self.compressor = compressor
        self.memory = memory
        self.bytes_sent = 0  # wire bytes this rank contributed (metrics)
        self.dense_bytes = 0

    @abstractmethod
    def send_receive(self, tensors, name, ctx) -> torch.Tensor:
        ...

    def step(self, tensor: torch.Tensor, name: str) -> torch.Tensor:
        tensor = self.memory.compensate(tensor, name)
        tensors_compressed, ctx = self.compressor.compress(tensor, name)
        self.memory.update(tensor, name, self.compressor, tensors_compressed, ctx)
        self.dense_bytes += tensor.numel() * tensor.element_size()
        return self.send_receive(tensors_compressed, name, ctx)
