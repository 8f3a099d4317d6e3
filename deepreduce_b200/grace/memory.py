"""Error-feedback memories (GRACE ``memory`` key).

``ResidualMemory``: ``compensate: g <- beta*r + gamma*g`` and
``update: r <- g - decompress(compress(g))`` (TF twin in
reference tensorflow/deepreduce.py:31-52).  Unlike GRACE the residual state is
checkpointable (``state_dict``), see SURVEY §5.
"""
from __future__ import annotations

import torch

from .base import Memory


class NoneMemory(Memory):
    def compensate(self, tensor, name):
        return tensor

    def update(self, tensor, name, compressor, tensor_compressed, ctx):
        pass


class ResidualMemory(Memory):
    def __init__(self, beta: float = 1.0, gamma: float = 1.0):
        self.residuals: dict[str, torch.Tensor] = {}
        self.beta = beta
        self.gamma = gamma

    def compensate(self, tensor, name):
        if name in self.residuals:
            tensor = self.beta * self.residuals[name] + self.gamma * tensor
        elif self.gamma != 1.0:
            tensor = self.gamma * tensor
        return tensor

    def update(self, tensor, name, compressor, tensor_compressed, ctx):
        tensor_decompressed = compressor.decompress(tensor_compressed, ctx)
        self.residuals[name] = tensor - tensor_decompressed.view_as(tensor)

    def state_dict(self):
        return {"beta": self.beta, "gamma": self.gamma,
                "residuals": {k: v.detach().cpu().clone() for k, v in self.residuals.items()}}

    def load_state_dict(self, state, device=None):
        self.beta = state.get("beta", self.beta)
        self.gamma = state.get("gamma", self.gamma)
        self.residuals = {k: (v.to(device) if device is not None else v.clone())
                          for k, v in state.get("residuals", {}).items()}
