"""Federated averaging with bidirectional DeepReduce compression (paper §6.2, Alg. 2 p.34:
FedAvg, Top-r 10 % with error feedback on both server→client and client→server links).
The reference repo ships no code for this; the paper's FedML harness is
reproduced here as a small in-process simulator so the S2C/C2S volume numbers of
Tables 2/4/5 can be regenerated with any DeepReduce configuration.
"""
from __future__ import annotations

import copy
import time
from typing import Callable, Dict, List

import torch
import torch.nn as nn

from ..grace import ResidualMemory, tensor_bits
from ..wrappers import deepreduce_from_params


class _Link:
    """One compressed link (compressor + its own error-feedback memory), no collective."""

    def __init__(self, params: dict):
        p = dict(params)
        p.setdefault('communicator', 'allgather')
        p['world_size'] = 1
        grc = deepreduce_from_params(p)
        self.compressor = grc.compressor
        self.memory = grc.memory if params.get('memory', 'residual') == 'residual' else ResidualMemory(0.0, 1.0)
        self.bits = 0
        self.dense_bits = 0
        self.encode_s = 0.0          # sender side: compensate + compress + memory update (paper Table 2 "encode")
        self.decode_s = 0.0          # receiver side: decompress ("decode")

    @staticmethod
    def _now(t: torch.Tensor) -> float:
        if t.is_cuda:
            torch.cuda.synchronize(t.device)
        return time.perf_counter()

    def send(self, tensor: torch.Tensor, name: str) -> torch.Tensor:
        t0 = self._now(tensor)
        t = self.memory.compensate(tensor, name)
        wire, ctx = self.compressor.compress(t, name)
        self.memory.update(t, name, self.compressor, wire, ctx)
        t1 = self._now(tensor)
        self.bits += tensor_bits(list(wire))
        self.dense_bits += tensor.numel() * 32
        out = self.compressor.decompress(wire, ctx).view_as(tensor)
        t2 = self._now(tensor)
        self.encode_s += t1 - t0
        self.decode_s += t2 - t1
        return out

    def relative_volume(self) -> float:
        return self.bits / max(self.dense_bits, 1)


class FederatedAveraging:
    def __init__(self, model: nn.Module, params: dict, n_clients: int, local_steps: int = 1, lr: float = 0.05):
        self.server = model
        self.n_clients = n_clients
        self.local_steps = local_steps
        self.lr = lr
        self.s2c = [_Link(params) for _ in range(n_clients)]      # per-client downlink memory
        self.c2s = [_Link(params) for _ in range(n_clients)]
        self.client_state: List[Dict[str, torch.Tensor]] = [
            {n: p.detach().clone() for n, p in model.named_parameters()} for _ in range(n_clients)]

    def round(self, client_batches: List, loss_fn: Callable) -> float:
        """One communication round; ``client_batches[c]`` is a list of (x, y) for client c."""
        server_params = {n: p.detach() for n, p in self.server.named_parameters()}
        agg = {n: torch.zeros_like(p) for n, p in server_params.items()}
        total_loss = 0.0
        for c in range(self.n_clients):
            # S2C: ship the compressed difference between the server model and the client's stale copy
            for n, p in server_params.items():
                delta = self.s2c[c].send(p - self.client_state[c][n], f"s2c.{n}")
                self.client_state[c][n] = self.client_state[c][n] + delta
            local = copy.deepcopy(self.server)
            with torch.no_grad():
                for n, p in local.named_parameters():
                    p.copy_(self.client_state[c][n])
            opt = torch.optim.SGD(local.parameters(), lr=self.lr)
            for x, y in client_batches[c][: self.local_steps]:
                opt.zero_grad()
                loss = loss_fn(local(x), y)
                loss.backward()
                opt.step()
                total_loss += float(loss.detach())
            # C2S: compressed model update
            for n, p in local.named_parameters():
                upd = self.c2s[c].send(p.detach() - self.client_state[c][n], f"c2s.{n}")
                agg[n] += upd / self.n_clients
        with torch.no_grad():
            for n, p in self.server.named_parameters():
                p.add_(agg[n])
        return total_loss / max(1, self.n_clients * self.local_steps)

    def volumes(self):
        s2c = sum(l.bits for l in self.s2c) / max(1, sum(l.dense_bits for l in self.s2c))
        c2s = sum(l.bits for l in self.c2s) / max(1, sum(l.dense_bits for l in self.c2s))
        return {"s2c_relative_volume": s2c, "c2s_relative_volume": c2s}

    def timings(self, rounds: int = 1):
        """Mean codec wall time per client per round (seconds): the rows of the paper's Table 2
        (client encode = C2S compress, client decode = S2C decompress; server side likewise)."""
        n = max(1, self.n_clients * max(1, rounds))
        return {"client_encode_s": sum(l.encode_s for l in self.c2s) / n,
                "client_decode_s": sum(l.decode_s for l in self.s2c) / n,
                "server_encode_s": sum(l.encode_s for l in self.s2c) / n,
                "server_decode_s": sum(l.decode_s for l in self.c2s) / n}
