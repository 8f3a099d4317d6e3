"""Metrics / observability.

The reference prints per-tensor relative volumes and codec wall times under
``'micro-benchmark'`` (pytorch/deepreduce.py:74-76,90-95,145-150,294-299) and
its C++ ops write per-step files ``<logs>/<rank>/step_<n>/<gradient_id>/
{fpr,policy_errors,stats}.txt`` (tensorflow/compression_utils.hpp:96-149),
``values.csv``/``coefficients.csv`` (:179-217).  This module provides both:
an in-process accumulator (``METRICS``) and the file loggers with the same
directory layout.
"""
from __future__ import annotations

import os
from collections import defaultdict
from typing import Iterable, Optional

import torch


class Metrics:
    def __init__(self):
        self.reset()

    def reset(self):
        self.times = defaultdict(float)
        self.counts = defaultdict(int)
        self.idx_volume = 0.0
        self.val_volume = 0.0
        self.n_volume = 0
        self.bytes = defaultdict(int)

    def add_time(self, label: str, seconds: float):
        self.times[label] += seconds
        self.counts[label] += 1

    def add_volume(self, idx_rel: float, val_rel: float):
        self.idx_volume += idx_rel
        self.val_volume += val_rel
        self.n_volume += 1

    def add_bytes(self, component: str, n: int):
        self.bytes[component] += int(n)

    def summary(self) -> dict:
        out = {"times_s": dict(self.times), "calls": dict(self.counts), "bytes": dict(self.bytes)}
        if self.n_volume:
            out["mean_idx_relative_volume"] = self.idx_volume / self.n_volume
            out["mean_val_relative_volume"] = self.val_volume / self.n_volume
        return out


METRICS = Metrics()


def relative_volume(wire_tensors: Iterable[torch.Tensor], dense_numel: int) -> float:
    """bits on the wire / (32 · d) — the paper's "relative data volume"."""
    from ..grace import tensor_bits
    return tensor_bits(list(wire_tensors)) / (32.0 * dense_numel)


# ----------------------------------------------------------------------------
# file loggers (layout of tensorflow/compression_utils.hpp)
# ----------------------------------------------------------------------------
def _step_dir(logs_path: str, rank: int, step: int, gradient_id: int) -> str:
    path = os.path.join(logs_path, str(rank), f"step_{step}", str(gradient_id))
    os.makedirs(path, exist_ok=True)
    return path


def log_compressor(logs_path: str, rank: int, step: int, gradient_id: int, *, N: int, K: int,
                   true_indices: torch.Tensor, selected_indices: torch.Tensor, bloom_bytes: int,
                   positives: Optional[int] = None, policy: str = "leftmost", verbosity: int = 1,
                   values: Optional[torch.Tensor] = None) -> dict:
    """Equivalent of ``CompressionUtilities::logging_compressor`` (:96-149):
    false-positive count over the universe, policy errors (selected indices that
    were not in the sparsifier's set), size stats."""
    t = set(true_indices.cpu().tolist())
    sel = selected_indices.cpu().tolist()
    policy_errors = sum(1 for i in sel[:K] if i not in t)
    false_pos = (positives - len(t)) if positives is not None else policy_errors
    path = _step_dir(logs_path, rank, step, gradient_id)
    with open(os.path.join(path, "fpr.txt"), "w") as f:
        f.write(f"FalsePositives: {false_pos}  Total: {N}\n")
    with open(os.path.join(path, "policy_errors.txt"), "w") as f:
        f.write(f"PolicyErrors: {policy_errors}  Total: {K}\n")
    with open(os.path.join(path, "stats.txt"), "w") as f:
        f.write(f"Initial_Size: {N}  Final_Size: {bloom_bytes * 8}\n")
    if verbosity > 1:
        with open(os.path.join(path, f"compressor_logs_{policy}.txt"), "w") as f:
            f.write(f"Indices: {sorted(t)}\n\nIndices Chosen: {sel}\n")
            if values is not None:
                f.write(f"Values-Sent: {values.cpu().tolist()}\n")
            f.write(f"Bloom size: = {bloom_bytes}\nFalsePositives: {false_pos}\nTotal: {N}\n")
    return {"false_positives": false_pos, "policy_errors": policy_errors}


def log_values(logs_path: str, rank: int, step: int, gradient_id: int,
               values: torch.Tensor, coefficients: torch.Tensor) -> None:
    """Equivalent of the ``Logger`` op / ``CompressionUtilities::logging``
    (logger.cc:37-51, compression_utils.hpp:179-217): values.csv + coefficients.csv."""
    path = _step_dir(logs_path, rank, step, gradient_id)
    from .. import ops
    if ops.has_cpu_native():
        ops.cpu.write_csv(os.path.join(path, "values.csv"), values.detach().cpu().double().numpy())
        ops.cpu.write_csv(os.path.join(path, "coefficients.csv"), coefficients.detach().cpu().double().numpy())
        return
    with open(os.path.join(path, "values.csv"), "w") as f:
        f.writelines(f"{v:.40g}\n" for v in values.detach().cpu().flatten().tolist())
    with open(os.path.join(path, "coefficients.csv"), "w") as f:
        f.writelines(f"{v:.40g}\n" for v in coefficients.detach().cpu().flatten().tolist())
