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


# ----------------------------------------------------------------------------
# remaining diagnostic writers of tensorflow/compression_utils.hpp, integer_compression.cc and logger.cc
# ----------------------------------------------------------------------------
def bitstream_str(buf) -> str:
    """Bit dump of a byte buffer, least-significant bit of every byte first — the
    layout ``CompressionUtilities::fprint`` prints (compression_utils.hpp:23-57)."""
    import numpy as np
    a = buf.detach().cpu().numpy() if torch.is_tensor(buf) else np.asarray(buf)
    bits = np.unpackbits(a.view(np.uint8).reshape(-1), bitorder="little")
    return "Bitstream Array: \n [ " + " ".join(map(str, bits.tolist())) + " ]\n"


def log_decompressor(logs_path: str, rank: int, step: int, gradient_id: int, *, N: int, bloom_words: torch.Tensor,
                     selected_indices: torch.Tensor, values: torch.Tensor, decompressed: torch.Tensor,
                     policy: str = "leftmost", suffix: int = 0, verbosity: int = 2) -> Optional[str]:
    """Receiver-side dump (``logging_decompressor``, compression_utils.hpp:151-176): written only at
    verbosity > 1, one file per sender ``suffix``."""
    if verbosity <= 1:
        return None
    path = os.path.join(_step_dir(logs_path, rank, step, gradient_id), f"decompressor_logs_{policy}_{suffix}.txt")
    with open(path, "w") as f:
        f.write(f"decompressed size: {N}\n\nBloom size: = {bloom_words.numel() * bloom_words.element_size()}\n")
        f.write(bitstream_str(bloom_words))
        f.write(f"\nIndices Chosen: {selected_indices.cpu().tolist()}\nValues Received: {values.cpu().tolist()}\n")
        f.write(f"Decompressed_tensor: {decompressed.detach().cpu().flatten().tolist()}\n{'#' * 88}\n\n")
    return path


def log_bitstream_compressor(logs_path: str, rank: int, step: int, gradient_id: int, *, indices: torch.Tensor,
                             encoded: torch.Tensor, initial_bits: int, runs: Optional[torch.Tensor] = None,
                             verbosity: int = 1) -> dict:
    """Run-length / bitstream index codec logs (``logging_bitstream_compressor``, compression_utils.hpp:220-256):
    ``stats.txt`` always (final size counts the 32-bit length word the worker sends), a verbose dump at verbosity > 1."""
    path = _step_dir(logs_path, rank, step, gradient_id)
    n_bytes = encoded.numel() * encoded.element_size()
    if verbosity > 1:
        with open(os.path.join(path, "RleCompressor_logs.txt"), "w") as f:
            f.write(f"indices_tensor: {indices.cpu().tolist()}\nOutput_concat_size: = {n_bytes}\n\n")
            f.write(bitstream_str(encoded))
            if runs is not None:
                f.write(f"Lengths:\n{runs.cpu().tolist()}\n")
            f.write(f"\n\n{'#' * 88}\n\n")
    stats = {"initial_bits": int(initial_bits), "final_bits": n_bytes * 8 + 32}
    with open(os.path.join(path, "stats.txt"), "w") as f:
        f.write(f"Initial_Size: {stats['initial_bits']}  Final_Size: {stats['final_bits']}\n")
    return stats


def log_bitstream_decompressor(logs_path: str, rank: int, step: int, gradient_id: int, *, encoded: torch.Tensor,
                               indices: torch.Tensor, runs: Optional[torch.Tensor] = None, suffix: int = 0,
                               verbosity: int = 2) -> Optional[str]:
    """``logging_bitstream_decompressor`` (compression_utils.hpp:258-290)."""
    if verbosity <= 1:
        return None
    path = os.path.join(_step_dir(logs_path, rank, step, gradient_id), f"RleDecompressor_logs_{suffix}.txt")
    with open(path, "w") as f:
        f.write(f"encoding_flat: {encoded.cpu().flatten().tolist()}\n")
        f.write(f"Output_concat_size: = {encoded.numel() * encoded.element_size()}\n\n")
        if runs is not None:
            f.write(f"Lengths:\n{runs.cpu().tolist()}\n")
        f.write(f"Indices: {indices.cpu().tolist()}\n\n\n{'#' * 88}\n\n")
    return path


def log_integer_codec(logs_root: str, step: int, suffix: int, *, input_words: torch.Tensor, encoded_words: torch.Tensor,
                      verbosity: int = 1) -> Optional[dict]:
    """Integer (FastPFor-class) codec logs (integer_compression.cc:100-127): every ``verbosity`` steps write
    ``<root>/step_<n>/<suffix>/intcompressor_logs_<suffix>.txt`` and ``stats<suffix>.txt`` (sizes in bits, +32 for the
    length word).  Returns the stats (incl. the compression rate the op prints to stdout, :71-77) or None when skipped."""
    if verbosity == 0 or step % verbosity != 0:
        return None
    path = os.path.join(logs_root, f"step_{step}", str(suffix))
    os.makedirs(path, exist_ok=True)
    n_in, n_out = int(input_words.numel()), int(encoded_words.numel())
    with open(os.path.join(path, f"intcompressor_logs_{suffix}.txt"), "w") as f:
        f.write(f"input_tensor: {input_words.cpu().tolist()}\nOutput_concat_size: = {n_out}\n\n")
        f.write(f"intcompressed_tensor: {encoded_words.cpu().tolist()}\n\n\n{'#' * 88}\n\n")
    stats = {"initial_bits": n_in * 32, "final_bits": n_out * 32 + 32, "rate": n_out / max(1, n_in)}
    with open(os.path.join(path, f"stats{suffix}.txt"), "w") as f:
        f.write(f"Initial_Size: {stats['initial_bits']}  Final_Size: {stats['final_bits']}\n")
    return stats


class StepLogger:
    """The reference's ``Logger`` op (tensorflow/logger.cc:25-55): dumps a dense tensor and its fit
    coefficients to ``values.csv`` / ``coefficients.csv`` every ``verbosity_frequency`` steps (0 = never)."""

    def __init__(self, logs_path: str, gradient_id: int, rank: int = 0, verbosity: int = 1, verbosity_frequency: int = 0):
        self.logs_path, self.gradient_id, self.rank = logs_path, int(gradient_id), int(rank)
        self.verbosity, self.verbosity_frequency = int(verbosity), int(verbosity_frequency)

    def __call__(self, initial_tensor: torch.Tensor, coefficients: torch.Tensor, step: int) -> bool:
        if self.verbosity_frequency == 0 or int(step) % self.verbosity_frequency != 0:
            return False
        log_values(self.logs_path, self.rank, int(step), self.gradient_id, initial_tensor.flatten(), coefficients.flatten())
        return True
