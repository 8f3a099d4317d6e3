"""Host bloom codec (``'index': 'bloom_cpu'``) and the TF-op blob format.

* ``BloomCPU`` — parity with reference pytorch/deepreduce.py:691-736, which
  drives ``pybloomfilter`` (an mmap'd file + Python loops over the universe).
  Here it is the native C++ filter (``ops/csrc/cpu/bloom_cpu.cpp``) with the same
  false-positive-aware leftmost fill; falls back to the torch oracle without
  the extension.
* ``bloom_compress_blob`` / ``bloom_decompress_blob`` — the single-blob wire of
  the TF custom ops ``BloomCompressor``/``BloomDecompressor`` (reference
  tensorflow/bloom_filter_compression.cc:72-233):
  ``int8[8 + 4K' + m] = [m_bytes:i32][h:i32][K' values bit-cast][m filter bytes]``
  with policies conflict_sets / leftmostK / randomK / policy_zero
  (policies.hpp:182-194) seeded by the training step.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import spec
from .base import SparseCompressor, register
from .bloom import apply_policy_oracle, bloom_insert_oracle, bloom_query_oracle, canonical_policy

_POLICY_ID = {"leftmost": 0, "random": 1, "p0": 2, "conflict_sets": 3}


def _native():
    from .. import ops
    return ops.cpu if ops.has_cpu_native() else None


@register("bloom_cpu")
class BloomCPU(SparseCompressor):
    order_preserving = False
    kind = "index"

    @staticmethod
    def compress(sparse_tensor, params):
        vals, idxs, shape = sparse_tensor
        d = shape.numel()
        K = int(idxs.numel())
        k, m_bits, _ = spec.bloom_layout(K, d, params.get('fpr', None))
        nat = _native()
        idx_cpu = idxs.detach().cpu().long().contiguous()
        if nat is not None:
            words = torch.from_numpy(nat.bloom_insert(idx_cpu.numpy(), k, m_bits, spec.DEFAULT_SEED).view(np.int32))
        else:
            words = bloom_insert_oracle(idx_cpu, k, m_bits)
        dense = params.get('dense_tensor', None)
        if dense is not None:
            if nat is not None:
                sel = torch.from_numpy(nat.bloom_select(words.numpy().view(np.uint32), d, K, k, m_bits,
                                                        spec.DEFAULT_SEED, 0, 0))
            else:
                sel = bloom_query_oracle(words, d, k, m_bits)[:K]
            vals = dense.flatten()[sel.to(dense.device)]
        return vals, words.to(vals.device), shape

    @staticmethod
    def decompress(sparse_tensor, params):
        vals, words, shape = sparse_tensor
        d = shape.numel()
        K = int(vals.numel())
        k, m_bits, _ = spec.bloom_layout(K, d, params.get('fpr', None))
        nat = _native()
        w = words.detach().cpu().contiguous()
        if nat is not None:
            idxs = torch.from_numpy(nat.bloom_select(w.numpy().view(np.uint32), d, K, k, m_bits,
                                                     spec.DEFAULT_SEED, 0, 0))
        else:
            idxs = bloom_query_oracle(w, d, k, m_bits)[:K]
        return vals[: idxs.numel()], idxs.to(vals.device), shape


def tf_bloom_sizes(K: int, fpr: float):
    """(m_bytes, h) with the C++ op's integer division (``m*8 / K``,
    reference bloom_filter_compression.cc:85-99)."""
    import math
    m = int((K * abs(math.log(fpr))) / (math.log(2) ** 2) / 8)
    if m % 8 != 0 or m == 0:
        m += 1
    h = int(math.ceil(((m * 8) // max(K, 1)) * math.log(2)))
    return m, max(h, 1)


def bloom_compress_blob(values: torch.Tensor, indices: torch.Tensor, dense: torch.Tensor, step: int = 0,
                        false_positives_aware: bool = True, policy: str = "conflict_sets",
                        fpr: float = 1e-3) -> torch.Tensor:
    """TF ``BloomCompressor`` op equivalent → int8 blob (CPU)."""
    policy = canonical_policy(policy)
    K = int(indices.numel())
    N = int(dense.numel())
    m_bytes, h = tf_bloom_sizes(K, fpr)
    m_bits = m_bytes * 8
    nat = _native()
    pseed = spec.policy_seed(int(step), 0)
    idx = indices.detach().cpu().long().contiguous()
    if nat is not None:
        words = nat.bloom_insert(idx.numpy(), h, m_bits, spec.DEFAULT_SEED)
        sel = nat.bloom_select(words, N, K, h, m_bits, spec.DEFAULT_SEED, _POLICY_ID[policy], pseed)
        sel = torch.from_numpy(sel)
        words = torch.from_numpy(words.view(np.int32))
    else:
        words = bloom_insert_oracle(idx, h, m_bits)
        pos = bloom_query_oracle(words, N, h, m_bits)
        sel = apply_policy_oracle(pos, K, policy, pseed, h, m_bits)
    flat = dense.detach().cpu().float().flatten()
    if false_positives_aware:
        vals = flat[sel]
    else:
        vals = values.detach().cpu().float().flatten()[: sel.numel()]
    head = torch.tensor([m_bits // 8, h], dtype=torch.int32)
    fbytes = words.view(torch.int8)[: m_bits // 8]
    return torch.cat([head.view(torch.int8), vals.contiguous().view(torch.int8), fbytes])


def bloom_decompress_blob(blob: torch.Tensor, N: int, step: int = 0, policy: str = "conflict_sets") -> torch.Tensor:
    """TF ``BloomDecompressor`` op equivalent → dense float32[N]."""
    policy = canonical_policy(policy)
    blob = blob.detach().cpu().contiguous()
    m_bytes, h = (int(x) for x in blob[:8].view(torch.int32).tolist())
    K = (blob.numel() - m_bytes) // 4 - 2
    vals = blob[8: 8 + 4 * K].contiguous().view(torch.float32)
    fbytes = blob[8 + 4 * K:]
    pad = (-m_bytes) % 4
    if pad:
        fbytes = torch.cat([fbytes, fbytes.new_zeros(pad)])
    words = fbytes.contiguous().view(torch.int32)
    m_bits = m_bytes * 8
    pseed = spec.policy_seed(int(step), 0)
    nat = _native()
    if nat is not None:
        sel = torch.from_numpy(nat.bloom_select(words.numpy().view(np.uint32), N, K, h, m_bits,
                                                spec.DEFAULT_SEED, _POLICY_ID[policy], pseed))
    else:
        pos = bloom_query_oracle(words, N, h, m_bits)
        sel = apply_policy_oracle(pos, K, policy, pseed, h, m_bits)
    out = torch.zeros(N, dtype=torch.float32)
    n = min(sel.numel(), K)
    out[sel[:n]] = vals[:n]
    return out
