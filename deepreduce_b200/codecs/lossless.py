"""Lossless byte-level codecs: Deflate for values, Huffman for indices.

Parity with reference pytorch/deepreduce.py:739-764 (``Gzip``: fp32 bytes →
zlib) and :767-802 (``Huffman``: a byte-level Huffman code whose model is the
byte histogram of ``arange(d)`` as int32, applied to the int32 index bytes).
Both are host codecs in the reference too (they exist as comparison points,
paper Table 3); ``dahuffman`` is not installable offline so the Huffman coder
is implemented here (canonical codes, numpy-vectorised bit assembly) and the
model is cached per d instead of being rebuilt on every call.
"""
from __future__ import annotations

import heapq
import zlib
from functools import lru_cache

import numpy as np
import torch

from .base import SparseCompressor, register


@register("gzip", "deflate")
class Gzip(SparseCompressor):
    order_preserving = True
    kind = "value"

    @staticmethod
    def compress(sparse_tensor, params):
        vals, idxs, shape = sparse_tensor
        raw = vals.detach().float().cpu().contiguous().numpy().tobytes()
        packed = zlib.compress(raw, int(params.get('gzip_level', 6)))
        wire = torch.frombuffer(bytearray(packed), dtype=torch.uint8).to(idxs.device)
        return wire, idxs, shape

    @staticmethod
    def decompress(gzip_sparse_tensor, params):
        wire, idxs, shape = gzip_sparse_tensor
        raw = zlib.decompress(wire.cpu().numpy().tobytes())
        if len(raw) == 0:                                  # empty selection (threshold sparsifier on a zero gradient)
            return torch.empty(0, dtype=torch.float32, device=wire.device), idxs, shape
        vals = torch.frombuffer(bytearray(raw), dtype=torch.float32).to(wire.device)
        return vals, idxs, shape


# ----------------------------------------------------------------------------
# canonical Huffman over bytes
# ----------------------------------------------------------------------------
def _code_lengths(freq: np.ndarray) -> np.ndarray:
    syms = [int(s) for s in np.nonzero(freq)[0]]
    lengths = np.zeros(256, dtype=np.int64)
    if len(syms) == 1:
        lengths[syms[0]] = 1
        return lengths
    heap = [(int(freq[s]), s, (s,)) for s in syms]
    heapq.heapify(heap)
    while len(heap) > 1:
        f1, t1, m1 = heapq.heappop(heap)
        f2, t2, m2 = heapq.heappop(heap)
        for s in m1 + m2:
            lengths[s] += 1
        heapq.heappush(heap, (f1 + f2, min(t1, t2), m1 + m2))
    return lengths


def _canonical(lengths: np.ndarray):
    order = sorted((int(l), s) for s, l in enumerate(lengths) if l > 0)
    codes = np.zeros(256, dtype=np.uint64)
    code, prev = 0, 0
    for l, s in order:
        code <<= (l - prev)
        codes[s] = code
        code += 1
        prev = l
    return codes


@lru_cache(maxsize=64)
def _model_for(d: int):
    """Byte histogram of arange(d).int32 little-endian, +1 smoothing so every
    byte value is encodable (the reference's model cannot encode unseen bytes)."""
    freq = np.ones(256, dtype=np.int64)
    step = 1 << 22
    for lo in range(0, d, step):
        b = np.arange(lo, min(d, lo + step), dtype=np.int32).view(np.uint8)
        freq += np.bincount(b, minlength=256)
    lengths = _code_lengths(freq)
    return lengths, _canonical(lengths)


def _native():
    from .. import ops
    return ops._cpu_mod if ops.has_cpu_native() and hasattr(ops._cpu_mod, "huffman_decode") else None


def huffman_encode(data: np.ndarray, lengths: np.ndarray, codes: np.ndarray, native: bool = True) -> np.ndarray:
    """uint8[n] -> uint8 bitstream (MSB-first), 4-byte LE symbol count header.  Uses the C++ coder of
    ``_dr_cpu`` when it is built (``native=False`` forces the numpy reference; both produce the same bytes)."""
    m = _native() if native else None
    if m is not None:
        return m.huffman_encode(np.ascontiguousarray(data, dtype=np.uint8), np.ascontiguousarray(lengths, dtype=np.int64),
                                np.ascontiguousarray(codes, dtype=np.uint64))
    n = data.size
    L = lengths[data]
    C = codes[data]
    total = int(L.sum())
    ends = np.cumsum(L)
    starts = ends - L
    sym = np.repeat(np.arange(n), L)
    bitpos = np.arange(total) - np.repeat(starts, L)
    shift = (np.repeat(L, L) - 1 - bitpos).astype(np.uint64)
    bits = ((C[sym] >> shift) & np.uint64(1)).astype(np.uint8)
    body = np.packbits(bits)
    head = np.array([n & 0xFF, (n >> 8) & 0xFF, (n >> 16) & 0xFF, (n >> 24) & 0xFF], dtype=np.uint8)
    return np.concatenate([head, body])


def huffman_decode(stream: np.ndarray, lengths: np.ndarray, codes: np.ndarray, native: bool = True) -> np.ndarray:
    m = _native() if native else None
    if m is not None:
        return m.huffman_decode(np.ascontiguousarray(stream, dtype=np.uint8), np.ascontiguousarray(lengths, dtype=np.int64),
                                np.ascontiguousarray(codes, dtype=np.uint64))
    n = int(stream[0]) | (int(stream[1]) << 8) | (int(stream[2]) << 16) | (int(stream[3]) << 24)
    bits = np.unpackbits(stream[4:])
    table = {(int(lengths[s]), int(codes[s])): s for s in range(256) if lengths[s] > 0}
    out = np.empty(n, dtype=np.uint8)
    code, length, j = 0, 0, 0
    for b in bits:
        code = (code << 1) | int(b)
        length += 1
        s = table.get((length, code))
        if s is not None:
            out[j] = s
            j += 1
            code, length = 0, 0
            if j == n:
                break
    return out


@register("huffman")
class Huffman(SparseCompressor):
    order_preserving = True
    kind = "index"

    @staticmethod
    def compress(sparse_tensor, params):
        vals, idxs, shape = sparse_tensor
        lengths, codes = _model_for(int(shape.numel()))
        data = idxs.detach().cpu().to(torch.int32).contiguous().numpy().view(np.uint8)
        enc = huffman_encode(data, lengths, codes)
        return vals, torch.from_numpy(enc).to(vals.device), shape

    @staticmethod
    def decompress(sparse_tensor, params):
        vals, enc, shape = sparse_tensor
        lengths, codes = _model_for(int(shape.numel()))
        data = huffman_decode(enc.cpu().numpy(), lengths, codes)
        idxs = torch.from_numpy(data.view(np.int32).astype(np.int64)).to(vals.device)
        return vals, idxs, shape
