"""Lossless integer-array index codec (``'index': 'integer'`` / ``'delta'``).

Capability parity with reference tensorflow/integer_compression.cc:45-207: a
uint32 array is coded by a codec picked **by name** (``code`` attribute →
FastPFor ``CODECFactory::getFromName``, :62).  FastPFor is an un-vendored
third-party library, so the codec family is implemented natively here
(``ops/csrc/cpu/intcodec.cpp``; CUDA warp-cooperative ``delta+bp128`` in
``ops/csrc/intpack.cu``):

    copy | vbyte (varint) | bp32 | bp128 (simdbinarypacking) | simple8b |
    pfor128 (fastpfor128, simdfastpfor128, fastpfor256)

Index arrays are sorted then **delta-coded** first (gaps of a K-sparse set are
small; this is the paper's "delta/RLE" index path, Table 3), controlled by
``params['delta']`` (default True).

Wire: ``int32[2 + n_words] = [N, codec_id | delta<<8, payload...]``.
"""
from __future__ import annotations

import numpy as np
import torch

from .base import SparseCompressor, register, use_cuda

CODECS = {"copy": 0, "vbyte": 1, "varint": 1, "bp32": 2, "bp128": 3, "simdbinarypacking": 3,
          "fastbinarypacking32": 2, "simple8b": 4, "pfor128": 5, "fastpfor128": 5,
          "simdfastpfor128": 5, "fastpfor256": 5}


# ---- numpy fallback (copy / vbyte / bp128) so the codec works without the .so ----
def _np_vbyte_encode(a: np.ndarray) -> np.ndarray:
    out = bytearray()
    for v in a.tolist():
        while v >= 128:
            out.append((v & 127) | 128)
            v >>= 7
        out.append(v)
    out += b"\0" * ((-len(out)) % 4)
    return np.frombuffer(bytes(out), dtype=np.uint32).copy()


def _np_vbyte_decode(w: np.ndarray, n: int) -> np.ndarray:
    b = w.view(np.uint8)
    out = np.empty(n, dtype=np.uint32)
    p = 0
    for i in range(n):
        v, s = 0, 0
        while True:
            c = int(b[p]); p += 1
            v |= (c & 127) << s
            s += 7
            if c < 128:
                break
        out[i] = v
    return out


def _np_bp_encode(a: np.ndarray, block: int) -> np.ndarray:
    """per-block bit width header word, then block*width bits little-endian."""
    words = []
    for lo in range(0, a.size, block):
        blk = a[lo:lo + block].astype(np.uint64)
        if blk.size < block:
            blk = np.concatenate([blk, np.zeros(block - blk.size, dtype=np.uint64)])
        width = int(blk.max()).bit_length()
        words.append(np.array([width], dtype=np.uint32))
        if width:
            sh = np.arange(width, dtype=np.uint64)
            bits = ((blk[:, None] >> sh[None, :]) & np.uint64(1)).astype(np.uint8).flatten()
            by = np.packbits(bits, bitorder="little")
            by = np.concatenate([by, np.zeros((-by.size) % 4, dtype=np.uint8)])
            words.append(by.view(np.uint32))
    return np.concatenate(words) if words else np.zeros(0, dtype=np.uint32)


def _np_bp_decode(w: np.ndarray, n: int, block: int) -> np.ndarray:
    out = np.zeros(((n + block - 1) // block) * block, dtype=np.uint32)
    p = 0
    for lo in range(0, n, block):
        width = int(w[p]); p += 1
        if width:
            nw = (block * width + 31) // 32
            bits = np.unpackbits(w[p:p + nw].view(np.uint8), bitorder="little")[: block * width]
            p += nw
            sh = (np.uint64(1) << np.arange(width, dtype=np.uint64))
            out[lo:lo + block] = (bits.reshape(block, width).astype(np.uint64) * sh).sum(axis=1).astype(np.uint32)
    return out[:n]


def int_encode(a: np.ndarray, code: str = "bp128") -> np.ndarray:
    from .. import ops
    cid = CODECS[code]
    a = np.ascontiguousarray(a, dtype=np.uint32)
    if ops.has_cpu_native():
        return ops.cpu.int_encode(cid, a)
    if cid == 0:
        return a.copy()
    if cid == 1:
        return _np_vbyte_encode(a)
    if cid in (2, 3):
        return _np_bp_encode(a, 32 if cid == 2 else 128)
    raise RuntimeError(f"integer codec '{code}' needs the native extension (run __graft_entry__.build())")


def int_decode(w: np.ndarray, n: int, code: str = "bp128") -> np.ndarray:
    from .. import ops
    cid = CODECS[code]
    w = np.ascontiguousarray(w, dtype=np.uint32)
    if ops.has_cpu_native():
        return ops.cpu.int_decode(cid, w, int(n))
    if cid == 0:
        return w[:n].copy()
    if cid == 1:
        return _np_vbyte_decode(w, n)
    if cid in (2, 3):
        return _np_bp_decode(w, n, 32 if cid == 2 else 128)
    raise RuntimeError(f"integer codec '{code}' needs the native extension")


@register("integer", "delta", "fastpfor")
class IntegerIndex(SparseCompressor):
    order_preserving = False   # sorts indices ascending (values are permuted with them)
    kind = "index"

    @staticmethod
    def compress(sparse_tensor, params):
        vals, idxs, shape = sparse_tensor
        code = params.get('code', 'bp128')
        delta = bool(params.get('delta', True))
        idxs, order = idxs.long().sort()
        vals = vals[order]
        n = idxs.numel()
        if use_cuda(idxs) and CODECS[code] == 3 and delta:
            from .. import ops
            payload = ops.delta_bp128_encode(idxs)
        else:
            a = idxs.cpu().numpy().astype(np.uint32)
            if delta and n:
                a = np.diff(a, prepend=np.uint32(0)).astype(np.uint32)
            payload = torch.from_numpy(int_encode(a, code).view(np.int32)).to(vals.device)
        head = torch.tensor([n, CODECS[code] | (int(delta) << 8)], dtype=torch.int32, device=vals.device)
        return vals, torch.cat([head, payload.to(torch.int32)]), shape

    @staticmethod
    def decompress(sparse_tensor, params):
        vals, wire, shape = sparse_tensor
        n, meta = (int(x) for x in wire[:2].cpu().tolist())
        cid, delta = meta & 0xFF, bool(meta >> 8)
        code = next(k for k, v in CODECS.items() if v == cid)
        payload = wire[2:]
        if use_cuda(wire) and cid == 3 and delta:
            from .. import ops
            idxs = ops.delta_bp128_decode(payload, n)
        else:
            a = int_decode(payload.cpu().numpy().view(np.uint32), n, code)
            if delta:
                a = np.cumsum(a.astype(np.uint64)).astype(np.int64)
            idxs = torch.from_numpy(a.astype(np.int64)).to(vals.device)
        return vals, idxs, shape
