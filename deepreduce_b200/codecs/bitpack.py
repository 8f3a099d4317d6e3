"""Arbitrary-bit-width integer packing.

Capability parity with reference pytorch/deepreduce.py:165-248
(``DeepReduce.pack_/unpack_``: 3×21-bit per int64; ``pack/unpack``: header
``[N:4B][bits:1B]`` + packed body, used for RLE run lengths (:831) and intended
for the 'both' mapping (:264-265)).  Wire layout here: 5-byte header then one
contiguous little-endian bit stream of ``N*bits`` bits (value i occupies bits
``[i*bits, (i+1)*bits)``, LSB first) — a single pass for the warp-cooperative
CUDA packer instead of the reference's byte-planes + cupy bit-planes.
"""
from __future__ import annotations

import math

import torch

from .base import use_cuda

HEADER_BYTES = 5


def packed_nbytes(n: int, bits: int) -> int:
    return (n * bits + 7) // 8


def pack_bits_oracle(vals: torch.Tensor, bits: int) -> torch.Tensor:
    """int tensor [N] (values < 2**bits) -> uint8[ceil(N*bits/8)]."""
    n = vals.numel()
    if n == 0:
        return torch.empty(0, dtype=torch.uint8, device=vals.device)
    v = vals.to(torch.int64).flatten()
    sh = torch.arange(bits, device=v.device, dtype=torch.int64)
    b = ((v[:, None] >> sh[None, :]) & 1).flatten()
    pad = (-b.numel()) % 8
    if pad:
        b = torch.cat([b, b.new_zeros(pad)])
    w = (1 << torch.arange(8, device=v.device, dtype=torch.int64))
    return (b.view(-1, 8) * w).sum(dim=1).to(torch.uint8)


def unpack_bits_oracle(buf: torch.Tensor, n: int, bits: int) -> torch.Tensor:
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=buf.device)
    sh8 = torch.arange(8, device=buf.device, dtype=torch.int64)
    b = ((buf.to(torch.int64)[:, None] >> sh8[None, :]) & 1).flatten()[: n * bits]
    w = (1 << torch.arange(bits, device=buf.device, dtype=torch.int64))
    return (b.view(n, bits) * w).sum(dim=1)


def pack_bits(vals, bits):
    if use_cuda(vals):
        from .. import ops
        return ops.pack_bits(vals, bits)
    return pack_bits_oracle(vals, bits)


def unpack_bits(buf, n, bits):
    if use_cuda(buf):
        from .. import ops
        return ops.unpack_bits(buf, n, bits)
    return unpack_bits_oracle(buf, n, bits)


def pack(mapping: torch.Tensor, max_val=None) -> torch.Tensor:
    """``[N:u32 LE][bits:u8][bitstream]`` as one uint8 tensor."""
    n = mapping.numel()
    if max_val is None:
        max_val = int(mapping.max().item()) if n else 0
    bits = max(1, int(max_val).bit_length())
    head = torch.tensor([n & 0xFF, (n >> 8) & 0xFF, (n >> 16) & 0xFF, (n >> 24) & 0xFF, bits],
                        dtype=torch.uint8, device=mapping.device)
    return torch.cat([head, pack_bits(mapping, bits)])


def unpack(encode: torch.Tensor) -> torch.Tensor:
    head = encode[:HEADER_BYTES].cpu().tolist()
    n = head[0] | (head[1] << 8) | (head[2] << 16) | (head[3] << 24)
    bits = head[4]
    return unpack_bits(encode[HEADER_BYTES:], n, bits).long()


def pack_(mapping: torch.Tensor, max_val=None) -> torch.Tensor:
    """Three 21-bit values per int64 (+ trailing N) — reference :165-180."""
    bits = 21
    mapping = mapping.long()
    n = mapping.numel()
    chunk = 63 // bits
    pad = (-n) % chunk
    padded = torch.cat([mapping, mapping.new_zeros(pad)]).view(-1, chunk)
    enc = (padded[:, 0] << (2 * bits)) | (padded[:, 1] << bits) | padded[:, 2]
    return torch.cat([enc, torch.tensor([n], dtype=torch.int64, device=mapping.device)])


def unpack_(encode: torch.Tensor) -> torch.Tensor:
    bits = 21
    n = int(encode[-1].item())
    e = encode[:-1]
    mask = (1 << bits) - 1
    out = torch.stack([(e >> (2 * bits)) & mask, (e >> bits) & mask, e & mask], dim=1).flatten()
    return out[:n].int()


def bits_needed(max_val: int) -> int:
    return max(1, int(math.ceil(math.log2(max_val + 1)))) if max_val > 0 else 1
