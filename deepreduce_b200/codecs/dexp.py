"""Double-exponential value codec, "Fit-DExp" (``'value': 'dexp'``).

Parity with reference tensorflow/deepreduce.py:66-144 (``double_exponential_fit``)
and :376-442 (``DoubleExpCompressor``): |values| sorted ascending are modelled as
``y ≈ a·e^{p x} + b·e^{q x}`` on x = 1..K using the integral-equation regression
(cumulative trapezoids S, SS → 4×4 normal system → p,q → 2×2 system for a,b);
only applied when the tensor has more than 9000 elements (:396); the sign rides
on the index as ``(idx+1)·sign`` (:399).

To keep e^{px} finite in fp32/fp64 for K up to 10^6 the abscissa is rescaled to
x ∈ (0, 1] (x_i = i/K); p and q on the wire are for that abscissa.
Wire: ``float32[4] = (a, b, p, q)`` + signed int32 indices ordered by ascending |v|.
"""
from __future__ import annotations

import torch

from .base import SparseCompressor, register

MIN_NUMEL = 9000


def _cumtrapz(y, dx):
    inc = 0.5 * (y[1:] + y[:-1]) * dx
    return torch.cat([y.new_zeros(1), torch.cumsum(inc, dim=0)])


def double_exponential_fit_oracle(y: torch.Tensor):
    """Plain-torch specification (always the library path; numerics reference of the kernel)."""
    return double_exponential_fit(y.detach().cpu())


def double_exponential_fit(y: torch.Tensor):
    """y: [K] (ascending |values|) -> (a, b, p, q) float64 on x_i = i/K."""
    K = y.numel()
    if y.is_cuda and K > 0:                                # hand-written kernel (ops/csrc/ops.cu::dexp_fit_kernel)
        from .. import ops
        if ops.require():
            return tuple(ops.dexp_fit(y.float()).unbind())
    y = y.double()
    if K == 0:                                             # empty selection: the zero curve
        z = torch.zeros((), dtype=torch.float64, device=y.device)
        return z, z, z, z
    x = torch.arange(1, K + 1, dtype=torch.float64, device=y.device) / K
    dx = 1.0 / K
    S = _cumtrapz(y, dx)
    SS = _cumtrapz(S, dx)
    cols = torch.stack([SS, S, x, torch.ones_like(x)], dim=1)      # y ~ A*SS + B*S + C*x + D
    G = cols.T @ cols
    rhs = cols.T @ y
    try:
        sol = torch.linalg.solve(G + 1e-18 * torch.eye(4, dtype=G.dtype, device=G.device), rhs)
    except RuntimeError:
        sol = torch.linalg.lstsq(G, rhs[:, None]).solution[:, 0]
    A, B = sol[0], sol[1]
    disc = torch.clamp(B * B + 4 * A, min=0.0)
    p = 0.5 * (B + torch.sqrt(disc))
    q = 0.5 * (B - torch.sqrt(disc))
    bk, ek = torch.exp(p * x), torch.exp(q * x)
    M = torch.stack([torch.stack([(bk * bk).sum(), (bk * ek).sum()]),
                     torch.stack([(bk * ek).sum(), (ek * ek).sum()])])
    r = torch.stack([(bk * y).sum(), (ek * y).sum()])
    if torch.abs(torch.linalg.det(M)) < 1e-300:
        ab = torch.stack([r[0] / M[0, 0], torch.zeros_like(r[0])])
    else:
        ab = torch.linalg.solve(M, r)
    return ab[0], ab[1], p, q


def double_exponential_eval(coef: torch.Tensor, K: int) -> torch.Tensor:
    a, b, p, q = coef.double().unbind()
    if K == 0:
        return torch.empty(0, dtype=torch.float32, device=coef.device)
    x = torch.arange(1, K + 1, dtype=torch.float64, device=coef.device) / K
    return (a * torch.exp(p * x) + b * torch.exp(q * x)).float()


@register("dexp", "double_exp")
class DoubleExp(SparseCompressor):
    order_preserving = False
    kind = "value"
    signed_mapping = True

    @staticmethod
    def compress(sparse_tensor, params):
        vals, idxs, shape = sparse_tensor
        if shape.numel() <= int(params.get('dexp_min_numel', MIN_NUMEL)):
            return vals, idxs, shape
        absv, order = vals.abs().sort(descending=False)
        sign = torch.where(vals[order] > 0, 1, -1)
        signed_idx = ((idxs[order].long() + 1) * sign).to(torch.int32)
        coef = torch.stack(double_exponential_fit(absv)).float()
        return coef, signed_idx, shape

    @staticmethod
    def decompress(sparse_tensor, params):
        coef, signed_idx, shape = sparse_tensor
        if shape.numel() <= int(params.get('dexp_min_numel', MIN_NUMEL)):
            return coef, signed_idx, shape
        K = signed_idx.numel()
        vals = double_exponential_eval(coef, K) * torch.sign(signed_idx).float()
        idxs = signed_idx.long().abs() - 1
        return vals, idxs, shape
