"""DeepReduce wrappers around any GRACE sparsifier + the ``from_params`` factory.

API parity with reference pytorch/deepreduce.py:28-302:
``ValueCompressor`` (:51-97), ``IndexCompressor`` (:100-153), ``DeepReduce``
('both', :156-302) and ``deepreduce_from_params`` (:28-48); usage per
reference README.md:30-48.

Decisions where the reference is ambiguous or buggy:

* 'both' is false-positive aware: the index codec sees the dense tensor, so the
  value codec fits ``g[S~]`` in the bloom order and ``mapping`` permutes it
  (the reference never sets ``dense_tensor`` in 'both' and scrambles values).
* ``mapping`` travels bit-packed at ⌈log2 K⌉ bits (``bitpack.pack``), not int64.
* no ``params`` side channels: per-call state (dense tensor) goes through a
  shallow copy of the user's dict; the user's dict is never mutated.
* timing uses CUDA events only when ``'micro-benchmark'`` is set — no
  unconditional ``torch.cuda.synchronize()`` (:71,86,120,140,256,278).
"""
from __future__ import annotations

import time

import torch

from . import spec
from .codecs import bitpack, compressor
from .grace import Compressor, grace_from_params, tensor_bits
from .utils.metrics import METRICS


class _Timer:
    """CUDA-event timer active only under 'micro-benchmark'."""

    def __init__(self, enabled: bool, label: str, device):
        self.enabled = enabled
        self.label = label
        self.cuda = enabled and torch.cuda.is_available() and getattr(device, "type", "cpu") == "cuda"

    def __enter__(self):
        if self.cuda:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        elif self.enabled:
            self.t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        if self.cuda:
            self.e1.record()
            self.e1.synchronize()
            dt = self.e0.elapsed_time(self.e1) * 1e-3
        elif self.enabled:
            dt = time.perf_counter() - self.t0
        else:
            return
        METRICS.add_time(self.label, dt)
        print(f'{self.label} time:{dt}')


def _report_volume(params, tensors, shape):
    if params.get('micro-benchmark', False):
        dense_bits = shape.numel() * 32
        idx_v = tensor_bits(list(tensors[1:])) / dense_bits
        val_v = tensor_bits([tensors[0]]) / dense_bits
        METRICS.add_volume(idx_v, val_v)
        print(f'idx_relative_volume: {idx_v:.4f}')
        print(f'val_relative_volume: {val_v:.4f}')


def _wire_idx(idxs: torch.Tensor, numel: int) -> torch.Tensor:
    """Indices travel as int32 whenever the tensor has < 2^31 elements (the paper's volume accounting, e.g. Table 2:
    Top-r 10 % = 0.2033, counts 32-bit keys; GRACE's sparsifier hands out int64)."""
    if idxs.dtype == torch.int64 and numel < 2 ** 31:
        return idxs.to(torch.int32)
    return idxs


def _host_idx(idxs: torch.Tensor) -> torch.Tensor:
    return idxs.long() if idxs.dtype == torch.int32 else idxs


class _Wrapper(Compressor):
    def __init__(self, sparsifier, params=None):
        super().__init__(average=getattr(sparsifier, "average", True),
                         tensors_size_are_same=sparsifier.tensors_size_are_same)
        self.sparsifier = sparsifier
        self.params = dict(params or {})
        self.min_numel = int(self.params.get('min_numel', spec.SMALL_TENSOR_NUMEL))
        self.bench = bool(self.params.get('micro-benchmark', False))
        # seeded index-selection policies ('random' / 'conflict_sets'): the seed is a function of (step, tensor) — the
        # C++ reference seeds with the training step (tensorflow/policies.hpp:171), the PyTorch one reseeds the GLOBAL
        # torch RNG with 42 on every call (pytorch/deepreduce.py:487).  Every rank compresses the same tensor the same
        # number of times, so the per-name step counters agree; decompress() runs right after compress() of the same
        # tensor (GRACE Communicator.step), so it uses the seed of the latest compress.  An explicit
        # params['policy_seed'] pins the seed instead.
        self._steps: dict = {}
        self._call_seed = int(self.params.get('policy_seed', 42))

    def _begin_call(self, name) -> dict:
        call = dict(self.params)
        if 'policy_seed' not in self.params:
            step = self._steps.get(name, 0)
            self._steps[name] = step + 1
            tid = sum(str(name).encode()) if name is not None else 0
            self._call_seed = spec.policy_seed(step, tid)
        call['policy_seed'] = self._call_seed
        return call

    def _decode_params(self) -> dict:
        call = dict(self.params)
        call['policy_seed'] = self._call_seed
        return call


class ValueCompressor(_Wrapper):
    def __init__(self, sparsifier, params=None):
        super().__init__(sparsifier, params)
        name = self.params.get('value', 'polyfit')
        self.val_compressor = compressor[name]
        if name not in ['polyfit', 'dexp', 'double_exp']:
            self.tensors_size_are_same = False

    def compress(self, tensor, name):
        tensors, ctx = self.sparsifier.compress(tensor, name)
        vals, idxs = tensors
        shape = ctx
        if shape.numel() > self.min_numel:
            with _Timer(self.bench, 'val_compression', tensor.device):
                vals, idxs, shape = self.val_compressor.compress((vals, idxs, tensor.size()), self.params)
            ctx = shape
        return (vals, _wire_idx(idxs, tensor.numel())), ctx

    def decompress(self, tensors, ctx):
        shape = ctx
        vals, idxs = tensors
        idxs = _host_idx(idxs)
        if shape.numel() > self.min_numel:
            with _Timer(self.bench, 'val_decompression', idxs.device):
                vals, idxs, shape = self.val_compressor.decompress((vals, idxs, shape), self.params)
            _report_volume(self.params, tensors, shape)
        return self.sparsifier.decompress((vals, _host_idx(idxs)), shape)


class IndexCompressor(_Wrapper):
    def __init__(self, sparsifier, params=None):
        super().__init__(sparsifier, params)
        name = self.params.get('index', 'bloom')
        self.idx_compressor = compressor[name]
        policy = self.params.get('policy', 'leftmost')
        if name not in ['bloom', 'bloom_cpu'] or policy in ('p0', 'policy_zero', 'P0'):
            self.tensors_size_are_same = False

    def compress(self, tensor, name):
        tensors, ctx = self.sparsifier.compress(tensor, name)
        vals, idxs = tensors
        shape = ctx
        if shape.numel() > self.min_numel:
            call = self._begin_call(name)
            call['dense_tensor'] = tensor     # FP-aware fill (reference :117) without mutating user params
            with _Timer(self.bench, 'idx_compression', tensor.device):
                vals, idxs, shape = self.idx_compressor.compress((vals, idxs, tensor.size()), call)
            ctx = shape
        else:
            idxs = _wire_idx(idxs, tensor.numel())       # small-tensor bypass: plain pairs, 32-bit keys
        return (vals, idxs), ctx

    def decompress(self, tensors, ctx):
        shape = ctx
        vals, idxs = tensors
        if shape.numel() > self.min_numel:
            with _Timer(self.bench, 'idx_decompression', vals.device):
                vals, idxs, shape = self.idx_compressor.decompress((vals, idxs, shape), self._decode_params())
            _report_volume(self.params, tensors, shape)
        else:
            idxs = _host_idx(idxs)
        return self.sparsifier.decompress((vals, idxs), shape)


class DeepReduce(_Wrapper):
    """'both': index codec first, then the value codec on ``(vals, arange)`` so its
    reorder permutation comes back as ``mapping`` (reference :250-302)."""

    pack = staticmethod(bitpack.pack)
    unpack = staticmethod(bitpack.unpack)
    pack_ = staticmethod(bitpack.pack_)
    unpack_ = staticmethod(bitpack.unpack_)

    def __init__(self, sparsifier, params=None):
        super().__init__(sparsifier, params)
        self.val_compressor = compressor[self.params.get('value', 'polyfit')]
        self.idx_compressor = compressor[self.params.get('index', 'bloom')]
        # dexp folds the sign into the (signed) mapping, which therefore cannot be bit-packed as unsigned
        self.pack_mapping = bool(self.params.get('pack_mapping', True)) and not getattr(
            self.val_compressor, 'signed_mapping', False)
        # value codecs other than the fixed-layout fits, and non-bloom index codecs, vary in size
        if (self.params.get('value', 'polyfit') not in ('polyfit', 'dexp')
                or self.params.get('index', 'bloom') not in ('bloom', 'bloom_cpu')
                or self.params.get('policy', 'leftmost') in ('p0', 'policy_zero', 'P0')):
            self.tensors_size_are_same = False

    def compress(self, tensor, name):
        tensors, ctx = self.sparsifier.compress(tensor, name)
        vals, idxs = tensors
        shape = ctx
        with _Timer(self.bench, '_compression', tensor.device):
            if shape.numel() > self.min_numel:
                call = self._begin_call(name)
                call['dense_tensor'] = tensor
                vals, idxs_c, _ = self.idx_compressor.compress((vals, idxs, tensor.size()), call)
                head = None
                if self.idx_compressor.kind == "index" and call.get('policy', 'leftmost') in ('p0', 'policy_zero', 'P0') \
                        and self.params.get('index', 'bloom') == 'bloom':
                    head, vals = vals[:1], vals[1:]          # keep K out of the value codec
                new_idxs = torch.arange(vals.numel(), device=vals.device)
                vals_c, mapping, shape = self.val_compressor.compress((vals, new_idxs, shape), self.params)
                n_map = mapping.numel()
                if self.val_compressor.order_preserving:
                    # the i-th value belongs to the i-th decoded index: there is no permutation to ship (the
                    # reference always sends `mapping`, :267; the paper's BF+QSGD volumes do not include one)
                    mapping = torch.empty(0, dtype=torch.uint8 if self.pack_mapping else mapping.dtype,
                                          device=mapping.device)
                elif self.pack_mapping:
                    mapping = bitpack.pack(mapping, max_val=max(n_map - 1, 1))
                if head is not None:
                    # K rides in front of the mapping blob as 4 extra bytes / one extra entry
                    extra = head.to(torch.float32).view(torch.uint8) if self.pack_mapping \
                        else head.to(mapping.dtype)
                    mapping = torch.cat([extra.to(mapping.device), mapping])
                ctx = shape
                tensors = (vals_c, idxs_c, mapping)
            else:
                tensors = (vals, _wire_idx(idxs, tensor.numel()))
        return tensors, ctx

    def decompress(self, tensors, ctx):
        shape = ctx
        dev = tensors[0].device
        with _Timer(self.bench, '_decompression', dev):
            if shape.numel() > self.min_numel:
                vals_c, idxs_c, mapping = tensors
                p0 = self.params.get('policy', 'leftmost') in ('p0', 'policy_zero', 'P0') \
                    and self.params.get('index', 'bloom') == 'bloom'
                head = None
                if p0:
                    if self.pack_mapping:
                        head, mapping = mapping[:4].contiguous().view(torch.float32), mapping[4:]
                    else:
                        head, mapping = mapping[:1].float(), mapping[1:]
                if self.val_compressor.order_preserving:
                    vals, _, _ = self.val_compressor.decompress((vals_c, None, shape), self.params)
                    mapping = torch.arange(vals.numel(), device=vals.device)
                else:
                    if self.pack_mapping:
                        mapping = bitpack.unpack(mapping)
                    vals, mapping, _ = self.val_compressor.decompress((vals_c, mapping, shape), self.params)
                carrier = vals.new_zeros(mapping.numel()) if head is None else torch.cat(
                    [head.to(vals.dtype), vals.new_zeros(mapping.numel())])
                _, idxs, _ = self.idx_compressor.decompress((carrier, idxs_c, shape), self._decode_params())
                if self.val_compressor.order_preserving:
                    pass                               # i-th value belongs to the i-th decoded index
                else:
                    idxs = idxs[mapping.long()]        # i-th sorted value ↔ mapping[i]-th index (:290)
                n = min(vals.numel(), idxs.numel())
                vals, idxs = vals[:n], idxs[:n]
            else:
                vals, idxs = tensors
                idxs = _host_idx(idxs)
        if shape.numel() > self.min_numel:
            _report_volume(self.params, tensors, shape)
        return self.sparsifier.decompress((vals, idxs), shape)


deepreduce_wrapper = {'value': ValueCompressor, 'index': IndexCompressor, 'both': DeepReduce}


def deepreduce_from_params(params):
    """Factory (reference :28-48).  No hash table is loaded: hashing is on the fly.
    The dict is validated once (``config.DeepReduceConfig``: types, ranges, codec names, typo'd keys) and never
    written to."""
    from .config import validate_params
    validate_params(params)
    grc = grace_from_params(params)
    deepreduce = params.get('deepreduce', None)
    if deepreduce:
        if deepreduce not in deepreduce_wrapper:
            raise ValueError(f"'deepreduce' must be one of None, 'value', 'index', 'both' (got {deepreduce!r})")
        grc.compressor = deepreduce_wrapper[deepreduce](grc.compressor, params)
    return grc


from_params = deepreduce_from_params
