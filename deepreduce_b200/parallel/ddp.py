"""Bucketed, overlapped data-parallel wrapper — the new entry point next to the
per-tensor ``grc.step(grad, name)``.

The reference runs strictly after backward, one Python call and 2-3 NCCL
all_gathers per tensor.  Here parameters are laid out in flat
fp32 buckets (``p.grad`` are views), autograd post-accumulate hooks mark
buckets ready during backward, and each ready bucket is handed to a C++
background thread that launches the fused exchange kernel on a high-priority
side stream; ``finish()`` makes the optimizer's stream wait on the done events.

CUDA + ('topk' [+ 'index': 'bloom' | plain])  -> fused engine (one kernel/bucket)
CUDA + 'none'/'allreduce'                     -> dense NCCL all-reduce of the flat bucket
anything else (CPU/gloo, other codecs)        -> GRACE-compatible per-tensor path
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import spec
from ..wrappers import deepreduce_from_params
from .engine import BucketEngine
from .plan import BucketPlan


def _is_dense(p: torch.Tensor) -> bool:
    """True if p's strides describe a permutation-dense layout (contiguous / channels_last / ...)."""
    sizes, strides = list(p.size()), list(p.stride())
    expect = 1
    for st, sz in sorted(zip(strides, sizes)):
        if sz == 1:
            continue
        if st != expect:
            return False
        expect *= sz
    return True


def _fused_supported(params: dict) -> bool:
    """Which ``params`` dicts the fused bucket engine serves (everything else takes the GRACE-compatible per-tensor
    path).  Covers every recipe of the reference's launch script (run_deepreduce.sh:35-107): top-k or threshold
    sparsifier x {no codec, index (bloom leftmost / random / p0, run-length), value (polyfit, QSGD int8/int16), both}.
    Not fused: bloom policy 'conflict_sets' (per-tensor GPU kernel), host codecs (Huffman, Deflate, dexp, the integer
    family), 'randomk', non-512 QSGD buckets."""
    if params.get('compressor') not in ('topk', 'threshold') or params.get('communicator', 'allgather') != 'allgather':
        return False
    dr = params.get('deepreduce', None)
    if dr is None:
        return True
    from ..codecs.bloom import canonical_policy
    pol_ok = canonical_policy(params.get('policy', 'leftmost')) in ('leftmost', 'random', 'p0')
    value_ok = (params.get('value', 'polyfit') == 'polyfit'
                or (params.get('value') == 'qsgd' and 1 <= int(params.get('quantum_num', 127)) <= 32767
                    and int(params.get('bucket_size', 512)) == 512))
    if dr == 'index' and params.get('index', 'bloom') == 'bloom':
        return pol_ok
    if dr == 'index' and params.get('index') == 'rle':
        return True                      # lossless tile-local run coding inside the fused kernel
    if dr == 'value':
        return value_ok                  # coded values + plain indices
    if dr == 'both' and params.get('index', 'bloom') == 'bloom':
        return pol_ok and value_ok
    return False


def plan_kwargs_from_params(params: dict) -> dict:
    """``params`` dict (the reference's ``--grace_config``) -> BucketPlan keyword arguments."""
    from ..codecs.bloom import canonical_policy
    dr = params.get('deepreduce')
    kw = dict(compress_ratio=params.get('compress_ratio', 0.01),
              index=(params.get('index', 'bloom') if dr in ('index', 'both') else None),
              value=(params.get('value', 'polyfit') if dr in ('value', 'both') else None),
              quantum_num=int(params.get('quantum_num', 127)),
              poly_degree=int(params.get('poly_degree', 5)),
              fpr=params.get('fpr', None),
              policy=canonical_policy(params.get('policy', 'leftmost')),
              min_numel=int(params.get('min_numel', spec.SMALL_TENSOR_NUMEL)),
              hint=bool(params.get('hint', True)))
    if params.get('compressor') == 'threshold':
        kw.update(sparsifier='threshold', threshold=float(params.get('threshold', 0.0)),
                  capacity_ratio=params.get('threshold_capacity', None))
    return kw


class DeepReduceDDP:
    def __init__(self, module: nn.Module, params: dict, *, bucket_cap_mb: float = 1e9, overlap: bool = True,
                 group=None, blocks_per_sm: int = 2, use_history: bool = True, background_thread: bool = True,
                 broadcast_parameters: bool = True, overlap_grid: int | None = None):
        self.module = module
        self.params = dict(params)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        self.device = self.named[0][1].device
        self.is_cuda = self.device.type == "cuda"
        self.dense = self.params.get('compressor', 'none') in ('none', None)
        self.fused = self.is_cuda and not self.dense and _fused_supported(self.params)
        self.overlap = overlap and self.is_cuda
        self.step_count = 0
        self.engines: List[BucketEngine] = []
        self.flat: List[torch.Tensor] = []
        self.bucket_of: Dict[int, int] = {}
        self.pending: List[int] = []
        self.sched = None
        self.grc = None
        self._handles = []
        self._exchange = True
        self._grad_views: Dict[int, torch.Tensor] = {}
        self._status_host = None          # pinned copies of the engines' status words (async check)
        self._status_event = None
        self.overlap_grid_cap = int(overlap_grid if overlap_grid is not None else (self.params.get('overlap_grid', 0) or 0))
        if self.world > 1 and broadcast_parameters:
            # replicas must start from the same weights: rank 0's parameters and buffers win (torch DDP does the same)
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t.data, 0, group=group)
        if self.fused or (self.dense and self.is_cuda):
            self._build_buckets(bucket_cap_mb, blocks_per_sm, use_history)
            if (self.fused and self.overlap and background_thread
                    and all(e.transport == "p2p" for e in self.engines)):     # the NCCL transport is issued from Python
                from .. import ops
                self.sched = ops.cuda_module().Scheduler(len(self.buckets))
            if self.overlap:
                self._install_hooks()
        else:
            self.grc = deepreduce_from_params(self.params)

    # ---- bucket construction ------------------------------------------------
    def _build_buckets(self, cap_mb, blocks_per_sm, use_history):
        cap = int(cap_mb * 1024 * 1024 / 4)
        order = list(reversed(self.named))           # roughly the order gradients become ready
        self.buckets: List[List] = []
        cur, cur_n = [], 0
        for n, p in order:
            if cur and cur_n + p.numel() > cap:
                self.buckets.append(cur)
                cur, cur_n = [], 0
            cur.append((n, p))
            cur_n += p.numel()
        if cur:
            self.buckets.append(cur)
        for b, items in enumerate(self.buckets):
            numels = [p.numel() for _, p in items]
            names = [n for n, _ in items]
            shapes = [tuple(p.shape) for _, p in items]
            owner = list(range(len(items)))
            if self.fused:
                # tensors whose bloom filter would not fit the kernel's SMEM staging buffer enter the plan as tile-aligned
                # chunks with their own top-k / filter ('split_numel': 'auto', the default; an int pins the chunk size,
                # 0 / None keeps every tensor whole and lets oversize filters be probed from L2)
                sn = self.params.get('split_numel', 'auto')
                if sn == 'auto':
                    uses_bloom = self.params.get('deepreduce') in ('index', 'both') and self.params.get('index', 'bloom') == 'bloom'
                    from .plan import auto_split_numel
                    sn = auto_split_numel(self.params.get('compress_ratio', 0.01), self.params.get('fpr', None),
                                          (160 if blocks_per_sm < 2 else 80) * 1024) if uses_bloom and self.params.get('compressor') == 'topk' else 0
                if sn:
                    from .plan import split_large
                    numels, names, shapes, owner = split_large(numels, names, shapes, int(sn))
            if self.fused:
                plan = BucketPlan(numels, names, shapes, **plan_kwargs_from_params(self.params))
                residual = self.params.get('memory', 'none') == 'residual'
                eng = BucketEngine(plan, device=self.device, group=self.group,
                                   beta=float(self.params.get('beta', 1.0)) if residual else 0.0,
                                   gamma=float(self.params.get('gamma', 1.0)), average=self.params.get('average', True),
                                   use_history=use_history, blocks_per_sm=blocks_per_sm)
                self.engines.append(eng)
                # re-cut the kernel's tile partitions from measured per-CTA phase times (collective; ~12 exchange steps
                # on synthetic gradients, state reset afterwards) — 'calibrate_partition': False keeps the static cut
                if self.params.get('calibrate_partition', True) and eng.cuts is not None:
                    try:
                        eng.calibrate_partition()
                    except (ValueError, ArithmeticError, IndexError) as e:      # host-side arithmetic only: the static
                        import warnings                                      # per-phase cut is a complete fallback
                        warnings.warn(f"deepreduce_b200: partition calibration skipped ({e!r}); using the static cut")
                        eng.cta_speeds = None
                        eng._set_cuts()
                        eng.resid.zero_(); eng.sel.zero_(); eng.grad.zero_()
                flat, views = eng.grad, eng.grad_views
            else:
                plan = BucketPlan(numels, names, shapes, index=None)
                flat = torch.zeros(plan.total_elems, dtype=torch.float32, device=self.device)
                views = plan.views(flat)
            self.flat.append(flat)
            first = {}
            for j, o in enumerate(owner):
                first.setdefault(o, plan.tensors[j])                # the first chunk of every parameter
            for i, (n, p) in enumerate(items):
                assert p.dtype == torch.float32, "flat buckets hold fp32 master gradients"
                t = first[i]
                seg = flat[t.elem_off:t.elem_off + p.numel()]        # chunks are tile multiples: contiguous
                # the gradient view mirrors the parameter's own (dense) layout — e.g. channels_last conv
                # weights — so fused optimizers see identical strides; the bucket is in storage order
                p.grad = seg.as_strided(p.size(), p.stride()) if _is_dense(p) else seg.view(p.shape)
                self._grad_views[id(p)] = p.grad
                self.bucket_of[id(p)] = b
        self._in_backward = False
        self._ready_count = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._bucket_size = [len(it) for it in self.buckets]

    def _install_hooks(self):
        for n, p in self.named:
            self._handles.append(p.register_post_accumulate_grad_hook(self._hook))

    def set_exchange_enabled(self, on: bool):
        """Gradient accumulation: hooks only launch the exchange on the last micro-step."""
        self._exchange = bool(on)

    def _hook(self, p):
        if not self._exchange:
            return
        b = self.bucket_of[id(p)]
        self._ready_count[b] += 1
        if self._ready_count[b] == self._bucket_size[b]:
            self._in_backward = True
            try:
                self._launch_bucket(b)
            finally:
                self._in_backward = False

    def _launch_bucket(self, b):
        self._ready_count[b] = 0
        self._launched[b] = True
        if self.fused:
            eng = self.engines[b]
            # the engine's own step counter, never reset: peer flags carry the epoch, so an epoch that was already used
            # (e.g. by calibrate_partition's synthetic steps, or by a self-check step between training steps) would
            # let the flag waits pass before the peers have written their slots
            eng.epoch = eng.epoch + 1
            # buckets that become ready while backward is still running are launched with a capped grid so that the
            # persistent exchange kernel does not take every SM from cuDNN / cuBLAS; the bucket launched from
            # finish() (nothing left to overlap with) gets the whole GPU
            more_to_come = not all(self._launched)
            if self.sched is not None and more_to_come:
                # backward is still running: hand the bucket to the C++ launch thread (high-priority side stream), with a
                # capped grid so that the persistent kernel does not take every SM from cuDNN / cuBLAS
                eng.ctx.set_grid_cap(self.overlap_grid_cap if self.overlap_grid_cap > 0 else 0)
                self.sched.submit(b, eng.ctx, eng.epoch)
            else:
                # the LAST bucket (or the only one): nothing is left to overlap with, so every microsecond until the
                # kernel starts is exposed — launch inline on the current stream with the whole GPU instead of paying
                # the thread hand-off + event round trip (measured: 54.30 -> 54.00 ms/step, profiles/overlap_sweep.md)
                eng.ctx.set_grid_cap(0)
                if self.sched is not None and len(self.buckets) > 1 and self.world > 1:
                    # cross-rank ordering: every rank must run its bucket kernels in the same order — a full-grid
                    # kernel that spins on peer flags would otherwise keep this rank's earlier (side-stream) buckets
                    # from starting while the peers wait for exactly those (deadlock until the peer watchdog fires).
                    # Make this stream wait for everything handed to the launch thread first.
                    self.sched.wait_all()
                eng.step(eng.epoch)
        else:
            if self.world > 1:
                self.pending.append(dist.all_reduce(self.flat[b], group=self.group, async_op=True))

    # ---- per-step API ---------------------------------------------------------
    def zero_grad(self):
        if self.flat:
            for f in self.flat:
                f.zero_()
        else:
            for _, p in self.named:
                p.grad = None

    def finish(self):
        """Call after backward, before optimizer.step(): gradients become the
        cross-rank aggregate."""
        if self.grc is not None:
            for n, p in self.named:
                if p.grad is not None:
                    p.grad = self.grc.step(p.grad, n).view_as(p)
        elif self.fused:
            self._check_grad_views()
            for b in range(len(self.buckets)):        # not overlapped, or a parameter received no gradient this step
                if not self._launched[b]:
                    self._launch_bucket(b)
            if self.sched is not None:
                self.sched.wait_all()
            self._launched = [False] * len(self.buckets)
        else:
            self._check_grad_views()
            for b in range(len(self.buckets)):
                if not self._launched[b]:
                    self._launch_bucket(b)
            self._launched = [False] * len(self.buckets)
            for w in self.pending:
                w.wait()
            self.pending = []
            if self.world > 1 and self.params.get('average', True):
                for f in self.flat:
                    f.div_(self.world)
        self.step_count += 1

    def _check_grad_views(self):
        """``p.grad`` must still be the view into the flat bucket: ``optimizer.zero_grad(set_to_none=True)`` (torch's
        default) or ``p.grad = None`` detaches it, autograd then allocates a fresh gradient and the exchange would
        silently run on stale bucket contents.  Re-attach: copy what autograd produced into the bucket and point
        ``p.grad`` back at the view."""
        for _, p in self.named:
            v = self._grad_views.get(id(p))
            if v is None or p.grad is v:
                continue
            if p.grad is None:
                v.zero_()                              # no gradient this step: the bucket slice must not keep old data
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            p.grad = v

    def check_async(self):
        """Per-step failure detection without a host sync: enqueue a copy of every engine's status word to pinned host
        memory, and inspect the copy enqueued by the PREVIOUS call once its event has completed (it has, a step
        later).  Raises like :meth:`check` one step after a watchdog fired."""
        if not self.engines:
            return
        if self._status_host is not None and self._status_event.query():
            st = self._status_host
            for b in range(len(self.engines)):
                if int(st[b, 0]) != 0:
                    from .engine import STATUS_NAMES
                    raise RuntimeError(f"[rank {self.rank}/{self.world}] bucket {b} (step {self.step_count}): deepreduce "
                                       f"engine error: {STATUS_NAMES.get(int(st[b, 0]), int(st[b, 0]))} (aux={int(st[b, 1])})")
        if self._status_host is None:
            self._status_host = torch.zeros(len(self.engines), 8, dtype=torch.int32).pin_memory()
            self._status_event = torch.cuda.Event()
        for b, e in enumerate(self.engines):
            self._status_host[b].copy_(e.status, non_blocking=True)
        self._status_event.record()

    def check(self):
        """Read the engines' device status words (one small D2H each); raises with rank and bucket on a watchdog."""
        for b, e in enumerate(self.engines):
            try:
                e.check_status()
            except RuntimeError as err:
                raise RuntimeError(f"[rank {self.rank}/{self.world}] bucket {b} "
                                   f"({len(self.buckets[b])} tensors, step {self.step_count}): {err}") from err

    # ---- accounting -------------------------------------------------------------
    def wire_bytes_per_step(self) -> int:
        if self.fused:
            return sum(e.plan.wire_bytes() for e in self.engines)
        if self.dense:
            return sum(f.numel() * 4 for f in self.flat)
        return int(self.grc.bytes_sent / max(self.step_count, 1))

    def stage2_bytes_per_step(self) -> int:
        """Sharded decode (W > 1): bytes of the second in-kernel exchange this rank sent last step — its decoded slice
        as (index, value) pairs to each of the W-1 peers — read from the live entry count in the stage-2 header.
        Together with ``wire_bytes_per_step`` this is everything a rank puts on NVLink per step."""
        if not self.fused or self.world == 1:
            return 0
        torch.cuda.synchronize(self.device)
        return sum(e.stage2_bytes() for e in self.engines)

    def dense_bytes(self) -> int:
        return sum(p.numel() * 4 for _, p in self.named)

    def exchange_stats(self) -> dict:
        """Device-side counters of the last exchanged step, summed over the buckets (fused path): shipped
        coordinates, bloom positives / false positives, value / index / header bytes (``BucketEngine.stats``).
        Call it after ``finish()`` (i.e. after a training step); it synchronises the device, so every N steps."""
        if not self.fused:
            return {"wire_bytes": self.wire_bytes_per_step(), "dense_bytes": self.dense_bytes()}
        torch.cuda.synchronize(self.device)
        tot: dict = {}
        for e in self.engines:
            for k, v in e.stats()["total"].items():
                tot[k] = tot.get(k, 0) + v
        tot["relative_volume"] = tot["wire_bytes"] / max(1, tot["dense_bytes"])
        return tot

    # ---- checkpoint / resume -------------------------------------------
    def state_dict(self):
        if self.fused:
            return {"step": self.step_count, "engines": [e.state_dict() for e in self.engines]}
        if self.grc is not None:
            return {"step": self.step_count, "memory": self.grc.memory.state_dict()}
        return {"step": self.step_count}

    def load_state_dict(self, state):
        self.step_count = int(state.get("step", 0))
        if self.fused:
            for e, s in zip(self.engines, state["engines"]):
                e.load_state_dict(s)
        elif self.grc is not None and "memory" in state:
            self.grc.memory.load_state_dict(state["memory"], device=self.device)

    def close(self):
        for h in self._handles:
            h.remove()
        if self.sched is not None:
            self.sched.shutdown()
            self.sched = None
        for e in self.engines:
            e.close()
