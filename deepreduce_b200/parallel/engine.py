"""Python driver + torch oracle for the fused bucket engine (``ops/csrc/engine.cu``).

``BucketEngine`` owns the flat gradient / residual buffers of one bucket, the
select state, the symmetric arena (CUDA IPC, or NVLS-capable symmetric memory
with ``DR_NVLS=1``, for W>1) and the C++ ``Engine`` context; ``step()`` launches
the single persistent kernel that does sparsify → encode → P2P push → sharded
decode → slice exchange for the whole bucket.

``engine_oracle`` is the plain-PyTorch specification of the same step,
including the wire format of the slot, used by the tests (GPU kernels vs
oracle: slots bit-exact, dense output allclose).

Parity with the reference (hangxu0304/DeepReduce), per tensor of the bucket:
  * residual compensate / update  — GRACE ResidualMemory, TF twin tensorflow/deepreduce.py:31-52;
  * top-k select                  — GRACE TopK (``torch.topk(abs(x), k)``), here a 22-bit-threshold radix select;
  * bloom insert / universe query / policy / FP-aware re-gather — pytorch/deepreduce.py:457-492,505-529;
  * 'both' (index first, value codec on the re-gathered values, mapping back) — :250-302;
    polyfit segments / fit / restore — :341-425; QSGD — :861-907;
  * per-rank decode + aggregate (+ average) — GRACE Allgather communicator, reference README.md:37;
  * lossless run-length index (``'index': 'rle'``) — :805-846, here tile-local (see ``ops/csrc/plan.h``).
"""
from __future__ import annotations

import os

from collections import OrderedDict
from typing import Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .. import spec
from ..codecs.bloom import bloom_insert_oracle, bloom_query_oracle
from .plan import update_cta_speeds
from .plan import (ARENA_HDR_WORDS, DYN_WORDS, HIST_BINS, MODE_BLOOM, MODE_RLE, NUM_HIST, POLICY_ID,
                   SLOT_HEADER_WORDS, BucketPlan, rle_stream_words)

(PH_ACCUM, PH_FALLBACK, PH_HIST2, PH_INSERT, PH_QUERY, PH_EMIT, PH_RANK_HIST, PH_RANK_SCAN, PH_RANK_SCATTER,
 PH_RANK_EXACT, PH_FIT, PH_FIX, PH_PUSH, PH_SIGNAL, PH_EXPAND, PH_DECODE, PH_COMPACT, PH_PUSH2, PH_SIGNAL2,
 PH_SCATTER, PH_END) = range(21)
MAGIC = 0xD33B2000
STATUS_NAMES = {0: "ok", 1: "(unused)", 2: "peer flag watchdog", 3: "select resolve failed", 4: "grid barrier watchdog", 5: "TMA mbarrier watchdog",
                6: "stage-2 slot overflow (sharded decode)"}


# ---------------------------------------------------------------------------
# oracle
# ---------------------------------------------------------------------------
def rle_pack12(pos: np.ndarray, cap: int) -> np.ndarray:
    """12-bit fields, LSB-first, entry j at bit 12*j (the kModeRle index stream)."""
    out = np.zeros(rle_stream_words(cap), dtype=np.uint64)
    j = np.arange(pos.size, dtype=np.int64)
    bit = 12 * j
    w, sh = bit >> 5, (bit & 31).astype(np.uint64)
    v = pos.astype(np.uint64) << sh                       # up to 43 bits
    np.bitwise_or.at(out, w, v & np.uint64(0xFFFFFFFF))
    np.bitwise_or.at(out, w + 1, v >> np.uint64(32))
    return out.astype(np.uint32)


def rle_unpack12(stream: np.ndarray, n: int) -> np.ndarray:
    j = np.arange(n, dtype=np.int64)
    bit = 12 * j
    w, sh = bit >> 5, (bit & 31).astype(np.uint64)
    s64 = stream.astype(np.uint64)
    both = s64[w] | (s64[np.minimum(w + 1, s64.size - 1)] << np.uint64(32))
    return ((both >> sh) & np.uint64(0xFFF)).astype(np.int64)


def _abs_keys(x: torch.Tensor) -> torch.Tensor:
    return x.contiguous().view(torch.int32).to(torch.int64) & 0x7FFFFFFF


def select_topk_oracle(acc: torch.Tensor, k: int):
    """The engine's selection rule (normative, see ops/csrc/plan.h): with key = |x| bit pattern and
    T22 = (K-th largest key) >> 9, select every element with (key >> 9) >= max(T22, 1).  At least K
    elements (when K non-zeros exist) plus the few sharing the threshold's 22-bit prefix; exact zeros
    never.  Returns (ascending indices, threshold key lower bound)."""
    keys = _abs_keys(acc)
    kth = int(torch.topk(keys, k, sorted=True).values[-1].item())
    T22 = max(kth >> 9, 1)
    sel = torch.nonzero((keys >> 9) >= T22).flatten()
    return sel, T22 << 9


def select_threshold_oracle(acc: torch.Tensor, fixed_thr: int):
    """'threshold' sparsifier (GRACE: |x| > threshold): every element whose |x| bit pattern is >= fixed_thr."""
    return torch.nonzero(_abs_keys(acc) >= fixed_thr).flatten(), int(fixed_thr)


def random_policy_filter(pos: torch.Tensor, n_ins: int, limit: int, epoch: int, salt: int, T: Optional[int] = None):
    """'random' policy of the fused engine: keep the positives x with policy_hash(x, policy_seed(step, tensor)) <= T,
    T = floor(2^32 * min(n_ins, limit) / n_pos) (0xFFFFFFFF = keep all when nothing has to go).  Sender: T from the
    counts; receiver: T from the header (pass it in).  Returns (surviving positives, T)."""
    if T is None:
        n_pos, target = int(pos.numel()), min(int(n_ins), int(limit))
        T = 0xFFFFFFFF if n_pos <= target else (target << 32) // n_pos
    if T != 0xFFFFFFFF:
        pos = pos[spec.policy_hash(pos, spec.policy_seed(epoch, salt)) <= T]
    return pos, T


def encode_tensor_oracle(tp, acc: torch.Tensor, slot: np.ndarray, t_index: int, policy: str, seed: int, epoch: int = 1):
    """Encode one tensor into `slot` (uint32 numpy view); returns new residual."""
    sel_topk, T = (select_threshold_oracle(acc, tp.fixed_thr) if getattr(tp, "fixed_thr", 0)
                   else select_topk_oracle(acc, tp.k))
    dyn = SLOT_HEADER_WORDS + DYN_WORDS * t_index
    resid = acc.clone()
    if tp.mode == MODE_BLOOM:
        words = bloom_insert_oracle(sel_topk, tp.n_hash, tp.m_bits, seed)
        pos = bloom_query_oracle(words, tp.numel, tp.n_hash, tp.m_bits, seed)
        if tp.off_hint:
            # occupancy hint: bit g of the tile's 128-bit field <=> the 32-element group g holds a selected element;
            # positives outside occupied groups are dropped on both sides (they can only be false positives)
            groups = torch.unique(sel_topk // 32)
            occ = torch.zeros((tp.numel + 31) // 32, dtype=torch.bool)
            occ[groups] = True
            pos = pos[occ[pos // 32]]
            hint = np.zeros(4 * tp.n_tiles, dtype=np.uint32)
            g = groups.numpy().astype(np.int64)
            tile, gi = g // 128, g % 128            # group gi of a tile covers elements [32*gi, 32*gi+32) of that tile...
            # ...in the kernel's (slot c, warp w) order: element e = c*512 + w*32 + lane  ->  gi = e // 32 = c*16 + w
            np.bitwise_or.at(hint, tile * 4 + gi // 32, (np.uint32(1) << (gi % 32).astype(np.uint32)))
            slot[tp.off_hint:tp.off_hint + 4 * tp.n_tiles] = hint
        limit = tp.val_cap if policy == "p0" else min(tp.k, tp.val_cap)
        n_pos = int(pos.numel())
        if policy == "random":
            # P1: seeded Bernoulli draw of rate target/n_pos over the positives (ops/csrc/engine.cu::policy_filter); the
            # acceptance threshold travels in header word 2, the receiver repeats the test on its own positives
            # (T now names the acceptance threshold: that is what header word 2 carries for this policy)
            pos, T = random_policy_filter(pos, int(sel_topk.numel()), limit, epoch, tp.salt)
        sel = pos[:limit]
        slot[tp.off_filter:tp.off_filter + tp.n_filter_words] = words.cpu().numpy().view(np.uint32)
        starts = torch.arange(tp.n_tiles, dtype=torch.int64) * spec.TILE
        slot[tp.off_prefix:tp.off_prefix + tp.n_tiles] = torch.searchsorted(sel.cpu(), starts).numpy().astype(np.uint32)
        cutoff = int(sel[-1].item()) if int(pos.numel()) >= limit and limit > 0 else 0xFFFFFFFF
    elif tp.mode == MODE_RLE:
        n_pos = int(sel_topk.numel())
        limit = tp.val_cap
        sel = sel_topk[:limit]
        s_np = sel.cpu().numpy().astype(np.int64)
        cnt = np.zeros(((tp.n_tiles + 1) // 2) * 2, dtype=np.uint16)
        cnt[:tp.n_tiles] = np.bincount(s_np // spec.TILE, minlength=tp.n_tiles).astype(np.uint16)
        slot[tp.off_prefix:tp.off_prefix + (tp.n_tiles + 1) // 2] = cnt.view(np.uint32)
        slot[tp.off_idx:tp.off_idx + rle_stream_words(tp.val_cap)] = rle_pack12(s_np % spec.TILE, tp.val_cap)
        cutoff = int(sel[-1].item()) if n_pos >= limit else 0xFFFFFFFF
    else:
        n_pos = int(sel_topk.numel())
        limit = tp.val_cap
        sel = sel_topk[:limit]                 # capacity K: the left-most of the (>= K) selected
        slot[tp.off_idx:tp.off_idx + sel.numel()] = sel.cpu().numpy().astype(np.uint32)
        cutoff = int(sel[-1].item()) if n_pos >= limit else 0xFFFFFFFF
    vals = acc[sel].float()
    if tp.vmode == 1:
        # 'both': stable descending sort -> rank map; per-segment Gram fit; fitted values are what is shipped,
        # and the residual keeps (value - fitted)
        from ..codecs.polyfit import MAX_SEGMENTS, get_segments, polyfit_eval_oracle, polyfit_fit_oracle
        n = int(sel.numel())
        order = torch.sort(vals, descending=True, stable=True).indices
        rank = torch.empty(n, dtype=torch.int64)
        rank[order] = torch.arange(n)
        num_pos = int((vals > 0).sum())
        segs = get_segments(n, num_pos)
        coef = polyfit_fit_oracle(vals[order], segs, tp.poly_degree)
        fitted = polyfit_eval_oracle(coef, segs, tp.poly_degree)[rank] if n else vals
        nc = MAX_SEGMENTS * (tp.poly_degree + 1)
        slot[tp.off_coef:tp.off_coef + nc] = coef.numpy().view(np.uint32)
        slot[tp.off_coef + nc] = num_pos
        slot[tp.off_coef + nc + 1] = n
        if tp.rank_u32:
            slot[tp.off_rankmap:tp.off_rankmap + n] = rank.numpy().astype(np.uint32)
        else:
            r16 = np.zeros(((n + 1) // 2) * 2, dtype=np.uint16)
            r16[:n] = rank.numpy().astype(np.uint16)
            slot[tp.off_rankmap:tp.off_rankmap + (n + 1) // 2] = r16.view(np.uint32)
        resid[sel] = vals - fitted
        vals = fitted
    elif tp.vmode == 2:
        from ..codecs.qsgd import qsgd_decode_oracle, qsgd_encode_oracle
        n = int(sel.numel())
        q = int(tp.poly_degree)
        lvl, norms = qsgd_encode_oracle(vals, q, 512, 0x51ED + epoch)
        dec = qsgd_decode_oracle(lvl, norms, q, 512) if n else vals
        nb = (n + 511) // 512
        slot[tp.off_coef:tp.off_coef + nb] = norms.float().numpy().view(np.uint32)
        if tp.rank_u32:                      # quantum_num >= 128: 16-bit levels
            l16 = np.zeros(((n + 1) // 2) * 2, dtype=np.int16)
            l16[:n] = lvl.numpy().astype(np.int16)
            slot[tp.off_rankmap:tp.off_rankmap + (n + 1) // 2] = l16.view(np.uint32)
        else:
            l8 = np.zeros(((n + 3) // 4) * 4, dtype=np.int8)
            l8[:n] = lvl.numpy().astype(np.int8)
            slot[tp.off_rankmap:tp.off_rankmap + (n + 3) // 4] = l8.view(np.uint32)
        resid[sel] = vals - dec
        vals = dec
    else:
        slot[tp.off_vals:tp.off_vals + sel.numel()] = vals.cpu().numpy().view(np.uint32)
        resid[sel] = 0
    slot[dyn + 0] = sel.numel()
    slot[dyn + 1] = cutoff
    slot[dyn + 2] = T
    slot[dyn + 3] = n_pos
    return resid, sel, vals


def engine_oracle(plan: BucketPlan, grads: Sequence[torch.Tensor], resids: Sequence[torch.Tensor], *, beta=1.0,
                  gamma=1.0, average=True, seed=spec.DEFAULT_SEED, epoch=1):
    """One bucket step for W ranks on the CPU.  grads/resids: per-rank flat
    buffers (plan.total_elems).  Returns (dense_out, new_resids, slots)."""
    W = len(grads)
    out = torch.zeros(plan.total_elems, dtype=torch.float32)
    new_resids, slots = [], []
    for r in range(W):
        g = grads[r].detach().cpu().float()
        res = resids[r].detach().cpu().float()
        acc_flat = beta * res + gamma * g if beta != 0.0 else gamma * g
        slot = np.zeros(plan.payload_words, dtype=np.uint32)
        slot[0:5] = [MAGIC, epoch, len(plan.tensors), plan.payload_words, r]
        nres = acc_flat.clone()
        for ti, tp in enumerate(plan.tensors):
            seg = slice(tp.elem_off, tp.elem_off + tp.numel)
            resid_t, sel, vals = encode_tensor_oracle(tp, acc_flat[seg], slot, ti, plan.policy, seed, epoch)
            nres[seg] = resid_t
            out[seg].index_add_(0, sel, vals)
        new_resids.append(nres)
        slots.append(slot)
    if average:
        out = out / W
    return out, new_resids, slots


def decode_slot_oracle(plan: BucketPlan, slot, *, seed=spec.DEFAULT_SEED) -> torch.Tensor:
    """Receiver-side specification: rebuild one sender's dense contribution (flat, ``plan.total_elems``, unscaled)
    from NOTHING but the plan and the words of its slot — what ``phase_decode`` does per (tile, sender).  Together
    with ``encode_tensor_oracle`` this pins the wire format from both ends: the tests check that the sum of the
    decoded slots equals the aggregate ``engine_oracle`` builds on the sender side."""
    a = slot.detach().cpu().numpy().view(np.uint32) if torch.is_tensor(slot) else np.asarray(slot, dtype=np.uint32)
    assert int(a[0]) == MAGIC and int(a[2]) == len(plan.tensors), "not a slot of this plan"
    out = torch.zeros(plan.total_elems, dtype=torch.float32)
    for ti, t in enumerate(plan.tensors):
        d0 = SLOT_HEADER_WORDS + DYN_WORDS * ti
        n_sel, cutoff = int(a[d0]), int(a[d0 + 1])
        if n_sel == 0:
            continue
        if t.mode == MODE_BLOOM:
            words = torch.from_numpy(a[t.off_filter:t.off_filter + t.n_filter_words].view(np.int32).copy())
            pos = bloom_query_oracle(words, t.numel, t.n_hash, t.m_bits, seed)          # ascending positives of the universe
            if t.off_hint:                                                               # only inside occupied 32-element groups
                hint = a[t.off_hint:t.off_hint + 4 * t.n_tiles]
                grp = pos // 32                                                          # group id: tile*128 + (e // 32)
                tile, gi = grp // 128, grp % 128
                bit = (torch.from_numpy(hint.astype(np.int64))[tile * 4 + gi // 32] >> (gi % 32)) & 1
                pos = pos[bit.bool()]
            if plan.policy == "random":
                pos, _ = random_policy_filter(pos, 0, 0, int(a[1]), t.salt, T=int(a[d0 + 2]))
            if cutoff != 0xFFFFFFFF:
                pos = pos[pos <= cutoff]
            idx = pos[:n_sel]
            # the per-tile prefix table must agree with what the receiver recomputes (the kernel starts ranks from it)
            starts = torch.arange(t.n_tiles, dtype=torch.int64) * spec.TILE
            pre = torch.from_numpy(a[t.off_prefix:t.off_prefix + t.n_tiles].astype(np.int64))
            assert torch.equal(torch.minimum(torch.searchsorted(pos, starts), torch.tensor(n_sel)), pre), t.name
        elif t.mode == MODE_RLE:
            cnt = a[t.off_prefix:t.off_prefix + (t.n_tiles + 1) // 2].view(np.uint16)[:t.n_tiles].astype(np.int64)
            assert int(cnt.sum()) == n_sel, t.name
            local = rle_unpack12(a[t.off_idx:t.off_idx + rle_stream_words(t.val_cap)], n_sel)
            idx = torch.from_numpy(np.repeat(np.arange(t.n_tiles, dtype=np.int64), cnt) * spec.TILE + local)
        else:
            idx = torch.from_numpy(a[t.off_idx:t.off_idx + n_sel].astype(np.int64))
        n = int(idx.numel())
        if t.vmode == 1:
            from ..codecs.polyfit import MAX_SEGMENTS, get_segments, polyfit_eval_oracle
            nc = MAX_SEGMENTS * (t.poly_degree + 1)
            coef = torch.from_numpy(a[t.off_coef:t.off_coef + nc].view(np.float32).copy())
            num_pos, n_fit = int(a[t.off_coef + nc]), int(a[t.off_coef + nc + 1])
            curve = polyfit_eval_oracle(coef, get_segments(n_fit, num_pos), t.poly_degree)
            if t.rank_u32:
                rank = a[t.off_rankmap:t.off_rankmap + n].astype(np.int64)
            else:
                rank = a[t.off_rankmap:t.off_rankmap + (n + 1) // 2].view(np.uint16)[:n].astype(np.int64)
            vals = curve[torch.from_numpy(rank)]
        elif t.vmode == 2:
            from ..codecs.qsgd import qsgd_decode_oracle
            norms = torch.from_numpy(a[t.off_coef:t.off_coef + (n + 511) // 512].view(np.float32).copy())
            if t.rank_u32:
                lvl = torch.from_numpy(a[t.off_rankmap:t.off_rankmap + (n + 1) // 2].view(np.int16)[:n].astype(np.int64))
            else:
                lvl = torch.from_numpy(a[t.off_rankmap:t.off_rankmap + (n + 3) // 4].view(np.int8)[:n].astype(np.int64))
            vals = qsgd_decode_oracle(lvl, norms, int(t.poly_degree), 512)
        else:
            vals = torch.from_numpy(a[t.off_vals:t.off_vals + n].view(np.float32).copy())
        out[t.elem_off:t.elem_off + t.numel].index_add_(0, idx, vals.float())
    return out


# ---------------------------------------------------------------------------
# engine
# ---------------------------------------------------------------------------
def stats_from_slot(plan: BucketPlan, slot) -> dict:
    """Per-step counters of one sender, read from the 4-word dynamic header every tensor carries in the slot
    (``{n_sel, cutoff, thr_bits, n_pos}``, written on the device by the emit phase — SURVEY §5 "per-step counters
    accumulated on device"): shipped coordinates, bloom positives / false positives among the universe, the magnitude
    threshold of the selection, and the wire bytes by component (static, from the plan)."""
    a = slot.detach().cpu().numpy().view(np.uint32) if torch.is_tensor(slot) else np.asarray(slot, dtype=np.uint32)
    per, tot = [], {"k": 0, "n_sel": 0, "n_pos": 0, "false_pos": 0, "value_bytes": 0, "index_bytes": 0}
    for ti, t in enumerate(plan.tensors):
        d0 = SLOT_HEADER_WORDS + DYN_WORDS * ti
        n_sel, cutoff, thr_bits, n_pos = (int(x) for x in a[d0:d0 + 4])
        thr = float(np.array([thr_bits], dtype=np.uint32).view(np.float32)[0])
        if t.vmode == 1:
            vbytes = 4 * (22 * (t.poly_degree + 1) + 2) + (4 if t.rank_u32 else 2) * t.val_cap
        elif t.vmode == 2:
            vbytes = 4 * ((t.val_cap + 511) // 512) + t.val_cap * (2 if t.rank_u32 else 1)
        else:
            vbytes = 4 * t.val_cap
        if t.mode == MODE_BLOOM:
            ibytes = 4 * (t.n_filter_words + t.n_tiles + (4 * t.n_tiles if t.off_hint else 0))
            false_pos = max(0, n_pos - min(t.k, n_pos))      # positives beyond the K inserted (upper bound under 22-bit ties)
        elif t.mode == MODE_RLE:
            ibytes = 4 * ((t.n_tiles + 1) // 2 + rle_stream_words(t.val_cap))
            false_pos = 0
        else:
            ibytes, false_pos = 4 * t.val_cap, 0
        row = {"name": t.name, "numel": t.numel, "k": t.k, "n_sel": n_sel, "n_pos": n_pos, "false_pos": false_pos,
               "threshold": thr, "cutoff": None if cutoff == 0xFFFFFFFF else cutoff,
               "value_bytes": vbytes, "index_bytes": ibytes}
        if plan.policy == "random" and t.mode == MODE_BLOOM:   # header word 2 is the policy's acceptance threshold here
            row["threshold"] = None
            row["accept_rate"] = 1.0 if thr_bits == 0xFFFFFFFF else thr_bits / 2.0 ** 32
        per.append(row)
        for key in tot:
            tot[key] += row[key]
    tot["header_bytes"] = 4 * (SLOT_HEADER_WORDS + DYN_WORDS * len(plan.tensors))
    tot["wire_bytes"] = plan.wire_bytes()
    tot["dense_bytes"] = plan.dense_bytes()
    tot["relative_volume"] = tot["wire_bytes"] / max(1, tot["dense_bytes"])
    return {"tensors": per, "total": tot}


class BucketEngine:
    """One flat bucket + its fused exchange kernel."""

    def __init__(self, plan: BucketPlan, device=None, group=None, *, beta: float = 1.0, gamma: float = 1.0,
                 average: bool = True, use_history: bool = True, blocks_per_sm: int = 2,
                 seed: int = spec.DEFAULT_SEED, spin_limit: int = 20_000_000, world: Optional[int] = None,
                 rank: Optional[int] = None, filter_smem_bytes: Optional[int] = None, use_tma: bool = True,
                 hist_shift: int = 22, shard: Optional[bool] = None, transport: Optional[str] = None,
                 peer_timeout_ms: Optional[int] = None, fault: int = 0):
        from .. import ops
        self.mod = ops.cuda_module()
        self.plan = plan
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.group = group
        if world is None:
            world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
            rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
        self.world, self.rank = int(world), int(rank or 0)
        self.beta, self.gamma, self.average = float(beta), float(gamma), bool(average)
        self.epoch = 0
        # sharded decode (W > 1): each rank decodes 1/W of the tiles for all senders, then the exact slices are
        # exchanged by a second in-kernel push.  DR_SHARD=0 restores the every-rank-decodes-everything path.
        self.shard = (os.environ.get("DR_SHARD", "1") != "0") if shard is None else bool(shard)
        # transport of the compressed slots between ranks: 'p2p' = in-kernel stores into peer-mapped arenas (one
        # NVLink/NVSwitch domain, the default), 'nccl' = encode phases -> ONE in-place NCCL all_gather of the slots per
        # bucket -> decode phases (ranks on different hosts, or no peer access).  None: DR_TRANSPORT, else auto.
        self.transport = self._pick_transport(transport)
        if self.transport == "nccl":
            self.shard = False
        dev = self.device
        nT, nt = len(plan.tensors), plan.n_tiles
        with torch.cuda.device(dev):
            self.grad = torch.zeros(plan.total_elems, dtype=torch.float32, device=dev)
            self.resid = torch.zeros(plan.total_elems, dtype=torch.float32, device=dev)
            self.tensor_table = plan.tensor_table().to(dev)
            self.tile_table = plan.tile_table().to(dev)
            self.hist = torch.zeros(NUM_HIST * nT * HIST_BINS, dtype=torch.int32, device=dev)
            self.hist_total = torch.zeros(NUM_HIST * nT, dtype=torch.int32, device=dev)
            self.sel = torch.zeros(nT * 8, dtype=torch.int32, device=dev)
            # per-tile positive counts + two per-tensor counters (positives, inserted) used by the random policy
            self.tile_count = torch.zeros(nt + 2 * len(plan.tensors), dtype=torch.int32, device=dev)
            # scratch of the candidate / bitmask pipeline (ops/csrc/engine.cu): one mask word per 32-element group
            # (own positives, decode scratch) and the candidate lists — (key, in-tile offset) of every element above
            # the history bound, 256 slots per (tile, warp) at a fixed place (8 B per element, touched sparsely)
            self.pos_mask = torch.zeros(nt * 128, dtype=torch.int32, device=dev)
            # decode scratch: one mask slot per (sender, tile of the slice this rank decodes)
            span_max = (nt // self.world + 1) if (self.shard and self.world > 1) else nt
            self.dec_mask = torch.zeros(self.world * span_max * 128, dtype=torch.int32, device=dev)
            self.cand = torch.empty(nt * 4096 * 2, dtype=torch.int32, device=dev)
            self.cand_cnt = torch.zeros(nt * 16, dtype=torch.int32, device=dev)
            self.barrier = torch.zeros(16, dtype=torch.int32, device=dev)
            self.status = torch.zeros(8, dtype=torch.int32, device=dev)
            self._setup_arena()
            self.ctx = self.mod.Engine(
                self.tensor_table.data_ptr(), self.tile_table.data_ptr(), nT, nt, plan.slot_words, plan.payload_words,
                self.grad.data_ptr(), self.resid.data_ptr(), self.hist.data_ptr(), self.hist_total.data_ptr(),
                self.sel.data_ptr(), self.tile_count.data_ptr(),
                self.barrier.data_ptr(), self.status.data_ptr(), self.arena_ptrs, self.rank, self.world)
            # equal-COST tile ranges per CTA (DR_BALANCE=0: equal counts)
            self.balanced = os.environ.get("DR_BALANCE", "1") != "0"
            if self.balanced:
                seg_c, single_c = (float(x) for x in os.environ.get("DR_SEG_COST", "6.0,2.0").split(","))
                self.cost_prefix = plan.cost_prefix(seg_c, single_c).to(dev)
                self.ctx.set_cost_prefix(self.cost_prefix.data_ptr())
            self.cta_speeds = None          # [4, grid] relative CTA speeds per phase class (calibrate_partition)
            self.cuts = None
            self.ctx.set_scratch(self.pos_mask.data_ptr(), self.dec_mask.data_ptr(), self.cand.data_ptr(),
                                 self.cand_cnt.data_ptr())
            # a peer that does not signal within this wall time is fatal (status 2, output poisoned, see wait_flags)
            if peer_timeout_ms is None:
                peer_timeout_ms = int(os.environ.get("DR_PEER_TIMEOUT_MS", "120000"))
            self.ctx.set_peer_timeout_ms(int(peer_timeout_ms))
            self.ctx.set_fault(int(fault))
            # DR_DETERMINISTIC=1: rank-ordered decode sums (bit-reproducible run to run); default: independent
            # (sender, tile) work items with RED.ADD.F32 (every rank still ends with identical bits: owner computes)
            self.deterministic = os.environ.get("DR_DETERMINISTIC", "0") == "1"
            self.ctx.set_deterministic(int(self.deterministic))
            scale = (1.0 / self.world) if average else 1.0
            if filter_smem_bytes is None:      # <1>: 128 regs, 1 CTA/SM; <2>: 64 regs, 2 CTAs/SM
                filter_smem_bytes = 160 * 1024 if blocks_per_sm < 2 else 80 * 1024
            if self.shard and self.world > 1:
                cap, s2w = plan.stage2_layout(self.world)
                self.ctx.set_shard(1, s2w, cap)
            if getattr(self, "multicast_ptr", 0):
                self.ctx.set_multicast(self.multicast_ptr)
            self.ctx.set_has_rle(int(any(t.mode == MODE_RLE for t in plan.tensors)))
            ids, n_poly, tasks, n_tasks = plan.poly_tables()
            self.poly_ids, self.poly_tasks = ids.to(dev), tasks.to(dev)
            from .plan import RANK_BINS
            tot = max(int(plan.poly_total), 1)
            self.poly_bins = torch.zeros(max(n_poly, 1) * 2 * RANK_BINS, dtype=torch.int32, device=dev)
            self.bucket_val = torch.zeros(tot, dtype=torch.float32, device=dev)
            self.bucket_pos = torch.zeros(tot, dtype=torch.int32, device=dev)
            self.expand_buf = torch.zeros(self.world * tot, dtype=torch.float32, device=dev)
            self.ctx.set_poly(self.poly_ids.data_ptr(), n_poly, self.poly_tasks.data_ptr(), n_tasks,
                              self.poly_bins.data_ptr(), self.bucket_val.data_ptr(), self.bucket_pos.data_ptr(),
                              self.expand_buf.data_ptr(), int(plan.poly_total))
            self.ctx.configure(self.beta, self.gamma, scale, int(seed), POLICY_ID[plan.policy], int(use_history),
                               int(spin_limit), int(blocks_per_sm), int(filter_smem_bytes), int(use_tma), int(hist_shift))
            # per-phase-class partitions (own cost weights per class; per-CTA speeds once calibrated); DR_CUTS=0: the
            # single cost prefix above for every phase
            if self.balanced and os.environ.get("DR_CUTS", "1") != "0":
                self._set_cuts()
        self.grad_views = plan.views(self.grad)

    def _set_cuts(self):
        g = self.grid()
        self.cuts = self.plan.phase_cuts(g, self.cta_speeds).contiguous().to(self.device)
        self.ctx.set_cuts(self.cuts.data_ptr(), g)

    @torch.no_grad()
    def calibrate_partition(self, steps: int = 3, rounds: int = 2, gain: float = 0.8, verbose: bool = False):
        """Measure how fast every CTA of the persistent kernel gets through its share of the accumulate / insert / query /
        emit phases (``%globaltimer`` stamps, ``set_debug_times``) on synthetic gradients and re-cut the four tile
        partitions so that slow CTAs get less work.  Why: the two co-resident CTAs of an SM do not run the issue-bound
        phases at the same speed (the per-CTA timeline shows the second-launched CTA of almost every SM 12 % slower in
        accumulate and 20 % in query; profiles/README.md section 2), and every phase ends at a grid barrier, i.e. lasts
        as long as its slowest CTA.  Collective at W > 1 (runs ``(rounds + 1) * (steps + 1)`` exchange steps).  Resets
        residual / select history / gradient afterwards; the step counter keeps running (flags are epoch-valued).
        Returns the per-round phase maxima / medians."""
        G = self.grid()
        dbg = torch.zeros((PH_END + 1) * G * 2, dtype=torch.int64, device=self.device)
        self.ctx.set_debug_times(dbg.data_ptr())
        gen = torch.Generator(device=self.device).manual_seed(977 + self.rank)
        speeds = np.ones((4, G)) if self.cta_speeds is None else np.asarray(self.cta_speeds, dtype=np.float64).copy()
        phases = (PH_ACCUM, PH_INSERT, PH_QUERY, PH_EMIT)
        log = []
        try:
            for rnd in range(rounds + 1):                  # the last round only measures the result
                dur = np.zeros((4, G))
                for i in range(steps + 1):
                    self.grad.normal_(generator=gen).mul_(1e-2)
                    dbg.zero_()
                    self.step()
                    torch.cuda.synchronize(self.device)
                    self.check_status()
                    if i == 0:
                        continue                           # first step of a round: no select history for the new cut
                    t = dbg.cpu().numpy().reshape(PH_END + 1, G, 2)
                    for c, ph in enumerate(phases):
                        dur[c] += (t[ph, :, 1] - t[ph, :, 0]) / 1e3 / steps
                log.append({"max_us": dur.max(axis=1).round(1).tolist(), "median_us": np.median(dur, axis=1).round(1).tolist()})
                if verbose and self.rank == 0:
                    print(f"[calibrate] round {rnd}: max {log[-1]['max_us']} median {log[-1]['median_us']}", flush=True)
                if rnd == rounds:
                    break
                for c in range(4):
                    speeds[c] = update_cta_speeds(speeds[c], dur[c], gain)   # finished early -> more work next cut
                self.cta_speeds = speeds
                self._set_cuts()
        finally:
            self.ctx.set_debug_times(0)
        self.resid.zero_(); self.sel.zero_(); self.grad.zero_()
        return log

    # ---- arena -------------------------------------------------------------
    def _setup_arena(self):
        words = self.plan.arena_words(self.world, self.shard)
        self._ipc = self.world > 1
        if not self._ipc:
            self.arena = torch.zeros(words, dtype=torch.int32, device=self.device)
            self.arena_ptrs = [self.arena.data_ptr()]
            return
        self.multicast_ptr = 0
        if self.transport == "nccl":          # peers' slots arrive by all_gather into MY arena: nothing to map
            self._ipc = False
            self.arena = torch.zeros(words, dtype=torch.int32, device=self.device)
            self.arena_ptrs = [self.arena.data_ptr()] * self.world      # peer entries are never dereferenced
            return
        if os.environ.get("DR_NVLS", "0") == "1" and self._setup_arena_nvls(words):
            return
        # peer-map only the GPUs of this group's ranks (never every device of the box)
        devs = [None] * self.world
        dist.all_gather_object(devs, int(self.device.index or 0), group=self.group)
        self.mod.enable_peer_access([int(d) for d in devs])
        self._arena_ptr = self.mod.arena_alloc(words * 4)
        self.arena = self.mod.arena_as_tensor(self._arena_ptr, words, self.device.index or 0)
        handle = self.mod.arena_export(self._arena_ptr)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=self.group)
        self.arena_ptrs = []
        self._imported = []
        for r in range(self.world):
            if r == self.rank:
                self.arena_ptrs.append(self._arena_ptr)
            else:
                p = self.mod.arena_import(handles[r])
                self._imported.append(p)
                self.arena_ptrs.append(p)
        dist.barrier(group=self.group)

    def _pick_transport(self, transport: Optional[str]) -> str:
        t = transport or os.environ.get("DR_TRANSPORT", "") or "auto"
        if t not in ("auto", "p2p", "nccl"):
            raise ValueError(f"transport must be 'p2p', 'nccl' or None/'auto' (got {t!r})")
        if self.world == 1:
            return "p2p"
        if t != "auto":
            return t
        # auto: peer-mapped arenas need every rank on the same host.  torchrun exports LOCAL_WORLD_SIZE: when it equals
        # the size of the (default) group the answer is known without a collective
        lws = os.environ.get("LOCAL_WORLD_SIZE", "")
        if self.group is None and lws.isdigit() and int(lws) == self.world:
            return "p2p"
        import socket
        hosts = [None] * self.world
        dist.all_gather_object(hosts, socket.gethostname(), group=self.group)
        return "p2p" if len(set(hosts)) == 1 else "nccl"

    def _setup_arena_nvls(self, words: int) -> bool:
        """Arena in NVLS-capable symmetric memory (cuMem + multicast object, torch's symmetric-memory rendezvous does the
        handle exchange): peers' arenas are mapped like the IPC path and a multicast VA lets ONE store reach all of
        them through the NVSwitch.  Every rank must agree, so the outcome is all-reduced; False -> IPC path."""
        ok, hdl, t = 1, None, None
        try:
            import torch.distributed._symmetric_memory as symm
            t = symm.empty(words, dtype=torch.int32, device=self.device)
            grp = self.group if self.group is not None else dist.group.WORLD
            hdl = symm.rendezvous(t, group=grp)
            if int(hdl.multicast_ptr) == 0 or hdl.world_size != self.world:
                ok = 0
        except Exception as e:  # noqa: BLE001
            self._nvls_error = repr(e)
            ok = 0
        flag = torch.tensor([ok], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            return False
        t.zero_()
        torch.cuda.synchronize(self.device)
        self._symm_tensor, self._symm_handle = t, hdl
        self.arena = t
        self.arena_ptrs = [int(p) for p in hdl.buffer_ptrs]
        self.multicast_ptr = int(hdl.multicast_ptr)
        self._ipc = False
        self._imported = []
        dist.barrier(group=self.group)
        return True

    def close(self):
        if getattr(self, "_ipc", False):
            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for p in self._imported:
                self.mod.arena_close(p)
            self._imported = []
            self.mod.arena_free(self._arena_ptr)
            self._ipc = False

    # ---- run ---------------------------------------------------------------
    def step(self, epoch: Optional[int] = None):
        """Launch the fused kernel on the current stream.  In: ``self.grad``
        (local dense grads).  Out: ``self.grad`` (aggregated), ``self.resid``."""
        self.epoch = self.epoch + 1 if epoch is None else int(epoch)
        if self.transport == "nccl" and self.world > 1:
            self._step_nccl()
        else:
            self.ctx.run(self.epoch, PH_ACCUM, PH_END)

    def _step_nccl(self):
        """Multi-host form of the step: the same kernel runs its encode phases, the slots travel by ONE in-place
        NCCL all_gather per bucket (the slot of rank r sits at index r of the parity's slot array in every arena, so
        the output buffer is the arena itself), then the kernel runs expand + decode.  Still no per-tensor
        collectives and no size exchange (static offsets), cf. the reference's 2-3 all_gathers per tensor."""
        W, sw = self.world, self.plan.slot_words
        self.ctx.run(self.epoch, PH_ACCUM, PH_PUSH)
        base = ARENA_HDR_WORDS + (self.epoch & 1) * W * sw
        out = self.arena[base:base + W * sw]
        dist.all_gather_into_tensor(out, out[self.rank * sw:(self.rank + 1) * sw], group=self.group)
        self.ctx.run(self.epoch, PH_EXPAND, PH_PUSH2)      # expand, probe pass, apply pass (unsharded: no stage 2)

    def run_phases(self, begin: int, end: int, epoch: Optional[int] = None):
        """Debug / unfused chain: run a sub-range of phases (one launch)."""
        if epoch is not None:
            self.epoch = int(epoch)
        self.ctx.run(self.epoch, begin, end)

    def run_unfused(self, epoch: Optional[int] = None):
        self.epoch = self.epoch + 1 if epoch is None else int(epoch)
        for ph in range(PH_ACCUM, PH_END):
            self.ctx.run(self.epoch, ph, ph + 1)

    def slot(self, src_rank: Optional[int] = None, epoch: Optional[int] = None) -> torch.Tensor:
        """int32 view of the local copy of `src_rank`'s slot for `epoch`."""
        e = self.epoch if epoch is None else epoch
        r = self.rank if src_rank is None else src_rank
        off = ARENA_HDR_WORDS + ((e & 1) * self.world + r) * self.plan.slot_words
        return self.arena[off:off + self.plan.payload_words]

    def stats(self, src_rank: Optional[int] = None, epoch: Optional[int] = None) -> dict:
        """Counters of the last step for one sender (default: this rank): see :func:`stats_from_slot`.  One small
        device-to-host copy of the slot's header region; call it off the critical path (e.g. every N steps)."""
        n_hdr = SLOT_HEADER_WORDS + DYN_WORDS * len(self.plan.tensors)
        return stats_from_slot(self.plan, self.slot(src_rank, epoch)[:n_hdr])

    def stage2_bytes(self) -> int:
        """Bytes this rank pushed in the stage-2 exchange of the last step (8 B per entry, to W-1 peers)."""
        if not (self.shard and self.world > 1):
            return 0
        cap, s2w = self.plan.stage2_layout(self.world)
        off = ARENA_HDR_WORDS + 2 * self.world * self.plan.slot_words + ((self.epoch & 1) * self.world + self.rank) * s2w
        n = min(int(self.arena[off].item()), cap)
        return 8 * n * (self.world - 1)

    def check_status(self):
        st = self.status.cpu().tolist()
        if st[0] != 0:
            raise RuntimeError(f"deepreduce engine error: {STATUS_NAMES.get(st[0], st[0])} (aux={st[1]})")

    def grid(self) -> int:
        return int(self.ctx.grid())

    # ---- state (checkpoint / resume) ----------------------------
    def state_dict(self):
        return {"resid": self.resid.detach().cpu().clone(), "epoch": self.epoch,
                "sel": self.sel.detach().cpu().clone()}

    def load_state_dict(self, state):
        self.resid.copy_(state["resid"].to(self.device))
        self.sel.copy_(state["sel"].to(self.device))
        # never move the step counter backwards inside a live process group: the peers' flags in the arena carry
        # epochs this engine has already used (e.g. during calibrate_partition)
        self.epoch = max(self.epoch, int(state["epoch"]))


# ---------------------------------------------------------------------------
# exact top-k via the engine's radix select (used by TopKCompressor on CUDA)
# ---------------------------------------------------------------------------
_TOPK_CACHE: "OrderedDict[tuple, BucketEngine]" = OrderedDict()
_TOPK_CACHE_MAX = 64


def topk_select_cuda(flat: torch.Tensor, k: int):
    """(values[k], indices int64[k] ascending) of the k largest |x| — exact, ties broken towards the smaller index
    (what ``torch.topk`` on a stable sort would give).

    The engine's radix select resolves the threshold to 22 bits and ships EVERY element sharing that prefix (>= k of
    them, plan.h "Selection rule"); the slot is provisioned with slack for those extra coordinates and the final k are
    picked here from that short list by (|x| descending, index ascending).  If the prefix class is larger than the
    slack (massive ties) the call falls back to ``torch.topk``.  With fewer than k non-zeros the result is padded with
    (index 0, value 0.0) — harmless for ``index_add_``-style desparsification."""
    d = flat.numel()
    k = max(1, min(int(k), d))
    key = (d, k, flat.device.index)
    eng = _TOPK_CACHE.get(key)
    if eng is None:
        plan = BucketPlan([d], index=None, ks=[k], raw_slack=max(64, k // 8), min_numel=d)   # min_numel=d: never value-coded
        eng = BucketEngine(plan, device=flat.device, beta=0.0, gamma=1.0, average=False, use_history=False,
                           world=1, rank=0)
        _TOPK_CACHE[key] = eng
        if len(_TOPK_CACHE) > _TOPK_CACHE_MAX:
            _TOPK_CACHE.popitem(last=False)
    else:
        _TOPK_CACHE.move_to_end(key)
    tp = eng.plan.tensors[0]
    x = flat.detach().float().flatten()
    with torch.cuda.device(flat.device):
        eng.grad[:d].copy_(x)
        eng.hist.zero_()
        eng.tile_count.zero_()
        eng.hist_total.zero_()
        eng.epoch += 1
        eng.ctx.run(eng.epoch, PH_ACCUM, PH_EMIT + 1)
        slot = eng.slot()
        n_sel, _, _, n_pos = (int(v) for v in slot[SLOT_HEADER_WORDS:SLOT_HEADER_WORDS + 4].tolist())
        n_sel, n_pos = n_sel & 0xFFFFFFFF, n_pos & 0xFFFFFFFF
        if n_pos > tp.val_cap:                    # more ties at the 22-bit prefix than the slack: exact fallback
            idxs = torch.topk(x.abs(), k, sorted=False).indices.sort().values
            return x[idxs].to(flat.dtype), idxs
        vals = slot[tp.off_vals:tp.off_vals + n_sel].view(torch.float32).clone()
        idxs = slot[tp.off_idx:tp.off_idx + n_sel].to(torch.int64)
        if n_sel > k:                             # drop the smallest of the prefix class; stable => smaller index wins ties
            order = torch.sort(vals.abs(), descending=True, stable=True).indices[:k]
            keep = torch.sort(order).values
            vals, idxs = vals[keep], idxs[keep]
        elif n_sel < k:                           # fewer than k non-zeros
            pad = k - n_sel
            vals = torch.cat([vals, vals.new_zeros(pad)])
            idxs = torch.cat([idxs, idxs.new_zeros(pad)])
    return vals.to(flat.dtype), idxs
