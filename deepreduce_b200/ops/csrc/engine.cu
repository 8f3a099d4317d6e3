// DeepReduce-B200 fused bucket engine (sm_100a).
//
// One persistent, cooperatively-launched kernel runs the whole per-bucket
// gradient exchange:
//
//   accumulate residual -> exact per-tensor top-k threshold (3-digit radix
//   select on |g| bits, history-guided lower bound) -> bloom insert ->
//   universe query + ordered compaction + FP-aware value gather + residual
//   update -> P2P store of the compressed slot into every peer's arena over
//   NVLink -> release/acquire flags -> membership-test decode of all W slots,
//   rank->value, sum, scale, one dense write.
//
// It replaces, per tensor, the reference's chain: GRACE residual add, torch.topk,
// Bloomfilter.add/query/policy (reference pytorch/deepreduce.py:457-492,506-533),
// cupy packbits, 2-3 NCCL all_gathers, W x Bloom.decompress (:536-555),
// W x zeros+scatter and the sum.  Phases can also be launched
// one at a time (phase_begin/phase_end) — the "unfused chain" debug mode.
//
// Work decomposition: the bucket is cut into 4096-element tiles that never
// cross a tensor; CTA b owns tiles b, b+G, b+2G, ... (increasing order, grid
// co-resident), which makes the decoupled look-back used for ordered ranks
// deadlock-free.
#include "common.cuh"
#include "plan.h"

#include <atomic>
#include <cstdio>

namespace dr {

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n); }
long long launch_count() { return g_launches.load(); }

namespace {

constexpr uint32_t kFlagAgg = 1u, kFlagInc = 2u;
constexpr uint32_t kErrLookback = 1u, kErrPeerWait = 2u, kErrResolve = 3u;

struct ScanSmem {
  uint32_t cnt[2][kPerThread * kWarps + 4];   // double-buffered tile_rank scratch (+ total)
  uint32_t warp_tot[kWarps];
  uint32_t res[4];                            // resolve results: bin, krem, bincount, spare
  uint32_t lb;                                // look-back result
  uint32_t buf;                               // which cnt buffer is next
};

struct Smem {
  union {
    uint32_t hist[kHistBins];
    float acc[kTile];
  } u;
  ScanSmem s;
  TensorDesc td;                              // current tensor
};

DR_D uint32_t* slot_ptr(uint32_t* arena, const EngineParams& P, uint32_t parity, int src) {
  return arena + kArenaHdrWords + (size_t)(parity * (uint32_t)P.world + (uint32_t)src) * P.slot_words;
}

DR_D uint64_t pack_desc(uint32_t epoch, uint32_t flag, uint32_t value) {
  return ((uint64_t)(epoch & 0x3FFFFFFFu) << 34) | ((uint64_t)flag << 32) | (uint64_t)value;
}

DR_D void load_tensor(const EngineParams& P, uint32_t t, Smem& sm) {
  // 16 threads copy the 16-word descriptor
  __syncthreads();
  if (threadIdx.x < 16) {
    reinterpret_cast<uint32_t*>(&sm.td)[threadIdx.x] =
        reinterpret_cast<const uint32_t*>(P.tensors + t)[threadIdx.x];
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// ordered in-tile ranks.  Thread `tid` owns elements tid + c*kThreads (c < 8);
// bit c of `flags` marks element c.  Returns the exclusive rank of every
// flagged element in element order and the tile total.  Two __syncthreads.
// ---------------------------------------------------------------------------
DR_D void tile_rank(uint32_t flags, ScanSmem& s, uint32_t (&rank)[kPerThread], uint32_t& total) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t buf = s.buf & 1u;            // uniform: read before the first sync of this call
  uint32_t ball[kPerThread];
#pragma unroll
  for (int c = 0; c < kPerThread; ++c) ball[c] = __ballot_sync(0xFFFFFFFFu, (flags >> c) & 1u);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < kPerThread; ++c) s.cnt[buf][c * kWarps + warp] = __popc(ball[c]);
  }
  __syncthreads();
  if (warp == 0) {
    // 128 counters, 4 per lane
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = s.cnt[buf][lane * 4 + i]; sum += v[i]; }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= (uint32_t)o) incl += n;
    }
    uint32_t run = incl - sum;
#pragma unroll
    for (int i = 0; i < 4; ++i) { s.cnt[buf][lane * 4 + i] = run; run += v[i]; }
    if (lane == 31) s.cnt[buf][kPerThread * kWarps] = incl;
    if (lane == 0) s.buf = buf ^ 1u;
  }
  __syncthreads();
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int c = 0; c < kPerThread; ++c) rank[c] = s.cnt[buf][c * kWarps + warp] + __popc(ball[c] & lt);
  total = s.cnt[buf][kPerThread * kWarps];
}

// ---------------------------------------------------------------------------
// decoupled look-back over the tiles of one tensor.  Returns the exclusive
// prefix (sum of `count` over earlier tiles of the same tensor).
// ---------------------------------------------------------------------------
DR_D uint32_t lookback(const EngineParams& P, uint64_t* desc, uint32_t tile, uint32_t first_tile,
                       uint32_t count, ScanSmem& s) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (warp == 0) {
    uint32_t excl = 0;
    if (tile == first_tile) {
      if (lane == 0) st_release_gpu64(desc + tile, pack_desc(P.epoch, kFlagInc, count));
    } else {
      if (lane == 0) st_release_gpu64(desc + tile, pack_desc(P.epoch, kFlagAgg, count));
      int j = (int)tile - 1;
      const uint32_t want = P.epoch & 0x3FFFFFFFu;
      while (true) {
        const int idx = j - (int)lane;
        const bool valid = idx >= (int)first_tile;
        uint32_t flag = 0, val = 0;
        if (valid) {
          uint64_t d;
          uint32_t spins = 0;
          while (true) {
            d = ld_acquire_gpu64(desc + idx);
            if ((uint32_t)(d >> 34) == want && ((uint32_t)(d >> 32) & 3u) != 0u) break;
            if (++spins > P.spin_limit) { atomicExch(P.status, kErrLookback); d = pack_desc(want, kFlagInc, 0); break; }
            __nanosleep(20);
          }
          flag = (uint32_t)(d >> 32) & 3u;
          val = (uint32_t)d;
        }
        const uint32_t inc_mask = __ballot_sync(0xFFFFFFFFu, valid && flag == kFlagInc);
        uint32_t contrib = valid ? val : 0u;
        if (inc_mask) {
          const uint32_t first = __ffs(inc_mask) - 1u;   // nearest tile holding an inclusive prefix
          if (lane > first) contrib = 0u;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, o);
        excl += contrib;
        if (inc_mask) break;
        j -= 32;
        if (j < (int)first_tile) break;
      }
      if (lane == 0) st_release_gpu64(desc + tile, pack_desc(P.epoch, kFlagInc, excl + count));
    }
    if (lane == 0) s.lb = excl;
  }
  __syncthreads();
  const uint32_t r = s.lb;
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------
// radix-select digit resolve: find the bin holding the k-th largest key.
// H has `nbins` counters (bin index = digit value).  Result in s.res:
//   res[0] = bin (0xFFFFFFFF if total < k), res[1] = k remaining inside bin,
//   res[2] = count in bin.
// ---------------------------------------------------------------------------
DR_D void resolve_bins(const uint32_t* __restrict__ H, int nbins, uint32_t k, ScanSmem& s) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) { s.res[0] = 0xFFFFFFFFu; s.res[1] = 0; s.res[2] = 0; }
  uint32_t c[4], sum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rb = (int)tid * 4 + i;            // reversed position: 0 = largest digit
    c[i] = (rb < nbins) ? __ldcg(H + (nbins - 1 - rb)) : 0u;
    sum += c[i];
  }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= (uint32_t)o) incl += n;
  }
  if (lane == 31) s.warp_tot[warp] = incl;
  __syncthreads();
  uint32_t base = 0;
  for (uint32_t w = 0; w < warp; ++w) base += s.warp_tot[w];
  uint32_t cum = base + incl - sum;             // keys strictly above my first bin
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rb = (int)tid * 4 + i;
    if (rb < nbins && cum < k && cum + c[i] >= k) {
      s.res[0] = (uint32_t)(nbins - 1 - rb);
      s.res[1] = k - cum;
      s.res[2] = c[i];
    }
    cum += c[i];
  }
  __syncthreads();
}

DR_D void flush_hist(uint32_t* __restrict__ gh, uint32_t* __restrict__ gtotal, Smem& sm, int nbins) {
  __syncthreads();
  uint32_t part = 0;
  for (int j = threadIdx.x; j < nbins; j += kThreads) {
    const uint32_t v = sm.u.hist[j];
    if (v) { atomicAdd(gh + j, v); part += v; sm.u.hist[j] = 0; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
  if ((threadIdx.x & 31u) == 0 && part) atomicAdd(gtotal, part);
  __syncthreads();
}

DR_D void clear_hist(Smem& sm) {
  for (int j = threadIdx.x; j < kHistBins; j += kThreads) sm.u.hist[j] = 0;
  __syncthreads();
}

DR_D uint32_t* hist_ptr(const EngineParams& P, int which, uint32_t t) {
  return P.hist + ((size_t)which * P.n_tensors + t) * kHistBins;
}

// ===========================================================================
// phase 0: accumulate + hist pass 1
// ===========================================================================
DR_D void phase_accum(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity, P.rank);
  // zero the outgoing slot (filters, headers, prefix tables)
  {
    uint4* p = reinterpret_cast<uint4*>(my_slot);
    const uint32_t n4 = (P.payload_words + 3u) >> 2;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n4; i += gridDim.x * kThreads) p[i] = z;
  }
  clear_hist(sm);
  uint32_t cur = 0xFFFFFFFFu, lower = 0;
  const bool has_resid = (P.beta != 0.0f);
  for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    const uint32_t t = P.tile_tensor[tile];
    if (t != cur) {
      if (cur != 0xFFFFFFFFu) flush_hist(hist_ptr(P, 0, cur), P.hist_total + cur, sm, kHistBins);
      load_tensor(P, t, sm);
      cur = t;
      const uint32_t prev = P.use_history ? P.sel[t].prev_thr : 0u;
      lower = (prev > (1u << 23)) ? prev - (1u << 23) : 0u;   // half of last step's threshold
    }
    const uint32_t local0 = (tile - sm.td.tile_begin) * kTile;
    const uint32_t n = min((uint32_t)kTile, sm.td.numel - local0);
    const size_t base = (size_t)sm.td.elem_off + local0;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t e = (c * kThreads + threadIdx.x) * 4u;
      if (e < n) {
        float4 g = ld_stream_f4(reinterpret_cast<const float4*>(P.grad + base + e));
        float4 a;
        if (has_resid) {
          const float4 r = ld_stream_f4(reinterpret_cast<const float4*>(P.resid + base + e));
          a.x = P.beta * r.x + P.gamma * g.x; a.y = P.beta * r.y + P.gamma * g.y;
          a.z = P.beta * r.z + P.gamma * g.z; a.w = P.beta * r.w + P.gamma * g.w;
        } else {
          a.x = P.gamma * g.x; a.y = P.gamma * g.y; a.z = P.gamma * g.z; a.w = P.gamma * g.w;
        }
        *reinterpret_cast<float4*>(P.resid + base + e) = a;
        const uint32_t k0 = __float_as_uint(a.x) & 0x7FFFFFFFu, k1 = __float_as_uint(a.y) & 0x7FFFFFFFu;
        const uint32_t k2 = __float_as_uint(a.z) & 0x7FFFFFFFu, k3 = __float_as_uint(a.w) & 0x7FFFFFFFu;
        if (k0 >= lower) atomicAdd(&sm.u.hist[k0 >> 20], 1u);
        if (e + 1 < n && k1 >= lower) atomicAdd(&sm.u.hist[k1 >> 20], 1u);
        if (e + 2 < n && k2 >= lower) atomicAdd(&sm.u.hist[k2 >> 20], 1u);
        if (e + 3 < n && k3 >= lower) atomicAdd(&sm.u.hist[k3 >> 20], 1u);
      }
    }
  }
  if (cur != 0xFFFFFFFFu) flush_hist(hist_ptr(P, 0, cur), P.hist_total + cur, sm, kHistBins);
}

// generic "histogram one digit of the keys matching a prefix" pass over resid
template <int kWhich>
DR_D void hist_tiles(const EngineParams& P, Smem& sm) {
  // kWhich: 1 = pass-1 fallback (all keys, digit = key>>20)
  //         2 = pass 2 (keys with key>>20 == bin1, digit = (key>>9)&0x7FF)
  //         3 = pass 3 (keys with key>>9 == prefix22, digit = key & 0x1FF)
  clear_hist(sm);
  uint32_t cur = 0xFFFFFFFFu;
  bool active = false;
  uint32_t prefix = 0;
  for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    const uint32_t t = P.tile_tensor[tile];
    if (t != cur) {
      if (cur != 0xFFFFFFFFu && active)
        flush_hist(hist_ptr(P, kWhich, cur), P.hist_total + (size_t)kWhich * P.n_tensors + cur, sm,
                   kWhich == 3 ? 512 : kHistBins);
      load_tensor(P, t, sm);
      cur = t;
      const bool unsafe = P.use_history && (__ldcg(P.hist_total + t) < sm.td.k);
      if (kWhich == 1) {
        active = unsafe;
      } else if (kWhich == 2) {
        resolve_bins(hist_ptr(P, unsafe ? 1 : 0, t), kHistBins, sm.td.k, sm.s);
        prefix = sm.s.res[0];
        if (tile == sm.td.tile_begin && threadIdx.x == 0) {
          P.sel[t].bin1 = sm.s.res[0]; P.sel[t].krem1 = sm.s.res[1];
          if (sm.s.res[0] == 0xFFFFFFFFu) atomicExch(P.status, kErrResolve);
        }
        active = true;
        __syncthreads();
      } else {
        const uint32_t bin1 = __ldcg(&P.sel[t].bin1), krem1 = __ldcg(&P.sel[t].krem1);
        resolve_bins(hist_ptr(P, 2, t), kHistBins, krem1, sm.s);
        prefix = (bin1 << 11) | sm.s.res[0];
        if (tile == sm.td.tile_begin && threadIdx.x == 0) {
          P.sel[t].bin2 = sm.s.res[0]; P.sel[t].krem2 = sm.s.res[1];
          if (sm.s.res[0] == 0xFFFFFFFFu) atomicExch(P.status, kErrResolve);
        }
        active = true;
        __syncthreads();
      }
    }
    if (!active) continue;
    const uint32_t local0 = (tile - sm.td.tile_begin) * kTile;
    const uint32_t n = min((uint32_t)kTile, sm.td.numel - local0);
    const size_t base = (size_t)sm.td.elem_off + local0;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t e = (c * kThreads + threadIdx.x) * 4u;
      if (e < n) {
        const uint4 q = ld_stream_u4(reinterpret_cast<const uint4*>(P.resid + base + e));
        const uint32_t key[4] = {q.x & 0x7FFFFFFFu, q.y & 0x7FFFFFFFu, q.z & 0x7FFFFFFFu, q.w & 0x7FFFFFFFu};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (e + i < n) {
            if (kWhich == 1) atomicAdd(&sm.u.hist[key[i] >> 20], 1u);
            else if (kWhich == 2) { if ((key[i] >> 20) == prefix) atomicAdd(&sm.u.hist[(key[i] >> 9) & 0x7FFu], 1u); }
            else { if ((key[i] >> 9) == prefix) atomicAdd(&sm.u.hist[key[i] & 0x1FFu], 1u); }
          }
        }
      }
    }
  }
  if (cur != 0xFFFFFFFFu && active)
    flush_hist(hist_ptr(P, kWhich, cur), P.hist_total + (size_t)kWhich * P.n_tensors + cur, sm,
               kWhich == 3 ? 512 : kHistBins);
}

// ===========================================================================
// phase 4: threshold resolve + bloom insert
// ===========================================================================
DR_D void phase_insert(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity, P.rank);
  uint32_t cur = 0xFFFFFFFFu, T = 0, need = 0, ties_total = 0;
  for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    const uint32_t t = P.tile_tensor[tile];
    if (t != cur) {
      load_tensor(P, t, sm);
      cur = t;
      const uint32_t bin1 = __ldcg(&P.sel[t].bin1), bin2 = __ldcg(&P.sel[t].bin2);
      const uint32_t krem2 = __ldcg(&P.sel[t].krem2);
      resolve_bins(hist_ptr(P, 3, t), 512, krem2, sm.s);
      T = (((bin1 << 11) | bin2) << 9) | sm.s.res[0];
      need = sm.s.res[1];
      ties_total = sm.s.res[2];
      if (tile == sm.td.tile_begin && threadIdx.x == 0) {
        P.sel[t].thr = T; P.sel[t].need = need; P.sel[t].ties_total = ties_total;
        if (sm.s.res[0] == 0xFFFFFFFFu) atomicExch(P.status, kErrResolve);
      }
      __syncthreads();
    }
    const uint32_t local0 = (tile - sm.td.tile_begin) * kTile;
    const uint32_t n = min((uint32_t)kTile, sm.td.numel - local0);
    const size_t base = (size_t)sm.td.elem_off + local0;
    const bool ordered_ties = (ties_total != need);
    uint32_t gt = 0, eq = 0;
#pragma unroll
    for (int c = 0; c < kPerThread; ++c) {
      const uint32_t e = c * kThreads + threadIdx.x;
      if (e < n) {
        const uint32_t key = __float_as_uint(__ldcg(P.resid + base + e)) & 0x7FFFFFFFu;
        if (key > T) gt |= 1u << c;
        else if (key == T) eq |= 1u << c;
      }
    }
    uint32_t take = gt;
    if (!ordered_ties) {
      take |= eq;
    } else {
      uint32_t rank[kPerThread], total;
      tile_rank(eq, sm.s, rank, total);
      const uint32_t excl = lookback(P, P.tie_desc, tile, sm.td.tile_begin, total, sm.s);
      if (threadIdx.x == 0) P.tie_prefix[tile] = excl;
#pragma unroll
      for (int c = 0; c < kPerThread; ++c)
        if (((eq >> c) & 1u) && excl + rank[c] < need) take |= 1u << c;
    }
    if (sm.td.mode == kModeBloom) {
      uint32_t* filter = my_slot + sm.td.off_filter;
#pragma unroll
      for (int c = 0; c < kPerThread; ++c)
        if ((take >> c) & 1u) bloom_set(filter, local0 + c * kThreads + threadIdx.x, P.seed, sm.td.n_hash, sm.td.m_bits);
    }
  }
}

// ===========================================================================
// phase 5: universe query + ordered compaction + value gather + residual update
// ===========================================================================
DR_D void phase_emit(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity, P.rank);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    my_slot[0] = kMagic; my_slot[1] = P.epoch; my_slot[2] = P.n_tensors; my_slot[3] = P.payload_words;
    my_slot[4] = (uint32_t)P.rank;
  }
  uint32_t cur = 0xFFFFFFFFu, T = 0, need = 0, ties_total = 0;
  for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    const uint32_t t = P.tile_tensor[tile];
    if (t != cur) {
      load_tensor(P, t, sm);
      cur = t;
      T = __ldcg(&P.sel[t].thr); need = __ldcg(&P.sel[t].need); ties_total = __ldcg(&P.sel[t].ties_total);
    }
    const uint32_t tile_local = tile - sm.td.tile_begin;
    const uint32_t local0 = tile_local * kTile;
    const uint32_t n = min((uint32_t)kTile, sm.td.numel - local0);
    const size_t base = (size_t)sm.td.elem_off + local0;
    DynHeader* dyn = reinterpret_cast<DynHeader*>(my_slot + kSlotHeaderWords) + t;
    const bool last_tile = (tile_local + 1 == sm.td.n_tiles);
    uint32_t flags = 0;
    if (sm.td.mode == kModeBloom) {
      const uint32_t* filter = my_slot + sm.td.off_filter;
      const uint32_t n_hash = sm.td.n_hash, m_bits = sm.td.m_bits;
#pragma unroll
      for (int c = 0; c < kPerThread; ++c) {
        const uint32_t e = c * kThreads + threadIdx.x;
        if (e < n && bloom_test(local0 + e, P.seed, n_hash, m_bits, [&](uint32_t w) { return filter[w]; }))
          flags |= 1u << c;
      }
    } else {
      uint32_t eq = 0;
#pragma unroll
      for (int c = 0; c < kPerThread; ++c) {
        const uint32_t e = c * kThreads + threadIdx.x;
        if (e < n) {
          const uint32_t key = __float_as_uint(__ldcg(P.resid + base + e)) & 0x7FFFFFFFu;
          if (key > T) flags |= 1u << c;
          else if (key == T) eq |= 1u << c;
        }
      }
      if (ties_total == need) {
        flags |= eq;
      } else {
        uint32_t trank[kPerThread], ttotal;
        tile_rank(eq, sm.s, trank, ttotal);
        const uint32_t texcl = __ldcg(P.tie_prefix + tile);
#pragma unroll
        for (int c = 0; c < kPerThread; ++c)
          if (((eq >> c) & 1u) && texcl + trank[c] < need) flags |= 1u << c;
      }
    }
    uint32_t rank[kPerThread], total;
    tile_rank(flags, sm.s, rank, total);
    const uint32_t excl = lookback(P, P.pos_desc, tile, sm.td.tile_begin, total, sm.s);
    const uint32_t limit = (sm.td.mode == kModeBloom && P.policy != kPolicyP0) ? min(sm.td.k, sm.td.val_cap)
                                                                               : sm.td.val_cap;
    float* vals = reinterpret_cast<float*>(my_slot + sm.td.off_vals);
    uint32_t* idxs = my_slot + sm.td.off_idx;
#pragma unroll
    for (int c = 0; c < kPerThread; ++c) {
      if ((flags >> c) & 1u) {
        const uint32_t rp = excl + rank[c];
        if (rp < limit) {
          const uint32_t e = c * kThreads + threadIdx.x;
          vals[rp] = P.resid[base + e];
          P.resid[base + e] = 0.0f;                      // residual is exactly 0 on the shipped set
          if (sm.td.mode == kModeRaw) idxs[rp] = local0 + e;
          if (rp == limit - 1u) dyn->cutoff = local0 + e;
        }
      }
    }
    if (threadIdx.x == 0) {
      if (sm.td.mode == kModeBloom) my_slot[sm.td.off_prefix + tile_local] = min(excl, limit);
      if (last_tile) {
        const uint32_t incl = excl + total;
        dyn->n_sel = min(incl, limit);
        dyn->n_pos = incl;
        dyn->thr_bits = T;
        if (incl < limit) dyn->cutoff = 0xFFFFFFFFu;
        P.sel[t].prev_thr = T;
      }
    }
  }
}

// ===========================================================================
// phase 6/7: push + flags
// ===========================================================================
DR_D void phase_push(const EngineParams& P) {
  const uint32_t parity = P.epoch & 1u;
  const uint4* src = reinterpret_cast<const uint4*>(slot_ptr(P.arena[P.rank], P, parity, P.rank));
  const uint32_t n4 = (P.payload_words + 3u) >> 2;
  for (int h = 1; h < P.world; ++h) {
    const int peer = (P.rank + h) % P.world;             // stagger so peers are not hit in lock-step
    uint4* dst = reinterpret_cast<uint4*>(slot_ptr(P.arena[peer], P, parity, P.rank));
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n4; i += gridDim.x * kThreads) {
      const uint4 v = __ldcg(src + i);
      asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                   :: "l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
  }
  __threadfence_system();
}

DR_D void phase_signal(const EngineParams& P) {
  const int p = threadIdx.x;
  if (blockIdx.x == 0 && p < P.world && p != P.rank) st_release_sys(P.arena[p] + P.rank, P.epoch);
  if (p < P.world && p != P.rank) {
    const uint32_t* flag = P.arena[P.rank] + p;
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(flag) - P.epoch) < 0) {
      if (++spins > P.spin_limit) { atomicExch(P.status, kErrPeerWait); atomicExch(P.status + 1, (uint32_t)p); break; }
      __nanosleep(100);
    }
  }
  __syncthreads();
}

// ===========================================================================
// phase 8: decode every rank's slot for my tiles, sum, scale, dense write
// ===========================================================================
DR_D uint32_t lower_bound_u32(const uint32_t* a, uint32_t n, uint32_t x) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldcg(a + mid) < x) lo = mid + 1; else hi = mid; }
  return lo;
}

DR_D void phase_decode(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* arena = P.arena[P.rank];
  // hist arrays are free after the insert phase: zero them for the next step
  {
    uint4* h = reinterpret_cast<uint4*>(P.hist);
    const size_t n4 = (size_t)4 * P.n_tensors * kHistBins / 4;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (size_t)gridDim.x * kThreads) h[i] = z;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < 4u * P.n_tensors; i += gridDim.x * kThreads)
      P.hist_total[i] = 0u;
  }
  uint32_t cur = 0xFFFFFFFFu;
  for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    const uint32_t t = P.tile_tensor[tile];
    if (t != cur) { load_tensor(P, t, sm); cur = t; }
    const uint32_t tile_local = tile - sm.td.tile_begin;
    const uint32_t local0 = tile_local * kTile;
    const uint32_t n = min((uint32_t)kTile, sm.td.numel - local0);
    const size_t base = (size_t)sm.td.elem_off + local0;
    if (sm.td.mode == kModeBloom) {
      float acc[kPerThread];
#pragma unroll
      for (int c = 0; c < kPerThread; ++c) acc[c] = 0.0f;
      const uint32_t n_hash = sm.td.n_hash, m_bits = sm.td.m_bits;
      for (int r = 0; r < P.world; ++r) {
        const uint32_t* slot = slot_ptr(arena, P, parity, r);
        const DynHeader* dyn = reinterpret_cast<const DynHeader*>(slot + kSlotHeaderWords) + t;
        const uint32_t n_sel = __ldcg(&dyn->n_sel), cutoff = __ldcg(&dyn->cutoff);
        const uint32_t pre = __ldcg(slot + sm.td.off_prefix + tile_local);
        const uint32_t* filter = slot + sm.td.off_filter;
        const float* vals = reinterpret_cast<const float*>(slot + sm.td.off_vals);
        if (!(pre < n_sel && local0 <= cutoff)) continue;   // tile-uniform: nothing of rank r lands in this tile
        uint32_t flags = 0;
#pragma unroll
        for (int c = 0; c < kPerThread; ++c) {
          const uint32_t e = c * kThreads + threadIdx.x;
          const uint32_t gi = local0 + e;
          if (e < n && gi <= cutoff &&
              bloom_test(gi, P.seed, n_hash, m_bits, [&](uint32_t w) { return filter[w]; }))
            flags |= 1u << c;
        }
        uint32_t rank[kPerThread], total;
        tile_rank(flags, sm.s, rank, total);
#pragma unroll
        for (int c = 0; c < kPerThread; ++c) {
          if ((flags >> c) & 1u) {
            const uint32_t rp = pre + rank[c];
            if (rp < n_sel) acc[c] += __ldcg(vals + rp);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kPerThread; ++c) {
        const uint32_t e = c * kThreads + threadIdx.x;
        if (e < n) P.grad[base + e] = acc[c] * P.scale;
      }
    } else {
      __syncthreads();
      for (int j = threadIdx.x; j < kTile; j += kThreads) sm.u.acc[j] = 0.0f;
      __syncthreads();
      for (int r = 0; r < P.world; ++r) {
        const uint32_t* slot = slot_ptr(arena, P, parity, r);
        const DynHeader* dyn = reinterpret_cast<const DynHeader*>(slot + kSlotHeaderWords) + t;
        const uint32_t n_sel = min(__ldcg(&dyn->n_sel), sm.td.val_cap);
        const uint32_t* idxs = slot + sm.td.off_idx;
        const float* vals = reinterpret_cast<const float*>(slot + sm.td.off_vals);
        const uint32_t lo = lower_bound_u32(idxs, n_sel, local0);
        const uint32_t hi = lower_bound_u32(idxs, n_sel, local0 + n);
        for (uint32_t j = lo + threadIdx.x; j < hi; j += kThreads)
          atomicAdd(&sm.u.acc[__ldcg(idxs + j) - local0], __ldcg(vals + j));
      }
      __syncthreads();
      for (uint32_t e = threadIdx.x; e < n; e += kThreads) P.grad[base + e] = sm.u.acc[e] * P.scale;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 2) dr_engine_kernel(const __grid_constant__ EngineParams P) {
  __shared__ Smem sm;
  if (threadIdx.x == 0) sm.s.buf = 0;
  __syncthreads();
  uint32_t bar_epoch = 0;
  for (int ph = P.phase_begin; ph < P.phase_end; ++ph) {
    bool ran = true;
    switch (ph) {
      case kPhAccum: phase_accum(P, sm); break;
      case kPhFallback: if (P.use_history) hist_tiles<1>(P, sm); else ran = false; break;
      case kPhHist2: hist_tiles<2>(P, sm); break;
      case kPhHist3: hist_tiles<3>(P, sm); break;
      case kPhInsert: phase_insert(P, sm); break;
      case kPhEmit: phase_emit(P, sm); break;
      case kPhPush: if (P.world > 1) phase_push(P); else ran = false; break;
      case kPhSignal: if (P.world > 1) phase_signal(P); else ran = false; break;
      case kPhDecode: phase_decode(P, sm); break;
      default: ran = false; break;
    }
    // a barrier separates dependent phases; signal->decode needs none (every CTA waits itself)
    if (ran && ph + 1 < P.phase_end && ph != kPhSignal) grid_barrier(P.barrier, bar_epoch, P.status, P.spin_limit);
  }
}

}  // namespace

// ---------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------
int engine_max_grid(int blocks_per_sm) {
  int dev = 0, sms = 0, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dr_engine_kernel, kThreads, 0);
  if (occ < 1) occ = 1;
  if (blocks_per_sm > 0 && blocks_per_sm < occ) occ = blocks_per_sm;
  return occ * sms;
}

cudaError_t engine_launch(const EngineParams& P, int grid, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(P.barrier, 0, sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  void* args[] = {const_cast<EngineParams*>(&P)};
  count_launch(1);
  return cudaLaunchCooperativeKernel((const void*)dr_engine_kernel, dim3(grid), dim3(kThreads), args, 0, stream);
}

}  // namespace dr
