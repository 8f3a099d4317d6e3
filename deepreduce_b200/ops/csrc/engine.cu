// DeepReduce-B200 fused bucket engine (sm_100a).
//
// One persistent, cooperatively-launched kernel runs the whole per-bucket
// gradient exchange:
//
//   accumulate residual -> per-tensor top-k threshold (2-digit radix select on
//   |g| bits, history-guided lower bound) -> bloom insert -> universe query +
//   ordered compaction + FP-aware value gather + residual update -> P2P store
//   of the compressed slot into every peer's arena over NVLink -> release /
//   acquire flags -> membership-test decode of all W slots, rank->value, sum,
//   scale, dense write.
//
// It replaces, per tensor, the reference's chain: GRACE residual add, torch.topk,
// Bloomfilter.add/query/policy (reference pytorch/deepreduce.py:457-492,506-533),
// cupy packbits, 2-3 NCCL all_gathers, W x Bloom.decompress (:536-555),
// W x zeros+scatter and the sum.  Phases can also be launched
// one at a time (phase_begin/phase_end) — the "unfused chain" debug mode.
//
// Work decomposition: the bucket is cut into 4096-element tiles that never
// cross a tensor; every phase gives CTA b the same contiguous tile range, so
// per-tensor work (histogram flush, threshold resolve, filter staging) is paid
// 1-3 times per CTA instead of once per tile (profiles/ v1->v3 notes).  Ordered
// ranks need a prefix over tiles: the query phase stores per-thread element
// flags + per-tile counts, and after one grid barrier the emit phase sums the
// counts it needs locally — no look-back chain.
//
// Latency structure: streaming passes software-prefetch the next tile while
// the current one is binned; bloom filters are staged in shared memory (every
// ResNet-50 tensor's filter fits) and probed from there.
#include "common.cuh"
#include "plan.h"

#include <atomic>
#include <cstdio>

namespace dr {

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n); }
long long launch_count() { return g_launches.load(); }

namespace {

constexpr uint32_t kErrLookback = 1u, kErrPeerWait = 2u, kErrResolve = 3u;
constexpr uint32_t kNoTensor = 0xFFFFFFFFu;

struct ScanSmem {
  alignas(16) uint32_t cnt[2][kPerThread * kWarps];   // double-buffered tile_rank scratch
  uint32_t warp_tot[kWarps];
  uint32_t res[4];                            // resolve results: bin, krem, bincount, spare
  uint32_t lb;                                // small CTA-wide scratch word
  uint32_t rle_pre[16];                       // kModeRle decode: running entry prefix per sender
};

struct Smem {
  union {
    uint32_t hist[kHistBins];
    float acc[kTile];
  } u;
  ScanSmem s;
  TensorDesc td;                              // current tensor
  uint64_t bar[8];                            // mbarriers of the TMA tile ring
  int seg_start[kMaxSeg + 2];                 // 'both': segment boundaries of the current (tensor, rank)
  int n_seg;
};

extern __shared__ __align__(16) uint32_t g_filter_smem[];   // dynamic: staged bloom filter

DR_D uint32_t* slot_ptr(uint32_t* arena, const EngineParams& P, uint32_t parity, int src) {
  return arena + kArenaHdrWords + (size_t)(parity * (uint32_t)P.world + (uint32_t)src) * P.slot_words;
}

DR_D uint32_t* s2_ptr(uint32_t* arena, const EngineParams& P, uint32_t parity, int src) {
  return arena + kArenaHdrWords + (size_t)2 * (uint32_t)P.world * P.slot_words +
         (size_t)(parity * (uint32_t)P.world + (uint32_t)src) * P.s2_words;
}

DR_D bool sharded(const EngineParams& P) { return P.shard && P.world > 1; }

// tiles this rank decodes: everything (W == 1 / unsharded) or its 1/W slice
DR_D void decode_span(const EngineParams& P, int owner, uint32_t& s_begin, uint32_t& s_end) {
  if (sharded(P)) {
    s_begin = (uint32_t)(((uint64_t)P.n_tiles * (uint32_t)owner) / (uint32_t)P.world);
    s_end = (uint32_t)(((uint64_t)P.n_tiles * ((uint32_t)owner + 1u)) / (uint32_t)P.world);
  } else { s_begin = 0; s_end = P.n_tiles; }
}

struct Tile { uint32_t tensor, base, n, local0, single; };   // `single`: the tensor has exactly one tile

DR_D Tile load_tile(const EngineParams& P, uint32_t tile) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(P.tiles) + tile);
  Tile t; t.tensor = q.x; t.base = q.y; t.n = q.z & 0xFFFFu; t.local0 = q.w; t.single = q.z >> 31;
  return t;
}

DR_D void load_tensor(const EngineParams& P, uint32_t t, Smem& sm) {
  __syncthreads();
  if (threadIdx.x < kDescWords) {
    reinterpret_cast<uint32_t*>(&sm.td)[threadIdx.x] =
        __ldg(reinterpret_cast<const uint32_t*>(P.tensors + t) + threadIdx.x);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// ordered in-tile ranks.  Thread `tid` owns elements tid + c*kThreads (c < 8);
// bit c of `flags` marks element c.  Returns the exclusive rank of every
// flagged element in element order and the tile total.  ONE __syncthreads:
// every warp publishes its 8 slot counts, then scans all 128 counts itself.
// `buf` is a per-thread toggle (double buffering makes a trailing barrier
// unnecessary: a buffer is rewritten two calls later, after the barrier of the
// call in between).
// ---------------------------------------------------------------------------
DR_D void tile_rank(uint32_t flags, ScanSmem& s, uint32_t& buf, uint32_t (&rank)[kPerThread], uint32_t& total) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t ball[kPerThread];
#pragma unroll
  for (int c = 0; c < kPerThread; ++c) ball[c] = __ballot_sync(0xFFFFFFFFu, (flags >> c) & 1u);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < kPerThread; ++c) s.cnt[buf][c * kWarps + warp] = __popc(ball[c]);
  }
  __syncthreads();
  // lane l holds counts 4l .. 4l+3 (element order = slot-major, warp-minor)
  const uint4 v = reinterpret_cast<const uint4*>(s.cnt[buf])[lane];
  const uint32_t sum = v.x + v.y + v.z + v.w;
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= (uint32_t)o) incl += n;
  }
  const uint32_t base = incl - sum;
  const uint32_t sub = warp & 3u;
  const uint32_t part = (sub > 0 ? v.x : 0u) + (sub > 1 ? v.y : 0u) + (sub > 2 ? v.z : 0u);   // my lane's partials, reused below
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int c = 0; c < kPerThread; ++c) {
    // prefix of index i = c*16 + warp lives in lane i>>2 = 4c + (warp>>2), at sub-position warp&3
    const int src = 4 * c + (int)(warp >> 2);
    const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, base, src);
    const uint32_t vx = __shfl_sync(0xFFFFFFFFu, v.x, src), vy = __shfl_sync(0xFFFFFFFFu, v.y, src),
                   vz = __shfl_sync(0xFFFFFFFFu, v.z, src);
    rank[c] = b0 + (sub > 0 ? vx : 0u) + (sub > 1 ? vy : 0u) + (sub > 2 ? vz : 0u) + __popc(ball[c] & lt);
  }
  (void)part;
  total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  buf ^= 1u;
}

// Same ranks from per-(slot, warp) counts that the query phase left in global memory (128 bytes per tile, index
// c*16 + warp): no shared memory, no CTA barrier — warps drift through their tiles independently.
DR_D void tile_rank_counts(uint32_t flags, const uint8_t* cnt, uint32_t (&rank)[kPerThread], uint32_t& total) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t w = __ldcg(reinterpret_cast<const uint32_t*>(cnt) + lane);       // counts 4l .. 4l+3
  const uint32_t vx = w & 0xFFu, vy = (w >> 8) & 0xFFu, vz = (w >> 16) & 0xFFu, vw = w >> 24;
  const uint32_t sum = vx + vy + vz + vw;
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= (uint32_t)o) incl += n;
  }
  const uint32_t base = incl - sum;
  const uint32_t sub = warp & 3u;
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int c = 0; c < kPerThread; ++c) {
    const int src = 4 * c + (int)(warp >> 2);
    const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, base, src);
    const uint32_t sx = __shfl_sync(0xFFFFFFFFu, vx, src), sy = __shfl_sync(0xFFFFFFFFu, vy, src),
                   sz = __shfl_sync(0xFFFFFFFFu, vz, src);
    const uint32_t ball = __ballot_sync(0xFFFFFFFFu, (flags >> c) & 1u);
    rank[c] = b0 + (sub > 0 ? sx : 0u) + (sub > 1 ? sy : 0u) + (sub > 2 ? sz : 0u) + __popc(ball & lt);
  }
  total = __shfl_sync(0xFFFFFFFFu, incl, 31);
}

// ---------------------------------------------------------------------------
// radix-select digit resolve: find the bin holding the k-th largest key.
// Result in s.res: [0] bin (0xFFFFFFFF if total < k), [1] k remaining inside
// the bin, [2] count in the bin.
// ---------------------------------------------------------------------------
DR_D uint32_t* hist_ptr(const EngineParams& P, int which, uint32_t t) {
  return P.hist + ((size_t)which * P.n_tensors + t) * kHistBins;
}

template <typename LoadFn>
DR_D void resolve_bins(LoadFn H, int nbins, uint32_t k, ScanSmem& s) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  __syncthreads();
  if (tid == 0) { s.res[0] = 0xFFFFFFFFu; s.res[1] = 0; s.res[2] = 0; }
  uint32_t c[4], sum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rb = (int)tid * 4 + i;            // reversed position: 0 = largest digit
    c[i] = (rb < nbins) ? H(nbins - 1 - rb) : 0u;
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

DR_D void clear_hist(Smem& sm) {
  __syncthreads();
  for (int j = threadIdx.x; j < kHistBins; j += kThreads) sm.u.hist[j] = 0;
  __syncthreads();
}

DR_D void flush_hist(uint32_t* __restrict__ gh, Smem& sm) {
  __syncthreads();
  for (int j = threadIdx.x; j < kHistBins; j += kThreads) {
    const uint32_t v = sm.u.hist[j];
    if (v) { atomicAdd(gh + j, v); sm.u.hist[j] = 0; }
  }
  __threadfence();                              // merged counts are visible before the ticket is taken
  __syncthreads();
}

// Finish one digit of one tensor for this CTA.  If the CTA owns every tile of the tensor the digit is
// resolved straight from the SMEM histogram; otherwise the histogram is merged into the global one and
// the CTA whose merge completes the tensor (ticket count == n_tiles) resolves it — once per tensor,
// inside the same phase, no extra grid barrier.  Returns true (CTA-uniform) if this CTA resolved; the
// result is then in sm.s.res.
DR_D bool finish_digit(const EngineParams& P, Smem& sm, int which, uint32_t t, uint32_t n_mine, uint32_t n_tiles,
                       uint32_t k) {
  if (n_mine == n_tiles) {
    resolve_bins([&](int b) { return sm.u.hist[b]; }, kHistBins, k, sm.s);
    clear_hist(sm);
    return true;
  }
  uint32_t* gh = hist_ptr(P, which, t);
  flush_hist(gh, sm);
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t before = atomicAdd(P.hist_total + (size_t)which * P.n_tensors + t, n_mine);
    sm.s.lb = (before + n_mine == n_tiles) ? 1u : 0u;
    __threadfence();
  }
  __syncthreads();
  const bool last = sm.s.lb != 0u;
  __syncthreads();
  if (last) resolve_bins([&](int b) { return __ldcg(gh + b); }, kHistBins, k, sm.s);
  return last;
}

// stage a bloom filter (global, 16-byte aligned) into the dynamic SMEM buffer
DR_D void stage_filter(const uint32_t* __restrict__ filter, uint32_t n_words) {
  __syncthreads();                              // previous users of the buffer are done
  const uint4* src = reinterpret_cast<const uint4*>(filter);
  uint4* dst = reinterpret_cast<uint4*>(g_filter_smem);
  const uint32_t n4 = (n_words + 3u) >> 2;
  for (uint32_t i = threadIdx.x; i < n4; i += kThreads) dst[i] = src[i];
  __syncthreads();
}

// Occupancy hint: 128 bits per tile, bit (c*16 + warp) covers the 32 consecutive elements that warp `warp`
// owns in slot c.  valid_from_hint() turns the 4 words into this thread's 8-bit element mask.
DR_D uint32_t valid_from_hint(const uint32_t (&h)[4]) {
  const uint32_t warp = threadIdx.x >> 5;
  uint32_t v = 0;
#pragma unroll
  for (int c = 0; c < kPerThread; ++c) v |= ((h[c >> 1] >> (((c & 1) << 4) + warp)) & 1u) << c;
  return v;
}

// every warp publishes which of its 8 slots contain a flagged element; 4 words land in `dst` (SMEM scratch in `s`)
DR_D void build_hint(uint32_t mask, ScanSmem& s, uint32_t* dst) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x < 4) s.warp_tot[threadIdx.x] = 0;
  __syncthreads();
  uint32_t occ = 0;
#pragma unroll
  for (int c = 0; c < kPerThread; ++c) if (__ballot_sync(0xFFFFFFFFu, (mask >> c) & 1u)) occ |= 1u << c;
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < kPerThread; ++c) if ((occ >> c) & 1u) atomicOr(&s.warp_tot[c >> 1], 1u << (((c & 1) << 4) + warp));
  }
  __syncthreads();
  if (threadIdx.x < 4) dst[threadIdx.x] = s.warp_tot[threadIdx.x];
}

DR_D void tile_range(const EngineParams& P, uint32_t& t_begin, uint32_t& t_end) {
  t_begin = (uint32_t)(((uint64_t)P.n_tiles * blockIdx.x) / gridDim.x);
  t_end = (uint32_t)(((uint64_t)P.n_tiles * (blockIdx.x + 1)) / gridDim.x);
}

// 8 membership tests per thread (elements idx0 + c*kThreads), early exit per element
template <typename LoadFn>
DR_D uint32_t bloom_test8(uint32_t idx0, uint32_t valid, uint32_t seed, uint32_t n_hash, uint32_t m_bits, LoadFn ld) {
  uint32_t flags = 0;
#pragma unroll
  for (int c = 0; c < kPerThread; ++c)
    if (((valid >> c) & 1u) && bloom_test(idx0 + c * kThreads, seed, n_hash, m_bits, ld)) flags |= 1u << c;
  return flags;
}

// ===========================================================================
// phase 0: accumulate + hist digit 1 (+ the whole select for single-tile tensors)
// ===========================================================================
constexpr uint32_t kUnsafe = 0xFFFFFFFFu;   // sel.bin1 marker: history bound hid the threshold -> fallback phase

DR_D void write_digit1(const EngineParams& P, Smem& sm, uint32_t t) {
  if (threadIdx.x == 0) { P.sel[t].bin1 = sm.s.res[0]; P.sel[t].krem1 = sm.s.res[1]; P.sel[t].done_epoch = 0; }
}

DR_D void write_final(const EngineParams& P, Smem& sm, uint32_t t, uint32_t bin1) {
  if (threadIdx.x == 0) {
    if (sm.s.res[0] == 0xFFFFFFFFu) atomicExch(P.status, kErrResolve);
    const uint32_t T22 = max((bin1 << 11) | sm.s.res[0], 1u);
    P.sel[t].bin2 = sm.s.res[0]; P.sel[t].krem2 = sm.s.res[1];
    P.sel[t].thr = T22 << 9; P.sel[t].n_ge = sm.s.res[2];
  }
}

DR_D void phase_accum(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity, P.rank);
  {  // zero the outgoing slot (filters, headers, prefix tables)
    uint4* p = reinterpret_cast<uint4*>(my_slot);
    const uint32_t n4 = (P.payload_words + 3u) >> 2;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n4; i += gridDim.x * kThreads) p[i] = z;
  }
  if (sharded(P) && blockIdx.x == 0 && threadIdx.x == 0) *s2_ptr(P.arena[P.rank], P, parity, P.rank) = 0u;
  clear_hist(sm);
  const bool has_resid = (P.beta != 0.0f);
  uint32_t cur = kNoTensor, lower = 0, n_mine = 0;
  uint32_t tile, t_end;
  tile_range(P, tile, t_end);
  if (tile >= t_end) return;
  Tile ti = load_tile(P, tile);
  float4 g[2], r[2];
  auto issue = [&](const Tile& t, float4 (&gg)[2], float4 (&rr)[2]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t e = (c * kThreads + threadIdx.x) * 4u;
      if (e < t.n) {
        gg[c] = ld_stream_f4(reinterpret_cast<const float4*>(P.grad + t.base + e));
        if (has_resid) rr[c] = ld_stream_f4(reinterpret_cast<const float4*>(P.resid + t.base + e));
      }
    }
  };
  auto finish = [&]() {      // digit 1 of tensor `cur` is complete for this CTA
    const uint32_t k = __ldg(&P.tensors[cur].k), nt = __ldg(&P.tensors[cur].n_tiles);
    if (finish_digit(P, sm, 0, cur, n_mine, nt, k)) write_digit1(P, sm, cur);
  };
  issue(ti, g, r);
  while (true) {
    const uint32_t next = tile + 1;
    Tile tn = ti;
    float4 gn[2], rn[2];
    const bool has_next = next < t_end;
    if (has_next) { tn = load_tile(P, next); issue(tn, gn, rn); }     // prefetch before binning the current tile
    if (ti.tensor != cur) {
      if (cur != kNoTensor) finish();
      cur = ti.tensor; n_mine = 0;
      const uint32_t prev = P.use_history ? __ldcg(&P.sel[cur].prev_thr) : 0u;
      lower = (prev > (1u << 23)) ? prev - (1u << P.hist_shift) : 0u;  // a fraction of last step's threshold
    }
    uint32_t key[8];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t e = (c * kThreads + threadIdx.x) * 4u;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < ti.n) {
        if (has_resid) {
          a.x = P.beta * r[c].x + P.gamma * g[c].x; a.y = P.beta * r[c].y + P.gamma * g[c].y;
          a.z = P.beta * r[c].z + P.gamma * g[c].z; a.w = P.beta * r[c].w + P.gamma * g[c].w;
        } else {
          a.x = P.gamma * g[c].x; a.y = P.gamma * g[c].y; a.z = P.gamma * g[c].z; a.w = P.gamma * g[c].w;
        }
        *reinterpret_cast<float4*>(P.resid + ti.base + e) = a;
      }
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = e + i < ti.n;
        key[c * 4 + i] = ok ? (__float_as_uint(av[i]) & 0x7FFFFFFFu) : 0xFFFFFFFFu;   // 0xFFFFFFFF = not an element
        if (ok && key[c * 4 + i] >= lower) atomicAdd(&sm.u.hist[key[c * 4 + i] >> 20], 1u);
      }
    }
    n_mine += 1;
    if (ti.single) {
      // one-tile tensor: finish the whole 2-digit select here, from the keys still in registers
      const uint32_t k = __ldg(&P.tensors[cur].k);
      resolve_bins([&](int b) { return sm.u.hist[b]; }, kHistBins, k, sm.s);
      const uint32_t bin1 = sm.s.res[0], krem1 = sm.s.res[1];
      write_digit1(P, sm, cur);
      clear_hist(sm);
      if (bin1 != kUnsafe) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (key[i] != 0xFFFFFFFFu && (key[i] >> 20) == bin1) atomicAdd(&sm.u.hist[(key[i] >> 9) & 0x7FFu], 1u);
        resolve_bins([&](int b) { return sm.u.hist[b]; }, kHistBins, krem1, sm.s);
        write_final(P, sm, cur, bin1);
        if (threadIdx.x == 0) P.sel[cur].done_epoch = P.epoch;
        clear_hist(sm);
      }
      cur = kNoTensor; n_mine = 0;
    }
    if (!has_next) break;
    tile = next; ti = tn;
#pragma unroll
    for (int c = 0; c < 2; ++c) { g[c] = gn[c]; r[c] = rn[c]; }
  }
  if (cur != kNoTensor) finish();
}

// generic "histogram one digit of the keys" pass over resid with tile prefetch
//   kWhich 1: digit-1 fallback (all keys, digit = key>>20) for tensors whose history bound was unsafe
//   kWhich 2: digit 2 (keys with key>>20 == bin1, digit = (key>>9) & 0x7FF)
template <int kWhich>
DR_D void hist_tiles(const EngineParams& P, Smem& sm) {
  clear_hist(sm);
  uint32_t cur = kNoTensor, n_mine = 0, k_cur = 0, nt_cur = 0;
  bool active = false;
  uint32_t prefix = 0;
  uint32_t tile, t_end;
  tile_range(P, tile, t_end);
  if (tile >= t_end) return;
  Tile ti = load_tile(P, tile);
  uint4 q[2];
  auto issue = [&](const Tile& t, uint4 (&qq)[2]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t e = (c * kThreads + threadIdx.x) * 4u;
      if (e < t.n) qq[c] = ld_stream_u4(reinterpret_cast<const uint4*>(P.resid + t.base + e));
    }
  };
  auto finish = [&]() {
    if (!active) return;
    if (finish_digit(P, sm, kWhich, cur, n_mine, nt_cur, k_cur)) {
      if (kWhich == 1) {
        if (threadIdx.x == 0 && sm.s.res[0] == 0xFFFFFFFFu) atomicExch(P.status, kErrResolve);
        write_digit1(P, sm, cur);
      } else {
        write_final(P, sm, cur, prefix);
      }
    }
  };
  if (kWhich == 2) issue(ti, q);
  while (true) {
    const uint32_t next = tile + 1;
    const bool has_next = next < t_end;
    Tile tn = ti;
    uint4 qn[2];
    if (has_next) { tn = load_tile(P, next); if (kWhich == 2) issue(tn, qn); }
    if (ti.tensor != cur) {
      if (cur != kNoTensor) finish();
      cur = ti.tensor; n_mine = 0;
      nt_cur = __ldg(&P.tensors[cur].n_tiles);
      const uint32_t bin1 = __ldcg(&P.sel[cur].bin1);
      const bool done = __ldcg(&P.sel[cur].done_epoch) == P.epoch;
      if (kWhich == 1) {
        active = (bin1 == kUnsafe) && !done;
        k_cur = __ldg(&P.tensors[cur].k);
      } else {
        active = !done;
        prefix = bin1;
        k_cur = __ldcg(&P.sel[cur].krem1);
        if (active && bin1 == kUnsafe && threadIdx.x == 0) atomicExch(P.status, kErrResolve);
      }
    }
    if (active) {
      if (kWhich == 1) issue(ti, q);            // rare path: no prefetch
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t e = (c * kThreads + threadIdx.x) * 4u;
        if (e < ti.n) {
          const uint32_t key[4] = {q[c].x & 0x7FFFFFFFu, q[c].y & 0x7FFFFFFFu, q[c].z & 0x7FFFFFFFu, q[c].w & 0x7FFFFFFFu};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (e + i < ti.n) {
              if (kWhich == 1) atomicAdd(&sm.u.hist[key[i] >> 20], 1u);
              else if ((key[i] >> 20) == prefix) atomicAdd(&sm.u.hist[(key[i] >> 9) & 0x7FFu], 1u);
            }
          }
        }
      }
      n_mine += 1;
    }
    if (!has_next) break;
    tile = next; ti = tn;
    q[0] = qn[0]; q[1] = qn[1];
  }
  if (cur != kNoTensor) finish();
}

// ===========================================================================
// phase 3: bloom insert of the selected set (queue-compacted: the ~1 % selected elements of a tile are
// gathered into an SMEM queue, then every thread sets one (element, hash) bit — no divergent tails)
// ===========================================================================
DR_D void phase_insert(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity, P.rank);
  uint32_t* queue = sm.u.hist;                 // 4096 entries: reuses the histogram/acc union
  uint32_t cur = kNoTensor, T22 = 1;
  uint32_t tile, t_end;
  tile_range(P, tile, t_end);
  if (tile >= t_end) return;
  Tile ti = load_tile(P, tile);
  uint32_t v[kPerThread];
  auto issue = [&](const Tile& t, uint32_t (&vv)[kPerThread]) {
#pragma unroll
    for (int c = 0; c < kPerThread; ++c) {
      const uint32_t e = c * kThreads + threadIdx.x;
      vv[c] = (e < t.n) ? (__float_as_uint(__ldcg(P.resid + t.base + e)) & 0x7FFFFFFFu) : 0u;
    }
  };
  issue(ti, v);
  while (true) {
    const uint32_t next = tile + 1;
    const bool has_next = next < t_end;
    Tile tn = ti;
    uint32_t vn[kPerThread];
    if (has_next) { tn = load_tile(P, next); issue(tn, vn); }
    if (ti.tensor != cur) {
      cur = ti.tensor;
      load_tensor(P, cur, sm);
      T22 = __ldcg(&P.sel[cur].thr) >> 9;
    }
    if (sm.td.mode == kModeBloom) {
      uint32_t mask = 0;
#pragma unroll
      for (int c = 0; c < kPerThread; ++c) if ((v[c] >> 9) >= T22) mask |= 1u << c;   // padding lanes hold 0 (< T22 >= 1)
      if (sm.td.off_hint) build_hint(mask, sm.s, my_slot + sm.td.off_hint + 4u * (tile - sm.td.tile_begin));
      // slot allocation: warp scan of the per-thread counts + one SMEM atomic per warp
      const uint32_t lane = threadIdx.x & 31u;
      const uint32_t cnt = __popc(mask);
      uint32_t incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += nb; }
      if (threadIdx.x == 0) sm.s.lb = 0;
      __syncthreads();
      uint32_t wbase = 0;
      if (lane == 31 && incl) wbase = atomicAdd(&sm.s.lb, incl);
      wbase = __shfl_sync(0xFFFFFFFFu, wbase, 31);
      uint32_t slot = wbase + incl - cnt;
#pragma unroll
      for (int c = 0; c < kPerThread; ++c) if ((mask >> c) & 1u) queue[slot++] = ti.local0 + c * kThreads + threadIdx.x;
      __syncthreads();
      const uint32_t total = sm.s.lb, n_hash = sm.td.n_hash, m_bits = sm.td.m_bits;
      uint32_t* filter = my_slot + sm.td.off_filter;
      for (uint32_t i = threadIdx.x; i < total * n_hash; i += kThreads) {
        const uint32_t ent = i / n_hash, j = i - ent * n_hash;
        const HashAB h = hash_ab(queue[ent], P.seed);
        const uint32_t pos = mulhi32(h.a + j * h.b, m_bits);
        atomicOr(filter + (pos >> 5), 1u << (pos & 31u));
      }
      __syncthreads();
    }
    if (!has_next) break;
    tile = next; ti = tn;
#pragma unroll
    for (int c = 0; c < kPerThread; ++c) v[c] = vn[c];
  }
}

// ===========================================================================
// TMA variants of the streaming phases.  The next tiles are fetched by the TMA unit
// (cp.async.bulk global->shared, completion on an mbarrier) into a ring carved out of
// the dynamic SMEM buffer, so the prefetch depth costs no registers (the 64-register
// build spilled its register prefetch, see profiles/).  One elected thread issues.
// ===========================================================================
struct Ring {
  uint8_t* buf;
  uint32_t stage_bytes;
  uint32_t n_stages;
};

DR_D Ring ring_setup(const EngineParams& P, Smem& sm, uint32_t stage_bytes, uint32_t max_stages) {
  Ring r;
  r.buf = reinterpret_cast<uint8_t*>(g_filter_smem);
  r.stage_bytes = stage_bytes;
  r.n_stages = min(max_stages, (P.filter_smem_words * 4u) / stage_bytes);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < r.n_stages; ++i) { mbar_inval(&sm.bar[i]); mbar_init(&sm.bar[i], 1); }
    mbar_fence_init();
    fence_proxy_async();
  }
  __syncthreads();
  return r;
}

DR_D uint32_t round16(uint32_t bytes) { return (bytes + 15u) & ~15u; }

DR_D void phase_accum_tma(const EngineParams& P, Smem& sm) {
  const uint32_t parity_slot = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity_slot, P.rank);
  {
    uint4* p = reinterpret_cast<uint4*>(my_slot);
    const uint32_t n4 = (P.payload_words + 3u) >> 2;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n4; i += gridDim.x * kThreads) p[i] = z;
  }
  if (sharded(P) && blockIdx.x == 0 && threadIdx.x == 0) *s2_ptr(P.arena[P.rank], P, parity_slot, P.rank) = 0u;
  clear_hist(sm);
  const bool has_resid = (P.beta != 0.0f);
  const Ring ring = ring_setup(P, sm, 2u * kTile * 4u, 8u);      // stage = g tile | r tile
  uint32_t t0, t_end;
  tile_range(P, t0, t_end);
  const uint32_t n_my = t_end - t0;
  auto issue = [&](uint32_t i) {                                  // thread 0 only
    const Tile t = load_tile(P, t0 + i);
    const uint32_t s = i % ring.n_stages, bytes = round16(t.n * 4u);
    uint8_t* dst = ring.buf + (size_t)s * ring.stage_bytes;
    mbar_expect_tx(&sm.bar[s], has_resid ? 2u * bytes : bytes);
    bulk_g2s(dst, P.grad + t.base, bytes, &sm.bar[s]);
    if (has_resid) bulk_g2s(dst + kTile * 4u, P.resid + t.base, bytes, &sm.bar[s]);
  };
  if (threadIdx.x == 0) for (uint32_t i = 0; i < min(ring.n_stages, n_my); ++i) issue(i);
  uint32_t cur = kNoTensor, lower = 0, n_mine = 0;
  auto finish = [&]() {
    const uint32_t k = __ldg(&P.tensors[cur].k), nt = __ldg(&P.tensors[cur].n_tiles);
    if (finish_digit(P, sm, 0, cur, n_mine, nt, k)) write_digit1(P, sm, cur);
  };
  for (uint32_t i = 0; i < n_my; ++i) {
    const Tile ti = load_tile(P, t0 + i);
    if (ti.tensor != cur) {
      if (cur != kNoTensor) finish();
      cur = ti.tensor; n_mine = 0;
      const uint32_t prev = P.use_history ? __ldcg(&P.sel[cur].prev_thr) : 0u;
      lower = (prev > (1u << 23)) ? prev - (1u << P.hist_shift) : 0u;
    }
    const uint32_t s = i % ring.n_stages;
    mbar_wait(&sm.bar[s], (i / ring.n_stages) & 1u, P.status);
    const float4* sg = reinterpret_cast<const float4*>(ring.buf + (size_t)s * ring.stage_bytes);
    const float4* sr = sg + kTile / 4;
    uint32_t key[8];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t v4 = c * kThreads + threadIdx.x, e = v4 * 4u;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < ti.n) {
        const float4 g = sg[v4];
        if (has_resid) {
          const float4 r = sr[v4];
          a.x = P.beta * r.x + P.gamma * g.x; a.y = P.beta * r.y + P.gamma * g.y;
          a.z = P.beta * r.z + P.gamma * g.z; a.w = P.beta * r.w + P.gamma * g.w;
        } else {
          a.x = P.gamma * g.x; a.y = P.gamma * g.y; a.z = P.gamma * g.z; a.w = P.gamma * g.w;
        }
        *reinterpret_cast<float4*>(P.resid + ti.base + e) = a;
      }
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = e + j < ti.n;
        key[c * 4 + j] = ok ? (__float_as_uint(av[j]) & 0x7FFFFFFFu) : 0xFFFFFFFFu;
        if (ok && key[c * 4 + j] >= lower) atomicAdd(&sm.u.hist[key[c * 4 + j] >> 20], 1u);
      }
    }
    n_mine += 1;
    __syncthreads();                                               // everyone is done with stage s
    if (threadIdx.x == 0 && i + ring.n_stages < n_my) { fence_proxy_async(); issue(i + ring.n_stages); }
    if (ti.single) {
      const uint32_t k = __ldg(&P.tensors[cur].k);
      resolve_bins([&](int b) { return sm.u.hist[b]; }, kHistBins, k, sm.s);
      const uint32_t bin1 = sm.s.res[0], krem1 = sm.s.res[1];
      write_digit1(P, sm, cur);
      clear_hist(sm);
      if (bin1 != kUnsafe) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (key[j] != 0xFFFFFFFFu && (key[j] >> 20) == bin1) atomicAdd(&sm.u.hist[(key[j] >> 9) & 0x7FFu], 1u);
        resolve_bins([&](int b) { return sm.u.hist[b]; }, kHistBins, krem1, sm.s);
        write_final(P, sm, cur, bin1);
        if (threadIdx.x == 0) P.sel[cur].done_epoch = P.epoch;
        clear_hist(sm);
      }
      cur = kNoTensor; n_mine = 0;
    }
  }
  if (cur != kNoTensor) finish();
}

DR_D void phase_hist2_tma(const EngineParams& P, Smem& sm) {
  clear_hist(sm);
  const Ring ring = ring_setup(P, sm, kTile * 4u, 8u);
  uint32_t t0, t_end;
  tile_range(P, t0, t_end);
  const uint32_t n_my = t_end - t0;
  // tiles of tensors that are already done are never fetched: the issue order follows the list of active tiles
  uint32_t cur = kNoTensor, n_mine = 0, k_cur = 0, nt_cur = 0, prefix = 0;
  bool active = false;
  auto tile_active = [&](uint32_t tensor) { return __ldcg(&P.sel[tensor].done_epoch) != P.epoch; };
  // thread 0 keeps its own cursor over the active tiles to issue; consumers walk the same sequence
  uint32_t issue_pos = 0, issued = 0;         // thread 0 state
  auto issue_next = [&]() {                    // thread 0: fetch the next active tile, if any
    while (issue_pos < n_my) {
      const Tile t = load_tile(P, t0 + issue_pos);
      ++issue_pos;
      if (!tile_active(t.tensor)) continue;
      const uint32_t s = issued % ring.n_stages, bytes = round16(t.n * 4u);
      mbar_expect_tx(&sm.bar[s], bytes);
      bulk_g2s(ring.buf + (size_t)s * ring.stage_bytes, P.resid + t.base, bytes, &sm.bar[s]);
      ++issued;
      return;
    }
  };
  if (threadIdx.x == 0) for (uint32_t i = 0; i < ring.n_stages; ++i) issue_next();
  auto finish = [&]() {
    if (!active) return;
    if (finish_digit(P, sm, 2, cur, n_mine, nt_cur, k_cur)) write_final(P, sm, cur, prefix);
  };
  uint32_t consumed = 0;
  for (uint32_t i = 0; i < n_my; ++i) {
    const Tile ti = load_tile(P, t0 + i);
    if (ti.tensor != cur) {
      if (cur != kNoTensor) finish();
      cur = ti.tensor; n_mine = 0;
      nt_cur = __ldg(&P.tensors[cur].n_tiles);
      active = tile_active(cur);
      prefix = __ldcg(&P.sel[cur].bin1);
      k_cur = __ldcg(&P.sel[cur].krem1);
      if (active && prefix == kUnsafe && threadIdx.x == 0) atomicExch(P.status, kErrResolve);
    }
    if (!active) continue;
    const uint32_t s = consumed % ring.n_stages;
    mbar_wait(&sm.bar[s], (consumed / ring.n_stages) & 1u, P.status);
    const uint4* sq = reinterpret_cast<const uint4*>(ring.buf + (size_t)s * ring.stage_bytes);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t v4 = c * kThreads + threadIdx.x, e = v4 * 4u;
      if (e < ti.n) {
        const uint4 q = sq[v4];
        const uint32_t key[4] = {q.x & 0x7FFFFFFFu, q.y & 0x7FFFFFFFu, q.z & 0x7FFFFFFFu, q.w & 0x7FFFFFFFu};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (e + j < ti.n && (key[j] >> 20) == prefix) atomicAdd(&sm.u.hist[(key[j] >> 9) & 0x7FFu], 1u);
      }
    }
    n_mine += 1;
    ++consumed;
    __syncthreads();
    if (threadIdx.x == 0) { fence_proxy_async(); issue_next(); }
  }
  if (cur != kNoTensor) finish();
}

DR_D void phase_insert_tma(const EngineParams& P, Smem& sm) {
  const uint32_t parity_slot = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity_slot, P.rank);
  uint32_t* queue = sm.u.hist;
  const Ring ring = ring_setup(P, sm, kTile * 4u, 8u);
  uint32_t t0, t_end;
  tile_range(P, t0, t_end);
  const uint32_t n_my = t_end - t0;
  auto issue = [&](uint32_t i) {
    const Tile t = load_tile(P, t0 + i);
    const uint32_t s = i % ring.n_stages, bytes = round16(t.n * 4u);
    mbar_expect_tx(&sm.bar[s], bytes);
    bulk_g2s(ring.buf + (size_t)s * ring.stage_bytes, P.resid + t.base, bytes, &sm.bar[s]);
  };
  if (threadIdx.x == 0) for (uint32_t i = 0; i < min(ring.n_stages, n_my); ++i) issue(i);
  uint32_t cur = kNoTensor, T22 = 1;
  for (uint32_t i = 0; i < n_my; ++i) {
    const Tile ti = load_tile(P, t0 + i);
    if (ti.tensor != cur) {
      cur = ti.tensor;
      load_tensor(P, cur, sm);
      T22 = __ldcg(&P.sel[cur].thr) >> 9;
    }
    const uint32_t s = i % ring.n_stages;
    mbar_wait(&sm.bar[s], (i / ring.n_stages) & 1u, P.status);
    const uint32_t* sv = reinterpret_cast<const uint32_t*>(ring.buf + (size_t)s * ring.stage_bytes);
    uint32_t mask = 0;
    if (sm.td.mode == kModeBloom) {
#pragma unroll
      for (int c = 0; c < kPerThread; ++c) {
        const uint32_t e = c * kThreads + threadIdx.x;
        if (e < ti.n && ((sv[e] & 0x7FFFFFFFu) >> 9) >= T22) mask |= 1u << c;
      }
    }
    __syncthreads();                                               // stage s consumed
    if (threadIdx.x == 0 && i + ring.n_stages < n_my) { fence_proxy_async(); issue(i + ring.n_stages); }
    if (sm.td.mode == kModeBloom && sm.td.off_hint)
      build_hint(mask, sm.s, my_slot + sm.td.off_hint + 4u * ((t0 + i) - sm.td.tile_begin));
    if (sm.td.mode == kModeBloom) {
      const uint32_t lane = threadIdx.x & 31u;
      const uint32_t cnt = __popc(mask);
      uint32_t incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += nb; }
      if (threadIdx.x == 0) sm.s.lb = 0;
      __syncthreads();
      uint32_t wbase = 0;
      if (lane == 31 && incl) wbase = atomicAdd(&sm.s.lb, incl);
      wbase = __shfl_sync(0xFFFFFFFFu, wbase, 31);
      uint32_t slot = wbase + incl - cnt;
#pragma unroll
      for (int c = 0; c < kPerThread; ++c) if ((mask >> c) & 1u) queue[slot++] = ti.local0 + c * kThreads + threadIdx.x;
      __syncthreads();
      const uint32_t total = sm.s.lb, n_hash = sm.td.n_hash, m_bits = sm.td.m_bits;
      uint32_t* filter = my_slot + sm.td.off_filter;
      for (uint32_t q = threadIdx.x; q < total * n_hash; q += kThreads) {
        const uint32_t ent = q / n_hash, j = q - ent * n_hash;
        const HashAB h = hash_ab(queue[ent], P.seed);
        const uint32_t pos = mulhi32(h.a + j * h.b, m_bits);
        atomicOr(filter + (pos >> 5), 1u << (pos & 31u));
      }
      __syncthreads();
    }
  }
}

// ===========================================================================
// kModeRle helpers: 12-bit fields, LSB-first, entry j at bit 12*j of the stream
// ===========================================================================
DR_D uint32_t rle_stream_words(uint32_t val_cap) { return (val_cap * 12u + 31u) / 32u + 1u; }

DR_D void rle_put(uint32_t* stream, uint32_t j, uint32_t pos) {
  const uint32_t bit = 12u * j, w = bit >> 5, sh = bit & 31u;
  atomicOr(stream + w, pos << sh);
  if (sh > 20u) atomicOr(stream + w + 1, pos >> (32u - sh));
}

DR_D uint32_t rle_get(const uint32_t* stream, uint32_t j) {
  const uint32_t bit = 12u * j, w = bit >> 5, sh = bit & 31u;
  uint32_t v = __ldcg(stream + w) >> sh;
  if (sh > 20u) v |= __ldcg(stream + w + 1) << (32u - sh);
  return v & 0xFFFu;
}

// the emit phase ORs fields into the stream: clear it first (any phase before emit; runs with the digit-2 pass)
DR_D void rle_zero_streams(const EngineParams& P) {
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, P.epoch & 1u, P.rank);
  for (uint32_t t = 0; t < P.n_tensors; ++t) {
    const TensorDesc* td = P.tensors + t;
    if (__ldg(&td->mode) != (uint32_t)kModeRle) continue;
    const uint32_t n = rle_stream_words(__ldg(&td->val_cap));
    uint32_t* dst = my_slot + __ldg(&td->off_idx);
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) dst[i] = 0u;
  }
}

// ===========================================================================
// phase 4: universe query — per-element flags + per-tile counts
// ===========================================================================
DR_D void phase_query(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity, P.rank);
  uint32_t cur = kNoTensor, T22 = 1;
  bool staged = false;
  uint32_t tile, t_end;
  tile_range(P, tile, t_end);
  for (; tile < t_end; ++tile) {
    const Tile ti = load_tile(P, tile);
    if (ti.tensor != cur) {
      cur = ti.tensor;
      load_tensor(P, cur, sm);
      T22 = __ldcg(&P.sel[cur].thr) >> 9;
      staged = false;
      if (sm.td.mode == kModeBloom && sm.td.n_filter_words <= P.filter_smem_words) {
        stage_filter(my_slot + sm.td.off_filter, sm.td.n_filter_words);
        staged = true;
      }
    }
    uint32_t valid = 0;
#pragma unroll
    for (int c = 0; c < kPerThread; ++c) if (c * kThreads + threadIdx.x < ti.n) valid |= 1u << c;
    uint32_t flags = 0;
    if (sm.td.mode == kModeBloom) {
      const uint32_t* filter = my_slot + sm.td.off_filter;
      if (sm.td.off_hint) {
        const uint4 hq = __ldcg(reinterpret_cast<const uint4*>(my_slot + sm.td.off_hint + 4u * (tile - sm.td.tile_begin)));
        const uint32_t h[4] = {hq.x, hq.y, hq.z, hq.w};
        valid &= valid_from_hint(h);
      }
      if (staged) flags = bloom_test8(ti.local0 + threadIdx.x, valid, P.seed, sm.td.n_hash, sm.td.m_bits,
                                      [&](uint32_t w) { return g_filter_smem[w]; });
      else flags = bloom_test8(ti.local0 + threadIdx.x, valid, P.seed, sm.td.n_hash, sm.td.m_bits,
                               [&](uint32_t w) { return filter[w]; });
    } else {
#pragma unroll
      for (int c = 0; c < kPerThread; ++c) {
        const uint32_t e = c * kThreads + threadIdx.x;
        if (e < ti.n && ((__float_as_uint(__ldcg(P.resid + ti.base + e)) & 0x7FFFFFFFu) >> 9) >= T22) flags |= 1u << c;
      }
    }
    P.flag_buf[(size_t)tile * kThreads + threadIdx.x] = (uint8_t)flags;
    // per-tile count: one fire-and-forget RED per warp (tile_count is zeroed by the previous step's decode
    // phase) — no CTA barrier in this loop, so warps run ahead through their tiles independently
    if (P.warp_count) {
      const uint32_t lane = threadIdx.x & 31u;
      uint32_t pc = 0, mine = 0;
#pragma unroll
      for (int c = 0; c < kPerThread; ++c) {
        const uint32_t n = __popc(__ballot_sync(0xFFFFFFFFu, (flags >> c) & 1u));
        pc += n;
        if (lane == (uint32_t)c) mine = n;
      }
      if (lane < (uint32_t)kPerThread) P.warp_count[(size_t)tile * 128u + lane * kWarps + (threadIdx.x >> 5)] = (uint8_t)mine;
      if (lane == 0 && pc) atomicAdd(P.tile_count + tile, pc);
      continue;
    }
    uint32_t pc = __popc(flags);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pc += __shfl_xor_sync(0xFFFFFFFFu, pc, o);
    if ((threadIdx.x & 31u) == 0 && pc) atomicAdd(P.tile_count + tile, pc);
  }
}

// ===========================================================================
// phase 5: ordered compaction + value gather + residual update
// ===========================================================================
// kFull = false compiles the value-codec / run-length branches out of the two hot loops (emit, decode): the
// index-only kernel (plain pairs, bloom) keeps its registers for the probe loop instead of spilling
template <bool kFull>
DR_D void phase_emit(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity, P.rank);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    my_slot[0] = kMagic; my_slot[1] = P.epoch; my_slot[2] = P.n_tensors; my_slot[3] = P.payload_words;
    my_slot[4] = (uint32_t)P.rank;
  }
  uint32_t cur = kNoTensor, T22 = 1, excl = 0, rank_buf = 0;
  uint32_t tile, t_end;
  tile_range(P, tile, t_end);
  for (; tile < t_end; ++tile) {
    const Tile ti = load_tile(P, tile);
    if (ti.tensor != cur) {
      cur = ti.tensor;
      load_tensor(P, cur, sm);
      T22 = __ldcg(&P.sel[cur].thr) >> 9;
      // exclusive prefix at my first tile of this tensor: sum of the counts of the tensor's earlier tiles
      uint32_t part = 0;
      for (uint32_t j = sm.td.tile_begin + threadIdx.x; j < tile; j += kThreads) part += __ldcg(P.tile_count + j);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
      if (threadIdx.x == 0) sm.s.lb = 0;
      __syncthreads();
      if ((threadIdx.x & 31u) == 0 && part) atomicAdd(&sm.s.lb, part);
      __syncthreads();
      excl = sm.s.lb;
      __syncthreads();
    }
    const uint32_t tile_local = tile - sm.td.tile_begin;
    const uint32_t local0 = ti.local0;
    const size_t base = ti.base;
    DynHeader* dyn = reinterpret_cast<DynHeader*>(my_slot + kSlotHeaderWords) + cur;
    const bool last_tile = (tile_local + 1 == sm.td.n_tiles);
    const uint32_t flags = P.flag_buf[(size_t)tile * kThreads + threadIdx.x];
    uint32_t rank[kPerThread], total;
    if (P.warp_count) tile_rank_counts(flags, P.warp_count + (size_t)tile * 128u, rank, total);
    else tile_rank(flags, sm.s, rank_buf, rank, total);
    const uint32_t limit = (sm.td.mode == kModeBloom && P.policy != kPolicyP0) ? min(sm.td.k, sm.td.val_cap)
                                                                               : sm.td.val_cap;
    float* vals = reinterpret_cast<float*>(my_slot + sm.td.off_vals);
    uint32_t* idxs = my_slot + sm.td.off_idx;
    if (excl < limit) {
#pragma unroll
      for (int c = 0; c < kPerThread; ++c) {
        if ((flags >> c) & 1u) {
          const uint32_t rp = excl + rank[c];
          if (rp < limit) {
            const uint32_t e = c * kThreads + threadIdx.x;
            vals[rp] = P.resid[base + e];
            P.resid[base + e] = 0.0f;                      // residual is exactly 0 on the shipped set
            if (sm.td.mode == kModeRaw) idxs[rp] = local0 + e;
            else if (kFull && sm.td.mode == kModeRle) rle_put(idxs, rp, e);
            if (kFull && sm.td.vmode) my_slot[sm.td.off_selidx + rp] = (uint32_t)(base + e);
            if (rp == limit - 1u) dyn->cutoff = local0 + e;
          }
        }
      }
    }
    if (threadIdx.x == 0) {
      if (sm.td.mode == kModeBloom) my_slot[sm.td.off_prefix + tile_local] = min(excl, limit);
      else if (kFull && sm.td.mode == kModeRle)
        reinterpret_cast<uint16_t*>(my_slot + sm.td.off_prefix)[tile_local] =
            (uint16_t)(excl >= limit ? 0u : min(total, limit - excl));
      if (last_tile) {
        const uint32_t incl = excl + total;
        dyn->n_sel = min(incl, limit);
        dyn->n_pos = incl;
        dyn->thr_bits = T22 << 9;
        if (incl < limit) dyn->cutoff = 0xFFFFFFFFu;
        P.sel[cur].prev_thr = T22 << 9;
      }
    }
    excl += total;
  }
}

// ===========================================================================
// 'both' (bloom index + polynomial value fit): phases rank / fit / fix and the decode-side evaluation.
// Replaces the reference's PolyFit (sort + per-segment Vandermonde normal equations with a CPU 6x6
// inverse per segment, reference pytorch/deepreduce.py:306-425) and the int64 `mapping` (:263-267).
// ===========================================================================
// segment table of reference get_segments (:362-377): fine segments at both steep ends of the descending curve
DR_D void build_segments(int n, int num_pos, int* start, int& n_seg) {
  const double ratios[10] = {1.0 / 5, 1.0 / 10, 1.0 / 30, 1.0 / 100, 1.0 / 300, 1.0 / 1000, 1.0 / 3000, 1.0 / 10000, 1.0 / 30000, 1.0 / 100000};
  int pos[10], neg[10], np = 0, nn = 0, sp = 0, sn = 0;
  const int num_neg = n - num_pos;
  for (int i = 0; i < 10; ++i) {
    const int a = (int)((double)num_pos * ratios[i]);
    if (a > 30) { pos[np++] = a; sp += a; }
    const int b = (int)((double)num_neg * ratios[i]);
    if (b > 30) { neg[nn++] = b; sn += b; }
  }
  int s = 0, acc = 0;
  for (int i = np - 1; i >= 0; --i) { start[s++] = acc; acc += pos[i]; }
  start[s++] = acc; acc += num_pos - sp;
  start[s++] = acc; acc += num_neg - sn;
  for (int i = 0; i < nn; ++i) { start[s++] = acc; acc += neg[i]; }
  start[s] = acc;                               // == n
  n_seg = s;
}

DR_D float poly_value(const float* __restrict__ coef, const int* start, int n_seg, int deg, uint32_t j) {
  int s = 0;
  for (int i = 0; i < n_seg; ++i) if (start[i + 1] > start[i] && (int)j >= start[i]) s = i;
  const int len = start[s + 1] - start[s];
  float p[kMaxDeg + 1];
  gram_eval<kMaxDeg + 1>((float)((int)j - start[s]), (float)(len - 1), min(deg, len - 1), p);
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k <= kMaxDeg; ++k) if (k <= deg) acc += __ldcg(coef + s * (deg + 1) + k) * p[k];
  return acc;
}

DR_D uint32_t load_rank(const uint32_t* slot, const TensorDesc& td, uint32_t p) {
  return td.rank_u32 ? __ldcg(slot + td.off_rankmap + p)
                     : (uint32_t)__ldcg(reinterpret_cast<const uint16_t*>(slot + td.off_rankmap) + p);
}

// ---- exact descending rank: counting sort on 13 bits of the order-preserving key + all-pairs inside a bin ----
// Bins are monotone non-increasing in the value and centred on the tensor's selection threshold T (31-bit
// key): per sign 1024 coarse bins above 4T (16 per octave), 2048 fine bins on [T, 4T) (relative width 2^-10)
// and 1024 coarse bins below T (false positives carry arbitrary small values).  Exactness never depends on
// the binning — phase 9 counts inside the bin — only the amount of in-bin work does.
DR_D uint32_t rank_bin(float v, uint32_t T) {
  const uint32_t bits = __float_as_uint(v), key = bits & 0x7FFFFFFFu;
  uint32_t pb;                                   // 0 = largest magnitude ... 4095 = smallest
  if (key >= T) {
    const uint32_t d = key - T;
    if (d < (2u << 23)) pb = 1024u + (2047u - (d >> 13));
    else pb = 1023u - min(1023u, (key >> 19) - ((T + (2u << 23)) >> 19));
  } else {
    pb = 3072u + min(1023u, (T >> 19) - (key >> 19));
  }
  return (bits & 0x80000000u) ? (4096u + (4095u - pb)) : pb;
}

// phase 6: bin populations
DR_D void phase_rank_hist(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity, P.rank);
  for (uint32_t task = blockIdx.x; task < P.n_poly_tasks; task += gridDim.x) {
    const uint32_t t = __ldg(P.poly_tasks + 2 * task), p0 = __ldg(P.poly_tasks + 2 * task + 1);
    load_tensor(P, t, sm);
    if (sm.td.vmode != 1) continue;
    const DynHeader* dyn = reinterpret_cast<const DynHeader*>(my_slot + kSlotHeaderWords) + t;
    const uint32_t n = __ldcg(&dyn->n_sel), p = p0 + threadIdx.x;
    if (p < n) {
      const float v = __ldcg(reinterpret_cast<const float*>(my_slot + sm.td.off_vals) + p);
      atomicAdd(P.poly_bins + (size_t)sm.td.poly_ord * 2 * kRankBins + rank_bin(v, __ldcg(&P.sel[t].thr)), 1u);
    }
  }
}

// phase 7: per tensor, exclusive prefix of the bin counts -> bin starts (second half of the bin table)
DR_D void phase_rank_scan(const EngineParams& P, Smem& sm) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t o = blockIdx.x; o < P.n_poly; o += gridDim.x) {
    uint32_t* cnt = P.poly_bins + (size_t)o * 2 * kRankBins;
    uint32_t* start = cnt + kRankBins;
    constexpr int kPer = kRankBins / kThreads;                    // 16 consecutive bins per thread
    uint32_t c[kPer], sum = 0;
#pragma unroll
    for (int i = 0; i < kPer; ++i) { c[i] = __ldcg(cnt + threadIdx.x * kPer + i); sum += c[i]; }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (uint32_t)d) incl += nb; }
    __syncthreads();
    if (lane == 31) sm.s.warp_tot[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (uint32_t w = 0; w < warp; ++w) base += sm.s.warp_tot[w];
    uint32_t run = base + incl - sum;
#pragma unroll
    for (int i = 0; i < kPer; ++i) { start[threadIdx.x * kPer + i] = run; run += c[i]; }
  }
}

// phase 8: group values by bin (order inside a bin is arbitrary here; phase 9 makes the rank exact)
DR_D void phase_rank_scatter(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity, P.rank);
  for (uint32_t task = blockIdx.x; task < P.n_poly_tasks; task += gridDim.x) {
    const uint32_t t = __ldg(P.poly_tasks + 2 * task), p0 = __ldg(P.poly_tasks + 2 * task + 1);
    load_tensor(P, t, sm);
    if (sm.td.vmode != 1) continue;
    const DynHeader* dyn = reinterpret_cast<const DynHeader*>(my_slot + kSlotHeaderWords) + t;
    const uint32_t n = __ldcg(&dyn->n_sel), p = p0 + threadIdx.x;
    if (p < n) {
      const float v = __ldcg(reinterpret_cast<const float*>(my_slot + sm.td.off_vals) + p);
      uint32_t* tab = P.poly_bins + (size_t)sm.td.poly_ord * 2 * kRankBins;
      const uint32_t b = rank_bin(v, __ldcg(&P.sel[t].thr));
      // the count array is re-used as the running cursor: it is decremented down to 0 while filling the bin
      const uint32_t within = atomicSub(tab + b, 1u) - 1u;
      const uint32_t pos = __ldcg(tab + kRankBins + b) + within;
      P.bucket_val[sm.td.poly_off + pos] = v;
      P.bucket_pos[sm.td.poly_off + pos] = p;
    }
  }
}

// phase 9: exact rank = bin start + #(bin mates that sort before me); writes rank map, sorted values, num_pos
DR_D void phase_rank_exact(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity, P.rank);
  for (uint32_t task = blockIdx.x; task < P.n_poly_tasks; task += gridDim.x) {
    const uint32_t t = __ldg(P.poly_tasks + 2 * task), p0 = __ldg(P.poly_tasks + 2 * task + 1);
    load_tensor(P, t, sm);
    if (sm.td.vmode != 1) continue;
    const DynHeader* dyn = reinterpret_cast<const DynHeader*>(my_slot + kSlotHeaderWords) + t;
    const uint32_t n = __ldcg(&dyn->n_sel), i = p0 + threadIdx.x;   // i = position in the grouped arrays
    float v = 0.f;
    if (i < n) {
      const float* bv = P.bucket_val + sm.td.poly_off;
      const uint32_t* bp = P.bucket_pos + sm.td.poly_off;
      v = __ldcg(bv + i);
      const uint32_t p = __ldcg(bp + i), b = rank_bin(v, __ldcg(&P.sel[t].thr));
      const uint32_t* start = P.poly_bins + (size_t)sm.td.poly_ord * 2 * kRankBins + kRankBins;
      const uint32_t lo = __ldcg(start + b), hi = (b + 1 < (uint32_t)kRankBins) ? __ldcg(start + b + 1) : n;
      uint32_t rank = lo;
      for (uint32_t q = lo; q < hi; ++q) {
        const float w = __ldcg(bv + q);
        rank += (w > v || (w == v && __ldcg(bp + q) < p)) ? 1u : 0u;
      }
      if (sm.td.rank_u32) my_slot[sm.td.off_rankmap + p] = rank;
      else reinterpret_cast<uint16_t*>(my_slot + sm.td.off_rankmap)[p] = (uint16_t)rank;
      reinterpret_cast<float*>(my_slot + sm.td.off_sorted)[rank] = v;
    }
    const uint32_t pc = __syncthreads_count(i < n && v > 0.f);
    if (threadIdx.x == 0) {
      uint32_t* tail = my_slot + sm.td.off_coef + kMaxSeg * (sm.td.poly_degree + 1);
      if (pc) atomicAdd(tail, pc);
      if (p0 == 0) tail[1] = n;
    }
  }
}

// phase 10: one warp per (tensor, segment): Gram least squares  c_k = sum p_k y / sum p_k^2
DR_D void phase_fit(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity, P.rank);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t gw = blockIdx.x * kWarps + (threadIdx.x >> 5), nw = gridDim.x * kWarps;
  for (uint32_t task = gw; task < P.n_poly * kMaxSeg; task += nw) {
    const uint32_t t = __ldg(P.poly_tensors + task / kMaxSeg), s = task % kMaxSeg;
    const TensorDesc* td = P.tensors + t;
    const uint32_t off_coef = __ldg(&td->off_coef), off_sorted = __ldg(&td->off_sorted);
    const int deg = (int)__ldg(&td->poly_degree);
    const uint32_t* tail = my_slot + off_coef + kMaxSeg * (deg + 1);
    const int num_pos = (int)__ldcg(tail), n = (int)__ldcg(tail + 1);
    int start[kMaxSeg + 2], n_seg;
    build_segments(n, num_pos, start, n_seg);
    if ((int)s >= n_seg) continue;
    const int len = start[s + 1] - start[s];
    if (len <= 0) continue;
    const float* y = reinterpret_cast<const float*>(my_slot + off_sorted) + start[s];
    const int deg_eff = min(deg, len - 1);
    // recurrence constants of this segment (no divisions in the inner loop):
    //   p_{k+1} = ra[k] * (N - 2x) * p_k - rb[k] * p_{k-1}
    const float N = (float)(len - 1), invN = len > 1 ? 1.f / N : 0.f;
    float ra[kMaxDeg], rb[kMaxDeg];
#pragma unroll
    for (int k = 1; k < kMaxDeg; ++k) {
      const float dnm = (k + 1.f) * (N - k);
      ra[k] = (k < deg_eff) ? (2.f * k + 1.f) / dnm : 0.f;
      rb[k] = (k < deg_eff) ? (float)k * (N + k + 1.f) / dnm : 0.f;
    }
    float num[kMaxDeg + 1], den[kMaxDeg + 1];
#pragma unroll
    for (int k = 0; k <= kMaxDeg; ++k) { num[k] = 0.f; den[k] = 0.f; }
    for (int x0 = lane; x0 < len; x0 += 32 * 8) {                  // 8 independent loads in flight per lane
      float yv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) yv[u] = (x0 + 32 * u < len) ? __ldcg(y + x0 + 32 * u) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int x = x0 + 32 * u;
        if (x < len) {
          float p[kMaxDeg + 1];
          const float uu = N - 2.f * (float)x;
          p[0] = 1.f;
          p[1] = deg_eff >= 1 ? uu * invN : 0.f;
#pragma unroll
          for (int k = 1; k < kMaxDeg; ++k) p[k + 1] = ra[k] * uu * p[k] - rb[k] * p[k - 1];
#pragma unroll
          for (int k = 0; k <= kMaxDeg; ++k) { num[k] += p[k] * yv[u]; den[k] += p[k] * p[k]; }
        }
      }
    }
#pragma unroll
    for (int k = 0; k <= kMaxDeg; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        num[k] += __shfl_xor_sync(0xFFFFFFFFu, num[k], o);
        den[k] += __shfl_xor_sync(0xFFFFFFFFu, den[k], o);
      }
    }
    float* coef = reinterpret_cast<float*>(my_slot + off_coef) + s * (deg + 1);
#pragma unroll
    for (int k = 0; k <= kMaxDeg; ++k)
      if ((int)lane == k && k <= deg) coef[k] = (k <= deg_eff && den[k] > 0.f) ? num[k] / den[k] : 0.f;
  }
}

// phase 11: error feedback sees the fit error: resid[idx_p] = value_p - fitted_p
DR_D void phase_fix(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* my_slot = slot_ptr(P.arena[P.rank], P, parity, P.rank);
  for (uint32_t task = blockIdx.x; task < P.n_poly_tasks; task += gridDim.x) {
    const uint32_t t = __ldg(P.poly_tasks + 2 * task), p0 = __ldg(P.poly_tasks + 2 * task + 1);
    load_tensor(P, t, sm);
    if (sm.td.vmode == 2) {
      // bucketed QSGD (reference QSGD, pytorch/deepreduce.py:849-907, which syncs the host once per bucket):
      // this CTA owns one 512-value bucket: L2 norm, stochastic rounding with a counter-based RNG, int8 level
      const DynHeader* dyn = reinterpret_cast<const DynHeader*>(my_slot + kSlotHeaderWords) + t;
      const uint32_t nq = __ldcg(&dyn->n_sel);
      if (p0 >= nq) continue;
      const uint32_t p = p0 + threadIdx.x;
      const float v = p < nq ? __ldcg(reinterpret_cast<const float*>(my_slot + sm.td.off_vals) + p) : 0.f;
      float ss = v * v;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
      __syncthreads();
      if ((threadIdx.x & 31u) == 0) sm.s.warp_tot[threadIdx.x >> 5] = __float_as_uint(ss);
      __syncthreads();
      float tot = 0.f;
      for (int w = 0; w < kWarps; ++w) tot += __uint_as_float(sm.s.warp_tot[w]);
      const float norm = sqrtf(tot), q = (float)sm.td.poly_degree;
      if (threadIdx.x == 0) reinterpret_cast<float*>(my_slot + sm.td.off_coef)[p0 >> 9] = norm;
      if (p < nq) {
        const float lf = (norm > 0.f ? q / norm : 0.f) * fabsf(v);
        const float prev = floorf(lf);
        const float u = (float)((double)policy_hash(p, 0x51EDu + P.epoch) / 4294967296.0);
        float l = prev + ((u < (lf - prev)) ? 1.f : 0.f);
        l = v > 0.f ? l : (v < 0.f ? -l : 0.f);
        reinterpret_cast<int8_t*>(my_slot + sm.td.off_rankmap)[p] = (int8_t)l;
        P.resid[__ldcg(my_slot + sm.td.off_selidx + p)] = v - norm / q * l;
      }
      continue;
    }
    const int deg = (int)sm.td.poly_degree;
    const uint32_t* tail = my_slot + sm.td.off_coef + kMaxSeg * (deg + 1);
    const int num_pos = (int)__ldcg(tail), n = (int)__ldcg(tail + 1);
    if ((int)p0 >= n) continue;
    if (threadIdx.x == 0) build_segments(n, num_pos, sm.seg_start, sm.n_seg);
    __syncthreads();
    const uint32_t p = p0 + threadIdx.x;
    if ((int)p < n) {
      const float fitted = poly_value(reinterpret_cast<const float*>(my_slot + sm.td.off_coef), sm.seg_start, sm.n_seg,
                                      deg, load_rank(my_slot, sm.td, p));
      const float v = __ldcg(reinterpret_cast<const float*>(my_slot + sm.td.off_vals) + p);
      P.resid[__ldcg(my_slot + sm.td.off_selidx + p)] = v - fitted;
    }
  }
}

// ===========================================================================
// phase 9/10: push + flags
// ===========================================================================
DR_D void phase_push(const EngineParams& P) {
  const uint32_t parity = P.epoch & 1u;
  const uint4* src = reinterpret_cast<const uint4*>(slot_ptr(P.arena[P.rank], P, parity, P.rank));
  const uint32_t n4 = (P.payload_words + 3u) >> 2;
  if (P.mc_arena) {                                      // NVLS: one multimem store lands in every GPU's arena (the switch replicates)
    uint4* dst = reinterpret_cast<uint4*>(slot_ptr(P.mc_arena, P, parity, P.rank));
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n4; i += gridDim.x * kThreads) multimem_st_v4(dst + i, __ldcg(src + i));
    __threadfence_system();
    return;
  }
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

// phase 14: evaluate every rank's fitted curve once (dense, all lanes busy); decode then gathers fitted[rank]
DR_D void phase_expand(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* arena = P.arena[P.rank];
  for (uint32_t wt = blockIdx.x; wt < P.n_poly_tasks * (uint32_t)P.world; wt += gridDim.x) {
    const uint32_t r = wt / P.n_poly_tasks, task = wt - r * P.n_poly_tasks;
    const uint32_t t = __ldg(P.poly_tasks + 2 * task), j0 = __ldg(P.poly_tasks + 2 * task + 1);
    load_tensor(P, t, sm);
    if (sm.td.vmode != 1) continue;
    const uint32_t* slot = slot_ptr(arena, P, parity, (int)r);
    const int deg = (int)sm.td.poly_degree;
    const uint32_t* tail = slot + sm.td.off_coef + kMaxSeg * (deg + 1);
    const int num_pos = (int)__ldcg(tail), n = (int)__ldcg(tail + 1);
    if ((int)j0 >= n) continue;
    if (threadIdx.x == 0) build_segments(n, num_pos, sm.seg_start, sm.n_seg);
    __syncthreads();
    const uint32_t j = j0 + threadIdx.x;
    if ((int)j < n)
      P.expand_buf[(size_t)r * P.poly_total + sm.td.poly_off + j] =
          poly_value(reinterpret_cast<const float*>(slot + sm.td.off_coef), sm.seg_start, sm.n_seg, deg, j);
  }
}

// ===========================================================================
// phase 15: decode.  Contiguous tile range per CTA; rank-major so one staged
// filter serves all of the CTA's tiles of that tensor; sparse RMW into the
// zero-filled dense output.
// ===========================================================================
DR_D uint32_t lower_bound_u32(const uint32_t* a, uint32_t n, uint32_t x) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldcg(a + mid) < x) lo = mid + 1; else hi = mid; }
  return lo;
}

template <bool kFull>
DR_D void phase_decode(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t* arena = P.arena[P.rank];
  {  // hist arrays are free after the insert phase: zero them for the next step
    uint4* h = reinterpret_cast<uint4*>(P.hist);
    const size_t n4 = (size_t)kNumHist * P.n_tensors * kHistBins / 4;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (size_t)gridDim.x * kThreads) h[i] = z;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < (uint32_t)kNumHist * P.n_tensors; i += gridDim.x * kThreads)
      P.hist_total[i] = 0u;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < P.n_tiles; i += gridDim.x * kThreads) P.tile_count[i] = 0u;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < P.n_poly * 2u * kRankBins; i += gridDim.x * kThreads)
      P.poly_bins[i] = 0u;
  }
  uint32_t tile, t_end, rank_buf = 0;
  {
    uint32_t s_begin, s_end;
    decode_span(P, P.rank, s_begin, s_end);
    const uint32_t span = s_end - s_begin;
    tile = s_begin + (uint32_t)(((uint64_t)span * blockIdx.x) / gridDim.x);
    t_end = s_begin + (uint32_t)(((uint64_t)span * (blockIdx.x + 1)) / gridDim.x);
  }
  while (tile < t_end) {
    const Tile t0 = load_tile(P, tile);
    const uint32_t t = t0.tensor;
    load_tensor(P, t, sm);
    const uint32_t seg_end = min(t_end, sm.td.tile_begin + sm.td.n_tiles);
    if (sm.td.mode == kModeBloom) {
      // 1) zero-fill my tiles of this tensor (contiguous in the flat buffer)
      {
        const Tile tl = load_tile(P, seg_end - 1);
        const uint32_t n_elems = (tl.base + tl.n) - t0.base;
        float4* dst = reinterpret_cast<float4*>(P.grad + t0.base);
        const uint32_t n4 = n_elems >> 2;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t i = threadIdx.x; i < n4; i += kThreads) dst[i] = z;
        for (uint32_t i = (n4 << 2) + threadIdx.x; i < n_elems; i += kThreads) P.grad[t0.base + i] = 0.f;
      }
      __syncthreads();
      const uint32_t n_hash = sm.td.n_hash, m_bits = sm.td.m_bits;
      const bool fits = sm.td.n_filter_words <= P.filter_smem_words;
      for (int r = 0; r < P.world; ++r) {
        const uint32_t* slot = slot_ptr(arena, P, parity, r);
        const DynHeader* dyn = reinterpret_cast<const DynHeader*>(slot + kSlotHeaderWords) + t;
        const uint32_t n_sel = __ldcg(&dyn->n_sel), cutoff = __ldcg(&dyn->cutoff);
        if (n_sel == 0) continue;
        const uint32_t* filter = slot + sm.td.off_filter;
        const float* vals = reinterpret_cast<const float*>(slot + sm.td.off_vals);
        const float* fitted = P.expand_buf + (size_t)r * P.poly_total + sm.td.poly_off;   // 'both': rank r's curve
        const bool own = P.own_flags && r == P.rank;               // my own positives are already in flag_buf (query phase)
        if (fits && !own) stage_filter(filter, sm.td.n_filter_words);
        for (uint32_t tl = tile; tl < seg_end; ++tl) {
          const Tile ti = load_tile(P, tl);
          const uint32_t pre = __ldcg(slot + sm.td.off_prefix + (tl - sm.td.tile_begin));
          if (!(pre < n_sel && ti.local0 <= cutoff)) continue;       // tile-uniform: nothing of rank r lands here
          uint32_t valid = 0;
#pragma unroll
          for (int c = 0; c < kPerThread; ++c) {
            const uint32_t e = c * kThreads + threadIdx.x;
            if (e < ti.n && ti.local0 + e <= cutoff) valid |= 1u << c;
          }
          if (sm.td.off_hint) {
            const uint4 hq = __ldcg(reinterpret_cast<const uint4*>(slot + sm.td.off_hint + 4u * (tl - sm.td.tile_begin)));
            if ((hq.x | hq.y | hq.z | hq.w) == 0u) continue;           // tile-uniform: rank r selected nothing here
            const uint32_t h[4] = {hq.x, hq.y, hq.z, hq.w};
            valid &= valid_from_hint(h);
          }
          uint32_t flags;
          if (own) flags = (uint32_t)P.flag_buf[(size_t)tl * kThreads + threadIdx.x] & valid;
          else if (fits) flags = bloom_test8(ti.local0 + threadIdx.x, valid, P.seed, n_hash, m_bits,
                                             [&](uint32_t w) { return g_filter_smem[w]; });
          else flags = bloom_test8(ti.local0 + threadIdx.x, valid, P.seed, n_hash, m_bits,
                                   [&](uint32_t w) { return filter[w]; });
          uint32_t rank[kPerThread], total;
          tile_rank(flags, sm.s, rank_buf, rank, total);
#pragma unroll
          for (int c = 0; c < kPerThread; ++c) {
            if ((flags >> c) & 1u) {
              const uint32_t rp = pre + rank[c];
              if (rp < n_sel) {
                float* o = P.grad + ti.base + c * kThreads + threadIdx.x;   // the same thread owns this element for every rank
                float val;
                if (kFull && sm.td.vmode == 1) val = __ldcg(fitted + load_rank(slot, sm.td, rp));
                else if (kFull && sm.td.vmode == 2)
                  val = __ldcg(reinterpret_cast<const float*>(slot + sm.td.off_coef) + (rp >> 9)) / (float)sm.td.poly_degree *
                        (float)__ldcg(reinterpret_cast<const int8_t*>(slot + sm.td.off_rankmap) + rp);
                else val = __ldcg(vals + rp);
                *o = *o + val * P.scale;
              }
            }
          }
        }
      }
      tile = seg_end;
    } else if (kFull && sm.td.mode == kModeRle) {
      // running entry prefix of every sender at my first tile of this tensor = sum of the earlier tiles' counts
      __syncthreads();
      if (threadIdx.x < 16) sm.s.rle_pre[threadIdx.x] = 0u;
      __syncthreads();
      const uint32_t first_local = tile - sm.td.tile_begin;
      for (int r = 0; r < P.world; ++r) {
        const uint16_t* cnt = reinterpret_cast<const uint16_t*>(slot_ptr(arena, P, parity, r) + sm.td.off_prefix);
        uint32_t part = 0;
        for (uint32_t j = threadIdx.x; j < first_local; j += kThreads) part += __ldcg(cnt + j);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
        if ((threadIdx.x & 31u) == 0 && part) atomicAdd(&sm.s.rle_pre[r], part);
      }
      for (; tile < seg_end; ++tile) {
        const Tile ti = load_tile(P, tile);
        __syncthreads();
        for (int j = threadIdx.x; j < kTile; j += kThreads) sm.u.acc[j] = 0.0f;
        __syncthreads();
        for (int r = 0; r < P.world; ++r) {
          const uint32_t* slot = slot_ptr(arena, P, parity, r);
          const uint32_t c = __ldcg(reinterpret_cast<const uint16_t*>(slot + sm.td.off_prefix) + (tile - sm.td.tile_begin));
          const uint32_t pre = sm.s.rle_pre[r];
          const float* vals = reinterpret_cast<const float*>(slot + sm.td.off_vals);
          for (uint32_t j = threadIdx.x; j < c; j += kThreads) {
            const uint32_t rp = pre + j;
            if (rp < sm.td.val_cap) sm.u.acc[rle_get(slot + sm.td.off_idx, rp)] += __ldcg(vals + rp) * P.scale;   // distinct positions per sender
          }
          __syncthreads();                                  // senders are added in rank order: deterministic sums
          if (threadIdx.x == 0) sm.s.rle_pre[r] = pre + c;
        }
        for (uint32_t e = threadIdx.x; e < ti.n; e += kThreads) P.grad[ti.base + e] = sm.u.acc[e];
      }
    } else {
      for (; tile < seg_end; ++tile) {
        const Tile ti = load_tile(P, tile);
        __syncthreads();
        for (int j = threadIdx.x; j < kTile; j += kThreads) sm.u.acc[j] = 0.0f;
        __syncthreads();
        for (int r = 0; r < P.world; ++r) {
          const uint32_t* slot = slot_ptr(arena, P, parity, r);
          const DynHeader* dyn = reinterpret_cast<const DynHeader*>(slot + kSlotHeaderWords) + t;
          const uint32_t n_sel = min(__ldcg(&dyn->n_sel), sm.td.val_cap);
          const uint32_t* idxs = slot + sm.td.off_idx;
          const float* vals = reinterpret_cast<const float*>(slot + sm.td.off_vals);
          const uint32_t lo = lower_bound_u32(idxs, n_sel, ti.local0);
          const uint32_t hi = lower_bound_u32(idxs, n_sel, ti.local0 + ti.n);
          for (uint32_t j = lo + threadIdx.x; j < hi; j += kThreads)
            atomicAdd(&sm.u.acc[__ldcg(idxs + j) - ti.local0], __ldcg(vals + j) * P.scale);
        }
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < ti.n; e += kThreads) P.grad[ti.base + e] = sm.u.acc[e];
      }
    }
  }
}

// ===========================================================================
// sharded decode, stage 2 (W > 1): my decoded slice -> exact (index, value) list -> peers
// ===========================================================================
DR_D void phase_compact(const EngineParams& P, Smem& sm) {
  const uint32_t parity = P.epoch & 1u;
  uint32_t my_b, my_e;
  decode_span(P, P.rank, my_b, my_e);
  // (a) zero every tile outside my slice (peers' lists are scattered into them later)
  for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    if (tile >= my_b && tile < my_e) continue;
    const Tile ti = load_tile(P, tile);
    float4* dst = reinterpret_cast<float4*>(P.grad + ti.base);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t i = threadIdx.x; i < ((ti.n + 3u) >> 2); i += kThreads) dst[i] = z;   // tensors are padded to 32 floats
  }
  // (b) non-zeros of my slice -> my stage-2 slot (unordered: slices are disjoint, receivers only write)
  uint32_t* s2 = s2_ptr(P.arena[P.rank], P, parity, P.rank);
  uint32_t* idx_out = s2 + 4;
  float* val_out = reinterpret_cast<float*>(s2 + 4 + P.s2_cap);
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t tile = my_b + blockIdx.x; tile < my_e; tile += gridDim.x) {
    const Tile ti = load_tile(P, tile);
    float v[kPerThread];
    uint32_t nz = 0;
#pragma unroll
    for (int c = 0; c < kPerThread; ++c) {
      const uint32_t e = c * kThreads + threadIdx.x;
      v[c] = e < ti.n ? __ldcg(P.grad + ti.base + e) : 0.f;
      if (v[c] != 0.f) nz |= 1u << c;
    }
    const uint32_t cnt = __popc(nz);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += nb; }
    uint32_t wbase = 0;
    if (lane == 31 && incl) wbase = atomicAdd(s2, incl);          // s2[0] = entry count (zeroed in the accumulate phase)
    wbase = __shfl_sync(0xFFFFFFFFu, wbase, 31);
    uint32_t pos = wbase + incl - cnt;
#pragma unroll
    for (int c = 0; c < kPerThread; ++c) {
      if ((nz >> c) & 1u) {
        if (pos < P.s2_cap) { idx_out[pos] = ti.base + c * kThreads + threadIdx.x; val_out[pos] = v[c]; }
        else atomicExch(P.status, 6u);                            // stage-2 capacity exceeded
        ++pos;
      }
    }
  }
}

DR_D void phase_push2(const EngineParams& P) {
  const uint32_t parity = P.epoch & 1u;
  const uint32_t* src = s2_ptr(P.arena[P.rank], P, parity, P.rank);
  const uint32_t n = min(__ldcg(src), P.s2_cap);
  if (P.mc_arena) {
    uint32_t* dst = s2_ptr(P.mc_arena, P, parity, P.rank);
    const uint32_t gtid = blockIdx.x * kThreads + threadIdx.x, gsz = gridDim.x * kThreads;
    if (gtid == 0) multimem_st_v4(reinterpret_cast<uint4*>(dst), make_uint4(n, P.epoch, 0u, 0u));
    const uint4* si = reinterpret_cast<const uint4*>(src + 4);
    const uint4* sv = reinterpret_cast<const uint4*>(src + 4 + P.s2_cap);
    uint4* di = reinterpret_cast<uint4*>(dst + 4);
    uint4* dv = reinterpret_cast<uint4*>(dst + 4 + P.s2_cap);
    const uint32_t n4 = (n + 3u) >> 2;
    for (uint32_t i = gtid; i < n4; i += gsz) { multimem_st_v4(di + i, __ldcg(si + i)); multimem_st_v4(dv + i, __ldcg(sv + i)); }
    __threadfence_system();
    return;
  }
  for (int h = 1; h < P.world; ++h) {
    const int peer = (P.rank + h) % P.world;
    uint32_t* dst = s2_ptr(P.arena[peer], P, parity, P.rank);
    const uint32_t gtid = blockIdx.x * kThreads + threadIdx.x, gsz = gridDim.x * kThreads;
    if (gtid < 4) dst[gtid] = gtid == 0 ? n : (gtid == 1 ? P.epoch : 0u);
    // idx block and val block, 16 bytes at a time
    const uint4* si = reinterpret_cast<const uint4*>(src + 4);
    const uint4* sv = reinterpret_cast<const uint4*>(src + 4 + P.s2_cap);
    uint4* di = reinterpret_cast<uint4*>(dst + 4);
    uint4* dv = reinterpret_cast<uint4*>(dst + 4 + P.s2_cap);
    const uint32_t n4 = (n + 3u) >> 2;
    for (uint32_t i = gtid; i < n4; i += gsz) {
      const uint4 a = __ldcg(si + i), b = __ldcg(sv + i);
      asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(di + i), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w) : "memory");
      asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dv + i), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
    }
  }
  __threadfence_system();
}

DR_D void phase_signal2(const EngineParams& P) {
  const int p = threadIdx.x;
  if (blockIdx.x == 0 && p < P.world && p != P.rank) st_release_sys(P.arena[p] + kArenaFlagWords + P.rank, P.epoch);
  if (p < P.world && p != P.rank) {
    const uint32_t* flag = P.arena[P.rank] + kArenaFlagWords + p;
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(flag) - P.epoch) < 0) {
      if (++spins > P.spin_limit) { atomicExch(P.status, kErrPeerWait); atomicExch(P.status + 1, 100u + (uint32_t)p); break; }
      __nanosleep(100);
    }
  }
  __syncthreads();
}

DR_D void phase_scatter(const EngineParams& P) {
  const uint32_t parity = P.epoch & 1u;
  const uint32_t gtid = blockIdx.x * kThreads + threadIdx.x, gsz = gridDim.x * kThreads;
  for (int h = 1; h < P.world; ++h) {
    const int r = (P.rank + h) % P.world;
    const uint32_t* s2 = s2_ptr(P.arena[P.rank], P, parity, r);
    const uint32_t n = min(__ldcg(s2), P.s2_cap);
    const uint32_t* idx = s2 + 4;
    const float* val = reinterpret_cast<const float*>(s2 + 4 + P.s2_cap);
    for (uint32_t i = gtid; i < n; i += gsz) P.grad[__ldcg(idx + i)] = __ldcg(val + i);
  }
}

template <int kMinBlocks, bool kFull>
__global__ void __launch_bounds__(kThreads, kMinBlocks) dr_engine_kernel(const __grid_constant__ EngineParams P) {
  __shared__ Smem sm;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&sm.bar[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t bar_epoch = 0;
  for (int ph = P.phase_begin; ph < P.phase_end; ++ph) {
    bool ran = true;
    switch (ph) {
      case kPhAccum: if (P.use_tma) phase_accum_tma(P, sm); else phase_accum(P, sm); break;
      case kPhFallback: if (P.use_history) hist_tiles<1>(P, sm); else ran = false; break;
      case kPhHist2:
        if (kFull && P.has_rle) rle_zero_streams(P);
        if (P.use_tma) phase_hist2_tma(P, sm); else hist_tiles<2>(P, sm);
        break;
      case kPhInsert: if (P.use_tma) phase_insert_tma(P, sm); else phase_insert(P, sm); break;
      case kPhQuery: phase_query(P, sm); break;
      case kPhEmit: phase_emit<kFull>(P, sm); break;
      case kPhRankHist: if constexpr (kFull) { if (P.n_poly) phase_rank_hist(P, sm); else ran = false; } else ran = false; break;
      case kPhRankScan: if constexpr (kFull) { if (P.n_poly) phase_rank_scan(P, sm); else ran = false; } else ran = false; break;
      case kPhRankScatter: if constexpr (kFull) { if (P.n_poly) phase_rank_scatter(P, sm); else ran = false; } else ran = false; break;
      case kPhRankExact: if constexpr (kFull) { if (P.n_poly) phase_rank_exact(P, sm); else ran = false; } else ran = false; break;
      case kPhFit: if constexpr (kFull) { if (P.n_poly) phase_fit(P, sm); else ran = false; } else ran = false; break;
      case kPhFix: if constexpr (kFull) { if (P.n_poly_tasks) phase_fix(P, sm); else ran = false; } else ran = false; break;
      case kPhExpand: if constexpr (kFull) { if (P.n_poly) phase_expand(P, sm); else ran = false; } else ran = false; break;
      case kPhPush: if (P.world > 1) phase_push(P); else ran = false; break;
      case kPhSignal: if (P.world > 1) phase_signal(P); else ran = false; break;
      case kPhDecode: phase_decode<kFull>(P, sm); break;
      case kPhCompact: if (sharded(P)) phase_compact(P, sm); else ran = false; break;
      case kPhPush2: if (sharded(P)) phase_push2(P); else ran = false; break;
      case kPhSignal2: if (sharded(P)) phase_signal2(P); else ran = false; break;
      case kPhScatter: if (sharded(P)) phase_scatter(P); else ran = false; break;
      default: ran = false; break;
    }
    // a barrier separates dependent phases; signal->decode needs none (every CTA waits itself)
    if (ran && ph + 1 < P.phase_end && !(ph == kPhSignal) && !(ph == kPhSignal2)) grid_barrier(P.barrier, bar_epoch, P.status, P.spin_limit);
  }
}

}  // namespace

// ---------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------
// Two register budgets of the same kernel: <1> = 128 regs, one CTA per SM, up to 200 KB of
// filter staging; <2> = 64 regs, two CTAs per SM, up to 88 KB each.
static bool g_attr_set = false;

// ... x two feature sets: <.., false> index-only (plain pairs / bloom), <.., true> + value codecs and run-length index
static const void* kernel_for(int blocks_per_sm, bool full) {
  if (blocks_per_sm >= 2) return full ? (const void*)dr_engine_kernel<2, true> : (const void*)dr_engine_kernel<2, false>;
  return full ? (const void*)dr_engine_kernel<1, true> : (const void*)dr_engine_kernel<1, false>;
}

static void ensure_attr() {
  if (!g_attr_set) {
    cudaFuncSetAttribute(dr_engine_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(dr_engine_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(dr_engine_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 88 * 1024);
    cudaFuncSetAttribute(dr_engine_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 88 * 1024);
    g_attr_set = true;
  }
}

int engine_max_grid(int blocks_per_sm, int dyn_smem_bytes) {
  int dev = 0, sms = 0, occ = 0;
  ensure_attr();
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel_for(blocks_per_sm, true), kThreads, (size_t)dyn_smem_bytes);
  int occ_plain = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_plain, kernel_for(blocks_per_sm, false), kThreads, (size_t)dyn_smem_bytes);
  if (occ_plain < occ) occ = occ_plain;
  if (occ < 1) occ = 1;
  if (blocks_per_sm > 0 && blocks_per_sm < occ) occ = blocks_per_sm;
  return occ * sms;
}

cudaError_t engine_launch(const EngineParams& P, int grid, int blocks_per_sm, int dyn_smem_bytes, cudaStream_t stream) {
  ensure_attr();
  cudaError_t e = cudaMemsetAsync(P.barrier, 0, sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  void* args[] = {const_cast<EngineParams*>(&P)};
  count_launch(1);
  const bool full = P.n_poly != 0 || P.n_poly_tasks != 0 || P.has_rle != 0;
  return cudaLaunchCooperativeKernel(kernel_for(blocks_per_sm, full), dim3(grid), dim3(kThreads), args,
                                     (size_t)dyn_smem_bytes, stream);
}

}  // namespace dr
