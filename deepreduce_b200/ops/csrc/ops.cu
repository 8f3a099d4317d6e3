// Stand-alone sm_100a kernels behind the GRACE-compatible per-tensor codec API
// (deepreduce_b200/codecs/*): bloom insert / universe query+select, QSGD,
// bit packing, Gram-polynomial fit/eval, delta+bp128 integer coding.
// Each has a plain-torch oracle in the codec module; tests compare them.
#include "common.cuh"
#include "ops.h"

namespace dr {
namespace {

// ---------------------------------------------------------------------------
// bloom
// ---------------------------------------------------------------------------
__global__ void bloom_insert_kernel(const int64_t* __restrict__ idx, int64_t n, uint32_t* filter,
                                    uint32_t n_hash, uint32_t m_bits, uint32_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    bloom_set(filter, (uint32_t)idx[i], seed, n_hash, m_bits);
}

// pass A: positives per tile;  pass C: emit ascending indices with rank < limit
template <bool kEmit>
__global__ void __launch_bounds__(kThreads) bloom_query_kernel(const uint32_t* __restrict__ filter, uint32_t d,
                                                               uint32_t n_hash, uint32_t m_bits, uint32_t seed,
                                                               uint32_t* __restrict__ tile_counts,
                                                               const uint32_t* __restrict__ tile_excl,
                                                               int64_t* __restrict__ out, uint32_t limit) {
  __shared__ uint32_t cnt[kPerThread * kWarps + 1];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t tile = blockIdx.x; tile * (uint32_t)kTile < d; tile += gridDim.x) {
    const uint32_t local0 = tile * kTile;
    uint32_t ball[kPerThread];
#pragma unroll
    for (int c = 0; c < kPerThread; ++c) {
      const uint32_t gi = local0 + c * kThreads + threadIdx.x;
      const bool hit = gi < d && bloom_test(gi, seed, n_hash, m_bits, [&](uint32_t w) { return __ldg(filter + w); });
      ball[c] = __ballot_sync(0xFFFFFFFFu, hit);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < kPerThread; ++c) cnt[c * kWarps + warp] = __popc(ball[c]);
    }
    __syncthreads();
    if (warp == 0) {
      uint32_t v[4], sum = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[i] = cnt[lane * 4 + i]; sum += v[i]; }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += nb; }
      uint32_t run = incl - sum;
#pragma unroll
      for (int i = 0; i < 4; ++i) { cnt[lane * 4 + i] = run; run += v[i]; }
      if (lane == 31) cnt[kPerThread * kWarps] = incl;
    }
    __syncthreads();
    if (!kEmit) {
      if (threadIdx.x == 0) tile_counts[tile] = cnt[kPerThread * kWarps];
    } else {
      const uint32_t excl = tile_excl[tile];
      const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
      for (int c = 0; c < kPerThread; ++c) {
        if ((ball[c] >> lane) & 1u) {
          const uint32_t rp = excl + cnt[c * kWarps + warp] + __popc(ball[c] & lt);
          if (rp < limit) out[rp] = (int64_t)(local0 + c * kThreads + threadIdx.x);
        }
      }
    }
  }
}

// single-CTA exclusive scan of tile counts (n up to a few 10k); total -> excl[n]
__global__ void __launch_bounds__(1024) scan_counts_kernel(const uint32_t* __restrict__ counts, uint32_t* __restrict__ excl, uint32_t n) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n ? counts[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += nb; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = wsum[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (lane >= (uint32_t)o) wi += nb; }
      wsum[lane] = wi - w;
    }
    __syncthreads();
    const uint32_t c = carry;
    if (i < n) excl[i] = c + wsum[warp] + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + wsum[warp] + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) excl[n] = carry;
}

// ---------------------------------------------------------------------------
// QSGD: one CTA per bucket
// ---------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(256) qsgd_encode_kernel(const float* __restrict__ v, int64_t K, int bucket, int q,
                                                          uint32_t seed, OutT* __restrict__ lvl, float* __restrict__ norms) {
  __shared__ float red[8];
  const int64_t b0 = (int64_t)blockIdx.x * bucket;
  const int n = (int)min((int64_t)bucket, K - b0);
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const float x = v[b0 + i]; ss += x * x; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
  const float norm = sqrtf(tot);
  if (threadIdx.x == 0) norms[blockIdx.x] = norm;
  const float scale = norm > 0.f ? (float)q / norm : 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = v[b0 + i];
    const float lf = scale * fabsf(x);
    const float prev = floorf(lf);
    const float u = (float)((double)policy_hash((uint32_t)(b0 + i), seed) / 4294967296.0);
    float l = prev + ((u < (lf - prev)) ? 1.f : 0.f);
    l = x > 0.f ? l : (x < 0.f ? -l : 0.f);
    lvl[b0 + i] = (OutT)l;
  }
}

template <typename InT>
__global__ void qsgd_decode_kernel(const InT* __restrict__ lvl, const float* __restrict__ norms, int64_t K, int bucket,
                                   int q, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < K; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = norms[i / bucket] / (float)q * (float)lvl[i];
}

// ---------------------------------------------------------------------------
// bit packing: thread per output word / per value
// ---------------------------------------------------------------------------
__global__ void pack_bits_kernel(const int64_t* __restrict__ vals, int64_t n, int bits, uint32_t* __restrict__ out, int64_t n_words) {
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bit0 = w * 32;
    int64_t i = bit0 / bits;
    uint32_t word = 0;
    for (; i < n && i * bits < bit0 + 32; ++i) {
      const uint64_t v = (uint64_t)vals[i] & ((bits >= 64) ? ~0ull : ((1ull << bits) - 1ull));
      const int64_t s = i * bits - bit0;               // position of value bit 0 relative to the word
      if (s >= 0) word |= (uint32_t)(v << s);
      else word |= (uint32_t)(v >> (-s));
    }
    out[w] = word;
  }
}

__global__ void unpack_bits_kernel(const uint32_t* __restrict__ in, int64_t n_words, int64_t n, int bits, int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bit0 = i * bits;
    const int64_t w = bit0 >> 5;
    const int s = (int)(bit0 & 31);
    uint64_t lo = in[w];
    uint64_t hi = (w + 1 < n_words) ? in[w + 1] : 0u;
    const uint64_t win = lo | (hi << 32);
    out[i] = (int64_t)((win >> s) & ((1ull << bits) - 1ull));
  }
}

// ---------------------------------------------------------------------------
// Gram-polynomial least squares: one CTA per segment.
// p_0=1, p_1=1-2x/N, (k+1)(N-k)p_{k+1} = (2k+1)(N-2x)p_k - k(N+k+1)p_{k-1}
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) polyfit_fit_kernel(const float* __restrict__ y, const int* __restrict__ seg_off,
                                                          const int* __restrict__ seg_len, int degree,
                                                          float* __restrict__ coeffs) {
  __shared__ float red[2][kMaxDeg + 1][8];
  const int s = blockIdx.x;
  const int n = seg_len[s], off = seg_off[s];
  const int deg_eff = min(degree, n - 1);
  float num[kMaxDeg + 1], den[kMaxDeg + 1];
#pragma unroll
  for (int k = 0; k <= kMaxDeg; ++k) { num[k] = 0.f; den[k] = 0.f; }
  const float N = (float)(n - 1);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float p[kMaxDeg + 1];
    gram_eval<kMaxDeg + 1>((float)i, N, deg_eff, p);
    const float yi = y[off + i];
#pragma unroll
    for (int k = 0; k <= kMaxDeg; ++k) { num[k] += p[k] * yi; den[k] += p[k] * p[k]; }
  }
#pragma unroll
  for (int k = 0; k <= kMaxDeg; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      num[k] += __shfl_xor_sync(0xFFFFFFFFu, num[k], o);
      den[k] += __shfl_xor_sync(0xFFFFFFFFu, den[k], o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][k][threadIdx.x >> 5] = num[k]; red[1][k][threadIdx.x >> 5] = den[k]; }
  }
  __syncthreads();
  if ((int)threadIdx.x <= degree) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][threadIdx.x][w]; b += red[1][threadIdx.x][w]; }
    coeffs[s * (degree + 1) + threadIdx.x] = (n > 0 && (int)threadIdx.x <= max(deg_eff, 0) && b > 0.f) ? a / b : 0.f;
  }
}

__global__ void polyfit_eval_kernel(const float* __restrict__ coeffs, const int* __restrict__ seg_off,
                                    const int* __restrict__ seg_len, int n_seg, int degree, int64_t total,
                                    float* __restrict__ out) {
  __shared__ int s_off[32], s_len[32];
  if ((int)threadIdx.x < n_seg) { s_off[threadIdx.x] = seg_off[threadIdx.x]; s_len[threadIdx.x] = seg_len[threadIdx.x]; }
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int s = 0;
    for (int j = 0; j < n_seg; ++j) if (s_len[j] > 0 && i >= s_off[j]) s = j;
    const int n = s_len[s];
    const int deg_eff = min(degree, n - 1);
    float p[kMaxDeg + 1];
    gram_eval<kMaxDeg + 1>((float)(i - s_off[s]), (float)(n - 1), deg_eff, p);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k <= kMaxDeg; ++k) if (k <= degree) acc += coeffs[s * (degree + 1) + k] * p[k];
    out[i] = acc;
  }
}

// ---------------------------------------------------------------------------
// delta + bp128: one warp per block of 128 sorted indices (4 per lane).
// block wire: [width][4*width words]; value j of the block occupies bits
// [j*width, (j+1)*width) of the block's little-endian bit stream.
// ---------------------------------------------------------------------------
__global__ void bp128_width_kernel(const int64_t* __restrict__ idx, int64_t n, uint32_t* __restrict__ widths) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_blocks = (n + 127) / 128;
  if (warp >= n_blocks) return;
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t i = warp * 128 + j * 32 + lane;
    if (i < n) {
      const uint32_t d = (uint32_t)(idx[i] - (i > 0 ? idx[i - 1] : 0));
      m |= d;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xFFFFFFFFu, m, o);
  if (lane == 0) widths[warp] = 32 - __clz(m);
}

__global__ void bp128_pack_kernel(const int64_t* __restrict__ idx, int64_t n, const uint32_t* __restrict__ widths,
                                  const int64_t* __restrict__ word_off, uint32_t* __restrict__ out) {
  __shared__ uint32_t sv[8][128];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_blocks = (n + 127) / 128;
  if (warp >= n_blocks) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t i = warp * 128 + j * 32 + lane;
    sv[wl][j * 32 + lane] = i < n ? (uint32_t)(idx[i] - (i > 0 ? idx[i - 1] : 0)) : 0u;
  }
  __syncwarp();
  const uint32_t width = widths[warp];
  uint32_t* dst = out + word_off[warp];
  if (lane == 0) dst[0] = width;
  for (uint32_t w = lane; w < 4 * width; w += 32) {
    const uint32_t bit0 = w * 32;
    uint32_t j = bit0 / width, word = 0;
    for (; j < 128 && j * width < bit0 + 32; ++j) {
      const uint64_t v = sv[wl][j];
      const int s = (int)(j * width) - (int)bit0;
      word |= s >= 0 ? (uint32_t)(v << s) : (uint32_t)(v >> (-s));
    }
    dst[1 + w] = word;
  }
}

// decode: one warp per block; deltas -> inclusive scan inside the block + block base (second kernel adds bases)
__global__ void bp128_unpack_kernel(const uint32_t* __restrict__ in, const int64_t* __restrict__ word_off, int64_t n,
                                    int64_t* __restrict__ deltas) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_blocks = (n + 127) / 128;
  if (warp >= n_blocks) return;
  const uint32_t* src = in + word_off[warp];
  const uint32_t width = src[0];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int v = j * 32 + lane;
    const int64_t i = warp * 128 + v;
    if (i < n) {
      uint64_t val = 0;
      if (width) {
        const uint32_t bit0 = v * width;
        const uint32_t w = bit0 >> 5, s = bit0 & 31u;
        const uint64_t lo = src[1 + w];
        const uint64_t hi = (w + 1 < 4 * width) ? src[2 + w] : 0u;
        val = ((lo | (hi << 32)) >> s) & ((1ull << width) - 1ull);
      }
      deltas[i] = (int64_t)val;
    }
  }
}

__global__ void bp128_header_scan_kernel(const uint32_t* __restrict__ in, int64_t n_blocks, int64_t* __restrict__ word_off) {
  // sequential walk over block headers (n_blocks is K/128: small); single thread.
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t off = 0;
    for (int64_t b = 0; b < n_blocks; ++b) { word_off[b] = off; off += 1 + 4 * (int64_t)in[off]; }
    word_off[n_blocks] = off;
  }
}

// ---------------------------------------------------------------------------
// bit-level run-length index coding (reference RunLength, pytorch/deepreduce.py:805-846, which loops over all d
// bits in Python).  Runs are derived from the sorted indices: a run of ones starts where idx[i] != idx[i-1]+1.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) rle_count_kernel(const int64_t* __restrict__ idx, int64_t n, uint32_t* __restrict__ counts) {
  const int64_t i = blockIdx.x * 1024ll + threadIdx.x;
  const bool start = i < n && (i == 0 || idx[i] != idx[i - 1] + 1);
  const int c = __syncthreads_count(start);
  if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024) rle_mark_kernel(const int64_t* __restrict__ idx, int64_t n, const uint32_t* __restrict__ excl,
                                                        int64_t* __restrict__ start_pos, int64_t* __restrict__ end_pos) {
  __shared__ uint32_t wsum[32];
  const int64_t i = blockIdx.x * 1024ll + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const bool start = i < n && (i == 0 || idx[i] != idx[i - 1] + 1);
  const bool end = i < n && (i == n - 1 || idx[i + 1] != idx[i] + 1);
  const uint32_t ball = __ballot_sync(0xFFFFFFFFu, start);
  if (lane == 0) wsum[warp] = __popc(ball);
  __syncthreads();
  if (warp == 0) {
    uint32_t w = wsum[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (lane >= (uint32_t)o) wi += nb; }
    wsum[lane] = wi - w;
  }
  __syncthreads();
  // inclusive count of starts up to and including element i
  const uint32_t incl = excl[blockIdx.x] + wsum[warp] + __popc(ball & ((2u << lane) - 1u));
  if (start) start_pos[incl - 1] = i;
  if (end) end_pos[incl - 1] = i;               // an end closes the run opened by the latest start
}

__global__ void rle_runs_kernel(const int64_t* __restrict__ idx, const int64_t* __restrict__ start_pos,
                                const int64_t* __restrict__ end_pos, int64_t n_runs, int64_t d, int64_t* __restrict__ runs) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_runs; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = start_pos[r], e = end_pos[r];
    const int64_t prev_end = r ? idx[end_pos[r - 1]] : -1;
    runs[2 * r] = idx[s] - prev_end - 1;
    runs[2 * r + 1] = e - s + 1;
    if (r == n_runs - 1) { const int64_t tail = d - 1 - idx[e]; if (tail > 0) runs[2 * n_runs] = tail; }
  }
}

// decode: thread j finds the run holding the j-th index (binary search over the cumulative ones counts)
__global__ void rle_expand_kernel(const int64_t* __restrict__ ones_excl, const int64_t* __restrict__ run_start, int64_t n_runs,
                                  int64_t total, int64_t* __restrict__ out) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = n_runs;                 // largest r with ones_excl[r] <= j
    while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (ones_excl[mid] <= j) lo = mid; else hi = mid; }
    out[j] = run_start[lo] + (j - ones_excl[lo]);
  }
}

inline int grid_for(int64_t n, int threads, int cap = 148 * 8) {
  int64_t g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

}  // namespace

// ---------------------------------------------------------------------------
// launchers (C ABI used by binding.cpp)
// ---------------------------------------------------------------------------
void launch_bloom_insert(const int64_t* idx, int64_t n, uint32_t* filter, uint32_t n_hash, uint32_t m_bits,
                         uint32_t seed, cudaStream_t st) {
  if (n == 0) return;
  count_launch();
  bloom_insert_kernel<<<grid_for(n, 256), 256, 0, st>>>(idx, n, filter, n_hash, m_bits, seed);
}

void launch_bloom_count(const uint32_t* filter, uint32_t d, uint32_t n_hash, uint32_t m_bits, uint32_t seed,
                        uint32_t* tile_counts, uint32_t* tile_excl, cudaStream_t st) {
  const uint32_t n_tiles = (d + kTile - 1) / kTile;
  count_launch(2);
  bloom_query_kernel<false><<<min(n_tiles, 148u * 4u), kThreads, 0, st>>>(filter, d, n_hash, m_bits, seed, tile_counts,
                                                                         nullptr, nullptr, 0);
  scan_counts_kernel<<<1, 1024, 0, st>>>(tile_counts, tile_excl, n_tiles);
}

void launch_bloom_emit(const uint32_t* filter, uint32_t d, uint32_t n_hash, uint32_t m_bits, uint32_t seed,
                       const uint32_t* tile_excl, int64_t* out, uint32_t limit, cudaStream_t st) {
  const uint32_t n_tiles = (d + kTile - 1) / kTile;
  count_launch();
  bloom_query_kernel<true><<<min(n_tiles, 148u * 4u), kThreads, 0, st>>>(filter, d, n_hash, m_bits, seed, nullptr,
                                                                        tile_excl, out, limit);
}

void launch_qsgd_encode(const float* v, int64_t K, int bucket, int q, uint32_t seed, void* lvl, bool i16, float* norms,
                        cudaStream_t st) {
  if (K == 0) return;
  const int nb = (int)((K + bucket - 1) / bucket);
  count_launch();
  if (i16) qsgd_encode_kernel<int16_t><<<nb, 256, 0, st>>>(v, K, bucket, q, seed, (int16_t*)lvl, norms);
  else qsgd_encode_kernel<int8_t><<<nb, 256, 0, st>>>(v, K, bucket, q, seed, (int8_t*)lvl, norms);
}

void launch_qsgd_decode(const void* lvl, bool i16, const float* norms, int64_t K, int bucket, int q, float* out,
                        cudaStream_t st) {
  if (K == 0) return;
  count_launch();
  if (i16) qsgd_decode_kernel<int16_t><<<grid_for(K, 256), 256, 0, st>>>((const int16_t*)lvl, norms, K, bucket, q, out);
  else qsgd_decode_kernel<int8_t><<<grid_for(K, 256), 256, 0, st>>>((const int8_t*)lvl, norms, K, bucket, q, out);
}

void launch_pack_bits(const int64_t* vals, int64_t n, int bits, uint32_t* out, int64_t n_words, cudaStream_t st) {
  if (n_words == 0) return;
  count_launch();
  pack_bits_kernel<<<grid_for(n_words, 256), 256, 0, st>>>(vals, n, bits, out, n_words);
}

void launch_unpack_bits(const uint32_t* in, int64_t n_words, int64_t n, int bits, int64_t* out, cudaStream_t st) {
  if (n == 0) return;
  count_launch();
  unpack_bits_kernel<<<grid_for(n, 256), 256, 0, st>>>(in, n_words, n, bits, out);
}

// ---------------------------------------------------------------------------
// Double-exponential fit ("Fit-DExp"): y_k ~ a e^{p x_k} + b e^{q x_k} on x_k = (k+1)/K, y = |values| ascending.
// Reference: tensorflow/deepreduce.py:66-144 (Jacquelin-style integral-equation regression: cumulative trapezoids
// S = int y, SS = int S; 4x4 normal system of  y ~ A SS + B S + C x + D;  p,q = (B +- sqrt(B^2 + 4A))/2;  then a
// 2x2 least squares for a, b) — a chain of TF GPU ops (cumsum, matmul, linalg.solve) upstream; here ONE CTA does the
// two scans, the 14 moment sums, both small solves and the second pass.  fp64 throughout like the reference (the sums
// span ~1e6 terms of very different magnitude); codecs/dexp.py::double_exponential_fit is the torch oracle.
// ---------------------------------------------------------------------------
constexpr int kDexpThreads = 1024;

__device__ __forceinline__ double warp_incl_scan_f64(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double n = __shfl_up_sync(0xFFFFFFFFu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// block-wide inclusive scan of one value per thread (blockDim = 1024); `tot` gets the block total
__device__ double block_incl_scan_f64(double v, double* warp_sums, double& tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_incl_scan_f64(v, lane);
  __syncthreads();
  if (lane == 31) warp_sums[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double w = warp_sums[lane];
    w = warp_incl_scan_f64(w, lane);
    warp_sums[lane] = w;
  }
  __syncthreads();
  const double base = warp ? warp_sums[warp - 1] : 0.0;
  tot = warp_sums[31];
  return v + base;
}

__device__ double block_sum_f64(double v, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kDexpThreads / 32; ++w) t += scratch[w];
  return t;
}

__global__ void __launch_bounds__(kDexpThreads) dexp_fit_kernel(const float* __restrict__ y, int64_t K, double* __restrict__ out) {
  __shared__ double ws[32];
  __shared__ double sums[14];
  __shared__ double pq[2];
  const int tid = threadIdx.x;
  if (K <= 0) { if (tid < 4) out[tid] = 0.0; return; }
  const double dx = 1.0 / (double)K;
  double acc[14];
#pragma unroll
  for (int i = 0; i < 14; ++i) acc[i] = 0.0;
  double carry_S = 0.0, carry_SS = 0.0;     // S, SS at the last element of the previous chunk
  for (int64_t c0 = 0; c0 < K; c0 += kDexpThreads) {
    const int64_t k = c0 + tid;
    const bool on = k < K;
    const double yk = on ? (double)y[k] : 0.0;
    const double ym = (on && k > 0) ? (double)y[k - 1] : 0.0;
    const double inc = (on && k > 0) ? 0.5 * (yk + ym) * dx : 0.0;
    double tot;
    const double S = carry_S + block_incl_scan_f64(inc, ws, tot);
    const double S_prev = S - inc;                                  // S_{k-1}
    const double inc2 = (on && k > 0) ? 0.5 * (S + S_prev) * dx : 0.0;
    double tot2;
    const double SS = carry_SS + block_incl_scan_f64(inc2, ws, tot2);
    carry_S += tot; carry_SS += tot2;
    if (on) {
      const double x = (double)(k + 1) * dx;
      const double col[4] = {SS, S, x, 1.0};
      int q = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) acc[q++] += col[i] * col[j];     // 10 entries of the Gram matrix
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[10 + i] += col[i] * yk;         // right-hand side
    }
  }
  for (int i = 0; i < 14; ++i) {
    const double t = block_sum_f64(acc[i], ws);
    if (tid == 0) sums[i] = t;
  }
  __syncthreads();
  if (tid == 0) {
    double G[4][5];
    int q = 0;
    for (int i = 0; i < 4; ++i) for (int j = i; j < 4; ++j) { G[i][j] = sums[q]; G[j][i] = sums[q]; ++q; }
    for (int i = 0; i < 4; ++i) { G[i][i] += 1e-18; G[i][4] = sums[10 + i]; }
    for (int c = 0; c < 4; ++c) {                                     // Gaussian elimination, partial pivoting
      int piv = c;
      for (int r = c + 1; r < 4; ++r) if (fabs(G[r][c]) > fabs(G[piv][c])) piv = r;
      if (piv != c) for (int j = 0; j < 5; ++j) { const double t = G[c][j]; G[c][j] = G[piv][j]; G[piv][j] = t; }
      const double d = G[c][c];
      if (d != 0.0) for (int r = c + 1; r < 4; ++r) { const double f = G[r][c] / d; for (int j = c; j < 5; ++j) G[r][j] -= f * G[c][j]; }
    }
    double sol[4];
    for (int i = 3; i >= 0; --i) {
      double t = G[i][4];
      for (int j = i + 1; j < 4; ++j) t -= G[i][j] * sol[j];
      sol[i] = G[i][i] != 0.0 ? t / G[i][i] : 0.0;
    }
    const double A = sol[0], B = sol[1];
    const double disc = fmax(B * B + 4.0 * A, 0.0);
    pq[0] = 0.5 * (B + sqrt(disc));
    pq[1] = 0.5 * (B - sqrt(disc));
  }
  __syncthreads();
  const double p = pq[0], qq = pq[1];
  double m[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int64_t k = tid; k < K; k += kDexpThreads) {
    const double x = (double)(k + 1) * dx, yk = (double)y[k];
    const double bk = exp(p * x), ek = exp(qq * x);
    m[0] += bk * bk; m[1] += bk * ek; m[2] += ek * ek; m[3] += bk * yk; m[4] += ek * yk;
  }
  for (int i = 0; i < 5; ++i) {
    const double t = block_sum_f64(m[i], ws);
    if (tid == 0) sums[i] = t;
  }
  __syncthreads();
  if (tid == 0) {
    const double det = sums[0] * sums[2] - sums[1] * sums[1];
    double a, b;
    if (fabs(det) < 1e-300) { a = sums[0] != 0.0 ? sums[3] / sums[0] : 0.0; b = 0.0; }
    else { a = (sums[3] * sums[2] - sums[4] * sums[1]) / det; b = (sums[0] * sums[4] - sums[1] * sums[3]) / det; }
    out[0] = a; out[1] = b; out[2] = p; out[3] = qq;
  }
}

void launch_dexp_fit(const float* y, int64_t K, double* out, cudaStream_t st) {
  count_launch();
  dexp_fit_kernel<<<1, kDexpThreads, 0, st>>>(y, K, out);
}

// ---------------------------------------------------------------------------
// P2 / conflict sets (paper Alg. 1; reference tensorflow/policies.hpp:43-146 — a host C++ routine upstream, and up to
// round 1 a device -> host -> device bounce here).  The draw is sequential BY DEFINITION (the r-th pick's random
// number and every set's "untouched since my last visit" test depend on all earlier picks), so the kernel is one
// warp that walks the conflict sets in (size, bit position) order:
//   * the chosen-flags of the positives live in shared memory (one bit per positive, indexed by its rank);
//   * 32 sets are fetched at a time (offset, size, last-visit count, first 4 member ranks per lane) and then visited
//     one by one with warp shuffles — no global-memory latency on the sequential path for sets of <= 4 members
//     (the vast majority: a set is the list of positives hashing to one filter bit);
//   * a set erases its chosen members implicitly (alive = not chosen), `last` keeps the alive count of the previous
//     visit, exactly the `cs.size() == before` test of the reference.
// Bit-exact with ops/csrc/cpu/native_cpu.cpp::conflict_sets_impl (tests/test_gpu_engine.py).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32) conflict_sets_pick_kernel(const uint32_t* __restrict__ set_off, const uint32_t* __restrict__ members,
                                                                 uint32_t* __restrict__ last, uint32_t n_sets, uint32_t n_pos, uint32_t K,
                                                                 uint32_t pseed, uint32_t* __restrict__ chosen_out) {
  extern __shared__ uint32_t chosen[];                     // ceil(n_pos / 32) words
  const uint32_t lane = threadIdx.x;
  const uint32_t n_words = (n_pos + 31u) >> 5;
  for (uint32_t i = lane; i < n_words; i += 32u) chosen[i] = 0u;
  __syncwarp();
  auto is_chosen = [&](uint32_t r) { return (chosen[r >> 5] >> (r & 31u)) & 1u; };
  uint32_t left = min(K, n_pos), draw = 0;
  while (left > 0) {
    bool picked = false;
    for (uint32_t base = 0; base < n_sets && left > 0; base += 32u) {
      const uint32_t i = base + lane;
      const bool valid = i < n_sets;
      const uint32_t off = valid ? set_off[i] : 0u;
      const uint32_t sz = valid ? set_off[i + 1] - off : 0u;
      uint32_t lc = valid ? last[i] : 0u;
      uint32_t m[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) m[j] = ((uint32_t)j < sz) ? members[off + j] : 0u;
      const uint32_t n_here = min(32u, n_sets - base);
      for (uint32_t sidx = 0; sidx < n_here && left > 0; ++sidx) {        // warp-uniform, sequential by definition
        const uint32_t ssz = __shfl_sync(0xFFFFFFFFu, sz, sidx), soff = __shfl_sync(0xFFFFFFFFu, off, sidx);
        const uint32_t slc = __shfl_sync(0xFFFFFFFFu, lc, sidx);
        uint32_t a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = __shfl_sync(0xFFFFFFFFu, m[j], sidx);
        uint32_t cnt = 0;
        if (ssz <= 4u) {
#pragma unroll
          for (int j = 0; j < 4; ++j) cnt += ((uint32_t)j < ssz && !is_chosen(a[j])) ? 1u : 0u;
        } else {
          for (uint32_t t = lane; t < ssz; t += 32u) cnt += is_chosen(members[soff + t]) ? 0u : 1u;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        }
        uint32_t nl = cnt;
        if (cnt == slc && cnt > 0u) {                                       // untouched since my last visit: draw one member
          uint32_t r = policy_hash(draw, pseed) % cnt;
          ++draw;
          uint32_t pick = 0;
          if (ssz <= 4u) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if ((uint32_t)j < ssz && !is_chosen(a[j])) { if (r == 0u) pick = a[j]; --r; }
            }
          } else {
            for (uint32_t t = 0; t < ssz; ++t) {                            // every lane walks the (rare) long set identically
              const uint32_t x = members[soff + t];
              if (!is_chosen(x)) { if (r == 0u) { pick = x; break; } --r; }
            }
          }
          __syncwarp();
          if (lane == 0) chosen[pick >> 5] |= 1u << (pick & 31u);
          __syncwarp();
          --left;
          picked = true;
          nl = cnt - 1u;
        }
        if (lane == sidx) lc = nl;
      }
      if (valid) last[i] = lc;
    }
    if (!picked && left > 0) {                                               // termination fallback: leftmost unchosen positives
      for (uint32_t r = 0; r < n_pos && left > 0; ++r) {
        if (!is_chosen(r)) {
          __syncwarp();
          if (lane == 0) chosen[r >> 5] |= 1u << (r & 31u);
          __syncwarp();
          --left;
        }
      }
    }
  }
  __syncwarp();
  for (uint32_t i = lane; i < n_words; i += 32u) chosen_out[i] = chosen[i];
}

cudaError_t launch_conflict_sets_pick(const uint32_t* set_off, const uint32_t* members, uint32_t* last, uint32_t n_sets, uint32_t n_pos,
                                      uint32_t K, uint32_t pseed, uint32_t* chosen_out, cudaStream_t st) {
  const size_t smem = (size_t)((n_pos + 31u) >> 5) * 4u;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;                       // > 1.6 M positives: the caller falls back to the host routine
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(conflict_sets_pick_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  count_launch();
  conflict_sets_pick_kernel<<<1, 32, smem, st>>>(set_off, members, last, n_sets, n_pos, K, pseed, chosen_out);
  return cudaGetLastError();
}

void launch_polyfit_fit(const float* y, const int* seg_off, const int* seg_len, int n_seg, int degree, float* coeffs,
                        cudaStream_t st) {
  count_launch();
  polyfit_fit_kernel<<<n_seg, 256, 0, st>>>(y, seg_off, seg_len, degree, coeffs);
}

void launch_polyfit_eval(const float* coeffs, const int* seg_off, const int* seg_len, int n_seg, int degree,
                         int64_t total, float* out, cudaStream_t st) {
  if (total == 0) return;
  count_launch();
  polyfit_eval_kernel<<<grid_for(total, 256), 256, 0, st>>>(coeffs, seg_off, seg_len, n_seg, degree, total, out);
}

void launch_bp128_widths(const int64_t* idx, int64_t n, uint32_t* widths, cudaStream_t st) {
  if (n == 0) return;
  const int64_t n_blocks = (n + 127) / 128;
  count_launch();
  bp128_width_kernel<<<(int)((n_blocks * 32 + 255) / 256), 256, 0, st>>>(idx, n, widths);
}

void launch_bp128_pack(const int64_t* idx, int64_t n, const uint32_t* widths, const int64_t* word_off, uint32_t* out,
                       cudaStream_t st) {
  if (n == 0) return;
  const int64_t n_blocks = (n + 127) / 128;
  count_launch();
  bp128_pack_kernel<<<(int)((n_blocks * 32 + 255) / 256), 256, 0, st>>>(idx, n, widths, word_off, out);
}

void launch_rle_count(const int64_t* idx, int64_t n, uint32_t* counts, uint32_t* excl, cudaStream_t st) {
  const uint32_t nb = (uint32_t)((n + 1023) / 1024);
  count_launch(2);
  rle_count_kernel<<<nb, 1024, 0, st>>>(idx, n, counts);
  scan_counts_kernel<<<1, 1024, 0, st>>>(counts, excl, nb);
}

void launch_rle_runs(const int64_t* idx, int64_t n, const uint32_t* excl, int64_t* start_pos, int64_t* end_pos,
                     int64_t n_runs, int64_t d, int64_t* runs, cudaStream_t st) {
  const uint32_t nb = (uint32_t)((n + 1023) / 1024);
  count_launch(2);
  rle_mark_kernel<<<nb, 1024, 0, st>>>(idx, n, excl, start_pos, end_pos);
  rle_runs_kernel<<<grid_for(n_runs, 256), 256, 0, st>>>(idx, start_pos, end_pos, n_runs, d, runs);
}

void launch_rle_expand(const int64_t* ones_excl, const int64_t* run_start, int64_t n_runs, int64_t total, int64_t* out,
                       cudaStream_t st) {
  if (total == 0) return;
  count_launch();
  rle_expand_kernel<<<grid_for(total, 256), 256, 0, st>>>(ones_excl, run_start, n_runs, total, out);
}

void launch_bp128_unpack(const uint32_t* in, int64_t n, int64_t* word_off, int64_t* deltas, cudaStream_t st) {
  if (n == 0) return;
  const int64_t n_blocks = (n + 127) / 128;
  count_launch(2);
  bp128_header_scan_kernel<<<1, 32, 0, st>>>(in, n_blocks, word_off);
  bp128_unpack_kernel<<<(int)((n_blocks * 32 + 255) / 256), 256, 0, st>>>(in, word_off, n, deltas);
}

}  // namespace dr
