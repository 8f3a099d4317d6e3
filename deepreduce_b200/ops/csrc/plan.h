// Bucket plan + engine parameter block shared by host (binding.cpp) and device
// (engine.cu).  The plan is built in Python (parallel/plan.py) and uploaded as
// int32 tensors; field order here is normative for that builder.
//
// Parity notes (reference = hangxu0304/DeepReduce):
//  * a TensorDesc carries what the reference recomputes per call from `params`: K = max(1, int(d * ratio))
//    (GRACE top-k), bloom m / hash count (pytorch/deepreduce.py:495-500,511-512), the 1000-element bypass
//    (:68,84,114), polyfit degree (:385), QSGD quantum_num / bucket 512 (:857-858);
//  * a slot replaces the per-tensor wire tuples `(vals, packed bit array)` (:529), `(coefficients, idxs)` (:413-414)
//    and `(vals', idxs', mapping)` (:267) — all tensors of a bucket in one buffer with static offsets, so the
//    size all_gather + pad-to-max of the GRACE communicator (tensors_size_are_same=False, :59,108) disappears;
//  * DynHeader.n_sel/cutoff is policy `leftmost` / `p0` (:479-492) expressed so the receiver needs no sort.
#pragma once
#include <stdint.h>

namespace dr {

constexpr int kMaxWorld = 16;
constexpr int kTile = 4096;           // elements per tile (wire-visible: per-tile prefix table)
constexpr uint32_t kDefaultSeed = 0x9747B28Cu;

enum TensorMode : uint32_t {
  kModeRaw = 0,     // plain (value, index) pairs — small-tensor bypass / plain top-k
  kModeBloom = 1,   // bloom-filter index codec, fp32 values
  kModeRle = 2,     // lossless tile-local run coding: u16 count per 4096-element tile (off_prefix) + the cumulative
                    // zero-run offset of every selected element inside its tile, 12 bits each, LSB-first (off_idx)
};

enum Policy : int { kPolicyLeftmost = 0, kPolicyRandom = 1, kPolicyP0 = 2 };

// 32 x uint32 per tensor (kDescWords)
struct TensorDesc {
  uint32_t elem_off;     // offset into the flat grad/residual buffers (elements, multiple of 4)
  uint32_t numel;        // d_i
  uint32_t k;            // K_i = max(1, int(d_i * ratio)), <= numel
  uint32_t tile_begin;   // first global tile id
  uint32_t n_tiles;      // ceil(numel / kTile)
  uint32_t mode;         // TensorMode
  uint32_t m_bits;       // bloom bits (multiple of 32)
  uint32_t n_hash;       // bloom hash count
  uint32_t off_vals;     // payload word offsets (within a slot)
  uint32_t off_filter;
  uint32_t off_prefix;   // per-tile exclusive prefix of selected counts (n_tiles words)
  uint32_t off_idx;      // raw mode: indices
  uint32_t val_cap;      // capacity of the value region (K, or K+slack for p0)
  uint32_t salt;         // tensor id (policy seeds)
  uint32_t n_filter_words;
  uint32_t off_hint;     // 0 = none; else 4 words per tile: bit g set <=> 32-element group g holds a selected element
  // ---- value codec ('both': bloom index + polynomial fit of the values) ----
  uint32_t vmode;        // 0 = fp32 values on the wire, 1 = piece-wise Gram-polynomial fit + rank map,
                         // 2 = bucketed QSGD (int8 levels, or int16 when rank_u32 is set: quantum_num >= 128)
  uint32_t off_coef;     // [kMaxSeg * (deg+1)] float coefficients, then {num_pos, n}
  uint32_t off_rankmap;  // rank of the p-th shipped value in the descending sort (u16 if val_cap <= 65536 else u32)
  uint32_t off_selidx;   // scratch (not shipped): element index of the p-th shipped value
  uint32_t off_sorted;   // scratch (not shipped): values in descending order
  uint32_t poly_degree;
  uint32_t rank_u32;     // 1: rank map entries are 32-bit
  uint32_t poly_off;     // offset of this tensor's values in the engine's per-value scratch arrays
  uint32_t poly_ord;     // ordinal among the vmode==1 tensors (selects its bin table)
  uint32_t fixed_thr;    // != 0: 'threshold' sparsifier — select key >= fixed_thr (31-bit |x| pattern), no radix select, variable K
  uint32_t reserved[6];
};
static_assert(sizeof(TensorDesc) == 128, "TensorDesc must be 32 words");
constexpr int kDescWords = 32;
constexpr int kRankBins = 8192;        // 'both': counting-sort bins = sign + 8 exponent + 4 mantissa bits
constexpr int kMaxSeg = 22;            // codecs/polyfit.py MAX_SEGMENTS
constexpr int kMaxDeg = 7;

// payload slot layout (uint32 words):
//   [0..8)                      : magic, epoch, n_tensors, payload_words, rank, 0,0,0
//   [8 .. 8+4*n_tensors)        : DynHeader per tensor
//   then per-tensor regions at the TensorDesc offsets
constexpr uint32_t kSlotHeaderWords = 8;
constexpr uint32_t kDynWords = 4;
constexpr uint32_t kMagic = 0xD33B2000u;
struct DynHeader {
  uint32_t n_sel;      // number of values actually shipped
  uint32_t cutoff;     // largest selected index (leftmost); 0xFFFFFFFF = no cut
  uint32_t thr_bits;   // |g| threshold bits chosen by the select
  uint32_t n_pos;      // filter positives in the universe (diagnostics / FP count)
};

// per-tensor select state (persists across steps; thr history drives the
// pass-1 lower bound): 8 words
struct SelState {
  uint32_t bin1, krem1, bin2, krem2;
  uint32_t thr;         // selection threshold as a key lower bound: (T22 << 9), T22 = 22-bit prefix
  uint32_t n_ge;        // diagnostics: keys in the threshold bin
  uint32_t done_epoch;  // == epoch when the accumulate phase already finished the select (one-tile tensors)
  uint32_t prev_thr;    // thr of the previous step (0 = none)
};

constexpr int kHistBins = 2048;
constexpr int kNumHist = 3;
// hist arrays: [3][n_tensors][kHistBins]  (pass1, pass1-fallback, pass2);  hist_total: [3][n_tensors] merge tickets

// arena layout per rank (uint32 words): [flags: 64][stage-2 flags: 64][slots: 2 * world * slot_words]
//                                       [stage-2 slots: 2 * world * s2_words]
constexpr uint32_t kArenaFlagWords = 64;
constexpr uint32_t kArenaHdrWords = 128;

// Selection rule (normative, mirrored by parallel/engine.py::select_topk_oracle):
//   key  = |x| bit pattern (31 bits);  T22 = (key of the K-th largest |x|) >> 9
//   selected  <=>  (key >> 9) >= max(T22, 1)
// i.e. the threshold is resolved to 22 bits (8 exponent + 14 mantissa): at least K elements are
// selected, plus the few that share the threshold's 22-bit prefix; exact zeros are never selected.
enum Phase : int {
  kPhAccum = 0,      // r = beta*r + gamma*g ; dense grad <- 0 ; zero slot ; candidate lists (keys >= history bound) ; hist digit 1
  kPhFallback = 1,   // (only if some bound was unsafe) digit 1 redone without the bound, candidate lists rebuilt in full
  kPhHist2 = 2,      // digit 2 of the candidate keys in the threshold bin
  kPhInsert = 3,     // selected candidates -> bloom filter + occupancy hint (bloom) / positive masks (raw, rle)
  kPhQuery = 4,      // hinted 32-element groups of the universe vs my filter -> positive masks + per-tile counts
  kPhEmit = 5,       // ordered compaction from the masks + value gather + residual zeroing (+ dense scatter when W == 1)
  // 'both' (bloom index + polynomial value fit): exact descending rank of every shipped value by a
  // counting sort on 13 key bits + an all-pairs count inside each bin, then the fit
  kPhRankHist = 6,   // bin populations
  kPhRankScan = 7,   // per-tensor exclusive prefix over the bins
  kPhRankScatter = 8,// group the values by bin
  kPhRankExact = 9,  // exact rank inside the bin -> rank map + sorted values + num_pos
  kPhFit = 10,       // per-segment Gram-polynomial least squares on the sorted values
  kPhFix = 11,       // residual <- value - fitted value (error feedback sees the fit error)
  kPhPush = 12,      // copy the finished slot into every peer's arena (P2P stores over NVLink); the CTA that finishes last
                     // (ticket) releases the flags — no grid barrier between the copy and the signal
  kPhSignal = 13,    // acquire peers' flags
  kPhExpand = 14,    // 'both': evaluate every rank's fitted curve once (dense), decode then only gathers
  kPhDecode = 15,    // membership test on every rank's filter, rank->value, sum, scale, dense write
                     // (sharded mode, W>1: only this rank's 1/W slice of the tiles, for all W senders)
  // sharded decode (W > 1): the decoded slice is exchanged as an exact (index, value) list — decode work per rank
  // no longer grows with W; NVLink carries the extra ~2 MB/rank
  kPhCompact = 16,   // compact the non-zeros of my decoded slice (same CTA that decoded the tile) and store them straight
                     // into every peer's stage-2 slot; last CTA (ticket) writes the count and releases the second flag set
  kPhPush2 = 17,     // (folded into kPhCompact; kept so phase numbers stay stable)
  kPhSignal2 = 18,   // acquire peers' second flags
  kPhScatter = 19,   // write every peer's slice list into the dense gradient
  kPhEnd = 20
};

// phase classes with their own tile -> CTA partition (the phases have different cost profiles, and the SMs of the two dies
// run the memory-heavy phases at different speeds): accumulate / fallback / hist2, insert, query, emit
enum Part : int { kPartAccum = 0, kPartInsert = 1, kPartQuery = 2, kPartEmit = 3, kNumParts = 4 };

// per-tile table (uint4): {tensor id, element offset of the tile in the flat buffers, valid count, offset inside tensor}
struct TileInfo { uint32_t tensor, base, n, local0; };

struct EngineParams {
  const TensorDesc* tensors;
  const TileInfo* tiles;         // [n_tiles]
  const uint32_t* cost_prefix;   // [n_tiles + 1] cumulative cost of the tiles (host plan: elements + per-tensor overheads); the
                                 // streaming phases cut the tile sequence into equal-COST ranges (nullptr: equal counts)
  uint32_t n_tensors;
  uint32_t n_tiles;
  uint32_t slot_words;           // words reserved per slot
  uint32_t payload_words;        // words actually used (pushed)
  float* grad;                   // in: local dense grad; out: aggregated dense grad
  float* resid;                  // residual accumulator (persists across steps)
  uint32_t* hist;                // [3][n_tensors][kHistBins]
  uint32_t* hist_total;          // [3][n_tensors] merge tickets (#tiles whose histogram was merged)
  SelState* sel;                 // [n_tensors]
  uint32_t* tile_count;          // [n_tiles] selected/positive count of every tile (insert / query phase)
  uint32_t* pos_mask;            // [n_tiles * 128] positives of this rank, one bit per element, one word per 32-element group
  uint32_t* dec_mask;            // [n_tiles * 128] decode scratch: positives of the sender being decoded
  uint2* cand;                   // [n_tiles * 4096] candidates {key = |x| pattern >= the history bound, in-tile element offset},
                                 // 256 slots per (tile, warp) at a fixed place
  uint32_t* cand_cnt;            // [n_tiles * 16] candidates per (tile, warp)
  uint32_t* barrier;             // [0] grid barrier counter, [1] push ticket, [2] stage-2 ticket (zeroed by the host per
                                 // launch); [8] number of tensors whose history bound hid the threshold (device-managed)
  uint32_t* status;              // [8] error / watchdog words (device-local)
  uint32_t* arena[kMaxWorld];    // peer-mapped arena base of every rank (arena[rank] is local)
  int rank;
  int world;
  uint32_t epoch;                // 1-based step counter; slot parity = epoch & 1
  float beta, gamma, scale;
  uint32_t seed;
  int policy;
  int use_history;               // pass-1 lower bound from prev_thr
  int phase_begin, phase_end;
  uint32_t spin_limit;           // watchdog for flag / look-back / barrier spins
  uint32_t filter_smem_words;    // capacity of the dynamic-SMEM buffer (filter staging / TMA tile ring)
  int use_tma;                   // streaming phases fetch tiles with cp.async.bulk into an SMEM ring
  uint32_t hist_shift;           // history bound = prev_thr - (1 << hist_shift): 23 -> x0.5, 22 -> ~x0.7
  const uint32_t* poly_tensors;  // ids of the tensors with vmode == 1 (largest K first)
  uint32_t n_poly;
  const uint32_t* poly_tasks;    // per-value tasks: {tensor id, first value of a 512-value chunk}
  uint32_t n_poly_tasks;
  uint32_t* poly_bins;           // [n_poly][2][kRankBins]: bin counts | bin starts/cursors (zeroed in decode)
  float* bucket_val;             // [sum K] values grouped by bin
  uint32_t* bucket_pos;          // [sum K] original position p of the grouped values
  float* expand_buf;             // [world][sum K] fitted curves of every rank
  uint32_t poly_total;           // sum K over the vmode==1 tensors
  unsigned long long* debug_times;  // optional [kPhEnd + 1][grid][2] globaltimer ns at phase entry / exit of every CTA (nullptr: off);
                                    // row kPhEnd: {%smid of the CTA, 0}
  const uint32_t* cuts;          // optional per-phase-class tile partitions [kNumParts][cuts_grid + 1] (first tile of every CTA,
  uint32_t cuts_grid;            // host-computed: per-phase cost weights x per-CTA speeds); used when gridDim.x == cuts_grid
  int deterministic;             // 1: decode adds the senders of a tile in rank order by one warp (bit-reproducible sums);
                                 // 0 (default): every (sender, tile) pair is an independent work item that adds with RED.ADD.F32 —
                                 // identical on all ranks (owner computes), order-dependent in the last ulp only where >= 3
                                 // senders hit the same element
  uint32_t peer_timeout_ms;      // peer-flag waits give up after this long (status 2, output poisoned with NaN, CTA exits)
  int fault;                     // fault injection (tests): 1 = this rank never releases its stage-1 flags
  uint32_t* mc_arena;            // NVLS multicast mapping of the symmetric arena (nullptr: per-peer P2P stores)
  int has_rle;                   // some tensor uses kModeRle (its bit stream is OR-ed, so it is zeroed every step)
  int shard;                     // 1: sharded decode + stage-2 exchange (when world > 1)
  uint32_t s2_words;             // words per stage-2 slot: [count, epoch, 0, 0][idx x cap][val x cap]
  uint32_t s2_cap;               // entries per stage-2 slot
};

}  // namespace dr
