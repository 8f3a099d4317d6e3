// Launcher declarations (implemented in ops.cu / engine.cu / p2p.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.h"

namespace dr {

// engine.cu
int engine_max_grid(int blocks_per_sm, int dyn_smem_bytes);
cudaError_t engine_launch(const EngineParams& P, int grid, int blocks_per_sm, int dyn_smem_bytes, cudaStream_t stream);
void count_launch(int n);
long long launch_count();

// ops.cu
void launch_bloom_insert(const int64_t* idx, int64_t n, uint32_t* filter, uint32_t n_hash, uint32_t m_bits,
                         uint32_t seed, cudaStream_t st);
void launch_bloom_count(const uint32_t* filter, uint32_t d, uint32_t n_hash, uint32_t m_bits, uint32_t seed,
                        uint32_t* tile_counts, uint32_t* tile_excl, cudaStream_t st);
void launch_bloom_emit(const uint32_t* filter, uint32_t d, uint32_t n_hash, uint32_t m_bits, uint32_t seed,
                       const uint32_t* tile_excl, int64_t* out, uint32_t limit, cudaStream_t st);
void launch_qsgd_encode(const float* v, int64_t K, int bucket, int q, uint32_t seed, void* lvl, bool i16, float* norms,
                        cudaStream_t st);
void launch_qsgd_decode(const void* lvl, bool i16, const float* norms, int64_t K, int bucket, int q, float* out,
                        cudaStream_t st);
void launch_pack_bits(const int64_t* vals, int64_t n, int bits, uint32_t* out, int64_t n_words, cudaStream_t st);
void launch_unpack_bits(const uint32_t* in, int64_t n_words, int64_t n, int bits, int64_t* out, cudaStream_t st);
void launch_polyfit_fit(const float* y, const int* seg_off, const int* seg_len, int n_seg, int degree, float* coeffs,
                        cudaStream_t st);
void launch_polyfit_eval(const float* coeffs, const int* seg_off, const int* seg_len, int n_seg, int degree,
                         int64_t total, float* out, cudaStream_t st);
cudaError_t launch_conflict_sets_pick(const uint32_t* set_off, const uint32_t* members, uint32_t* last, uint32_t n_sets, uint32_t n_pos,
                                      uint32_t K, uint32_t pseed, uint32_t* chosen_out, cudaStream_t st);
void launch_dexp_fit(const float* y, int64_t K, double* out_abpq, cudaStream_t st);
void launch_bp128_widths(const int64_t* idx, int64_t n, uint32_t* widths, cudaStream_t st);
void launch_bp128_pack(const int64_t* idx, int64_t n, const uint32_t* widths, const int64_t* word_off, uint32_t* out,
                       cudaStream_t st);
void launch_bp128_unpack(const uint32_t* in, int64_t n, int64_t* word_off, int64_t* deltas, cudaStream_t st);
void launch_rle_count(const int64_t* idx, int64_t n, uint32_t* counts, uint32_t* excl, cudaStream_t st);
void launch_rle_runs(const int64_t* idx, int64_t n, const uint32_t* excl, int64_t* start_pos, int64_t* end_pos,
                     int64_t n_runs, int64_t d, int64_t* runs, cudaStream_t st);
void launch_rle_expand(const int64_t* ones_excl, const int64_t* run_start, int64_t n_runs, int64_t total, int64_t* out,
                       cudaStream_t st);

// p2p.cu — symmetric arena over CUDA IPC
struct ArenaHandle { unsigned char bytes[64]; };
void* arena_alloc(size_t bytes);                       // cudaMalloc + zero
void arena_free(void* p);
ArenaHandle arena_export(void* p);
void* arena_import(const ArenaHandle& h);              // cudaIpcOpenMemHandle
void arena_close(void* p);
int arena_enable_peer_access(const int* devices, int n);   // peers = the listed devices only; returns #enabled
void launch_u8_to_nhwc_norm(const uint8_t* in, void* out_bf16, int64_t n_pix, const float* mean, const float* inv_std,
                            cudaStream_t st);

}  // namespace dr
