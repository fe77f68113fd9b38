// irr_kernels.h — internal launch functions shared between the translation units of
// libirr_b200.so.  The C ABI in include/irr_b200.h (irr_cabi.cu) validates arguments and calls
// these; none of them allocates or synchronises.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/irr_b200.h"

namespace irr {

// irr_cabi.cu: event pair armed by irr_profile_next_topk (thread-local, one-shot)
void profile_mark_start(cudaStream_t st);
void profile_mark_stop(cudaStream_t st);

// row_norms.cu
irr_status row_inv_norms(const void* x, int64_t N, int32_t D, irr_dtype dt, float eps, float* out,
                         cudaStream_t st);
irr_status pair_cosine(const void* x1, int64_t x1_rows, const void* x2, int64_t N, int32_t D,
                       irr_dtype dt, float eps, float* out, cudaStream_t st);

// topk_merge.cu
// partial lists [S][Q][k] (score without the query norm, local int32 index) -> final [Q,k]
irr_status merge_partials(const float* part_val, const int32_t* part_idx, int32_t S, int64_t Q,
                          int32_t k, const void* q, int32_t D, irr_dtype dt, float eps,
                          int64_t idx_offset, float* out_val, int64_t* out_idx, cudaStream_t st);
irr_status merge_candidates(const float* cand_val, int64_t val_rank_stride, const int64_t* cand_idx,
                            int64_t idx_rank_stride, int32_t G, int64_t Q, int32_t k,
                            float* out_val, int64_t* out_idx, cudaStream_t st);
// every slot (-inf, -1): the result of searching an empty shard
irr_status fill_padding(float* out_val, int64_t* out_idx, int64_t n, cudaStream_t st);
irr_status topk_hits(const int64_t* idx, int64_t Q, int32_t k, const int64_t* q_label,
                     const int64_t* g_label, int64_t N, int64_t instance_offset, int64_t* out_hits,
                     cudaStream_t st);

// topk_exchange.cu (peer-memory exchange fused with the cross-shard merge)
size_t topk_exchange_bytes(int32_t G, int64_t Q, int32_t k);
irr_status topk_exchange_merge(const float* local_val, const int64_t* local_idx,
                               void* const* peer_bufs, int32_t G, int32_t rank, int64_t Q, int32_t k,
                               size_t buf_bytes, int32_t mode, float* out_val, int64_t* out_idx,
                               cudaStream_t st);

// cosine_topk_bf16.cu (tcgen05 / TMA)
size_t bf16_topk_workspace_bytes(int64_t Q, int64_t N, int32_t D, int32_t k);
irr_status bf16_cosine_topk(const void* q, const void* g, const float* g_inv_norm, int64_t Q,
                            int64_t N, int32_t D, int32_t k, float eps, int64_t idx_offset,
                            float* out_val, int64_t* out_idx, void* ws, size_t ws_bytes,
                            cudaStream_t st, irr_dtype dt = IRR_BF16 /* or IRR_F16 */);
irr_status bf16_cosine_scores(const void* q, const void* g, int64_t Q, int64_t N, int32_t D,
                              float eps, float* out_scores, void* ws, size_t ws_bytes,
                              cudaStream_t st);

irr_status bf16_scores_block(const void* q, const void* g, const float* g_inv_norm,
                             const float* q_inv_norm, int64_t Q, int64_t N, int32_t D, float eps,
                             float* out_scores, cudaStream_t st, irr_dtype dt = IRR_BF16);

// topk_select.cu (large k: select from a dense score block; class de-duplication)
irr_status topk_select(const float* scores, int64_t Q, int64_t N, int32_t k, int64_t idx_offset,
                       float* out_val, int64_t* out_idx, cudaStream_t st);
irr_status merge_candidates_large(const float* cand_val, int64_t val_rank_stride,
                                  const int64_t* cand_idx, int64_t idx_rank_stride, int32_t G,
                                  int64_t Q, int32_t k, float* out_val, int64_t* out_idx,
                                  cudaStream_t st);
irr_status merge_candidates_large_exchange(const uint8_t* buf, size_t state_off, size_t data_off,
                                           size_t half_bytes, size_t slot_bytes, size_t idx_off,
                                           int32_t G, int64_t Q, int32_t k, float* out_val,
                                           int64_t* out_idx, cudaStream_t st);
irr_status class_dedup(const float* val, const int64_t* idx, int64_t Q, int32_t k,
                       const int64_t* g_label, int64_t N, int32_t n_distinct, const int64_t* q_label,
                       int64_t* out_label, int64_t* out_idx, float* out_val, int64_t* out_hits,
                       cudaStream_t st);

// cosine_topk_f32.cu (fp32 FFMA, exactness path)
size_t f32_topk_workspace_bytes(int64_t Q, int64_t N, int32_t k);
irr_status f32_cosine_topk(const void* q, const void* g, const float* g_inv_norm, int64_t Q,
                           int64_t N, int32_t D, int32_t k, float eps, int64_t idx_offset,
                           float* out_val, int64_t* out_idx, void* ws, size_t ws_bytes,
                           cudaStream_t st);

irr_status f32_cosine_scores(const void* q, const void* g, const float* g_inv_norm,
                             const float* q_inv_norm, int64_t Q, int64_t N, int32_t D,
                             float* out_scores, cudaStream_t st);

// producer_consumer.cu
irr_status avgpool_fwd(const void* fm, int in_dt, int64_t rows, int32_t hw, void* out, int out_dt,
                       cudaStream_t st);
irr_status avgpool_bwd(const void* grad_out, int go_dt, int64_t rows, int32_t hw, void* grad_fm,
                       int gf_dt, cudaStream_t st);
size_t ce_pair_workspace_bytes(int64_t B);
irr_status ce_pair_fwd(const void* a, const void* b, const int64_t* target, int64_t B, int32_t C,
                       int dt, int64_t ignore_index, float* out_loss, void* ws, size_t ws_bytes,
                       cudaStream_t st);
irr_status ce_pair_bwd(const void* a, const void* b, const int64_t* target, int64_t B, int32_t C,
                       int dt, int64_t ignore_index, const float* grad_out, const void* ws, void* da,
                       void* db, cudaStream_t st);

// triplet_loss.cu
struct LossArgs {
  const void *q, *p, *n;       // n == nullptr: pair form (a = q, b = p)
  const float* label;          // pair form only
  int64_t label_count;
  int64_t B;
  int32_t D;
  irr_dtype dt;
  int32_t kind;                // pair form: IRR_LOSS_*
  float margin_cos, margin_con;
  int32_t reduce_mean;
  float pair_eps;
  float* losses;               // device fp32[4] (triplet) / [1] (pair)
  float* pair_cos;             // optional [2*B]
  float* row_stats;            // optional [B*IRR_ROW_STATS]
  void *dq, *dp, *dn;          // optional
  float grad_scale[4];
};
size_t loss_workspace_bytes(int64_t B, int32_t D, irr_dtype dt);
irr_status loss_fwd_bwd(const LossArgs& a, void* ws, size_t ws_bytes, cudaStream_t st);
irr_status loss_bwd(const LossArgs& a, const float* grad_out, cudaStream_t st);

}  // namespace irr
