// irr_cabi.cu — the extern "C" boundary declared in include/irr_b200.h: argument validation and
// dispatch to the kernels.  Nothing here allocates, synchronises or falls back to the host.
#include <string.h>

#include "irr_common.cuh"
#include "irr_kernels.h"

namespace irr {

int num_sms() {
  static int cached = []() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      return 148;
    }
    return n;
  }();
  return cached;
}

int device_cc() {
  static int cached = []() {
    int dev = 0, mj = 0, mn = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&mn, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    return mj * 10 + mn;
  }();
  return cached;
}

namespace {
thread_local cudaEvent_t g_prof_start = nullptr;
thread_local cudaEvent_t g_prof_stop = nullptr;
}  // namespace

void profile_mark_start(cudaStream_t st) {
  if (g_prof_start) cudaEventRecord(g_prof_start, st);
}
void profile_mark_stop(cudaStream_t st) {
  if (g_prof_stop) cudaEventRecord(g_prof_stop, st);
  g_prof_start = g_prof_stop = nullptr;
}

namespace {

bool any_float(irr_dtype dt) { return dt == IRR_F32 || dt == IRR_BF16 || dt == IRR_F16; }

// allow_f16: the search / similarity entry points take fp16 rows (autocast embeddings) on the
// tensor-core path; the loss kernels do not
irr_status check_rows(const void* p, int32_t D, irr_dtype dt, bool allow_f16 = false) {
  if (dt != IRR_F32 && dt != IRR_BF16 && !(allow_f16 && dt == IRR_F16))
    return IRR_ERR_UNSUPPORTED_DTYPE;
  if (D <= 0) return IRR_ERR_INVALID_ARG;
  if ((D * dtype_bytes(dt)) % 16 != 0) return IRR_ERR_ALIGNMENT;
  if (p && !aligned16(p)) return IRR_ERR_ALIGNMENT;
  return IRR_OK;
}

// ---- large-k path (IRR_MAX_K_FUSED < k <= IRR_MAX_K): dense score blocks + row-wise selection ----
constexpr size_t kScoreBlockBudget = size_t(1) << 30;   // bytes of scores materialised at a time

int64_t large_k_block_rows(int64_t Q, int64_t N) {
  int64_t qb = static_cast<int64_t>(kScoreBlockBudget / (static_cast<size_t>(N > 0 ? N : 1) * 4));
  if (qb < 1) qb = 1;
  if (qb >= 128) qb = qb / 128 * 128;
  return qb < Q ? qb : Q;
}

size_t large_k_workspace_bytes(int64_t Q, int64_t N) {
  return align_up(static_cast<size_t>(N) * 4, 256) + align_up(static_cast<size_t>(Q) * 4, 256) +
         align_up(static_cast<size_t>(large_k_block_rows(Q, N)) * N * 4, 256) + 256;
}

irr_status large_k_cosine_topk(const void* q, const void* g, const float* g_inv_norm, int64_t Q,
                               int64_t N, int32_t D, int32_t k, irr_dtype dt, float eps,
                               int64_t idx_offset, float* out_val, int64_t* out_idx, void* ws,
                               size_t ws_bytes, cudaStream_t st) {
  if (N > 0x7fffff00ll) return IRR_ERR_INVALID_ARG;
  if (ws_bytes < large_k_workspace_bytes(Q, N)) return IRR_ERR_WORKSPACE_TOO_SMALL;
  uint8_t* w = static_cast<uint8_t*>(ws);
  float* gin_ws = reinterpret_cast<float*>(w);
  w += align_up(static_cast<size_t>(N) * 4, 256);
  float* qin = reinterpret_cast<float*>(w);
  w += align_up(static_cast<size_t>(Q) * 4, 256);
  float* scores = reinterpret_cast<float*>(w);
  const float* gin = g_inv_norm;
  irr_status s;
  if (!gin) {
    s = row_inv_norms(g, N, D, dt, eps, gin_ws, st);
    if (s != IRR_OK) return s;
    gin = gin_ws;
  }
  s = row_inv_norms(q, Q, D, dt, eps, qin, st);
  if (s != IRR_OK) return s;
  const int64_t qb = large_k_block_rows(Q, N);
  const size_t row_bytes = static_cast<size_t>(D) * dtype_bytes(dt);
  for (int64_t b0 = 0; b0 < Q; b0 += qb) {
    const int64_t rows = b0 + qb <= Q ? qb : Q - b0;
    const void* qblk = static_cast<const uint8_t*>(q) + b0 * row_bytes;
    if (dt != IRR_F32)
      s = bf16_scores_block(qblk, g, gin, qin + b0, rows, N, D, eps, scores, st, dt);
    else
      s = f32_cosine_scores(qblk, g, gin, qin + b0, rows, N, D, scores, st);
    if (s != IRR_OK) return s;
    s = topk_select(scores, rows, N, k, idx_offset, out_val + b0 * k, out_idx + b0 * k, st);
    if (s != IRR_OK) return s;
  }
  return IRR_OK;
}

// test aid (irr_debug_occupy_sms): CTAs that hold shared memory and spin for a fixed time
__global__ void occupy_kernel(unsigned long long ns) {
  extern __shared__ uint8_t occupy_smem[];
  if (threadIdx.x == 0) occupy_smem[0] = 1;
  const uint64_t t0 = global_timer_ns();
  while (global_timer_ns() - t0 < ns) __nanosleep(2000);
}

}  // namespace
}  // namespace irr

using namespace irr;

extern "C" {

int32_t irr_version(void) { return 100; }

void irr_profile_next_topk(void* ev_start, void* ev_stop) {
  g_prof_start = static_cast<cudaEvent_t>(ev_start);
  g_prof_stop = static_cast<cudaEvent_t>(ev_stop);
}

irr_status irr_debug_occupy_sms(int32_t ctas, int32_t smem_bytes, int64_t nanoseconds,
                                irr_stream_t stream) {
  if (ctas < 1 || smem_bytes < 0 || smem_bytes > 227 * 1024 || nanoseconds < 0 ||
      nanoseconds > 2000000000ll)
    return IRR_ERR_INVALID_ARG;
  IRR_CUDA_TRY(cudaFuncSetAttribute(occupy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    smem_bytes));
  occupy_kernel<<<ctas, 32, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<unsigned long long>(nanoseconds));
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

const char* irr_status_string(irr_status s) {
  switch (s) {
    case IRR_OK: return "ok";
    case IRR_ERR_INVALID_ARG: return "invalid argument";
    case IRR_ERR_UNSUPPORTED_DTYPE: return "unsupported dtype";
    case IRR_ERR_ALIGNMENT: return "pointer or row length violates the 16-byte alignment contract";
    case IRR_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case IRR_ERR_K_TOO_LARGE: return "k exceeds IRR_MAX_K";
    case IRR_ERR_UNSUPPORTED_DEVICE: return "device is not sm_100 or the driver lacks tensor maps";
    case IRR_ERR_ROW_TOO_LONG: return "row too long for the shared-memory staged loss kernel";
    default: return s > 0 ? cudaGetErrorString(static_cast<cudaError_t>(s)) : "unknown status";
  }
}

size_t irr_cosine_topk_workspace_bytes(int64_t Q, int64_t N, int32_t D, int32_t k, irr_dtype dt) {
  (void)D;
  if (Q < 0 || N < 0 || k < 1) return 0;
  if (k > IRR_MAX_K_FUSED) return large_k_workspace_bytes(Q, N);
  return dt != IRR_F32 ? bf16_topk_workspace_bytes(Q, N, D, k) : f32_topk_workspace_bytes(Q, N, k);
}

irr_status irr_cosine_topk(const void* q, const void* g, const float* g_inv_norm, int64_t Q,
                           int64_t N, int32_t D, int32_t k, irr_dtype dt, float eps,
                           int64_t idx_offset, float* out_val, int64_t* out_idx, void* workspace,
                           size_t workspace_bytes, irr_stream_t stream) {
  if (Q < 0 || N < 0 || k < 1 || !out_val || !out_idx) return IRR_ERR_INVALID_ARG;
  if (k > IRR_MAX_K) return IRR_ERR_K_TOO_LARGE;
  if (Q == 0) return IRR_OK;
  // an empty gallery shard (total rows < ranks) is not an error: all k slots are padding
  if (N == 0)
    return check_rows(q, D, dt, true) != IRR_OK
               ? check_rows(q, D, dt, true)
               : fill_padding(out_val, out_idx, Q * k, reinterpret_cast<cudaStream_t>(stream));
  if (!q || !g || !workspace) return IRR_ERR_INVALID_ARG;
  irr_status s = check_rows(q, D, dt, true);
  if (s != IRR_OK) return s;
  s = check_rows(g, D, dt, true);
  if (s != IRR_OK) return s;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (k > IRR_MAX_K_FUSED)
    return large_k_cosine_topk(q, g, g_inv_norm, Q, N, D, k, dt, eps, idx_offset, out_val, out_idx,
                               workspace, workspace_bytes, st);
  if (dt != IRR_F32)
    return bf16_cosine_topk(q, g, g_inv_norm, Q, N, D, k, eps, idx_offset, out_val, out_idx,
                            workspace, workspace_bytes, st, dt);
  return f32_cosine_topk(q, g, g_inv_norm, Q, N, D, k, eps, idx_offset, out_val, out_idx, workspace,
                         workspace_bytes, st);
}

irr_status irr_cosine_scores_bf16(const void* q, const void* g, int64_t Q, int64_t N, int32_t D,
                                  float eps, float* out_scores, void* workspace,
                                  size_t workspace_bytes, irr_stream_t stream) {
  if (Q <= 0 || N <= 0 || !q || !g || !out_scores || !workspace) return IRR_ERR_INVALID_ARG;
  irr_status s = check_rows(q, D, IRR_BF16);
  if (s != IRR_OK) return s;
  s = check_rows(g, D, IRR_BF16);
  if (s != IRR_OK) return s;
  return bf16_cosine_scores(q, g, Q, N, D, eps, out_scores, workspace, workspace_bytes,
                            reinterpret_cast<cudaStream_t>(stream));
}

irr_status irr_row_inv_norms(const void* x, int64_t N, int32_t D, irr_dtype dt, float eps,
                             float* out, irr_stream_t stream) {
  if (N < 0 || (N > 0 && (!x || !out))) return IRR_ERR_INVALID_ARG;
  irr_status s = check_rows(x, D, dt, true);
  if (s != IRR_OK) return s;
  return row_inv_norms(x, N, D, dt, eps, out, reinterpret_cast<cudaStream_t>(stream));
}

irr_status irr_topk_merge(const float* cand_val, const int64_t* cand_idx, int32_t G, int64_t Q,
                          int32_t k, float* out_val, int64_t* out_idx, irr_stream_t stream) {
  if (G < 1 || Q < 0 || k < 1 || !cand_val || !cand_idx || !out_val || !out_idx)
    return IRR_ERR_INVALID_ARG;
  if (k > IRR_MAX_K) return IRR_ERR_K_TOO_LARGE;
  return merge_candidates(cand_val, Q * k, cand_idx, Q * k, G, Q, k, out_val, out_idx,
                          reinterpret_cast<cudaStream_t>(stream));
}

irr_status irr_topk_merge_strided(const float* cand_val, int64_t val_rank_stride,
                                  const int64_t* cand_idx, int64_t idx_rank_stride, int32_t G,
                                  int64_t Q, int32_t k, float* out_val, int64_t* out_idx,
                                  irr_stream_t stream) {
  if (G < 1 || Q < 0 || k < 1 || !cand_val || !cand_idx || !out_val || !out_idx)
    return IRR_ERR_INVALID_ARG;
  if (val_rank_stride < Q * k || idx_rank_stride < Q * k) return IRR_ERR_INVALID_ARG;
  if (k > IRR_MAX_K) return IRR_ERR_K_TOO_LARGE;
  return merge_candidates(cand_val, val_rank_stride, cand_idx, idx_rank_stride, G, Q, k, out_val,
                          out_idx, reinterpret_cast<cudaStream_t>(stream));
}

size_t irr_topk_exchange_bytes(int32_t G, int64_t Q, int32_t k) {
  if (G < 1 || G > IRR_MAX_PEERS || Q < 0 || k < 1 || k > IRR_MAX_K) return 0;
  return topk_exchange_bytes(G, Q, k);
}

irr_status irr_topk_exchange_merge(const float* local_val, const int64_t* local_idx,
                                   void* const* peer_bufs, int32_t G, int32_t rank, int64_t Q,
                                   int32_t k, size_t buf_bytes, int32_t mode, float* out_val,
                                   int64_t* out_idx, irr_stream_t stream) {
  if (G < 1 || G > IRR_MAX_PEERS || rank < 0 || rank >= G || Q < 0 || k < 1 || !peer_bufs)
    return IRR_ERR_INVALID_ARG;
  if (k > IRR_MAX_K) return IRR_ERR_K_TOO_LARGE;
  if (mode != IRR_XCHG_FUSED && mode != IRR_XCHG_PUSH && mode != IRR_XCHG_MERGE)
    return IRR_ERR_INVALID_ARG;
  if ((mode == IRR_XCHG_FUSED || mode == IRR_XCHG_PUSH) && Q > 0 && (!local_val || !local_idx))
    return IRR_ERR_INVALID_ARG;
  if (mode != IRR_XCHG_PUSH && Q > 0 && (!out_val || !out_idx)) return IRR_ERR_INVALID_ARG;
  return topk_exchange_merge(local_val, local_idx, peer_bufs, G, rank, Q, k, buf_bytes, mode,
                             out_val, out_idx, reinterpret_cast<cudaStream_t>(stream));
}

irr_status irr_peer_export(const void* dev_ptr, uint8_t handle_out[64], uint64_t* offset_out) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
  if (!dev_ptr || !handle_out || !offset_out) return IRR_ERR_INVALID_ARG;
  typedef int (*RangeFn)(unsigned long long*, size_t*, unsigned long long);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  IRR_CUDA_TRY(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return IRR_ERR_UNSUPPORTED_DEVICE;
  unsigned long long base = 0;
  size_t size = 0;
  const unsigned long long p = reinterpret_cast<unsigned long long>(dev_ptr);
  if (reinterpret_cast<RangeFn>(fn)(&base, &size, p) != 0 || base == 0 || p < base)
    return IRR_ERR_INVALID_ARG;
  cudaIpcMemHandle_t h;
  IRR_CUDA_TRY(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  memcpy(handle_out, &h, 64);
  *offset_out = p - base;
  return IRR_OK;
}

irr_status irr_peer_import(const uint8_t handle[64], void** mapped_base) {
  if (!handle || !mapped_base) return IRR_ERR_INVALID_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  IRR_CUDA_TRY(cudaIpcOpenMemHandle(mapped_base, h, cudaIpcMemLazyEnablePeerAccess));
  return IRR_OK;
}

irr_status irr_peer_close(void* mapped_base) {
  if (!mapped_base) return IRR_ERR_INVALID_ARG;
  IRR_CUDA_TRY(cudaIpcCloseMemHandle(mapped_base));
  return IRR_OK;
}

size_t irr_cosine_topk_sharded_workspace_bytes(int64_t Q, int64_t N_local, int32_t D, int32_t k,
                                               irr_dtype dt) {
  const size_t inner = irr_cosine_topk_workspace_bytes(Q, N_local, D, k, dt);
  if (inner == 0) return 0;
  const size_t n = static_cast<size_t>(Q) * k;
  return align_up(inner, 256) + align_up(n * 4, 256) + align_up(n * 8, 256);
}

irr_status irr_cosine_topk_sharded(const void* q, const void* g_local, const float* g_inv_norm,
                                   int64_t Q, int64_t N_local, int32_t D, int32_t k, irr_dtype dt,
                                   float eps, int64_t idx_offset, void* const* peer_bufs, int32_t G,
                                   int32_t rank, size_t buf_bytes, float* out_val, int64_t* out_idx,
                                   void* workspace, size_t workspace_bytes, irr_stream_t stream) {
  if (Q < 0 || N_local < 0 || k < 1 || !workspace) return IRR_ERR_INVALID_ARG;
  const size_t inner = irr_cosine_topk_workspace_bytes(Q, N_local, D, k, dt);
  if (inner == 0 || workspace_bytes < irr_cosine_topk_sharded_workspace_bytes(Q, N_local, D, k, dt))
    return IRR_ERR_WORKSPACE_TOO_SMALL;
  uint8_t* w = static_cast<uint8_t*>(workspace) + align_up(inner, 256);
  const size_t n = static_cast<size_t>(Q) * k;
  float* lv = reinterpret_cast<float*>(w);
  int64_t* li = reinterpret_cast<int64_t*>(w + align_up(n * 4, 256));
  irr_status s = irr_cosine_topk(q, g_local, g_inv_norm, Q, N_local, D, k, dt, eps, idx_offset, lv,
                                 li, workspace, inner, stream);
  if (s != IRR_OK) return s;
  return irr_topk_exchange_merge(lv, li, peer_bufs, G, rank, Q, k, buf_bytes, IRR_XCHG_FUSED,
                                 out_val, out_idx, stream);
}

irr_status irr_topk_hits(const int64_t* idx, int64_t Q, int32_t k, const int64_t* q_label,
                         const int64_t* g_label, int64_t N, int64_t instance_offset,
                         int64_t* out_hits, irr_stream_t stream) {
  if (Q < 0 || k < 1 || !out_hits || (Q > 0 && !idx)) return IRR_ERR_INVALID_ARG;
  if ((q_label == nullptr) != (g_label == nullptr)) return IRR_ERR_INVALID_ARG;
  return topk_hits(idx, Q, k, q_label, g_label, N, instance_offset, out_hits,
                   reinterpret_cast<cudaStream_t>(stream));
}

irr_status irr_topk_class_dedup(const float* val, const int64_t* idx, int64_t Q, int32_t k,
                                const int64_t* g_label, int64_t N, int32_t n_distinct,
                                const int64_t* q_label, int64_t* out_label, int64_t* out_idx,
                                float* out_val, int64_t* out_hits, irr_stream_t stream) {
  if (Q < 0 || k < 1 || N < 0 || !g_label || !out_label || !out_idx || !out_val ||
      (Q > 0 && (!val || !idx)))
    return IRR_ERR_INVALID_ARG;
  if (out_hits && !q_label) return IRR_ERR_INVALID_ARG;
  return class_dedup(val, idx, Q, k, g_label, N, n_distinct, q_label, out_label, out_idx, out_val,
                     out_hits, reinterpret_cast<cudaStream_t>(stream));
}

irr_status irr_pair_cosine(const void* x1, int64_t x1_rows, const void* x2, int64_t N, int32_t D,
                           irr_dtype dt, float eps, float* out, irr_stream_t stream) {
  if (N < 0 || (x1_rows != N && x1_rows != 1)) return IRR_ERR_INVALID_ARG;
  if (N > 0 && (!x1 || !x2 || !out)) return IRR_ERR_INVALID_ARG;
  irr_status s = check_rows(x1, D, dt, true);
  if (s != IRR_OK) return s;
  s = check_rows(x2, D, dt, true);
  if (s != IRR_OK) return s;
  return pair_cosine(x1, x1_rows, x2, N, D, dt, eps, out, reinterpret_cast<cudaStream_t>(stream));
}

size_t irr_triplet_loss_workspace_bytes(int64_t B, int32_t D, irr_dtype dt) {
  return loss_workspace_bytes(B, D, dt);
}
size_t irr_pair_loss_workspace_bytes(int64_t B, int32_t D, irr_dtype dt) {
  return loss_workspace_bytes(B, D, dt);
}

irr_status irr_triplet_loss_fwd_bwd(const void* q, const void* p, const void* n, int64_t B,
                                    int32_t D, irr_dtype dt, float margin_cos, float margin_con,
                                    int32_t reduce_mean, float pair_eps, float* losses,
                                    float* pair_cos, float* row_stats, void* dq, void* dp, void* dn,
                                    const float grad_scale[4], void* workspace,
                                    size_t workspace_bytes, irr_stream_t stream) {
  if (B <= 0 || !q || !p || !n || !losses || !workspace) return IRR_ERR_INVALID_ARG;
  const int ng = (dq != nullptr) + (dp != nullptr) + (dn != nullptr);
  if (ng != 0 && ng != 3) return IRR_ERR_INVALID_ARG;
  const void* ptrs[6] = {q, p, n, dq, dp, dn};
  for (const void* x : ptrs) {
    irr_status s = check_rows(x, D, dt, true);
    if (s != IRR_OK) return s;
  }
  if (row_stats && !aligned16(row_stats)) return IRR_ERR_ALIGNMENT;
  LossArgs a = {};
  a.q = q; a.p = p; a.n = n;
  a.B = B; a.D = D; a.dt = dt;
  a.margin_cos = margin_cos; a.margin_con = margin_con;
  a.reduce_mean = reduce_mean; a.pair_eps = pair_eps;
  a.losses = losses; a.pair_cos = pair_cos; a.row_stats = row_stats;
  a.dq = dq; a.dp = dp; a.dn = dn;
  for (int j = 0; j < 4; ++j) a.grad_scale[j] = grad_scale ? grad_scale[j] : 1.0f;
  return loss_fwd_bwd(a, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

irr_status irr_triplet_loss_bwd(const void* q, const void* p, const void* n,
                                const float* row_stats, const float* grad_out, int64_t B, int32_t D,
                                irr_dtype dt, float margin_cos, float margin_con,
                                int32_t reduce_mean, void* dq, void* dp, void* dn,
                                irr_stream_t stream) {
  if (B <= 0 || !q || !p || !n || !row_stats || !grad_out || !dq || !dp || !dn)
    return IRR_ERR_INVALID_ARG;
  const void* ptrs[6] = {q, p, n, dq, dp, dn};
  for (const void* x : ptrs) {
    irr_status s = check_rows(x, D, dt, true);
    if (s != IRR_OK) return s;
  }
  if (!aligned16(row_stats)) return IRR_ERR_ALIGNMENT;
  LossArgs a = {};
  a.q = q; a.p = p; a.n = n;
  a.B = B; a.D = D; a.dt = dt;
  a.margin_cos = margin_cos; a.margin_con = margin_con;
  a.reduce_mean = reduce_mean;
  a.row_stats = const_cast<float*>(row_stats);
  a.dq = dq; a.dp = dp; a.dn = dn;
  return loss_bwd(a, grad_out, reinterpret_cast<cudaStream_t>(stream));
}

irr_status irr_pair_loss_fwd_bwd(const void* a_, const void* b_, const float* label,
                                 int64_t label_count, int64_t B, int32_t D, irr_dtype dt,
                                 int32_t kind, float margin, int32_t reduce_mean, float* loss,
                                 float* row_stats, void* da, void* db, float grad_scale,
                                 void* workspace, size_t workspace_bytes, irr_stream_t stream) {
  if (B <= 0 || !a_ || !b_ || !label || !loss || !workspace) return IRR_ERR_INVALID_ARG;
  if (label_count != 1 && label_count != B) return IRR_ERR_INVALID_ARG;
  if (kind != IRR_LOSS_CONTRASTIVE && kind != IRR_LOSS_COSINE_EMBEDDING) return IRR_ERR_INVALID_ARG;
  if ((da != nullptr) != (db != nullptr)) return IRR_ERR_INVALID_ARG;
  const void* ptrs[4] = {a_, b_, da, db};
  for (const void* x : ptrs) {
    irr_status s = check_rows(x, D, dt, true);
    if (s != IRR_OK) return s;
  }
  if (row_stats && !aligned16(row_stats)) return IRR_ERR_ALIGNMENT;
  LossArgs a = {};
  a.q = a_; a.p = b_; a.n = nullptr;
  a.label = label; a.label_count = label_count;
  a.B = B; a.D = D; a.dt = dt; a.kind = kind;
  a.margin_cos = margin; a.margin_con = margin;
  a.reduce_mean = reduce_mean; a.pair_eps = 1e-6f;
  a.losses = loss; a.row_stats = row_stats;
  a.dq = da; a.dp = db;
  a.grad_scale[0] = grad_scale;
  return loss_fwd_bwd(a, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

irr_status irr_pair_loss_bwd(const void* a_, const void* b_, const float* label,
                             int64_t label_count, const float* row_stats, const float* grad_out,
                             int64_t B, int32_t D, irr_dtype dt, int32_t kind, float margin,
                             int32_t reduce_mean, void* da, void* db, irr_stream_t stream) {
  if (B <= 0 || !a_ || !b_ || !label || !row_stats || !grad_out || !da || !db)
    return IRR_ERR_INVALID_ARG;
  if (label_count != 1 && label_count != B) return IRR_ERR_INVALID_ARG;
  if (kind != IRR_LOSS_CONTRASTIVE && kind != IRR_LOSS_COSINE_EMBEDDING) return IRR_ERR_INVALID_ARG;
  const void* ptrs[4] = {a_, b_, da, db};
  for (const void* x : ptrs) {
    irr_status s = check_rows(x, D, dt, true);
    if (s != IRR_OK) return s;
  }
  if (!aligned16(row_stats)) return IRR_ERR_ALIGNMENT;
  LossArgs a = {};
  a.q = a_; a.p = b_; a.n = nullptr;
  a.label = label; a.label_count = label_count;
  a.B = B; a.D = D; a.dt = dt; a.kind = kind;
  a.margin_cos = margin; a.margin_con = margin;
  a.reduce_mean = reduce_mean;
  a.row_stats = const_cast<float*>(row_stats);
  a.dq = da; a.dp = db;
  return loss_bwd(a, grad_out, reinterpret_cast<cudaStream_t>(stream));
}

irr_status irr_avgpool_fwd(const void* fm, irr_dtype in_dt, int64_t rows, int32_t hw, void* out,
                           irr_dtype out_dt, irr_stream_t stream) {
  if (rows < 0 || hw < 1 || (rows > 0 && (!fm || !out))) return IRR_ERR_INVALID_ARG;
  if (!any_float(in_dt) || (out_dt != IRR_F32 && out_dt != IRR_BF16)) return IRR_ERR_UNSUPPORTED_DTYPE;
  return avgpool_fwd(fm, in_dt, rows, hw, out, out_dt, reinterpret_cast<cudaStream_t>(stream));
}

irr_status irr_avgpool_bwd(const void* grad_out, irr_dtype go_dt, int64_t rows, int32_t hw,
                           void* grad_fm, irr_dtype gf_dt, irr_stream_t stream) {
  if (rows < 0 || hw < 1 || (rows > 0 && (!grad_out || !grad_fm))) return IRR_ERR_INVALID_ARG;
  if ((go_dt != IRR_F32 && go_dt != IRR_BF16) || !any_float(gf_dt)) return IRR_ERR_UNSUPPORTED_DTYPE;
  return avgpool_bwd(grad_out, go_dt, rows, hw, grad_fm, gf_dt, reinterpret_cast<cudaStream_t>(stream));
}

size_t irr_ce_pair_workspace_bytes(int64_t B) { return B < 0 ? 0 : ce_pair_workspace_bytes(B); }

irr_status irr_ce_pair_fwd(const void* a, const void* b, const int64_t* target, int64_t B, int32_t C,
                           irr_dtype dt, int64_t ignore_index, float* out_loss, void* workspace,
                           size_t workspace_bytes, irr_stream_t stream) {
  if (B <= 0 || C < 1 || !a || !b || !target || !out_loss || !workspace) return IRR_ERR_INVALID_ARG;
  if (!any_float(dt)) return IRR_ERR_UNSUPPORTED_DTYPE;
  return ce_pair_fwd(a, b, target, B, C, dt, ignore_index, out_loss, workspace, workspace_bytes,
                     reinterpret_cast<cudaStream_t>(stream));
}

irr_status irr_ce_pair_bwd(const void* a, const void* b, const int64_t* target, int64_t B, int32_t C,
                           irr_dtype dt, int64_t ignore_index, const float* grad_out,
                           const void* workspace, void* da, void* db, irr_stream_t stream) {
  if (B <= 0 || C < 1 || !a || !b || !target || !grad_out || !workspace || !da || !db)
    return IRR_ERR_INVALID_ARG;
  if (!any_float(dt)) return IRR_ERR_UNSUPPORTED_DTYPE;
  return ce_pair_bwd(a, b, target, B, C, dt, ignore_index, grad_out, workspace, da, db,
                     reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
