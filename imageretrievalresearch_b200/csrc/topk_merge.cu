// topk_merge.cu — K3: candidate-list merge, used twice:
//   (1) inside one GPU: fold the per-(query tile, gallery chunk) partial lists of the top-k
//       kernels, apply 1/max(|q|,eps) (computed here from the query row: CosineSimilarity's
//       per-operand clamp, train/train_efficient_cos_con_ce_loss.py:89) and widen to int64;
//   (2) across GPUs: merge the all-gathered [G,Q,k] lists of a row-sharded gallery.
// plus the top-1 / top-k hit counters that replace the per-row Python label tests
// (train/train_efficient_cos_con_ce_loss.py:279-281,390-392; inference/inference.py:237,242).
//
// One warp per query: each lane keeps a sorted top-k of its strided candidates in registers, then
// k rounds of shuffle-argmax pop the winners.  Order: score descending, ties -> lower index.
#include "irr_common.cuh"
#include "irr_kernels.h"

namespace irr {
namespace {

constexpr int THREADS = 128;
constexpr int WARPS = THREADS / 32;

template <int KMAX>
__global__ void __launch_bounds__(THREADS)
merge_partials_kernel(const float* __restrict__ part_val, const int32_t* __restrict__ part_idx,
                      int S, int64_t Q, int k, const void* __restrict__ q, int D, int elt,
                      float eps, int64_t idx_offset, float* __restrict__ out_val,
                      int64_t* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int64_t qi = static_cast<int64_t>(blockIdx.x) * WARPS + (threadIdx.x >> 5);
  // programmatic dependent launch: scheduled while the top-k kernel drains; its partial lists are
  // complete and visible once this returns (a no-op when launched without the attribute)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (qi >= Q) return;

  // query norm
  float ss = 0.f;
  if (elt != IRR_F32) {   // bf16 / fp16 rows: two elements per 32-bit word
    const uint32_t* r = reinterpret_cast<const uint32_t*>(static_cast<const uint16_t*>(q) + qi * D);
    for (int v = lane; v < D / 2; v += 32) {
      const float2 f = unpack16x2(__ldg(r + v), elt == IRR_F16);
      ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
    }
  } else {
    const float* r = static_cast<const float*>(q) + qi * D;
    for (int v = lane; v < D; v += 32) {
      const float a = __ldg(r + v);
      ss = fmaf(a, a, ss);
    }
  }
  const float qn = 1.0f / fmaxf(sqrtf(warp_sum(ss)), eps);

  TopKList<KMAX, int32_t> L;
  L.reset();
  const int total = S * k;
  for (int c = lane; c < total; c += 32) {
    const int s = c / k, j = c - s * k;
    const size_t o = (static_cast<size_t>(s) * Q + qi) * k + j;
    L.push_any(__ldg(part_val + o), __ldg(part_idx + o));
  }
  warp_merge_topk<KMAX, int32_t>(L, k, [&](int j, float v, long long i) {
    if (lane == 0) {
      out_val[qi * k + j] = i >= 0 ? v * qn : kNegInf;
      out_idx[qi * k + j] = i >= 0 ? i + idx_offset : -1;
    }
  });
}

template <int KMAX>
__global__ void __launch_bounds__(THREADS)
merge_candidates_kernel(const float* __restrict__ cand_val, int64_t val_rank_stride,
                        const int64_t* __restrict__ cand_idx, int64_t idx_rank_stride, int G,
                        int64_t Q, int k, float* __restrict__ out_val,
                        int64_t* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int64_t qi = static_cast<int64_t>(blockIdx.x) * WARPS + (threadIdx.x >> 5);
  if (qi >= Q) return;
  TopKList<KMAX, long long> L;
  L.reset();
  const int total = G * k;
  for (int c = lane; c < total; c += 32) {
    const int g = c / k, j = c - g * k;
    const size_t o = static_cast<size_t>(qi) * k + j;
    L.push_any(__ldg(cand_val + g * val_rank_stride + o),
               static_cast<long long>(__ldg(cand_idx + g * idx_rank_stride + o)));
  }
  warp_merge_topk<KMAX, long long>(L, k, [&](int j, float v, long long i) {
    if (lane == 0) {
      out_val[qi * k + j] = i >= 0 ? v : kNegInf;
      out_idx[qi * k + j] = i >= 0 ? i : -1;
    }
  });
}

__global__ void __launch_bounds__(256)
topk_hits_kernel(const int64_t* __restrict__ idx, int64_t Q, int k,
                 const int64_t* __restrict__ q_label, const int64_t* __restrict__ g_label,
                 int64_t N, int64_t instance_offset, unsigned long long* __restrict__ out_hits) {
  unsigned long long h1 = 0, hk = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < Q;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    bool any = false, first = false;
    const int64_t want = q_label ? q_label[i] : i + instance_offset;
    for (int j = 0; j < k; ++j) {
      const int64_t g = idx[i * k + j];
      if (g < 0 || (g_label && g >= N)) continue;
      const bool hit = (g_label ? g_label[g] : g) == want;
      any |= hit;
      if (j == 0) first = hit;
    }
    h1 += first;
    hk += any;
  }
  // integer adds: order-independent, so the atomics keep the result deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    h1 += __shfl_xor_sync(0xffffffffu, h1, o);
    hk += __shfl_xor_sync(0xffffffffu, hk, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (h1) atomicAdd(out_hits + 0, h1);
    if (hk) atomicAdd(out_hits + 1, hk);
  }
}

// an empty gallery shard: every slot is padding
__global__ void __launch_bounds__(256)
fill_padding_kernel(float* __restrict__ out_val, int64_t* __restrict__ out_idx, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    out_val[i] = kNegInf;
    out_idx[i] = -1;
  }
}

}  // namespace

irr_status fill_padding(float* out_val, int64_t* out_idx, int64_t n, cudaStream_t st) {
  if (n <= 0) return IRR_OK;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  fill_padding_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(out_val, out_idx, n);
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

irr_status merge_partials(const float* part_val, const int32_t* part_idx, int32_t S, int64_t Q,
                          int32_t k, const void* q, int32_t D, irr_dtype dt, float eps,
                          int64_t idx_offset, float* out_val, int64_t* out_idx, cudaStream_t st) {
  if (Q == 0) return IRR_OK;
  const int grid = static_cast<int>((Q + WARPS - 1) / WARPS);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const int iS = S, ik = k, iD = D, idt = static_cast<int>(dt);
  const cudaError_t e =
      k <= 4 ? cudaLaunchKernelEx(&cfg, merge_partials_kernel<4>, part_val, part_idx, iS, Q, ik, q, iD,
                                  idt, eps, idx_offset, out_val, out_idx)
             : cudaLaunchKernelEx(&cfg, merge_partials_kernel<16>, part_val, part_idx, iS, Q, ik, q, iD,
                                  idt, eps, idx_offset, out_val, out_idx);
  if (e != cudaSuccess) return static_cast<irr_status>(static_cast<int>(e));
  return IRR_OK;
}

irr_status merge_candidates(const float* cand_val, int64_t val_rank_stride, const int64_t* cand_idx,
                            int64_t idx_rank_stride, int32_t G, int64_t Q, int32_t k,
                            float* out_val, int64_t* out_idx, cudaStream_t st) {
  if (Q == 0) return IRR_OK;
  if (k > IRR_MAX_K_FUSED)
    return merge_candidates_large(cand_val, val_rank_stride, cand_idx, idx_rank_stride, G, Q, k,
                                  out_val, out_idx, st);
  const int grid = static_cast<int>((Q + WARPS - 1) / WARPS);
  if (k <= 4)
    merge_candidates_kernel<4><<<grid, THREADS, 0, st>>>(cand_val, val_rank_stride, cand_idx,
                                                         idx_rank_stride, G, Q, k, out_val, out_idx);
  else
    merge_candidates_kernel<16><<<grid, THREADS, 0, st>>>(cand_val, val_rank_stride, cand_idx,
                                                          idx_rank_stride, G, Q, k, out_val, out_idx);
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

irr_status topk_hits(const int64_t* idx, int64_t Q, int32_t k, const int64_t* q_label,
                     const int64_t* g_label, int64_t N, int64_t instance_offset, int64_t* out_hits,
                     cudaStream_t st) {
  IRR_CUDA_TRY(cudaMemsetAsync(out_hits, 0, 2 * sizeof(int64_t), st));
  if (Q == 0) return IRR_OK;
  int64_t blocks = (Q + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  topk_hits_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(
      idx, Q, k, q_label, g_label, N, instance_offset,
      reinterpret_cast<unsigned long long*>(out_hits));
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

}  // namespace irr
