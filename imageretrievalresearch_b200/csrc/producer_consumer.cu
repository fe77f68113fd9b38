// producer_consumer.cu — the two steps either side of the retrieval-ranking path:
//
//   get_fm: global average pool [B,C,H,W] -> [B,C]       train/train_efficient_cos_con_ce_loss.py:103-122
//           (AvgPool2d((H,W)) + reshape), the producer of every embedding the path consumes;
//           written straight into the embedding row in fp32 or bf16.
//   cross-entropy on the classifier logits                train/train_efficient_cos_con_ce_loss.py:160,240-242
//           loss_ce = CrossEntropyLoss()(lbl_ims, clss) + CrossEntropyLoss()(lbl_poss, clss):
//           both terms, forward and backward, in one launch.
//
// Both are HBM-bound streaming kernels.  The pool stages a contiguous slab of 128 (b,c) rows with one
// bulk async copy (cp.async.bulk) and lets each thread sum its H*W elements from shared memory in a
// rotated order (bank-conflict-free for any H*W); the cross-entropy is one warp per logits row.
#include <cuda_fp16.h>

#include "irr_common.cuh"
#include "irr_kernels.h"

namespace irr {
namespace {

constexpr int POOL_ROWS = 128;
constexpr int POOL_SMEM_MAX = 160 * 1024;

template <int DT>
__device__ __forceinline__ float load_elem(const void* base, size_t i) {
  if (DT == IRR_F32) return static_cast<const float*>(base)[i];
  if (DT == IRR_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(base)[i]);
  return __half2float(static_cast<const __half*>(base)[i]);
}
template <int DT>
__device__ __forceinline__ void store_elem(void* base, size_t i, float v) {
  if (DT == IRR_F32) static_cast<float*>(base)[i] = v;
  else if (DT == IRR_BF16) static_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
  else static_cast<__half*>(base)[i] = __float2half_rn(v);
}
__host__ __device__ inline int elem_bytes(int dt) { return dt == IRR_F32 ? 4 : 2; }

template <int IN, int OUT>
__global__ void __launch_bounds__(POOL_ROWS)
avgpool_fwd_kernel(const void* __restrict__ fm, int64_t rows, int hw, int rows_per_cta,
                   void* __restrict__ out, float inv_hw) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
  const int64_t left = rows - r0;
  const int nrows = left < rows_per_cta ? static_cast<int>(left) : rows_per_cta;
  const size_t es = elem_bytes(IN);
  const size_t bytes = static_cast<size_t>(nrows) * hw * es;
  const uint8_t* src = static_cast<const uint8_t*>(fm) + static_cast<size_t>(r0) * hw * es;
  const bool bulk = (bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (bulk) {
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
      mbar_init(b, 1);
      fence_mbar_init();
      mbar_arrive_expect_tx(b, static_cast<uint32_t>(bytes));
      bulk_load_1d(smem_u32(smem), src, static_cast<uint32_t>(bytes), b);
    }
    __syncthreads();
    mbar_wait(b, 0, 900);
  } else {  // ragged tail / unaligned slab: plain cooperative copy
    for (size_t i = threadIdx.x; i < bytes; i += POOL_ROWS) smem[i] = src[i];
    __syncthreads();
  }
  const int t = threadIdx.x;
  if (t < nrows) {
    const size_t base = static_cast<size_t>(t) * hw;
    float s0 = 0.f, s1 = 0.f;
    int j = t % hw;  // rotated start: lanes hit different banks whatever hw is
    int i = 0;
    for (; i + 1 < hw; i += 2) {
      s0 += load_elem<IN>(smem, base + j);
      if (++j == hw) j = 0;
      s1 += load_elem<IN>(smem, base + j);
      if (++j == hw) j = 0;
    }
    if (i < hw) s0 += load_elem<IN>(smem, base + j);
    store_elem<OUT>(out, static_cast<size_t>(r0) + t, (s0 + s1) * inv_hw);
  }
}

template <int GO, int GF>
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const void* __restrict__ grad_out, int64_t rows, int hw, void* __restrict__ grad_fm,
                   float inv_hw) {
  const int64_t total = rows * hw;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x)
    store_elem<GF>(grad_fm, e, load_elem<GO>(grad_out, e / hw) * inv_hw);
}

// ---- cross-entropy: one warp per logits row, two logits tensors sharing the targets -------------
// loss = mean_i(lse(a_i) - a_i[t_i]) + mean_i(lse(b_i) - b_i[t_i]);  d/da = (softmax(a_i) - onehot)/B * g
template <int DT>
__global__ void __launch_bounds__(256)
ce_pair_kernel(const void* __restrict__ a, const void* __restrict__ b, const int64_t* __restrict__ target,
               int64_t B, int C, int64_t ignore_index, float* __restrict__ row_loss /*[2,B]*/,
               void* __restrict__ da, void* __restrict__ db, const float* __restrict__ grad_out,
               const float* __restrict__ inv_count) {
  const int lane = threadIdx.x & 31;
  const int64_t w = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= 2 * B) return;
  const bool second = w >= B;
  const int64_t row = second ? w - B : w;
  const void* x = second ? b : a;
  void* dx = second ? db : da;
  const int64_t tgt = target[row];
  const bool ignored = tgt == ignore_index;
  const size_t base = static_cast<size_t>(row) * C;
  float m = kNegInf;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, load_elem<DT>(x, base + c));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(load_elem<DT>(x, base + c) - m);
  se = warp_sum(se);
  const float lse = m + logf(se);
  if (row_loss && lane == 0)
    row_loss[w] = ignored ? 0.f : lse - load_elem<DT>(x, base + (tgt >= 0 && tgt < C ? tgt : 0));
  if (dx) {
    const float g = ignored ? 0.f : __ldg(grad_out) * __ldg(inv_count);
    const float inv_se = 1.0f / se;
    for (int c = lane; c < C; c += 32) {
      const float p = expf(load_elem<DT>(x, base + c) - m) * inv_se;
      store_elem<DT>(dx, base + c, g * (p - (c == tgt ? 1.f : 0.f)));
    }
  }
}

// deterministic finish: fixed-order sum of the 2*B row losses by one CTA, divided by the number of
// non-ignored targets (torch's 'mean'); also publishes 1/count for the backward
__global__ void __launch_bounds__(256)
ce_finish_kernel(const float* __restrict__ row_loss, const int64_t* __restrict__ target, int64_t B,
                 int64_t ignore_index, float* __restrict__ out_loss /*[3]: sum, a, b*/,
                 float* __restrict__ inv_count) {
  __shared__ float sa[256], sb[256];
  __shared__ int sc[256];
  float la = 0.f, lb = 0.f;
  int cnt = 0;
  for (int64_t i = threadIdx.x; i < B; i += 256) {
    la += row_loss[i];
    lb += row_loss[B + i];
    cnt += target[i] != ignore_index;
  }
  sa[threadIdx.x] = la; sb[threadIdx.x] = lb; sc[threadIdx.x] = cnt;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sa[threadIdx.x] += sa[threadIdx.x + s];
      sb[threadIdx.x] += sb[threadIdx.x + s];
      sc[threadIdx.x] += sc[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float inv = sc[0] > 0 ? 1.0f / static_cast<float>(sc[0]) : __int_as_float(0x7fc00000);
    out_loss[1] = sa[0] * inv;
    out_loss[2] = sb[0] * inv;
    out_loss[0] = out_loss[1] + out_loss[2];
    *inv_count = sc[0] > 0 ? inv : 0.f;
  }
}

}  // namespace

irr_status avgpool_fwd(const void* fm, int in_dt, int64_t rows, int32_t hw, void* out, int out_dt,
                       cudaStream_t st) {
  if (rows == 0) return IRR_OK;
  const size_t row_bytes = static_cast<size_t>(hw) * elem_bytes(in_dt);
  int rpc = POOL_ROWS;
  if (row_bytes * rpc > POOL_SMEM_MAX) rpc = static_cast<int>(POOL_SMEM_MAX / row_bytes);
  if (rpc < 1) return IRR_ERR_ROW_TOO_LONG;
  const size_t smem = row_bytes * rpc;
  const unsigned grid = static_cast<unsigned>((rows + rpc - 1) / rpc);
  const float inv = 1.0f / static_cast<float>(hw);
#define IRR_POOL(I, O)                                                                              \
  do {                                                                                              \
    auto kern = avgpool_fwd_kernel<I, O>;                                                           \
    static std::atomic<uint64_t> attr_done{0};   /* the budget, once per instantiation and device */ \
    if (attr_needed(attr_done)) {                                                                   \
      IRR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                        static_cast<int>(POOL_SMEM_MAX)));                          \
      attr_set(attr_done);                                                                          \
    }                                                                                               \
    kern<<<grid, POOL_ROWS, smem, st>>>(fm, rows, hw, rpc, out, inv);                               \
  } while (0)
  if (out_dt == IRR_F32) {
    if (in_dt == IRR_F32) IRR_POOL(IRR_F32, IRR_F32);
    else if (in_dt == IRR_BF16) IRR_POOL(IRR_BF16, IRR_F32);
    else IRR_POOL(IRR_F16, IRR_F32);
  } else if (out_dt == IRR_BF16) {
    if (in_dt == IRR_F32) IRR_POOL(IRR_F32, IRR_BF16);
    else if (in_dt == IRR_BF16) IRR_POOL(IRR_BF16, IRR_BF16);
    else IRR_POOL(IRR_F16, IRR_BF16);
  } else {
    return IRR_ERR_UNSUPPORTED_DTYPE;
  }
#undef IRR_POOL
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

irr_status avgpool_bwd(const void* grad_out, int go_dt, int64_t rows, int32_t hw, void* grad_fm,
                       int gf_dt, cudaStream_t st) {
  if (rows == 0) return IRR_OK;
  const int64_t total = rows * hw;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  const float inv = 1.0f / static_cast<float>(hw);
  const unsigned grid = static_cast<unsigned>(blocks);
#define IRR_POOLB(G, F) avgpool_bwd_kernel<G, F><<<grid, 256, 0, st>>>(grad_out, rows, hw, grad_fm, inv)
  if (go_dt == IRR_F32) {
    if (gf_dt == IRR_F32) IRR_POOLB(IRR_F32, IRR_F32);
    else if (gf_dt == IRR_BF16) IRR_POOLB(IRR_F32, IRR_BF16);
    else IRR_POOLB(IRR_F32, IRR_F16);
  } else if (go_dt == IRR_BF16) {
    if (gf_dt == IRR_F32) IRR_POOLB(IRR_BF16, IRR_F32);
    else if (gf_dt == IRR_BF16) IRR_POOLB(IRR_BF16, IRR_BF16);
    else IRR_POOLB(IRR_BF16, IRR_F16);
  } else {
    return IRR_ERR_UNSUPPORTED_DTYPE;
  }
#undef IRR_POOLB
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

size_t ce_pair_workspace_bytes(int64_t B) { return static_cast<size_t>(2 * B + 4) * sizeof(float); }

irr_status ce_pair_fwd(const void* a, const void* b, const int64_t* target, int64_t B, int32_t C,
                       int dt, int64_t ignore_index, float* out_loss, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  if (ws_bytes < ce_pair_workspace_bytes(B)) return IRR_ERR_WORKSPACE_TOO_SMALL;
  float* row_loss = static_cast<float*>(ws);
  float* inv_count = row_loss + 2 * B;
  const unsigned grid = static_cast<unsigned>((2 * B + 7) / 8);
  if (dt == IRR_F32)
    ce_pair_kernel<IRR_F32><<<grid, 256, 0, st>>>(a, b, target, B, C, ignore_index, row_loss, nullptr,
                                                  nullptr, nullptr, nullptr);
  else if (dt == IRR_BF16)
    ce_pair_kernel<IRR_BF16><<<grid, 256, 0, st>>>(a, b, target, B, C, ignore_index, row_loss,
                                                   nullptr, nullptr, nullptr, nullptr);
  else
    ce_pair_kernel<IRR_F16><<<grid, 256, 0, st>>>(a, b, target, B, C, ignore_index, row_loss, nullptr,
                                                  nullptr, nullptr, nullptr);
  IRR_LAUNCH_CHECK();
  ce_finish_kernel<<<1, 256, 0, st>>>(row_loss, target, B, ignore_index, out_loss, inv_count);
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

irr_status ce_pair_bwd(const void* a, const void* b, const int64_t* target, int64_t B, int32_t C,
                       int dt, int64_t ignore_index, const float* grad_out, const void* ws, void* da,
                       void* db, cudaStream_t st) {
  const float* inv_count = static_cast<const float*>(ws) + 2 * B;
  const unsigned grid = static_cast<unsigned>((2 * B + 7) / 8);
  if (dt == IRR_F32)
    ce_pair_kernel<IRR_F32><<<grid, 256, 0, st>>>(a, b, target, B, C, ignore_index, nullptr, da, db,
                                                  grad_out, inv_count);
  else if (dt == IRR_BF16)
    ce_pair_kernel<IRR_BF16><<<grid, 256, 0, st>>>(a, b, target, B, C, ignore_index, nullptr, da, db,
                                                   grad_out, inv_count);
  else
    ce_pair_kernel<IRR_F16><<<grid, 256, 0, st>>>(a, b, target, B, C, ignore_index, nullptr, da, db,
                                                  grad_out, inv_count);
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

}  // namespace irr
