// row_norms.cu — HBM-streaming row reductions: inverse row norms (what a gallery handle caches and
// what the top-k kernels scale scores with) and row-wise cosine similarity, the literal
// CosineSimilarity(dim=1, eps) of train/train_efficient_cos_con_ce_loss.py:89 applied to pairs
// (:377,381) or to one query against the gallery (:273).
//
// One warp per row, 128-bit L1-bypassing loads, four loads in flight per lane, shuffle reduction.
#include "irr_common.cuh"
#include "irr_kernels.h"

namespace irr {
namespace {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;

// ELT: element kind of the rows = irr_dtype (0 fp32, 1 bf16, 2 fp16)
template <int ELT>
__device__ __forceinline__ float sumsq_vec(const uint4& u) {
  if (ELT != IRR_F32) {
    float s = 0.f;
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack16x2(w[j], ELT == IRR_F16);
      s = fmaf(f.x, f.x, s);
      s = fmaf(f.y, f.y, s);
    }
    return s;
  } else {
    const float a = __uint_as_float(u.x), b = __uint_as_float(u.y), c = __uint_as_float(u.z),
                d = __uint_as_float(u.w);
    return fmaf(a, a, fmaf(b, b, fmaf(c, c, d * d)));
  }
}

template <int ELT>
__device__ __forceinline__ void dot3_vec(const uint4& ua, const uint4& ub, float& aa, float& bb,
                                         float& ab) {
  if (ELT != IRR_F32) {
    const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w};
    const uint32_t wb[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = unpack16x2(wa[j], ELT == IRR_F16), fb = unpack16x2(wb[j], ELT == IRR_F16);
      const float a0 = fa.x, a1 = fa.y, b0 = fb.x, b1 = fb.y;
      aa = fmaf(a0, a0, fmaf(a1, a1, aa));
      bb = fmaf(b0, b0, fmaf(b1, b1, bb));
      ab = fmaf(a0, b0, fmaf(a1, b1, ab));
    }
  } else {
    const float a[4] = {__uint_as_float(ua.x), __uint_as_float(ua.y), __uint_as_float(ua.z),
                        __uint_as_float(ua.w)};
    const float b[4] = {__uint_as_float(ub.x), __uint_as_float(ub.y), __uint_as_float(ub.z),
                        __uint_as_float(ub.w)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      aa = fmaf(a[j], a[j], aa);
      bb = fmaf(b[j], b[j], bb);
      ab = fmaf(a[j], b[j], ab);
    }
  }
}

template <int ELT>
__global__ void __launch_bounds__(THREADS)
row_inv_norm_kernel(const uint4* __restrict__ x, int64_t N, int vec_per_row, float eps,
                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * WARPS + (threadIdx.x >> 5);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * WARPS;
  for (int64_t row = warp0; row < N; row += stride) {
    const uint4* r = x + row * vec_per_row;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int v = lane;
    for (; v + 96 < vec_per_row; v += 128) {
      const uint4 u0 = ldg_stream(r + v), u1 = ldg_stream(r + v + 32), u2 = ldg_stream(r + v + 64),
                  u3 = ldg_stream(r + v + 96);
      s0 += sumsq_vec<ELT>(u0);
      s1 += sumsq_vec<ELT>(u1);
      s2 += sumsq_vec<ELT>(u2);
      s3 += sumsq_vec<ELT>(u3);
    }
    for (; v < vec_per_row; v += 32) s0 += sumsq_vec<ELT>(ldg_stream(r + v));
    const float ss = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) out[row] = 1.0f / fmaxf(sqrtf(ss), eps);
  }
}

template <int ELT>
__global__ void __launch_bounds__(THREADS)
pair_cosine_kernel(const uint4* __restrict__ x1, int64_t x1_row_stride_vec,
                   const uint4* __restrict__ x2, int64_t N, int vec_per_row, float eps,
                   float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * WARPS + (threadIdx.x >> 5);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * WARPS;
  for (int64_t row = warp0; row < N; row += stride) {
    const uint4* a = x1 + row * x1_row_stride_vec;  // stride 0: one query row broadcast
    const uint4* b = x2 + row * vec_per_row;
    float aa = 0.f, bb = 0.f, ab = 0.f;
    int v = lane;
    for (; v + 32 < vec_per_row; v += 64) {
      const uint4 a0 = __ldg(a + v), a1 = __ldg(a + v + 32);
      const uint4 b0 = ldg_stream(b + v), b1 = ldg_stream(b + v + 32);
      dot3_vec<ELT>(a0, b0, aa, bb, ab);
      dot3_vec<ELT>(a1, b1, aa, bb, ab);
    }
    for (; v < vec_per_row; v += 32) dot3_vec<ELT>(__ldg(a + v), ldg_stream(b + v), aa, bb, ab);
    aa = warp_sum(aa);
    bb = warp_sum(bb);
    ab = warp_sum(ab);
    if (lane == 0) out[row] = ab / (fmaxf(sqrtf(aa), eps) * fmaxf(sqrtf(bb), eps));
  }
}

int grid_for_rows(int64_t N) {
  const int64_t want = (N + WARPS - 1) / WARPS;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8 * 4;  // 8 resident CTAs/SM, 4 waves
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

irr_status row_inv_norms(const void* x, int64_t N, int32_t D, irr_dtype dt, float eps, float* out,
                         cudaStream_t st) {
  if (N == 0) return IRR_OK;
  const int vec = D * dtype_bytes(dt) / 16;
  const int grid = grid_for_rows(N);
  const uint4* xv = static_cast<const uint4*>(x);
  if (dt == IRR_BF16)
    row_inv_norm_kernel<IRR_BF16><<<grid, THREADS, 0, st>>>(xv, N, vec, eps, out);
  else if (dt == IRR_F16)
    row_inv_norm_kernel<IRR_F16><<<grid, THREADS, 0, st>>>(xv, N, vec, eps, out);
  else
    row_inv_norm_kernel<IRR_F32><<<grid, THREADS, 0, st>>>(xv, N, vec, eps, out);
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

irr_status pair_cosine(const void* x1, int64_t x1_rows, const void* x2, int64_t N, int32_t D,
                       irr_dtype dt, float eps, float* out, cudaStream_t st) {
  if (N == 0) return IRR_OK;
  const int vec = D * dtype_bytes(dt) / 16;
  const int64_t s1 = x1_rows == 1 ? 0 : vec;
  const int grid = grid_for_rows(N);
  const uint4 *a = static_cast<const uint4*>(x1), *b = static_cast<const uint4*>(x2);
  if (dt == IRR_BF16)
    pair_cosine_kernel<IRR_BF16><<<grid, THREADS, 0, st>>>(a, s1, b, N, vec, eps, out);
  else if (dt == IRR_F16)
    pair_cosine_kernel<IRR_F16><<<grid, THREADS, 0, st>>>(a, s1, b, N, vec, eps, out);
  else
    pair_cosine_kernel<IRR_F32><<<grid, THREADS, 0, st>>>(a, s1, b, N, vec, eps, out);
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

}  // namespace irr
