// triplet_loss.cu — K2: fused forward(+backward) of the two embedding losses over (q, p, n) rows.
// Replaces, in ONE pass over the rows,
//   ContrastiveLoss.forward                 utils/contrastive_loss.py:56-61   (eps 1e-9, :34)
//   torch.nn.CosineEmbeddingLoss(margin)    train/train_efficient_cos_con_ce_loss.py:158,230-231
//   the pos/neg composition                 train/train_efficient_cos_con_ce_loss.py:230-237
//   the logged paired cosine scores         train/train_efficient_cos_con_ce_loss.py:377-382
// and their autograd backward (closed forms: SURVEY.md §A.1).
//
// HBM-bound design: every input byte crosses L2->SM exactly once.  One warp owns one row triplet at
// a time; lane 0 stages the three rows into the warp's private shared-memory slot with 1-D bulk
// async copies (cp.async.bulk, completion on an mbarrier) one row ahead of the arithmetic; the
// warp reduces seven sums with 128-bit shared loads + shuffles, turns them into the four loss terms
// and seven gradient coefficients, and streams dq/dp/dn straight from the staged rows with 128-bit
// stores.  Each gradient is a per-row linear combination of the three rows:
//   dq = aqq*q + aqp*p + aqn*n     dp = app*p + aqp*q     dn = ann*n + aqn*q
// Loss scalars: fixed row->warp assignment, per-warp partials, last-CTA-done fixed-order reduction
// (deterministic; the sync word resets itself, see irr_b200.h).
#include "irr_common.cuh"
#include "irr_kernels.h"

namespace irr {
namespace {

constexpr int LSTAGES = 2;
constexpr int MAX_WARPS = 16;
constexpr int SMEM_BUDGET = 220 * 1024;

struct Coef {
  float aqq, aqp, aqn, app, ann;
};

template <bool BF16>
struct Vec {
  static constexpr int N = BF16 ? 8 : 4;
  __device__ static __forceinline__ void unpack(const uint4& u, float (&f)[N]) {
    if constexpr (BF16) {
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[2 * j] = bf16lo(w[j]);
        f[2 * j + 1] = bf16hi(w[j]);
      }
    } else {
      f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y);
      f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    }
  }
  __device__ static __forceinline__ uint4 pack(const float (&f)[N]) {
    uint4 u;
    if constexpr (BF16) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        w[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      u = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      u = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                     __float_as_uint(f[3]));
    }
    return u;
  }
};

struct RowSums {
  float qq, pp, nn, qp, qn, dp, dn;
};

// Loss terms and gradient coefficients of one row from its seven sums.
// w[j]: weight of loss j in the differentiated scalar (already divided by B for 'mean').
struct RowOut {
  float l[4];
  Coef c;
};

__device__ __forceinline__ RowOut triplet_row(const RowSums& s, float m_cos, float m_con,
                                              const float (&w)[4]) {
  RowOut o;
  const float A = s.qq + kCosEmbEps, P = s.pp + kCosEmbEps, Nn = s.nn + kCosEmbEps;
  // same operation order as ATen: cos = prod / sqrt(mag1 * mag2), IEEE sqrt and divide
  const float dqp = sqrtf(A * P), dqn = sqrtf(A * Nn);
  const float cqp = s.qp / dqp, cqn = s.qn / dqn;
  const float rqp = 1.0f / dqp, rqn = 1.0f / dqn;
  o.l[IRR_L_COS_POS] = 1.0f - cqp;
  const float over = cqn - m_cos;
  o.l[IRR_L_COS_NEG] = fmaxf(over, 0.0f);
  o.l[IRR_L_CON_POS] = 0.5f * s.dp;
  const float sn = sqrtf(s.dn + kContrastiveEps);
  const float gap = fmaxf(m_con - sn, 0.0f);
  o.l[IRR_L_CON_NEG] = 0.5f * gap * gap;
  const float w1a = over >= 0.0f ? w[1] : 0.0f;  // clamp_min passes the gradient at equality
  const float w3c = w[3] * (gap / sn);
  o.c.aqq = w[0] * cqp / A - w1a * cqn / A + w[2] - w3c;
  o.c.aqp = -w[0] * rqp - w[2];
  o.c.aqn = w1a * rqn + w3c;
  o.c.app = w[0] * cqp / P + w[2];
  o.c.ann = -w1a * cqn / Nn - w3c;
  return o;
}

// Pair form: one loss of `kind` with label y (contrastive) / target t (cosine embedding).
__device__ __forceinline__ RowOut pair_row(const RowSums& s, int kind, float y, float margin,
                                           float w) {
  RowOut o;
  o.l[1] = o.l[2] = o.l[3] = 0.f;
  o.c.aqn = o.c.ann = 0.f;
  if (kind == IRR_LOSS_CONTRASTIVE) {
    const float sd = sqrtf(s.dp + kContrastiveEps);
    const float gap = fmaxf(margin - sd, 0.0f);
    o.l[0] = 0.5f * (y * s.dp + (1.0f - y) * gap * gap);
    const float gcoef = w * (y - (1.0f - y) * (gap / sd));  // d/db = gcoef * (b - a)
    o.c.aqq = gcoef; o.c.aqp = -gcoef; o.c.app = gcoef;
  } else {
    const float A = s.qq + kCosEmbEps, P = s.pp + kCosEmbEps;
    const float den = sqrtf(A * P);
    const float c = s.qp / den;
    const float r = 1.0f / den;
    float sign = 0.f;  // d loss / d c
    if (y == 1.0f) { o.l[0] = 1.0f - c; sign = -1.0f; }
    else if (y == -1.0f) { o.l[0] = fmaxf(c - margin, 0.0f); sign = (c - margin >= 0.0f) ? 1.0f : 0.0f; }
    else o.l[0] = 0.0f;
    const float ws = w * sign;
    o.c.aqq = -ws * c / A; o.c.aqp = ws * r; o.c.app = -ws * c / P;
  }
  return o;
}

struct KParams {
  const uint4 *q, *p, *n;
  const float* label;
  int64_t label_count, B;
  int vec_per_row;  // 16-byte vectors per row
  int kind;
  float m_cos, m_con, pair_eps;
  float w[4];       // loss weights incl. 1/B
  float red_scale;  // 1/B or 1
  float *losses, *pair_cos, *row_stats;
  uint4 *dq, *dp, *dn;
  unsigned int* sync_word;
  float* partials;  // [gridDim.x * warps][4]
};

template <bool BF16, bool TRIPLET>
__global__ void __launch_bounds__(MAX_WARPS * 32, 1)
loss_fwd_bwd_kernel(const KParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ROWS = TRIPLET ? 3 : 2;
  const uint32_t row_bytes = static_cast<uint32_t>(P.vec_per_row) * 16u;
  const uint32_t slot_bytes = ROWS * row_bytes;
  // [warps][LSTAGES][ROWS][row_bytes] then the mbarriers
  uint8_t* my_slots = smem + static_cast<size_t>(warp) * LSTAGES * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(warps) * LSTAGES * slot_bytes);
  const uint32_t bar0 = smem_u32(bars + warp * LSTAGES);

  if (lane == 0) {
    for (int s = 0; s < LSTAGES; ++s) mbar_init(bar0 + 8u * s, 1);
    fence_mbar_init();
  }
  __syncthreads();

  const int64_t gw = static_cast<int64_t>(blockIdx.x) * warps + warp;
  const int64_t tw = static_cast<int64_t>(gridDim.x) * warps;

  auto issue = [&](int64_t row, int s) {
    const uint32_t dst = smem_u32(my_slots + static_cast<size_t>(s) * slot_bytes);
    const uint32_t bar = bar0 + 8u * s;
    mbar_arrive_expect_tx(bar, slot_bytes);
    bulk_load_1d(dst, P.q + row * P.vec_per_row, row_bytes, bar);
    bulk_load_1d(dst + row_bytes, P.p + row * P.vec_per_row, row_bytes, bar);
    if (TRIPLET) bulk_load_1d(dst + 2 * row_bytes, P.n + row * P.vec_per_row, row_bytes, bar);
  };

  if (lane == 0) {
    for (int s = 0; s < LSTAGES; ++s) {
      const int64_t row = gw + s * tw;
      if (row < P.B) issue(row, s);
    }
  }
  __syncwarp();

  float lsum[4] = {0.f, 0.f, 0.f, 0.f};
  constexpr int VN = Vec<BF16>::N;
  int it = 0;
  for (int64_t row = gw; row < P.B; row += tw, ++it) {
    const int s = it % LSTAGES;
    const uint32_t parity = (it / LSTAGES) & 1;
    mbar_wait(bar0 + 8u * s, parity, 500 + s);
    const uint4* sq = reinterpret_cast<const uint4*>(my_slots + static_cast<size_t>(s) * slot_bytes);
    const uint4* sp = sq + P.vec_per_row;
    const uint4* sn = sp + P.vec_per_row;

    RowSums S = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int v = lane; v < P.vec_per_row; v += 32) {
      float fq[VN], fp[VN], fn[VN];
      Vec<BF16>::unpack(sq[v], fq);
      Vec<BF16>::unpack(sp[v], fp);
      if (TRIPLET) Vec<BF16>::unpack(sn[v], fn);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        S.qq = fmaf(fq[j], fq[j], S.qq);
        S.pp = fmaf(fp[j], fp[j], S.pp);
        S.qp = fmaf(fq[j], fp[j], S.qp);
        const float d = fp[j] - fq[j];
        S.dp = fmaf(d, d, S.dp);
        if (TRIPLET) {
          S.nn = fmaf(fn[j], fn[j], S.nn);
          S.qn = fmaf(fq[j], fn[j], S.qn);
          const float e = fn[j] - fq[j];
          S.dn = fmaf(e, e, S.dn);
        }
      }
    }
    S.qq = warp_sum(S.qq); S.pp = warp_sum(S.pp); S.qp = warp_sum(S.qp); S.dp = warp_sum(S.dp);
    if (TRIPLET) { S.nn = warp_sum(S.nn); S.qn = warp_sum(S.qn); S.dn = warp_sum(S.dn); }

    RowOut o;
    if (TRIPLET) {
      o = triplet_row(S, P.m_cos, P.m_con, P.w);
    } else {
      const float y = __ldg(P.label + (P.label_count == 1 ? 0 : row));
      o = pair_row(S, P.kind, y, P.kind == IRR_LOSS_CONTRASTIVE ? P.m_con : P.m_cos, P.w[0]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) lsum[j] += o.l[j];

    if (lane == 0) {
      if (P.pair_cos) {
        const float nq = fmaxf(sqrtf(S.qq), P.pair_eps);
        P.pair_cos[row] = S.qp / (nq * fmaxf(sqrtf(S.pp), P.pair_eps));
        if (TRIPLET) P.pair_cos[P.B + row] = S.qn / (nq * fmaxf(sqrtf(S.nn), P.pair_eps));
      }
      if (P.row_stats) {
        float4* rs = reinterpret_cast<float4*>(P.row_stats + row * IRR_ROW_STATS);
        rs[0] = make_float4(S.qq, S.pp, S.nn, S.qp);
        rs[1] = make_float4(S.qn, S.dp, S.dn, 0.f);
      }
    }

    if (P.dq) {
      uint4* gq = P.dq + row * P.vec_per_row;
      uint4* gp = P.dp + row * P.vec_per_row;
      uint4* gn = TRIPLET ? P.dn + row * P.vec_per_row : nullptr;
      for (int v = lane; v < P.vec_per_row; v += 32) {
        float fq[VN], fp[VN], fn[VN], oq[VN], op[VN], on[VN];
        Vec<BF16>::unpack(sq[v], fq);
        Vec<BF16>::unpack(sp[v], fp);
        if (TRIPLET) Vec<BF16>::unpack(sn[v], fn);
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          float a = fmaf(o.c.aqq, fq[j], o.c.aqp * fp[j]);
          op[j] = fmaf(o.c.app, fp[j], o.c.aqp * fq[j]);
          if (TRIPLET) {
            a = fmaf(o.c.aqn, fn[j], a);
            on[j] = fmaf(o.c.ann, fn[j], o.c.aqn * fq[j]);
          }
          oq[j] = a;
        }
        gq[v] = Vec<BF16>::pack(oq);
        gp[v] = Vec<BF16>::pack(op);
        if (TRIPLET) gn[v] = Vec<BF16>::pack(on);
      }
    }

    __syncwarp();  // every lane is done with the slot before it is refilled
    const int64_t next = row + LSTAGES * tw;
    if (lane == 0 && next < P.B) issue(next, s);
  }

  // ---- deterministic reduction of the loss scalars ----
  constexpr int NL = TRIPLET ? 4 : 1;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) P.partials[gw * 4 + j] = lsum[j];
  }
  __threadfence();
  __syncthreads();
  __shared__ int is_last;
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(P.sync_word, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && warp == 0) {
    __threadfence();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t i = lane; i < tw; i += 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += __ldcg(P.partials + i * 4 + j);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < NL; ++j) P.losses[j] = acc[j] * P.red_scale;
      *P.sync_word = 0u;  // self-reset for the next call on this workspace
    }
  }
}

// Backward from the saved row sums: no reduction, one read and one write per element.
struct BParams {
  const uint4 *q, *p, *n;
  const float* label;
  int64_t label_count, B;
  int vec_per_row;
  int kind;
  float m_cos, m_con, red_scale;
  const float* row_stats;
  const float* grad_out;
  uint4 *dq, *dp, *dn;
};

template <bool BF16, bool TRIPLET>
__global__ void __launch_bounds__(256)
loss_bwd_kernel(const BParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t tw = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  constexpr int VN = Vec<BF16>::N;
  float w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) w[j] = (TRIPLET || j == 0) ? __ldg(P.grad_out + j) * P.red_scale : 0.f;
  for (int64_t row = gw; row < P.B; row += tw) {
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(P.row_stats + row * IRR_ROW_STATS));
    const float4 r1 = __ldg(reinterpret_cast<const float4*>(P.row_stats + row * IRR_ROW_STATS) + 1);
    RowSums S = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z};
    Coef c;
    if (TRIPLET) {
      c = triplet_row(S, P.m_cos, P.m_con, w).c;
    } else {
      const float y = __ldg(P.label + (P.label_count == 1 ? 0 : row));
      c = pair_row(S, P.kind, y, P.kind == IRR_LOSS_CONTRASTIVE ? P.m_con : P.m_cos, w[0]).c;
    }
    const uint4* sq = P.q + row * P.vec_per_row;
    const uint4* sp = P.p + row * P.vec_per_row;
    const uint4* sn = TRIPLET ? P.n + row * P.vec_per_row : nullptr;
    uint4* gq = P.dq + row * P.vec_per_row;
    uint4* gp = P.dp + row * P.vec_per_row;
    uint4* gn = TRIPLET ? P.dn + row * P.vec_per_row : nullptr;
    for (int v = lane; v < P.vec_per_row; v += 32) {
      float fq[VN], fp[VN], fn[VN], oq[VN], op[VN], on[VN];
      Vec<BF16>::unpack(ldg_stream(sq + v), fq);
      Vec<BF16>::unpack(ldg_stream(sp + v), fp);
      if (TRIPLET) Vec<BF16>::unpack(ldg_stream(sn + v), fn);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        float a = fmaf(c.aqq, fq[j], c.aqp * fp[j]);
        op[j] = fmaf(c.app, fp[j], c.aqp * fq[j]);
        if (TRIPLET) {
          a = fmaf(c.aqn, fn[j], a);
          on[j] = fmaf(c.ann, fn[j], c.aqn * fq[j]);
        }
        oq[j] = a;
      }
      gq[v] = Vec<BF16>::pack(oq);
      gp[v] = Vec<BF16>::pack(op);
      if (TRIPLET) gn[v] = Vec<BF16>::pack(on);
    }
  }
}

struct LaunchShape {
  int warps, grid;
  size_t smem;
};

// warps per CTA from the shared-memory budget, then spread the rows over the SMs
bool shape_for(int64_t B, int32_t D, irr_dtype dt, bool triplet, LaunchShape* s) {
  const size_t row_bytes = static_cast<size_t>(D) * dtype_bytes(dt);
  const size_t per_warp = LSTAGES * (triplet ? 3 : 2) * row_bytes + LSTAGES * 8;
  int wmax = static_cast<int>(SMEM_BUDGET / per_warp);
  if (wmax < 1) return false;
  if (wmax > MAX_WARPS) wmax = MAX_WARPS;
  const int sms = num_sms();
  int64_t want = (B + sms - 1) / sms;  // rows per SM if every SM takes part
  int warps = static_cast<int>(want < 1 ? 1 : (want > wmax ? wmax : want));
  int64_t grid = (B + warps - 1) / warps;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  s->warps = warps;
  s->grid = static_cast<int>(grid);
  s->smem = static_cast<size_t>(warps) * per_warp;
  return true;
}

}  // namespace

size_t loss_workspace_bytes(int64_t, int32_t, irr_dtype) {
  // sync word (padded) + per-warp partials for the largest launch shape
  return 256 + static_cast<size_t>(num_sms()) * MAX_WARPS * 4 * sizeof(float);
}

irr_status loss_fwd_bwd(const LossArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (ws_bytes < loss_workspace_bytes(a.B, a.D, a.dt)) return IRR_ERR_WORKSPACE_TOO_SMALL;
  const bool triplet = a.n != nullptr;
  LaunchShape sh;
  if (!shape_for(a.B, a.D, a.dt, triplet, &sh)) return IRR_ERR_ROW_TOO_LONG;
  KParams P;
  P.q = static_cast<const uint4*>(a.q);
  P.p = static_cast<const uint4*>(a.p);
  P.n = static_cast<const uint4*>(a.n);
  P.label = a.label;
  P.label_count = a.label_count;
  P.B = a.B;
  P.vec_per_row = a.D * dtype_bytes(a.dt) / 16;
  P.kind = a.kind;
  P.m_cos = a.margin_cos;
  P.m_con = a.margin_con;
  P.pair_eps = a.pair_eps;
  P.red_scale = a.reduce_mean ? 1.0f / static_cast<float>(a.B) : 1.0f;
  for (int j = 0; j < 4; ++j) P.w[j] = a.grad_scale[j] * P.red_scale;
  P.losses = a.losses;
  P.pair_cos = a.pair_cos;
  P.row_stats = a.row_stats;
  P.dq = static_cast<uint4*>(a.dq);
  P.dp = static_cast<uint4*>(a.dp);
  P.dn = static_cast<uint4*>(a.dn);
  P.sync_word = static_cast<unsigned int*>(ws);
  P.partials = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 256);

#define IRR_LAUNCH_LOSS(BF, TR)                                                                   \
  do {                                                                                            \
    auto kern = loss_fwd_bwd_kernel<BF, TR>;                                                      \
    IRR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                      static_cast<int>(sh.smem)));                                \
    kern<<<sh.grid, sh.warps * 32, sh.smem, st>>>(P);                                             \
  } while (0)
  if (a.dt == IRR_BF16) {
    if (triplet) IRR_LAUNCH_LOSS(true, true); else IRR_LAUNCH_LOSS(true, false);
  } else {
    if (triplet) IRR_LAUNCH_LOSS(false, true); else IRR_LAUNCH_LOSS(false, false);
  }
#undef IRR_LAUNCH_LOSS
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

irr_status loss_bwd(const LossArgs& a, const float* grad_out, cudaStream_t st) {
  const bool triplet = a.n != nullptr;
  BParams P;
  P.q = static_cast<const uint4*>(a.q);
  P.p = static_cast<const uint4*>(a.p);
  P.n = static_cast<const uint4*>(a.n);
  P.label = a.label;
  P.label_count = a.label_count;
  P.B = a.B;
  P.vec_per_row = a.D * dtype_bytes(a.dt) / 16;
  P.kind = a.kind;
  P.m_cos = a.margin_cos;
  P.m_con = a.margin_con;
  P.red_scale = a.reduce_mean ? 1.0f / static_cast<float>(a.B) : 1.0f;
  P.row_stats = a.row_stats;
  P.grad_out = grad_out;
  P.dq = static_cast<uint4*>(a.dq);
  P.dp = static_cast<uint4*>(a.dp);
  P.dn = static_cast<uint4*>(a.dn);
  const int64_t want = (a.B + 7) / 8;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  const int grid = static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
  if (a.dt == IRR_BF16) {
    if (triplet) loss_bwd_kernel<true, true><<<grid, 256, 0, st>>>(P);
    else loss_bwd_kernel<true, false><<<grid, 256, 0, st>>>(P);
  } else {
    if (triplet) loss_bwd_kernel<false, true><<<grid, 256, 0, st>>>(P);
    else loss_bwd_kernel<false, false><<<grid, 256, 0, st>>>(P);
  }
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

}  // namespace irr
