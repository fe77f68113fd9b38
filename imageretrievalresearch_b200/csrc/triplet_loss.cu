// triplet_loss.cu — K2: fused forward(+backward) of the two embedding losses over (q, p, n) rows.
// Replaces, in ONE pass over the rows,
//   ContrastiveLoss.forward                 utils/contrastive_loss.py:56-61   (eps 1e-9, :34)
//   torch.nn.CosineEmbeddingLoss(margin)    train/train_efficient_cos_con_ce_loss.py:158,230-231
//   the pos/neg composition                 train/train_efficient_cos_con_ce_loss.py:230-237
//   the logged paired cosine scores         train/train_efficient_cos_con_ce_loss.py:377-382
// and their autograd backward (closed forms: SURVEY.md §A.1).
//
// HBM-bound design: every input byte crosses L2->SM exactly once.  A group of GW warps (1, 2 or 4:
// as many as give every thread about three 128-bit vectors per row) owns one row triplet at a time
// (up to 15 groups per CTA, one CTA per SM); the group's first lane stages the three rows into the
// group's shared-memory ring with 1-D bulk async copies (cp.async.bulk, completion on an mbarrier)
// up to three rows ahead of the arithmetic; each thread reduces seven sums over its vectors
// (LDS.128 + shuffles), the warps' partials meet in a double-buffered scratch line behind ONE named
// barrier per row, every warp adds them in the same fixed order and derives the four loss terms
// and the gradient coefficients itself (a few dozen scalar instructions, cheaper than a second
// barrier and a shared-memory round trip), and the group streams dq/dp/dn straight from the staged
// rows with 128-bit stores.  (One warp per row — the first version — left 1.5 warps per scheduler;
// four warps per row with three barriers — the second — spent a bf16 row's 9 KB on a latency chain:
// 0.52 of HBM.)  Each gradient is a per-row linear combination of the three rows:
//   dq = aqq*q + aqp*p + aqn*n     dp = app*p + aqp*q     dn = ann*n + aqn*q
// Loss scalars: fixed row->group assignment, per-group partials, last-CTA-done fixed-order reduction
// (deterministic; the sync word resets itself, see irr_b200.h).
#include <stdlib.h>

#include <atomic>

#include "irr_common.cuh"
#include "irr_kernels.h"

namespace irr {
namespace {

constexpr int MAX_STAGES = 4;    // rows a group keeps in flight (ring depth, chosen at launch)
constexpr int MAX_GROUPS = 15;   // groups per CTA: one named barrier each (ids 1..15)
constexpr int MAX_THREADS = 1024;
constexpr int SMEM_BUDGET = 216 * 1024;   // dynamic; ~5 KB of static shared memory sit next to it
constexpr int DYN_MAX_ROWS = 256;         // rows per CTA up to which groups take rows on demand

struct Coef {
  float aqq, aqp, aqn, app, ann;
};

// One 16-byte vector as packed fp32 pairs: the arithmetic below runs on Blackwell's packed
// FFMA2 / FMUL2 (two fp32 lanes per instruction), which halves the issue slots of a kernel whose
// limit next to HBM is instruction issue.  KIND = irr_dtype of the rows (IRR_F32 / IRR_BF16 / IRR_F16).
template <int KIND>
struct Vec {
  static constexpr int N = KIND == IRR_F32 ? 4 : 8;   // elements per vector
  static constexpr int H = N / 2;                     // float2 pairs per vector
  __device__ static __forceinline__ void unpack(const uint4& u, float2 (&f)[H]) {
    if constexpr (KIND == IRR_F32) {
      f[0] = make_float2(__uint_as_float(u.x), __uint_as_float(u.y));
      f[1] = make_float2(__uint_as_float(u.z), __uint_as_float(u.w));
    } else {
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        f[j] = KIND == IRR_F16 ? f16x2(w[j]) : make_float2(bf16lo(w[j]), bf16hi(w[j]));
    }
  }
  __device__ static __forceinline__ uint4 pack(const float2 (&f)[H]) {
    uint4 u;
    if constexpr (KIND == IRR_F32) {
      u = make_uint4(__float_as_uint(f[0].x), __float_as_uint(f[0].y), __float_as_uint(f[1].x),
                     __float_as_uint(f[1].y));
    } else {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if constexpr (KIND == IRR_F16) {
          const __half2 h = __floats2half2_rn(f[j].x, f[j].y);
          w[j] = *reinterpret_cast<const uint32_t*>(&h);
        } else {
          const __nv_bfloat162 h = __floats2bfloat162_rn(f[j].x, f[j].y);
          w[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
      }
      u = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return u;
  }
  // b - a for the contrastive distance.  fp32 / bf16 rows: the fp32 difference of the (exactly
  // widened) values.  fp16 rows: the reference runs this path under autocast (precision=16,
  // train/train_efficient_cos_con_ce_loss.py:465), where `fm2 - fm1` (utils/contrastive_loss.py:56)
  // is an fp16 subtraction — rounded to fp16 — and only the following pow / sum are widened to
  // fp32: the same rounding is applied here (one HSUB2 on the raw words), so the loss is the
  // reference's value, not a more exact one.
  __device__ static __forceinline__ void diff(const uint4& ua, const uint4& ub, const float2 (&fa)[H],
                                              const float2 (&fb)[H], float2 (&d)[H]) {
    if constexpr (KIND == IRR_F16) {
      const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        d[j] = __half22float2(__hsub2(*reinterpret_cast<const __half2*>(&wb[j]),
                                      *reinterpret_cast<const __half2*>(&wa[j])));
    } else {
#pragma unroll
      for (int j = 0; j < H; ++j) d[j] = __ffma2_rn(fa[j], make_float2(-1.0f, -1.0f), fb[j]);  // exact like a subtraction
    }
  }
};

__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// Sum eight values across the warp with 15 shuffles instead of 40: each butterfly step halves the
// number of values a lane carries (the lane keeps the half selected by its own lane bit and sends
// the other half).  Returns in lane l the warp total of value ((l>>4)&1)*4 + ((l>>3)&1)*2 + ((l>>2)&1);
// the four lanes that share l>>2 hold the same number.  Fixed order: deterministic.
__device__ __forceinline__ float warp_reduce8(const float (&v)[8], int lane) {
  float a[4], b[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = h16 ? v[i] : v[i + 4];
    const float keep = h16 ? v[i + 4] : v[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = h8 ? a[i] : a[i + 2];
    const float keep = h8 ? a[i + 2] : a[i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const float send = h4 ? b[0] : b[1];
  float r = (h4 ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, send, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

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
  float* partials;  // [gridDim.x][4], 16-byte aligned
  int hints;        // measurement knob IRR_LOSS_HINTS: 1 = evict-first loads, 2 = streaming stores
  int stages;       // ring depth per group (2..MAX_STAGES)
  int dynamic;      // groups take the CTA's rows on demand when the CTA has few of them
};

template <int KIND, bool TRIPLET, int GW>
__global__ void __launch_bounds__(MAX_THREADS, 1)
loss_fwd_bwd_kernel(const KParams P) {
  constexpr int GT = GW * 32;
  extern __shared__ __align__(128) uint8_t smem[];
  const int groups = blockDim.x / GT;
  const int grp = threadIdx.x / GT;   // row-triplet ring this thread's group owns
  const int gt = threadIdx.x % GT;    // thread within the group
  const int gwarp = gt >> 5, lane = gt & 31;
  const int stages = P.stages;
  constexpr int ROWS = TRIPLET ? 3 : 2;
  const uint32_t row_bytes = static_cast<uint32_t>(P.vec_per_row) * 16u;
  const uint32_t slot_bytes = ROWS * row_bytes;
  // [groups][stages][ROWS][row_bytes] | mbarriers [groups][stages] (padded to 16 B) |
  // scratch [groups][2][GW][8] (read as float4)
  uint8_t* my_slots = smem + static_cast<size_t>(grp) * stages * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(groups) * stages * slot_bytes);
  float* scratch = reinterpret_cast<float*>(bars + ((groups * stages + 1) & ~1)) + grp * (2 * GW * 8);
  const uint32_t bar0 = smem_u32(bars + grp * stages);

  if (gt == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(bar0 + 8u * s, 1);
    fence_mbar_init();
  }
  __syncthreads();
  // Programmatic dependent launch: everything above overlapped the tail of the previous kernel in
  // the stream; nothing below (the first global read included) runs before that kernel has
  // completed and flushed.  A no-op when the launch does not carry the attribute.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // Rows of this CTA: r_i = blockIdx + i * gridDim, i < cnt (CTA-interleaved, so every SM gets
  // B / gridDim rows give or take one: 27-28 of 4096 on each of 148 SMs).  Inside the CTA a group
  // takes item i = grp, grp + groups, ... (static) or — with few rows per CTA, where "three rows
  // for some groups, two for the others" would cost a fifth of the time — whichever item is next
  // when one of its ring slots frees up (a shared-memory ticket).  The loss terms of a row taken
  // on demand go to a per-item slot and are added in item order, so the sums do not depend on
  // which group happened to take which row: deterministic either way.
  const int64_t G = gridDim.x;
  const int cnt = blockIdx.x < P.B ? static_cast<int>((P.B - blockIdx.x + G - 1) / G) : 0;
  const bool dynamic = P.dynamic && cnt <= DYN_MAX_ROWS;
  __shared__ int next_item;
  __shared__ int slot_item[MAX_GROUPS][MAX_STAGES];
  __shared__ float4 item_loss[DYN_MAX_ROWS];
  if (threadIdx.x == 0) next_item = groups * stages;
  int* my_items = slot_item[grp];

  auto issue = [&](int64_t row, int s) {
    const uint32_t dst = smem_u32(my_slots + static_cast<size_t>(s) * slot_bytes);
    const uint32_t bar = bar0 + 8u * s;
    mbar_arrive_expect_tx(bar, slot_bytes);
    if (P.hints & 1) {
      bulk_load_1d_hint(dst, P.q + row * P.vec_per_row, row_bytes, bar, kPolicyEvictFirst);
      bulk_load_1d_hint(dst + row_bytes, P.p + row * P.vec_per_row, row_bytes, bar, kPolicyEvictFirst);
      if (TRIPLET)
        bulk_load_1d_hint(dst + 2 * row_bytes, P.n + row * P.vec_per_row, row_bytes, bar, kPolicyEvictFirst);
    } else {
      bulk_load_1d(dst, P.q + row * P.vec_per_row, row_bytes, bar);
      bulk_load_1d(dst + row_bytes, P.p + row * P.vec_per_row, row_bytes, bar);
      if (TRIPLET) bulk_load_1d(dst + 2 * row_bytes, P.n + row * P.vec_per_row, row_bytes, bar);
    }
  };

  // a ring slot's phase completes once per round: with a row's bytes, or — no item left — with a
  // bare arrive; the item number travels in shared memory, ordered by that same barrier
  auto fill = [&](int item, int s) {
    if (item < cnt) {
      my_items[s] = item;
      issue(blockIdx.x + static_cast<int64_t>(item) * G, s);
    } else {
      my_items[s] = -1;
      mbar_arrive(bar0 + 8u * s);
    }
  };
  __syncthreads();   // next_item initialised
  if (gt == 0) {
    for (int s = 0; s < stages; ++s) fill(grp + s * groups, s);
  }

  float lsum[4] = {0.f, 0.f, 0.f, 0.f};   // static assignment: kept by the group's first warp
  constexpr int VH = Vec<KIND>::H;
  int it = 0, s = 0;
  uint32_t parity = 0;
  for (;; ++it) {
    mbar_wait_parked(bar0 + 8u * s, parity, 500 + s);
    const int item = my_items[s];
    if (item < 0) break;
    const int64_t row = blockIdx.x + static_cast<int64_t>(item) * G;
    const uint4* sq = reinterpret_cast<const uint4*>(my_slots + static_cast<size_t>(s) * slot_bytes);
    const uint4* sp = sq + P.vec_per_row;
    const uint4* sn = sp + P.vec_per_row;

    // ---- seven sums: this thread's 16-byte vectors (even / odd elements in the two packed lanes),
    // then the warp (warp_reduce8), then the group's warps (every warp, same fixed order) ----
    float2 aqq = splat2(0.f), app = aqq, ann = aqq, aqp = aqq, aqn = aqq, adp = aqq, adn = aqq;
    for (int v = gt; v < P.vec_per_row; v += GT) {
      float2 fq[VH], fp[VH], fn[VH], d[VH], e[VH];
      const uint4 uq = sq[v], up = sp[v];
      Vec<KIND>::unpack(uq, fq);
      Vec<KIND>::unpack(up, fp);
      Vec<KIND>::diff(uq, up, fq, fp, d);          // p - q
      if (TRIPLET) {
        const uint4 un = sn[v];
        Vec<KIND>::unpack(un, fn);
        Vec<KIND>::diff(uq, un, fq, fn, e);        // n - q
      }
#pragma unroll
      for (int j = 0; j < VH; ++j) {
        aqq = __ffma2_rn(fq[j], fq[j], aqq);
        app = __ffma2_rn(fp[j], fp[j], app);
        aqp = __ffma2_rn(fq[j], fp[j], aqp);
        adp = __ffma2_rn(d[j], d[j], adp);
        if (TRIPLET) {
          ann = __ffma2_rn(fn[j], fn[j], ann);
          aqn = __ffma2_rn(fq[j], fn[j], aqn);
          adn = __ffma2_rn(e[j], e[j], adn);
        }
      }
    }
    // scratch line of this row (double-buffered by row parity: a warp that is already on the next
    // row must not overwrite what a slower warp of the group is still adding up)
    float* line = scratch + (it & 1) * (GW * 8);
    {
      const float part[8] = {aqq.x + aqq.y, app.x + app.y, ann.x + ann.y, aqp.x + aqp.y,
                             aqn.x + aqn.y, adp.x + adp.y, adn.x + adn.y, 0.f};
      const float r = warp_reduce8(part, lane);
      if ((lane & 3) == 0) line[gwarp * 8 + (lane >> 2)] = r;
    }
    named_bar_sync(1 + grp, GT);   // the ONE barrier per row
    // every thread of the group is past the previous row: its ring slot can be refilled
    if (gt == 0 && it > 0) {
      const int sp = s == 0 ? stages - 1 : s - 1;
      fill(dynamic ? atomicAdd(&next_item, 1) : my_items[sp] + stages * groups, sp);
    }

    // order of the eight slots = warp_reduce8's value index: qq pp nn qp | qn dp dn -
    RowOut o;
    {
      const float4* sc = reinterpret_cast<const float4*>(line);
      float4 a = sc[0], b = sc[1];
#pragma unroll
      for (int w = 1; w < GW; ++w) {
        const float4 c = sc[2 * w], d = sc[2 * w + 1];
        a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
        b.x += d.x; b.y += d.y; b.z += d.z;
      }
      const RowSums S = {a.x, a.y, a.z, a.w, b.x, b.y, b.z};
      if (TRIPLET) {
        o = triplet_row(S, P.m_cos, P.m_con, P.w);
      } else {
        const float y = __ldg(P.label + (P.label_count == 1 ? 0 : row));
        o = pair_row(S, P.kind, y, P.kind == IRR_LOSS_CONTRASTIVE ? P.m_con : P.m_cos, P.w[0]);
      }
      if (gwarp == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) lsum[j] += o.l[j];
      }
      if (gt == 0) {
        if (dynamic) item_loss[item] = make_float4(o.l[0], o.l[1], o.l[2], o.l[3]);
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
    }

    if (P.dq) {
      const float2 kqq = splat2(o.c.aqq), kqp = splat2(o.c.aqp), kqn = splat2(o.c.aqn),
                   kpp = splat2(o.c.app), knn = splat2(o.c.ann);
      uint4* gq = P.dq + row * P.vec_per_row;
      uint4* gp = P.dp + row * P.vec_per_row;
      uint4* gn = TRIPLET ? P.dn + row * P.vec_per_row : nullptr;
      for (int v = gt; v < P.vec_per_row; v += GT) {
        float2 fq[VH], fp[VH], fn[VH], oq[VH], op[VH], on[VH];
        Vec<KIND>::unpack(sq[v], fq);
        Vec<KIND>::unpack(sp[v], fp);
        if (TRIPLET) Vec<KIND>::unpack(sn[v], fn);
#pragma unroll
        for (int j = 0; j < VH; ++j) {
          float2 a = __ffma2_rn(kqq, fq[j], __fmul2_rn(kqp, fp[j]));
          op[j] = __ffma2_rn(kpp, fp[j], __fmul2_rn(kqp, fq[j]));
          if (TRIPLET) {
            a = __ffma2_rn(kqn, fn[j], a);
            on[j] = __ffma2_rn(knn, fn[j], __fmul2_rn(kqn, fq[j]));
          }
          oq[j] = a;
        }
        if (P.hints & 2) {
          __stcs(gq + v, Vec<KIND>::pack(oq));
          __stcs(gp + v, Vec<KIND>::pack(op));
          if (TRIPLET) __stcs(gn + v, Vec<KIND>::pack(on));
        } else {
          gq[v] = Vec<KIND>::pack(oq);
          gp[v] = Vec<KIND>::pack(op);
          if (TRIPLET) gn[v] = Vec<KIND>::pack(on);
        }
      }
    }
    if (++s == stages) { s = 0; parity ^= 1u; }
  }

  // the next kernel in the stream may start its prologue now (it waits for this grid to complete
  // before it touches memory, see above)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // ---- deterministic reduction of the loss scalars ----
  // group partials (fixed row -> group map) -> one partial per CTA, added in group order by the
  // CTA's first warp -> the last CTA to check in adds the CTAs' partials in a fixed order.  Only
  // warp 0 takes part after the barrier; the ticket is ONE acq_rel atomic (its release covers the
  // partial this warp just wrote, its acquire the other CTAs' partials) instead of two fences
  // around a relaxed one — this tail runs on a single SM while the other 147 idle.
  constexpr int NL = TRIPLET ? 4 : 1;
  __shared__ float4 cta_part[MAX_GROUPS];
  if (gt == 0) cta_part[grp] = make_float4(lsum[0], lsum[1], lsum[2], lsum[3]);
  __syncthreads();
  if (threadIdx.x < 32) {
    const int l = threadIdx.x;
    if (dynamic) {   // the CTA's rows in item order, whoever processed them
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = l; i < cnt; i += 32) {
        const float4 v = item_loss[i];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
      if (l == 0) reinterpret_cast<float4*>(P.partials)[blockIdx.x] = acc;
    } else if (l == 0) {
      float4 acc = cta_part[0];
      for (int g = 1; g < groups; ++g) {
        const float4 v = cta_part[g];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      reinterpret_cast<float4*>(P.partials)[blockIdx.x] = acc;
    }
    unsigned int prev = 0;
    if (l == 0)
      asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(P.sync_word) : "memory");
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev == gridDim.x - 1) {       // last CTA: every other CTA's partial is visible
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = l; i < static_cast<int>(gridDim.x); i += 32) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(P.partials) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
      if (l == 0) {
        const float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int j = 0; j < NL; ++j) P.losses[j] = r[j] * P.red_scale;
        *P.sync_word = 0u;  // self-reset for the next call on this workspace
      }
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

template <int KIND, bool TRIPLET>
__global__ void __launch_bounds__(256)
loss_bwd_kernel(const BParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t tw = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  constexpr int VH = Vec<KIND>::H;
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
    const float2 kqq = splat2(c.aqq), kqp = splat2(c.aqp), kqn = splat2(c.aqn), kpp = splat2(c.app),
                 knn = splat2(c.ann);
    for (int v = lane; v < P.vec_per_row; v += 32) {
      float2 fq[VH], fp[VH], fn[VH], oq[VH], op[VH], on[VH];
      Vec<KIND>::unpack(ldg_stream(sq + v), fq);
      Vec<KIND>::unpack(ldg_stream(sp + v), fp);
      if (TRIPLET) Vec<KIND>::unpack(ldg_stream(sn + v), fn);
#pragma unroll
      for (int j = 0; j < VH; ++j) {
        float2 a = __ffma2_rn(kqq, fq[j], __fmul2_rn(kqp, fp[j]));
        op[j] = __ffma2_rn(kpp, fp[j], __fmul2_rn(kqp, fq[j]));
        if (TRIPLET) {
          a = __ffma2_rn(kqn, fn[j], a);
          on[j] = __ffma2_rn(knn, fn[j], __fmul2_rn(kqn, fq[j]));
        }
        oq[j] = a;
      }
      gq[v] = Vec<KIND>::pack(oq);
      gp[v] = Vec<KIND>::pack(op);
      if (TRIPLET) gn[v] = Vec<KIND>::pack(on);
    }
  }
}

struct LaunchShape {
  int gw, stages, groups, grid;
  size_t smem;
};

// measurement knobs for profiles/ (environment, read once): IRR_LOSS_GW, IRR_LOSS_STAGES,
// IRR_LOSS_HINTS — not an API
struct LossKnobs { int gw, stages, hints, pdl, dynamic; };
const LossKnobs& loss_knobs() {
  static const LossKnobs k = []() {
    auto num = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
    LossKnobs r;
    r.gw = num("IRR_LOSS_GW", 0);
    r.stages = num("IRR_LOSS_STAGES", 0);
    r.hints = num("IRR_LOSS_HINTS", 1);   // evict-first loads measured +3 % (profiles/r01_notes.md)
    r.dynamic = num("IRR_LOSS_DYNAMIC", 1);   // rows on demand inside a CTA (see the kernel)
    r.pdl = num("IRR_LOSS_PDL", 1);       // programmatic dependent launch (prologue overlaps the predecessor's tail)
    return r;
  }();
  return k;
}

// Warps per row: as many as leave every thread about three 128-bit vectors per row (a bf16 row of
// 1536 elements is 192 vectors: two warps; the fp32 row 384: four).  Ring depth and groups per CTA
// from the shared-memory budget, then spread the rows over the SMs.
bool shape_for(int64_t B, int32_t D, irr_dtype dt, bool triplet, LaunchShape* s) {
  const size_t row_bytes = static_cast<size_t>(D) * dtype_bytes(dt);
  const int vec_per_row = static_cast<int>(row_bytes / 16);
  int gw = vec_per_row >= 4 * 96 ? 4 : vec_per_row >= 2 * 96 ? 2 : 1;
  if (loss_knobs().gw == 1 || loss_knobs().gw == 2 || loss_knobs().gw == 4) gw = loss_knobs().gw;
  const int sms = num_sms();
  const int64_t rows_per_sm = (B + sms - 1) / sms;  // if every SM takes part
  const int gcap = MAX_THREADS / (gw * 32) < MAX_GROUPS ? MAX_THREADS / (gw * 32) : MAX_GROUPS;
  auto per_group = [&](int stages) {
    return static_cast<size_t>(stages) * ((triplet ? 3 : 2) * row_bytes + 8) + 2 * gw * 8 * sizeof(float);
  };
  constexpr size_t BUDGET = SMEM_BUDGET - 16;
  // deepest ring that still leaves room for enough groups to keep every row of the SM's share in
  // some group's hands (few rows per SM: more groups matter more than depth)
  int stages = 3;
  if (loss_knobs().stages >= 2 && loss_knobs().stages <= MAX_STAGES) stages = loss_knobs().stages;
  while (stages > 2 && static_cast<int64_t>(BUDGET / per_group(stages)) < (rows_per_sm < gcap ? rows_per_sm : gcap))
    --stages;
  int gmax = static_cast<int>(BUDGET / per_group(stages));
  if (gmax < 1) return false;
  if (gmax > gcap) gmax = gcap;
  int groups = static_cast<int>(rows_per_sm < 1 ? 1 : (rows_per_sm > gmax ? gmax : rows_per_sm));
  int64_t grid = (B + groups - 1) / groups;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  s->gw = gw;
  s->stages = stages;
  s->groups = groups;
  s->grid = static_cast<int>(grid);
  s->smem = static_cast<size_t>(groups) * per_group(stages) + 8;   // mbarrier array padded to 16 B
  return true;
}

}  // namespace

size_t loss_workspace_bytes(int64_t, int32_t, irr_dtype) {
  // sync word (padded) + one float4 partial per group for the largest launch shape
  return 256 + static_cast<size_t>(num_sms()) * 16 * 4 * sizeof(float);
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
  P.hints = loss_knobs().hints;
  P.stages = sh.stages;
  P.dynamic = loss_knobs().dynamic;

  // the attribute is set to the budget once per instantiation and device, not per call
#define IRR_LAUNCH_LOSS(BF, TR, GWV)                                                              \
  do {                                                                                            \
    auto kern = loss_fwd_bwd_kernel<BF, TR, GWV>;                                                 \
    static std::atomic<uint64_t> attr_done{0};                                                    \
    int dev = 0;                                                                                  \
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 ||                             \
        !(attr_done.load(std::memory_order_relaxed) >> dev & 1ull)) {                             \
      IRR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                        SMEM_BUDGET));                                            \
      if (dev >= 0 && dev < 64) attr_done.fetch_or(1ull << dev, std::memory_order_relaxed);       \
    }                                                                                             \
    cudaLaunchConfig_t cfg = {};                                                                  \
    cfg.gridDim = dim3(sh.grid);                                                                  \
    cfg.blockDim = dim3(sh.groups * (GWV) * 32);                                                  \
    cfg.dynamicSmemBytes = sh.smem;                                                               \
    cfg.stream = st;                                                                              \
    cudaLaunchAttribute at[1];                                                                    \
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                \
    at[0].val.programmaticStreamSerializationAllowed = 1;                                         \
    cfg.attrs = at;                                                                               \
    cfg.numAttrs = loss_knobs().pdl ? 1 : 0;                                                      \
    IRR_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, P));                                              \
  } while (0)
#define IRR_LAUNCH_LOSS_GW(BF, TR)                                                                \
  do {                                                                                            \
    if (sh.gw == 4) IRR_LAUNCH_LOSS(BF, TR, 4);                                                   \
    else if (sh.gw == 2) IRR_LAUNCH_LOSS(BF, TR, 2);                                              \
    else IRR_LAUNCH_LOSS(BF, TR, 1);                                                              \
  } while (0)
  if (a.dt == IRR_BF16) {
    if (triplet) IRR_LAUNCH_LOSS_GW(IRR_BF16, true); else IRR_LAUNCH_LOSS_GW(IRR_BF16, false);
  } else if (a.dt == IRR_F16) {
    if (triplet) IRR_LAUNCH_LOSS_GW(IRR_F16, true); else IRR_LAUNCH_LOSS_GW(IRR_F16, false);
  } else {
    if (triplet) IRR_LAUNCH_LOSS_GW(IRR_F32, true); else IRR_LAUNCH_LOSS_GW(IRR_F32, false);
  }
#undef IRR_LAUNCH_LOSS_GW
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
    if (triplet) loss_bwd_kernel<IRR_BF16, true><<<grid, 256, 0, st>>>(P);
    else loss_bwd_kernel<IRR_BF16, false><<<grid, 256, 0, st>>>(P);
  } else if (a.dt == IRR_F16) {
    if (triplet) loss_bwd_kernel<IRR_F16, true><<<grid, 256, 0, st>>>(P);
    else loss_bwd_kernel<IRR_F16, false><<<grid, 256, 0, st>>>(P);
  } else {
    if (triplet) loss_bwd_kernel<IRR_F32, true><<<grid, 256, 0, st>>>(P);
    else loss_bwd_kernel<IRR_F32, false><<<grid, 256, 0, st>>>(P);
  }
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

}  // namespace irr
