// cosine_topk_bf16.cu — K1: query x gallery cosine contraction on tcgen05 tensor cores with a
// per-row top-k epilogue.  Replaces the reference's per-query
//     sim = cos(q[i][None], G); vals, inds = torch.topk(sim, k)
// loops (train/train_efficient_cos_con_ce_loss.py:270-281,374-392;
// inference/training_analysis.ipynb:231-251) for all query rows in one launch.
//
// Layout / roles of the single-CTA kernel (one persistent CTA per SM, 384 threads; the CTA-pair
// kernel further down is the cta_group::2 variant for more than 128 queries):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D tiles of Q (128 x 64) and G (256 x 64),
//               128-byte swizzle, into a 4-stage shared-memory ring (48 KB per stage)
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma (M=128, N=256, K=16) x4 per stage,
//               fp32 accumulators in TMEM, two accumulator stages (2 x 256 of the 512 columns);
//               M=64 for a query tile of at most 64 rows (the padding rows of a 128-row MMA cost
//               energy, and even the HBM-bound searches run at the power cap when sustained)
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: tcgen05.ld the accumulator (thread = query row, 32 columns at a time),
//               scale by 1/max(|g|,eps), keep a running sorted top-k per row in registers.
//               The Q x N scores never leave the SM.
//   warps 8-11  (single-query-tile searches) gallery-norm warps: L2-normalisation fused into the
//               load — they square-sum each gallery row from the shared-memory stages the MMA reads
// Work unit = (query tile of 128 rows, chunk of consecutive 256-row gallery tiles); units are
// numbered query-tile-fastest so that CTAs running at the same time share gallery tiles in L2.
// Every unit writes a [128, k] partial list; topk_merge.cu folds the partials, applies
// 1/max(|q|,eps) and widens the indices.
#include <cuda.h>
#include <stdlib.h>

#include <atomic>

#include "irr_common.cuh"
#include "irr_kernels.h"

namespace irr {

namespace {

constexpr int BLOCK_M = 128;   // query rows per tile  (UMMA M)
constexpr int BLOCK_N = 256;   // gallery rows per tile (UMMA N)
constexpr int BLOCK_K = 64;    // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;  // 512
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;  // 32 KB
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 128;
constexpr int NORM_WARP0 = 8;   // warps 8-11: gallery-norm warps of the fused-norm variant
constexpr int NORM_THREADS = 128;

// dynamic shared memory carve-up (base aligned to 1024 B for the 128-byte swizzle)
constexpr int SMEM_GN = ACC_STAGES * BLOCK_N * 4;                // inverse gallery norms per tile

struct Plan {
  int m_tiles, n_tiles, tiles_per_chunk, n_chunks, grid;
};

// The two tensor maps of a query batch (see encode_queries)
struct QueryMaps {
  CUtensorMap full, tail;
  int tail_tile;    // index of the ragged tile, -1 if Q is a multiple of 128
  int tail_bytes;   // bytes one k-block of the ragged tile brings in
};

// Measurement knobs for profiles/ (environment, read ONCE per process) — not an API.
struct Knobs {
  int tiles_per_chunk;       // IRR_TILES_PER_CHUNK: override the planner's chunk length
  bool no_pair, force_pair;  // IRR_NO_PAIR=1 / IRR_FORCE_PAIR=1: kernel choice
  bool producers;            // IRR_NORMS_INSIDE=0: no in-kernel norm producers
  int producers_min_pairs;   // IRR_NORMS_MIN_PAIRS: query-tile pairs from which the producers are used
  int norm_ahead;            // IRR_NORM_AHEAD: tiles the producers may run ahead (negative = unpaced)
  bool fused_pair;           // IRR_FUSED_PAIR=0: no fused norms in the pair kernel (pre-pass instead)
  bool pdl;                  // IRR_PDL=0: no programmatic dependent launches along a search's kernels
  bool m64;                  // IRR_M64=0: 128-row MMAs also for up to 64 queries
};
const Knobs& knobs() {
  static const Knobs k = []() {
    auto num = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
    auto flag = [](const char* name, char c) { const char* e = getenv(name); return e && e[0] == c; };
    Knobs r;
    r.tiles_per_chunk = num("IRR_TILES_PER_CHUNK", 0);
    r.no_pair = flag("IRR_NO_PAIR", '1');
    r.force_pair = flag("IRR_FORCE_PAIR", '1');
    r.producers = !flag("IRR_NORMS_INSIDE", '0');
    r.producers_min_pairs = num("IRR_NORMS_MIN_PAIRS", 3);
    r.norm_ahead = num("IRR_NORM_AHEAD", 1);   // one tile: same speed as two, 4.59 instead of 5.44 GB of DRAM reads
    r.fused_pair = !flag("IRR_FUSED_PAIR", '0');
    r.pdl = !flag("IRR_PDL", '0');
    r.m64 = !flag("IRR_M64", '0');
    return r;
  }();
  return k;
}

// Chunk the gallery tiles so that (query tiles x chunks) spreads evenly over the SMs.
Plan make_plan(int64_t Q, int64_t N) {
  Plan p;
  const int sms = num_sms();
  p.m_tiles = static_cast<int>((Q + BLOCK_M - 1) / BLOCK_M);
  p.n_tiles = static_cast<int>((N + BLOCK_N - 1) / BLOCK_N);
  if (p.m_tiles < 1) p.m_tiles = 1;
  if (p.n_tiles < 1) p.n_tiles = 1;
  // candidates: enough chunks that every SM gets work, few enough that partial lists stay small
  int best_tpc = 1;
  double best_cost = 1e300;
  const int max_tpc = p.n_tiles < 64 ? p.n_tiles : 64;
  for (int tpc = 1; tpc <= max_tpc; ++tpc) {
    const int chunks = (p.n_tiles + tpc - 1) / tpc;
    const long long units = 1ll * chunks * p.m_tiles;
    const long long waves = (units + sms - 1) / sms;
    // time ~ waves * tpc tiles (+ a per-unit start-up worth roughly a third of a tile)
    const double cost = static_cast<double>(waves) * (tpc + 0.35);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best_tpc = tpc;
    }
  }
  const int forced = knobs().tiles_per_chunk;
  if (forced >= 1 && forced <= p.n_tiles) best_tpc = forced;
  p.tiles_per_chunk = best_tpc;
  p.n_chunks = (p.n_tiles + best_tpc - 1) / best_tpc;
  const long long units = 1ll * p.n_chunks * p.m_tiles;
  p.grid = static_cast<int>(units < sms ? units : sms);
  return p;
}

// One accumulator tile (this thread = one query row, 256 gallery columns): TMEM -> registers 32
// columns at a time, scale by the inverse gallery norms, and fold into the row's running top-k.
// Columns arrive in increasing gallery index, so the strict '>' insert keeps the lower index on ties.
// v[j] for a runtime (warp-uniform) j without spilling the array to local memory
__device__ __forceinline__ float pick32(const float (&v)[32], int j) {
  float a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
  return (j & 16) ? d[1] : d[0];
}

// three-input maximum that propagates NaN (one FMNMX3.NAN on sm_100): a NaN score is the LARGEST
// in torch.topk's order, so the group maximum must not hide it
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float max2(float a, float b) {
  float d;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
// largest of 32 register values (NaN if any is NaN) in 17 instructions
__device__ __forceinline__ float max32(const float (&v)[32]) {
  float a[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) a[i] = max3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
  a[10] = max2(v[30], v[31]);
  const float b0 = max3(a[0], a[1], a[2]), b1 = max3(a[3], a[4], a[5]), b2 = max3(a[6], a[7], a[8]);
  return max2(max3(b0, b1, b2), max2(a[9], a[10]));
}
// "score s still enters a list whose k-th best is kth, given the shared floor": kth must be a
// number (a list full of NaNs is closed: later columns have higher indices), s larger than kth or
// NaN, and not strictly below the floor.  Written with unordered compares: no extra instructions
// for the NaN rule on the hot path.
__device__ __forceinline__ bool wants(float s, float kth, float floor) {
  return kth == kth && !(s <= kth) && !(s < floor);
}

// `floor` is a lower bound on this row's final k-th best score published by other gallery chunks
// (row_floor[], see below): anything strictly below it can be dropped without looking at the list.
template <int KMAX, bool WRITE_SCORES>
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, const float* gn, int n0, int n_valid,
                                              int row, int Q, int N, float qn,
                                              float* __restrict__ scores_out,
                                              TopKList<KMAX, int32_t>& top, float floor) {
#pragma unroll 1
  for (int c = 0; c < BLOCK_N; c += 32) {
    float v[32];
    tmem_ld_32x32(taddr + c, v);
    tmem_ld_wait();
    if (c >= n_valid) continue;  // warp-uniform
    const float4* gn4 = reinterpret_cast<const float4*>(gn + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) {   // packed FMUL2: two columns per instruction
      const float4 g4 = gn4[j];
      const float2 lo = __fmul2_rn(make_float2(v[4 * j + 0], v[4 * j + 1]), make_float2(g4.x, g4.y));
      const float2 hi = __fmul2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(g4.z, g4.w));
      v[4 * j + 0] = lo.x; v[4 * j + 1] = lo.y;
      v[4 * j + 2] = hi.x; v[4 * j + 3] = hi.y;
    }
    if (WRITE_SCORES) {
      if (row < Q) {
        float* dst = scores_out + static_cast<size_t>(row) * N + n0 + c;
        if ((N & 3) == 0 && c + 32 <= n_valid) {   // 16-byte stores (n0 and c are multiples of 32)
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) =
                make_float4(v[j] * qn, v[j + 1] * qn, v[j + 2] * qn, v[j + 3] * qn);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c + j < n_valid) dst[j] = v[j] * qn;
        }
      }
    } else {
      if (c + 32 > n_valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c + j >= n_valid) v[j] = kNegInf;
      }
      // Does ANY row of this warp still want ANY of the 32 columns?  Once the thresholds are warm
      // the answer is almost always no, and it costs 17 FMNMX3 + one vote instead of the ~100
      // instructions of the per-column masks (a column qualifies iff the row's maximum does).
      const float kth = top.v[KMAX - 1];
      const float vmax = max32(v);
      if (!__any_sync(0xffffffffu, wants(vmax, kth, floor))) continue;
      // Which columns?  Every lane builds its own 32-bit take-mask with independent compares (no
      // per-column vote/branch latency chain — there is a single epilogue warp per scheduler), one
      // REDUX ORs the masks, and the insert code runs only for the set bits.
      uint32_t mine = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) mine |= (wants(v[j], kth, floor) ? 1u : 0u) << j;
      const uint32_t any = __reduce_or_sync(0xffffffffu, mine);
      // ONE copy of the insert code (it is ~120 instructions for KMAX=16; unrolled per column it
      // would not fit the instruction cache): walk the set bits, fetching column j's score from
      // the register array with a 5-level select tree on the warp-uniform j.
      uint32_t todo = any;
#pragma unroll 1
      while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        const float sj = pick32(v, j);
        // re-evaluated: an insert earlier in this group may have raised this row's threshold
        const bool take = wants(sj, top.v[KMAX - 1], floor);
        top.insert_ranked(take, sj, n0 + c + j);
      }
    }
  }
}

// Fused-norm warps: the running square-sum of one gallery row, four partial sums — (even, odd)
// elements of the even and of the odd 16-byte chunks — so that one packed FFMA2 squares and adds
// the two elements of a 32-bit word (the unpack + FMA instructions of these warps are what the
// fused norms cost next to cached ones; packed, they are a quarter fewer).
struct RowSq {
  float2 a, b;
  __device__ __forceinline__ void reset() { a = make_float2(0.f, 0.f); b = a; }
  __device__ __forceinline__ float total() const { return (a.x + a.y) + (b.x + b.y); }
};
// One 128-byte slice (64 elements) of ONE gallery row out of a staged tile, in logical chunk order
// (position j ^ (row & 7) under the 128-byte swizzle: conflict-free and independent of where the
// row sits, so duplicate rows get bit-identical norms — in every fused-norm variant).
template <bool F16>
__device__ __forceinline__ void norm_row_sums(const uint4* r, int nt, RowSq& s) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    const uint4 u0 = r[j ^ (nt & 7)], u1 = r[(j + 1) ^ (nt & 7)];
    const uint32_t x0[4] = {u0.x, u0.y, u0.z, u0.w}, x1[4] = {u1.x, u1.y, u1.z, u1.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f0 = unpack16x2(x0[e], F16), f1 = unpack16x2(x1[e], F16);
      s.a = __ffma2_rn(f0, f0, s.a);
      s.b = __ffma2_rn(f1, f1, s.b);
    }
  }
}

// row_floor[Q]: orderable(k-th best score) each row has reached in ANY gallery chunk so far
// (0 = none yet; zeroed before every launch).  Chunks of the same rows run on other CTAs at the same
// time or earlier; reading the floor once per tile lets every chunk start with a warm threshold
// instead of re-learning it, which removes almost all insert work for larger k.  Exactness: the
// floor is the KMAX-th best of a subset, hence <= the final k-th best; only scores strictly below
// it are dropped.
__device__ __forceinline__ float read_floor(const uint32_t* row_floor, int row, int Q) {
  if (row >= Q) return kNegInf;
  const uint32_t u = __ldcg(row_floor + row);
  return u ? from_orderable(u) : kNegInf;
}
template <int KMAX>
__device__ __forceinline__ void publish_floor(uint32_t* row_floor, int row, int Q,
                                              const TopKList<KMAX, int32_t>& top, float floor) {
  if (row < Q && top.v[KMAX - 1] > floor) atomicMax(row_floor + row, orderable(top.v[KMAX - 1]));
}

// Geometry of the single-CTA kernel: 4 stages x (16 KB A + 32 KB B), two accumulator stages
// (2 x 256 TMEM columns).
struct SC {
  static constexpr int STAGES = 4;
  static constexpr int ACC = ACC_STAGES;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_STAGE_BYTES;
  static constexpr int TILES = STAGES * STAGE_BYTES;
  static constexpr int GN = ACC * BLOCK_N * 4;
  static constexpr int BARS = (2 * STAGES + 3 * ACC) * 8;
  static constexpr int ALLOC = TILES + GN + BARS + 16 + 1024;
  static_assert(ACC * BLOCK_N <= TMEM_COLS, "accumulators exceed TMEM");
  static_assert(TILES <= 196608, "stage ring exceeds the shared-memory budget");
};

// FUSE_NORM: four extra warps square-sum the gallery rows out of the SAME shared-memory stages the
// MMA reads, so the gallery crosses HBM once per search and no inverse-norm pre-pass exists (used
// when a gallery tile has a single consumer; with several query tiles the norms come from
// g_inv_norm instead).
template <int KMAX, bool WRITE_SCORES, bool FUSE_NORM>
__global__ void __launch_bounds__(NUM_THREADS, 1)
cosine_topk_bf16_kernel(const __grid_constant__ CUtensorMap tmap_q,
                        const __grid_constant__ CUtensorMap tmap_q_tail,
                        const __grid_constant__ CUtensorMap tmap_g,
                        const float* __restrict__ g_inv_norm, const float* __restrict__ q_inv_norm,
                        int Q, int N, int num_kb, int k, int m_tiles, int n_tiles,
                        int tiles_per_chunk, int n_chunks, float* __restrict__ part_val,
                        int32_t* __restrict__ part_idx, float* __restrict__ scores_out,
                        uint64_t g_policy, float eps, int tail_tile, int tail_bytes,
                        uint32_t* __restrict__ row_floor, int is_f16, int mma_m) {
  using G = SC;
  constexpr int STAGES = G::STAGES;
  constexpr int ACC = G::ACC;
  constexpr int STAGE_BYTES = G::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t bars = smem_base + G::TILES + G::GN;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + ACC + s); };
  auto gnfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + 2 * ACC + s); };
  const uint32_t tmem_slot = bars + G::BARS;
  float* gn_smem = reinterpret_cast<float*>(smem_gen + G::TILES);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + G::TILES + G::GN + G::BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_q_tail);
    tma_prefetch_desc(&tmap_g);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      // a stage is free once its MMAs retired (tcgen05.commit) and, when the norms are fused,
      // once each of the four norm warps has read it
      mbar_init(empty_bar(s), FUSE_NORM ? 1 + NORM_THREADS / 32 : 1);
    }
    for (int s = 0; s < ACC; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), EPI_THREADS / 32);
      mbar_init(gnfull_bar(s), NORM_THREADS / 32);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  // The last query tile may be ragged (Q % 128 rows): it is loaded through a tensor map whose box
  // only covers the rows that exist, rounded up to the 8-row swizzle atom — out-of-bounds rows
  // cost TMA time (Q=1 used to run slower than Q=64).  The MMA still reads 128 rows; whatever the
  // rest of the A stage holds only reaches accumulator rows >= Q, which nobody reads.
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // Programmatic dependent launch: the set-up above (barriers, TMEM, descriptor prefetch) may have
  // overlapped the tail of the previous kernel in the stream (the workspace zeroing); nothing
  // below touches global memory before that kernel has completed.  A no-op without the attribute.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int total_units = m_tiles * n_chunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int chunk = u / m_tiles, mt = u - chunk * m_tiles;
      const int t0 = chunk * tiles_per_chunk;
      const int t1 = min(t0 + tiles_per_chunk, n_tiles);
      for (int t = t0; t < t1; ++t) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 100 + stage);
          if (lane == 0) {
            const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
            const uint32_t b_dst = a_dst + G::A_BYTES;
            const bool tail = mt == tail_tile;
            mbar_arrive_expect_tx(full_bar(stage), (tail ? tail_bytes : G::A_BYTES) + B_STAGE_BYTES);
            tma_load_2d(a_dst, tail ? &tmap_q_tail : &tmap_q, kb * BLOCK_K, mt * BLOCK_M,
                        full_bar(stage), kPolicyEvictLast);
            tma_load_2d(b_dst, &tmap_g, kb * BLOCK_K, t * BLOCK_N, full_bar(stage), g_policy);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // mma_m = 64 allows 64-row MMAs for query tiles of up to 64 rows: half the multipliers of the 128-row datapath stay idle —
    // same MMA time (the kernel is HBM-bound there anyway), less energy, and a sustained loop of
    // small batches runs at the power cap (profiles/r02_notes.md)
    const uint32_t idesc_full = umma_idesc_16(BLOCK_M, BLOCK_N, is_f16 != 0);
    const uint32_t idesc_half = umma_idesc_16(64, BLOCK_N, is_f16 != 0);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t it = 0;  // accumulator tiles issued by this CTA
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int chunk = u / m_tiles;
      // a query tile with at most 64 rows (the only tile of a small batch, or a ragged last one)
      const bool half = mma_m == 64 && Q - (u - chunk * m_tiles) * BLOCK_M <= 64;
      const uint32_t idesc = half ? idesc_half : idesc_full;
      const int t0 = chunk * tiles_per_chunk;
      const int t1 = min(t0 + tiles_per_chunk, n_tiles);
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t as = it % ACC, aphase = (it / ACC) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u, 200 + as);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, 300 + stage);
          tcgen05_fence_after();
          if (lane == 0) {
            const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
            const uint64_t adesc = umma_desc_k128(a_addr);
            const uint64_t bdesc = umma_desc_k128(a_addr + G::A_BYTES);
#pragma unroll
            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
              // advance 16 elements = 32 bytes inside the swizzle row: +2 in 16-byte units
              umma_bf16_ss(tmem_d, adesc + 2u * kk, bdesc + 2u * kk, idesc,
                           (kb > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));                 // frees the smem stage when MMAs finish
            if (kb == num_kb - 1) umma_commit(tfull_bar(as));  // accumulator(s) ready
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (FUSE_NORM && warp >= NORM_WARP0) {
    // ===================== gallery-norm warps (fused-norm variant) =====================
    // thread nt owns gallery rows nt and nt+128 of the tile.  A row is 128 bytes per stage; the
    // TMA swizzle stores logical 16-byte chunk c of row r at position c ^ (r & 7).  Reading the
    // chunks in LOGICAL order (position j ^ (r & 7)) makes the accumulation order independent of
    // where a row sits in the gallery (duplicate rows get bit-identical norms), and because the
    // eight lanes of a quarter-warp own eight consecutive rows they hit eight different bank
    // groups: conflict-free LDS.128.
    const int nt = threadIdx.x - NORM_WARP0 * 32;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int chunk = u / m_tiles;
      const int t0 = chunk * tiles_per_chunk;
      const int t1 = min(t0 + tiles_per_chunk, n_tiles);
      for (int t = t0; t < t1; ++t, ++it) {
        RowSq s0, s1;
        s0.reset();
        s1.reset();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, 600 + stage);
          const uint8_t* b = smem_gen + stage * STAGE_BYTES + G::A_BYTES;
          const uint4* r0 = reinterpret_cast<const uint4*>(b + nt * 128);
          const uint4* r1 = reinterpret_cast<const uint4*>(b + (nt + NORM_THREADS) * 128);
          // one branch per stage, not one select per element: these four warps have to keep pace
          // with the HBM stream
          if (is_f16) { norm_row_sums<true>(r0, nt, s0); norm_row_sums<true>(r1, nt, s1); }
          else        { norm_row_sums<false>(r0, nt, s0); norm_row_sums<false>(r1, nt, s1); }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        const uint32_t as = it % ACC, aphase = (it / ACC) & 1u;
        // the epilogue must be done with the norms it last read from this buffer
        mbar_wait(tempty_bar(as), aphase ^ 1u, 700 + as);
        float* gn = gn_smem + as * BLOCK_N;
        gn[nt] = 1.0f / fmaxf(sqrtf(s0.total()), eps);
        gn[nt + NORM_THREADS] = 1.0f / fmaxf(sqrtf(s1.total()), eps);
        __syncwarp();
        if (lane == 0) mbar_arrive(gnfull_bar(as));
      }
    }
  } else if (warp >= EPI_WARP0 && warp < NORM_WARP0) {
    // ===================== epilogue: scale + running top-k =====================
    const int ew = warp - EPI_WARP0;          // == warp % 4: TMEM lane quarter this warp may read
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..127
    // accumulator rows -> TMEM lanes: M = 128: row r in lane r; M = 64: row r in lane
    // 32 (r / 16) + r % 16 (every warp's lane quarter holds 16 rows, its lanes 16-31 nothing)
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    TopKList<KMAX, int32_t> top;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int chunk = u / m_tiles, mt = u - chunk * m_tiles;
      const int t0 = chunk * tiles_per_chunk;
      const int t1 = min(t0 + tiles_per_chunk, n_tiles);
      const bool half = mma_m == 64 && Q - mt * BLOCK_M <= 64;   // same rule as the MMA issuer
      const int row_in_tile = half ? (lane < 16 ? ew * 16 + lane : BLOCK_M) : ew * 32 + lane;
      // (a lane without a row gets a row index past the tile, hence >= Q: a closed list)
      const int row = half && lane >= 16 ? Q : mt * BLOCK_M + row_in_tile;
      top.reset();
      // rows past the last query hold whatever the A stage held: their list starts closed (a NaN
      // k-th entry admits nothing), so they never drag their warp into the insert path
      if (row >= Q) top.v[KMAX - 1] = __uint_as_float(0x7fffffffu);
      float qn = 1.0f;
      if (WRITE_SCORES && row < Q) qn = q_inv_norm[row];
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t as = it % ACC, aphase = (it / ACC) & 1u;
        const int n0 = t * BLOCK_N;
        float* gn = gn_smem + as * BLOCK_N;
        if (FUSE_NORM) {
          mbar_wait(gnfull_bar(as), aphase, 800 + as);   // norm warps published this tile's norms
        } else {
          const int c0 = n0 + et, c1 = n0 + et + EPI_THREADS;
          gn[et] = c0 < N ? __ldg(g_inv_norm + c0) : 0.0f;
          gn[et + EPI_THREADS] = c1 < N ? __ldg(g_inv_norm + c1) : 0.0f;
          named_bar_sync(1, EPI_THREADS);
        }
        mbar_wait(tfull_bar(as), aphase, 400 + as);
        tcgen05_fence_after();
        const int n_valid = min(BLOCK_N, N - n0);
        const float floor = WRITE_SCORES ? kNegInf : read_floor(row_floor, row, Q);
        epilogue_tile<KMAX, WRITE_SCORES>(tmem_base + lane_base + as * BLOCK_N, gn, n0, n_valid, row,
                                          Q, N, qn, scores_out, top, floor);
        if (!WRITE_SCORES) publish_floor<KMAX>(row_floor, row, Q, top, floor);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(as));
      }
      if (!WRITE_SCORES && row < Q) {
        const size_t o = (static_cast<size_t>(chunk) * Q + row) * k;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          if (j < k) {
            part_val[o + j] = top.v[j];
            part_idx[o + j] = top.i[j];
          }
        }
      }
    }
  }

  // the partial-list merge may be scheduled now (it waits for this grid to complete and flush
  // before it reads anything)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2) for batches of several query tiles.
//
// Two CTAs on the SMs of one TPC form a cluster and compute a 256-query x 256-gallery-row tile:
// each CTA stages ITS 128 query rows (A) and ONE HALF (128 rows) of the gallery tile (B), the
// leader issues M=256 MMAs that read both CTAs' shared memory, and each CTA ends up with its own
// 128 x 256 accumulator in its own TMEM.  Per SM and k-block this moves 32 KB through shared
// memory instead of 48 KB (the single-CTA kernel is bound by exactly that), which also buys a
// 6-stage ring in the same 192 KB.
//   - full barriers live in the leader; both CTAs' TMA loads complete_tx on them
//   - tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs
//   - both epilogues arrive on the leader's "accumulator drained" barrier
// ---------------------------------------------------------------------------------------------
constexpr int P_STAGES = 6;
constexpr int P_B_ROWS = BLOCK_N / 2;                          // gallery rows staged per CTA
constexpr int P_A_BYTES = BLOCK_M * BLOCK_K * 2;               // 16 KB
constexpr int P_B_BYTES = P_B_ROWS * BLOCK_K * 2;              // 16 KB
constexpr int P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;           // 32 KB
constexpr int P_SMEM_TILES = P_STAGES * P_STAGE_BYTES;         // 196608
// full / empty / norm-done per stage, accumulator full / empty / norms-published per accumulator stage
constexpr int P_SMEM_BARS = (3 * P_STAGES + 3 * ACC_STAGES) * 8;
constexpr int P_SMEM_TOTAL = P_SMEM_TILES + SMEM_GN + P_SMEM_BARS + 16;
constexpr int P_SMEM_ALLOC = P_SMEM_TOTAL + 1024;
constexpr int P_THREADS = 256;
// Warps of the variants that make their own gallery norms: 2, 3, 10 and 11 — the two schedulers
// (warp id mod 4) that do NOT host the TMA producer (warp 0) and the MMA issuer (warp 1), whose
// single-thread issue loops are the latency-critical part of the CTA; warps 8 and 9 stay idle.
constexpr int P_THREADS_NORM = 384;
__device__ __forceinline__ bool is_norm_warp(int warp) { return (warp & 2) && (warp < 4 || warp >= 8); }
__device__ __forceinline__ int norm_warp_index(int warp) { return warp < 4 ? warp - 2 : warp - 8; }
// where the pair kernel's inverse gallery norms come from
constexpr int NORMS_CACHED = 0;      // g_inv_norm (the caller's Gallery cache)
constexpr int NORMS_PRODUCERS = 1;   // grid-wide in-kernel producers reading global memory (>= 3 pairs)
constexpr int NORMS_FUSED = 2;       // square-summed from the staged gallery tiles (1-2 pairs)
constexpr int OCTET = 8;                   // rows a norm warp finishes between two publications
constexpr int OCTETS_PER_TILE = BLOCK_N / OCTET;

// 128-bit streaming load that also leaves L2 first (the norm producers run far ahead of the tile
// stream: what they touch must not displace the query tiles and the tiles in flight)
__device__ __forceinline__ uint4 ldg_stream_evict_first(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(kPolicyEvictFirst));
  return r;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
static __device__ __noinline__ void wait_tile_norms_slow(const uint32_t* cnt, uint32_t need, int tile) {
  const uint64_t t0 = global_timer_ns();
  while (ld_acquire_gpu(cnt) < need) {
    if (global_timer_ns() - t0 > 4000000000ull) {
      printf("irr_b200: norm watchdog: block %d tile %d has %u of %u rows\n", (int)blockIdx.x, tile,
             ld_acquire_gpu(cnt), need);
      __trap();
    }
    __nanosleep(500);
  }
}

// One octet of gallery rows -> 1/max(|g|,eps), by one warp of the in-kernel norm producers.
// PACED producers read with the default L2 policy (the TMA loads of the same tile follow within a
// few tiles and should hit those lines); unpaced ones run far ahead and read evict-first.
template <bool F16, bool PACED>
__device__ __forceinline__ void norm_octet(const uint4* __restrict__ g_rows, int vec_per_row, int row0,
                                           int rows, float eps, float* __restrict__ norm_out, int lane) {
  // four rows at a time, four 16-byte vectors per row and lane in flight (16 independent loads
  // per lane = 8 KB per warp): the producers have to keep pace with the tile stream, which for a
  // few hundred queries consumes the gallery at several TB/s
  for (int r = 0; r < rows; r += 4) {
    const uint4* base = g_rows + static_cast<size_t>(row0 + r) * vec_per_row;
    const int nr = min(4, rows - r);
    float ss[4] = {0.f, 0.f, 0.f, 0.f};
    int v = lane;
    for (; v + 96 < vec_per_row; v += 128) {
      uint4 u[4][4];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          u[rr][c] = rr >= nr ? make_uint4(0u, 0u, 0u, 0u)
                     : PACED ? ldg_stream(base + rr * vec_per_row + v + 32 * c)
                             : ldg_stream_evict_first(base + rr * vec_per_row + v + 32 * c);
      }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t w[4] = {u[rr][c].x, u[rr][c].y, u[rr][c].z, u[rr][c].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack16x2(w[e], F16);
            ss[rr] = fmaf(f.x, f.x, ss[rr]);
            ss[rr] = fmaf(f.y, f.y, ss[rr]);
          }
        }
      }
    }
    for (; v < vec_per_row; v += 32) {
      uint4 u[4];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
        u[rr] = rr >= nr ? make_uint4(0u, 0u, 0u, 0u)
                : PACED ? ldg_stream(base + rr * vec_per_row + v)
                        : ldg_stream_evict_first(base + rr * vec_per_row + v);
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const uint32_t w[4] = {u[rr].x, u[rr].y, u[rr].z, u[rr].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack16x2(w[e], F16);
          ss[rr] = fmaf(f.x, f.x, ss[rr]);
          ss[rr] = fmaf(f.y, f.y, ss[rr]);
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) ss[rr] = warp_sum(ss[rr]);
    if (lane < nr) {
      const float mine = lane == 0 ? ss[0] : lane == 1 ? ss[1] : lane == 2 ? ss[2] : ss[3];
      norm_out[row0 + r + lane] = 1.0f / fmaxf(sqrtf(mine), eps);
    }
  }
}

// Where the inverse gallery norms of the pair kernel come from (template parameter NORMS):
//
// NORMS_FUSED (no cached norms, one or two query-tile pairs): L2-normalisation fused into the load.
// Four extra warps per CTA square-sum the CTA's half of every gallery tile out of the SAME
// shared-memory stages the MMAs read — after the stage's MMAs retired (they wait on the stage's
// "empty" barrier, which tcgen05.commit multicasts to both CTAs; the "full" barriers live in the
// leader only, and forwarding them to the partner in software measured 2x slower) and before the
// TMA producer may refill it (it also waits on the stage's "norm done" barrier) — and publish
// 1/max(|g|,eps) of
// their 128 rows into BOTH CTAs' norm buffers (own shared memory + a DSMEM store to the partner),
// arriving on both CTAs' "norms published" barrier.  The gallery crosses HBM once, nothing is
// exchanged through global memory and no CTA waits for a CTA outside its own cluster.  With many
// query-tile pairs every pair would recompute the norms of the tiles it streams (16x at Q=4096),
// which is why three pairs and up use the producers below instead.
//
// NORMS_PRODUCERS (no cached norms, three or more pairs): instead of a streaming pre-pass kernel, four extra warps per CTA
// compute 1/max(|g|,eps) for the WHOLE gallery cooperatively across the grid, straight from global
// memory, in the order the tile stream will need the tiles (the chunks that start together are
// interleaved tile by tile), publish each finished octet of rows on a per-tile counter (fence +
// atomic), and are done after about the time the pre-pass used to take — but concurrently with the
// MMAs, which leave most of the DRAM bandwidth idle.  The epilogue waits (acquire) for its tile's
// counter, which after the first few tiles is always complete already.  Producers wait for nothing
// and the grid is launched COOPERATIVELY (the driver guarantees every CTA is resident, or refuses
// the launch and the host falls back to the pre-pass), so the waits cannot deadlock; they trap on a
// 4 s watchdog like the mbarriers.  Measured: neutral at Q=4096 (the chip is at its power cap: the
// same joules take the same time wherever they are spent), a clear win for 257..2048 queries.
template <int KMAX, int NORMS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NORMS != NORMS_CACHED ? P_THREADS_NORM : P_THREADS, 1)
cosine_topk_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmap_q,
                             const __grid_constant__ CUtensorMap tmap_q_tail,
                             const __grid_constant__ CUtensorMap tmap_g,
                             const float* __restrict__ g_inv_norm, int Q, int N, int num_kb, int k,
                             int m_pairs, int n_tiles, int tiles_per_chunk, int n_chunks,
                             float* __restrict__ part_val, int32_t* __restrict__ part_idx,
                             uint32_t* __restrict__ row_floor, int is_f16,
                             const uint4* __restrict__ g_rows, int vec_per_row, float eps,
                             float* __restrict__ norm_out, uint32_t* __restrict__ tile_rows_done,
                             int norm_ahead, int tail_tile, int tail_bytes) {
  constexpr bool NORMS_INSIDE = NORMS == NORMS_PRODUCERS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t bars = smem_base + P_SMEM_TILES + SMEM_GN;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (P_STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * P_STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * P_STAGES + ACC_STAGES + s); };
  auto gnfull_bar = [&](int s) { return bars + 8u * (2 * P_STAGES + 2 * ACC_STAGES + s); };
  auto normdone_bar = [&](int s) { return bars + 8u * (2 * P_STAGES + 3 * ACC_STAGES + s); };
  const uint32_t tmem_slot = bars + P_SMEM_BARS;
  float* gn_smem = reinterpret_cast<float*>(smem_gen + P_SMEM_TILES);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + P_SMEM_TILES + SMEM_GN + P_SMEM_BARS);

  // NORMS_INSIDE: this CTA's position on the tile stream's timeline (wave * tiles_per_chunk + tile
  // in chunk), published by the epilogue, read by the CTA's norm producers to pace themselves
  volatile int* stream_pos = reinterpret_cast<volatile int*>(smem_gen + P_SMEM_TILES + SMEM_GN + P_SMEM_BARS + 8);
  if (NORMS_INSIDE && threadIdx.x == 0) *stream_pos = 0;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();   // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_q_tail);
    tma_prefetch_desc(&tmap_g);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(full_bar(s), 1);    // leader's own arrive.expect_tx; bytes come from both CTAs
      mbar_init(empty_bar(s), 1);   // one multicast tcgen05.commit
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(tfull_bar(s), 1);                        // one multicast tcgen05.commit
      mbar_init(tempty_bar(s), 2 * (EPI_THREADS / 32));  // 4 epilogue warps of each CTA (leader's copy)
      mbar_init(gnfull_bar(s), 2 * (NORM_THREADS / 32)); // NORMS_FUSED: 4 norm warps of each CTA
    }
    for (int s = 0; s < P_STAGES; ++s) mbar_init(normdone_bar(s), NORM_THREADS / 32);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  cluster_sync_all();   // barrier inits + TMEM allocation visible to both CTAs
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int total_units = m_pairs * n_chunks;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    // bytes a query tile contributes to a stage: a whole tile, the ragged last tile's box (rows
    // that exist, rounded up to the swizzle atom), or nothing for a tile past the last query — an
    // odd number of query tiles leaves the last pair's second CTA without rows: no load at all,
    // its MMA half multiplies whatever the stage holds into accumulator rows nobody reads
    auto a_bytes = [&](int tile) {
      return tile * BLOCK_M >= Q ? 0 : (tile == tail_tile ? tail_bytes : P_A_BYTES);
    };
    for (int u = cluster_id; u < total_units; u += num_clusters) {
      const int chunk = u / m_pairs, mp = u - chunk * m_pairs;
      const int mt = mp * 2 + static_cast<int>(rank);
      const int t0 = chunk * tiles_per_chunk;
      const int t1 = min(t0 + tiles_per_chunk, n_tiles);
      const uint32_t stage_tx = a_bytes(mp * 2) + a_bytes(mp * 2 + 1) + 2 * P_B_BYTES;
      const int my_a = a_bytes(mt);
      for (int t = t0; t < t1; ++t) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 1100 + stage);
          // fused norms: this CTA's norm warps have square-summed the stage's previous contents
          if (NORMS == NORMS_FUSED) mbar_wait(normdone_bar(stage), phase ^ 1u, 1150 + stage);
          if (lane == 0) {
            const uint32_t a_dst = smem_base + stage * P_STAGE_BYTES;
            const uint32_t b_dst = a_dst + P_A_BYTES;
            const uint32_t lead_full = mapa_rank(full_bar(stage), 0);
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), stage_tx);
            if (my_a)
              tma_load_2d_pair(a_dst, mt == tail_tile ? &tmap_q_tail : &tmap_q, kb * BLOCK_K,
                               mt * BLOCK_M, lead_full, kPolicyEvictLast);
            tma_load_2d_pair(b_dst, &tmap_g, kb * BLOCK_K, t * BLOCK_N + static_cast<int>(rank) * P_B_ROWS,
                             lead_full, kPolicyEvictNormal);
          }
          __syncwarp();
          if (++stage == P_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_16(2 * BLOCK_M, BLOCK_N, is_f16 != 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int u = cluster_id; u < total_units; u += num_clusters) {
        const int chunk = u / m_pairs;
        const int t0 = chunk * tiles_per_chunk;
        const int t1 = min(t0 + tiles_per_chunk, n_tiles);
        for (int t = t0; t < t1; ++t, ++it) {
          const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
          mbar_wait(tempty_bar(as), aphase ^ 1u, 1200 + as);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + as * BLOCK_N;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(full_bar(stage), phase, 1300 + stage);
            tcgen05_fence_after();
            if (lane == 0) {
              const uint32_t a_addr = smem_base + stage * P_STAGE_BYTES;
              const uint64_t adesc = umma_desc_k128(a_addr);
              const uint64_t bdesc = umma_desc_k128(a_addr + P_A_BYTES);
#pragma unroll
              for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
                umma_bf16_ss_pair(tmem_d, adesc + 2u * kk, bdesc + 2u * kk, idesc,
                                  (kb > 0 || kk > 0) ? 1u : 0u);
              umma_commit_pair(empty_bar(stage), 0x3);
              if (kb == num_kb - 1) umma_commit_pair(tfull_bar(as), 0x3);
            }
            __syncwarp();
            if (++stage == P_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (NORMS == NORMS_FUSED && is_norm_warp(warp)) {
    // ===================== fused gallery norms (this CTA's half of every tile) ===============
    // thread nt owns gallery row rank*128 + nt of the tile: one 128-byte slice per stage, read in
    // logical chunk order (conflict-free, position independent — see the single-CTA kernel)
    const int nt = norm_warp_index(warp) * 32 + lane;
    const uint32_t peer = rank ^ 1u;
    const bool f16 = is_f16 != 0;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t it = 0;
    for (int u = cluster_id; u < total_units; u += num_clusters) {
      const int chunk = u / m_pairs;
      const int t0 = chunk * tiles_per_chunk;
      const int t1 = min(t0 + tiles_per_chunk, n_tiles);
      for (int t = t0; t < t1; ++t, ++it) {
        RowSq sq;
        sq.reset();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase, 1500 + stage);   // the stage's MMAs have retired
          const uint4* r = reinterpret_cast<const uint4*>(smem_gen + stage * P_STAGE_BYTES + P_A_BYTES + nt * 128);
          if (f16) norm_row_sums<true>(r, nt, sq);
          else     norm_row_sums<false>(r, nt, sq);
          __syncwarp();
          if (lane == 0) mbar_arrive(normdone_bar(stage));   // the producer may refill the stage
          if (++stage == P_STAGES) { stage = 0; phase ^= 1u; }
        }
        // gn[as] (in BOTH CTAs) is free once both epilogues drained the tile that used accumulator
        // stage `as` and its norms before: in this variant they arrive on both CTAs' copies of the
        // "accumulator drained" barrier (the norm warps may be several short tiles ahead of the MMAs)
        const uint32_t as = it & 1u;
        mbar_wait_cluster(tempty_bar(as), ((it >> 1) & 1u) ^ 1u, 1700 + as);
        const float inv = 1.0f / fmaxf(sqrtf(sq.total()), eps);
        float* mine = gn_smem + as * BLOCK_N + static_cast<int>(rank) * P_B_ROWS + nt;
        *mine = inv;
        st_shared_cluster_f32(mapa_rank(smem_u32(mine), peer), inv);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(mapa_rank(gnfull_bar(as), 0));
          mbar_arrive_cluster(mapa_rank(gnfull_bar(as), 1));
        }
      }
    }
  } else if (NORMS_INSIDE && is_norm_warp(warp)) {
    // ===================== gallery-norm producers (whole grid, need order) =====================
    const int nw = static_cast<int>(blockIdx.x) * 4 + norm_warp_index(warp);
    const int NW = static_cast<int>(gridDim.x) * 4;
    // Slots (one octet of rows each) in the order the tile stream needs them: wave by wave (a wave
    // = one unit per cluster), inside a wave tile by tile, inside a tile position the chunks whose
    // first unit runs in that wave.  key = wave * tiles_per_chunk + tile-in-chunk is the slot's
    // place on the same timeline as stream_pos; a producer does not run more than norm_ahead tiles
    // ahead of its own CTA's epilogue, so that its DRAM reads are the ones the TMA loads of the
    // same tile then find in L2 (one DRAM read of the gallery serves both) — all CTAs move along
    // the timeline at the same pace, and the slowest CTA's needs are never held back (its key is
    // <= everybody's position), so pacing cannot deadlock.
    const long long slots_per_chunk = 1ll * tiles_per_chunk * OCTETS_PER_TILE;
    const long long total_slots = slots_per_chunk * n_chunks;
    const bool f16 = is_f16 != 0;
    const bool paced = norm_ahead >= 0;
    for (long long o = nw; o < total_slots; o += NW) {
      const long long cq = o / slots_per_chunk;
      const int w = static_cast<int>(cq * m_pairs / num_clusters);
      const int c_lo = static_cast<int>((1ll * w * num_clusters + m_pairs - 1) / m_pairs);
      const int c_hi = min(static_cast<int>((1ll * (w + 1) * num_clusters + m_pairs - 1) / m_pairs), n_chunks);
      const int nch = c_hi - c_lo;
      const int rem = static_cast<int>(o - c_lo * slots_per_chunk);
      const int j = rem / (nch * OCTETS_PER_TILE);
      const int r2 = rem - j * (nch * OCTETS_PER_TILE);
      const int chunk = c_lo + r2 / OCTETS_PER_TILE;
      const int oct = r2 % OCTETS_PER_TILE;
      const int tile = chunk * tiles_per_chunk + j;
      if (paced) {
        // The CTA's four producer warps take four consecutive slots per round — same tile position,
        // same key — so ONE of them polls the position and the others wait for it at a named
        // barrier, which costs no issue slots (polled by all four, this loop alone was a third of
        // the kernel's executed instructions: nanosleep returns after ~80 ns whatever it is asked).
        const int key = w * tiles_per_chunk + j;
        if (norm_warp_index(warp) == 0)
          while (key > *stream_pos + norm_ahead) __nanosleep(1500);
        named_bar_sync(2, NORM_THREADS);
      }
      if (tile >= n_tiles) continue;
      const int row0 = tile * BLOCK_N + oct * OCTET;
      if (row0 >= N) continue;
      const int rows = min(OCTET, N - row0);
      if (f16) { if (paced) norm_octet<true, true>(g_rows, vec_per_row, row0, rows, eps, norm_out, lane);
                 else       norm_octet<true, false>(g_rows, vec_per_row, row0, rows, eps, norm_out, lane); }
      else     { if (paced) norm_octet<false, true>(g_rows, vec_per_row, row0, rows, eps, norm_out, lane);
                 else       norm_octet<false, false>(g_rows, vec_per_row, row0, rows, eps, norm_out, lane); }
      __syncwarp();
      if (lane == 0) {
        __threadfence();                                  // the octet's norms before its count
        atomicAdd(tile_rows_done + tile, static_cast<uint32_t>(rows));
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4) {
    // ===================== epilogue (both CTAs: own 128 query rows x 256 columns) ==========
    const int ew = warp - EPI_WARP0;
    const int et = threadIdx.x - EPI_WARP0 * 32;
    const int row_in_tile = ew * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    TopKList<KMAX, int32_t> top;
    uint32_t it = 0;
    for (int u = cluster_id; u < total_units; u += num_clusters) {
      const int chunk = u / m_pairs, mp = u - chunk * m_pairs;
      const int mt = mp * 2 + static_cast<int>(rank);
      const int t0 = chunk * tiles_per_chunk;
      const int t1 = min(t0 + tiles_per_chunk, n_tiles);
      const int row = mt * BLOCK_M + row_in_tile;
      top.reset();
      if (row >= Q) top.v[KMAX - 1] = __uint_as_float(0x7fffffffu);   // closed list (see above)
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
        const int n0 = t * BLOCK_N;
        float* gn = gn_smem + as * BLOCK_N;
        // One epilogue warp waits on the mbarriers (norms published, accumulator ready); the other
        // three wait for it at the named barrier below, where a waiting warp issues nothing — the
        // epilogue idles ~80 % of a tile's time, and four warps polling was a sixth of the kernel's
        // executed instructions.
        if (NORMS == NORMS_FUSED) {
          if (ew == 0) mbar_wait_cluster(gnfull_bar(as), aphase, 1600 + as);   // both halves' norms have landed
        } else {
          const int c0 = n0 + et, c1 = n0 + et + EPI_THREADS;
          if (NORMS_INSIDE) {
            if (et == 0) *stream_pos = (u / num_clusters) * tiles_per_chunk + (t - t0);
            // producers publish rows in octets; the tile is usable once all of its rows are counted
            if (lane == 0) {
              const uint32_t need = static_cast<uint32_t>(min(BLOCK_N, N - n0));
              if (ld_acquire_gpu(tile_rows_done + t) < need) wait_tile_norms_slow(tile_rows_done + t, need, t);
            }
            __syncwarp();
            gn[et] = c0 < N ? __ldcg(g_inv_norm + c0) : 0.0f;
            gn[et + EPI_THREADS] = c1 < N ? __ldcg(g_inv_norm + c1) : 0.0f;
          } else {
            gn[et] = c0 < N ? __ldg(g_inv_norm + c0) : 0.0f;
            gn[et + EPI_THREADS] = c1 < N ? __ldg(g_inv_norm + c1) : 0.0f;
          }
        }
        if (ew == 0) mbar_wait(tfull_bar(as), aphase, 1400 + as);
        named_bar_sync(1, EPI_THREADS);
        tcgen05_fence_after();
        const int n_valid = min(BLOCK_N, N - n0);
        const float floor = read_floor(row_floor, row, Q);
        epilogue_tile<KMAX, false>(tmem_base + lane_base + as * BLOCK_N, gn, n0, n_valid, row, Q, N,
                                   1.0f, nullptr, top, floor);
        publish_floor<KMAX>(row_floor, row, Q, top, floor);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(mapa_rank(tempty_bar(as), 0));
          if (NORMS == NORMS_FUSED) mbar_arrive_cluster(mapa_rank(tempty_bar(as), 1));
        }
      }
      if (row < Q) {
        const size_t o = (static_cast<size_t>(chunk) * Q + row) * k;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          if (j < k) {
            part_val[o + j] = top.v[j];
            part_idx[o + j] = top.i[j];
          }
        }
      }
    }
    // done with the tile stream: this CTA's producers finish whatever is left unpaced
    if (NORMS_INSIDE && et == 0) *stream_pos = 0x3fffffff;
  }

  // neither CTA may leave (or free TMEM) while its partner can still read its shared memory,
  // multicast to its barriers or arrive remotely
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_pair<TMEM_COLS>(tmem_base);
  }
}

__global__ void __launch_bounds__(256) zero_words_kernel(uint32_t* __restrict__ p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0u;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// chunking for the pair kernel: units = (query-tile pair, gallery chunk) over sms/2 clusters
// The m_pairs clusters that walk a chunk together fetch each of its gallery tiles from DRAM once
// (the other m_pairs - 1 reads hit L2).  clusters / m_pairs is not whole, so one chunk per wave
// straddles into the next wave, where its late units stream it AGAIN — one wave after the first
// pass, and a wave of long chunks moves more bytes than L2 holds: at 4096 queries (16 pairs on 74
// clusters, 4.6 chunks per wave of 53 tiles = 193 MB) every fifth chunk was read twice, 3.77 GB of
// DRAM reads for 3.08 GB of gallery.  (a) From 12 pairs on (a straddler is >= 15 % of a wave) the
// chunks are therefore kept short enough that a wave's tiles stay in L2 until the late units
// come by: 8 tiles at 4096 x 1536 -> 3.12 GB, same kernel time (profiles/r02_notes.md).  (b) Only
// for k <= 4: every unit restarts its rows' lists, which is free for three entries and was
// measured 14 % slower for config 5's k = 10 (16-entry lists, 611 instead of 111 chunks).
constexpr size_t WAVE_L2_BUDGET = 32u << 20;   // a quarter of the 126 MB L2
Plan make_plan_pair(int64_t Q, int64_t N, int32_t D, int32_t k) {
  Plan p;
  const int clusters = num_sms() / 2;
  p.m_tiles = static_cast<int>((Q + 2 * BLOCK_M - 1) / (2 * BLOCK_M));  // pairs of query tiles
  p.n_tiles = static_cast<int>((N + BLOCK_N - 1) / BLOCK_N);
  if (p.n_tiles < 1) p.n_tiles = 1;
  int best_tpc = 1;
  double best_cost = 1e300;
  int max_tpc = p.n_tiles < 64 ? p.n_tiles : 64;
  // (c) only for launches long enough to run at the power cap (about 5 ms of tile stream and up):
  // the DRAM reads saved are energy, not time — uncapped (a 1 ms launch on one of 8 GPUs) the
  // extra units cost 1.5 % and save nothing.
  const bool long_launch = 1ll * p.n_tiles * p.m_tiles * D >= 512ll * clusters * 1536;
  if (p.m_tiles >= 12 && k <= 4 && long_launch) {
    const size_t chunks_per_wave = static_cast<size_t>((clusters + p.m_tiles - 1) / p.m_tiles);
    const size_t tile_bytes = static_cast<size_t>(BLOCK_N) * static_cast<size_t>(D > 0 ? D : 1) * 2;
    const int cap = static_cast<int>(WAVE_L2_BUDGET / (chunks_per_wave * tile_bytes));
    max_tpc = std::min(max_tpc, std::max(cap, 4));
  }
  for (int tpc = 1; tpc <= max_tpc; ++tpc) {
    const int chunks = (p.n_tiles + tpc - 1) / tpc;
    const long long units = 1ll * chunks * p.m_tiles;
    const long long waves = (units + clusters - 1) / clusters;
    const double cost = static_cast<double>(waves) * (tpc + 0.35);
    if (cost < best_cost - 1e-9) { best_cost = cost; best_tpc = tpc; }
  }
  const int forced = knobs().tiles_per_chunk;
  if (forced >= 1 && forced <= p.n_tiles) best_tpc = forced;
  p.tiles_per_chunk = best_tpc;
  p.n_chunks = (p.n_tiles + best_tpc - 1) / best_tpc;
  const long long units = 1ll * p.n_chunks * p.m_tiles;
  p.grid = 2 * static_cast<int>(units < clusters ? units : clusters);
  return p;
}

// Which kernel serves a batch of Q queries (measured on B200 at N=1M, D=1536, see profiles/):
//   Q <= 128          single-CTA kernel (norm warps fused into the tile stream when uncached)
//   129..256          CTA-pair kernel: one pair streams the gallery once
//   257..384 cached   single-CTA kernel, three query tiles (a second pair would multiply 128 rows
//                     of zeros)
//   257..384 uncached CTA-pair kernel with fused norms (beats the norm pre-pass + three tiles)
//   >= 385            CTA-pair kernel (it moves a third less data per flop)
bool use_pair(int64_t Q, bool cached_norms) {
  if (Q <= BLOCK_M) return false;
  if (knobs().no_pair) return false;
  if (knobs().force_pair) return true;
  if (Q > 2 * BLOCK_M && Q <= 3 * BLOCK_M) return !cached_norms && knobs().fused_pair;
  return true;
}

// norm source of an uncached pair launch: grid-wide producers from three query-tile pairs on
// (every pair recomputing the norms of the tiles it streams would cost 16x the arithmetic at
// Q=4096), norms fused into the tile stream below (the tile stream outruns four producer warps per
// CTA there: measured 512 queries 1.69 vs 1.66 ms with the pre-pass, 768 queries 2.14 vs 2.54 ms)
int pair_norm_mode(bool cached, int m_pairs) {
  if (cached) return NORMS_CACHED;
  if (knobs().producers && m_pairs >= knobs().producers_min_pairs) return NORMS_PRODUCERS;
  return knobs().fused_pair ? NORMS_FUSED : NORMS_CACHED;   // CACHED here = streaming pre-pass first
}

// NORMS_PRODUCERS: gin is the (not yet filled) fp32[N] buffer the in-kernel producers write and the
// epilogues read; tile_done is the zeroed per-tile row counter array.  That variant's epilogues
// wait for producers in OTHER CTAs, so it is launched cooperatively: the driver either makes the
// whole grid resident (also under MPS partitions, green contexts or with SMs held by kernels of
// other streams) or refuses the launch — reported to the caller as *refused = true, which then
// takes the pre-pass instead.  The other variants only wait inside their own cluster.
template <int KMAX, int NORMS>
irr_status launch_pair(const CUtensorMap& tq, const CUtensorMap& tg, const float* gin, int64_t Q,
                       int64_t N, int32_t D, int32_t k, const Plan& p, float* pv, int32_t* pi,
                       uint32_t* row_floor, bool f16, const void* g, float eps, uint32_t* tile_done,
                       cudaStream_t st, bool* refused, const QueryMaps& qm) {
  const CUtensorMap& tqt = qm.tail;
  auto kern = cosine_topk_bf16_pair_kernel<KMAX, NORMS>;
  static std::atomic<uint64_t> attr_done{0};
  if (attr_needed(attr_done)) {
    IRR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_ALLOC));
    attr_set(attr_done);
  }
  const int num_kb = (D + BLOCK_K - 1) / BLOCK_K;
  const int threads = NORMS != NORMS_CACHED ? P_THREADS_NORM : P_THREADS;
  if (NORMS == NORMS_PRODUCERS) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = P_SMEM_ALLOC;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    profile_mark_start(st);
    const cudaError_t e = cudaLaunchKernelEx(
        &cfg, kern, tq, tqt, tg, gin, static_cast<int>(Q), static_cast<int>(N), num_kb, static_cast<int>(k),
        p.m_tiles, p.n_tiles, p.tiles_per_chunk, p.n_chunks, pv, pi, row_floor, f16 ? 1 : 0,
        static_cast<const uint4*>(g), D * 2 / 16, eps, const_cast<float*>(gin), tile_done,
        knobs().norm_ahead, qm.tail_tile, qm.tail_bytes);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources ||
        e == cudaErrorNotSupported) {
      cudaGetLastError();
      profile_mark_stop(st);
      *refused = true;
      return IRR_OK;
    }
    profile_mark_stop(st);
    if (e != cudaSuccess) return static_cast<irr_status>(static_cast<int>(e));
    return IRR_OK;
  }
  profile_mark_start(st);
  kern<<<p.grid, threads, P_SMEM_ALLOC, st>>>(
      tq, tqt, tg, gin, static_cast<int>(Q), static_cast<int>(N), num_kb, k, p.m_tiles, p.n_tiles,
      p.tiles_per_chunk, p.n_chunks, pv, pi, row_floor, f16 ? 1 : 0, static_cast<const uint4*>(g),
      D * 2 / 16, eps, const_cast<float*>(gin), tile_done, knobs().norm_ahead, qm.tail_tile,
      qm.tail_bytes);
  profile_mark_stop(st);
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
            cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// [rows, cols] bf16 / fp16 row-major, tile = box_rows x 64 columns, 128-byte swizzle, zero fill OOB.
// A tensor map is a pure function of (base, rows, cols, box, dtype): the last few are kept per host
// thread, so a resident gallery searched again and again (and a static query buffer replayed from a
// serving loop) does not pay the driver's encode call per search.
struct TmapKey {
  const void* base;
  int64_t rows, cols;
  uint32_t box_rows;
  bool f16;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && box_rows == o.box_rows && f16 == o.f16;
  }
};
bool encode_bf16_rows(CUtensorMap* m, const void* base, int64_t rows, int64_t cols,
                      uint32_t box_rows, bool f16 = false) {
  constexpr int SLOTS = 8;
  struct Slot { TmapKey key; CUtensorMap map; bool used; };
  thread_local Slot cache[SLOTS] = {};
  thread_local int next = 0;
  const TmapKey key = {base, rows, cols, box_rows, f16};
  for (int i = 0; i < SLOTS; ++i)
    if (cache[i].used && cache[i].key == key) { *m = cache[i].map; return true; }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  if (fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
         const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  Slot& s = cache[next];
  next = (next + 1) % SLOTS;
  s.key = key; s.map = *m; s.used = true;
  return true;
}

// The two tensor maps of a query batch: whole 128-row tiles, and the ragged last tile (Q % 128
// rows) whose box only covers the rows that exist, rounded up to the 8-row swizzle atom —
// out-of-bounds rows cost TMA time.
bool encode_queries(QueryMaps* m, const void* q, int64_t Q, int64_t D, bool f16) {
  const int rem = static_cast<int>(Q % BLOCK_M);
  const int tail_rows = rem ? (rem + 7) / 8 * 8 : BLOCK_M;
  m->tail_tile = rem ? static_cast<int>(Q / BLOCK_M) : -1;
  m->tail_bytes = tail_rows * BLOCK_K * 2;
  if (!encode_bf16_rows(&m->full, q, Q, D, BLOCK_M, f16)) return false;
  if (rem) return encode_bf16_rows(&m->tail, q, Q, D, tail_rows, f16);
  m->tail = m->full;
  return true;
}

template <int KMAX, bool WS, bool FN>
irr_status launch(const QueryMaps& qm, const CUtensorMap& tg, const float* gin, const float* qin,
                  int64_t Q, int64_t N, int32_t D, int32_t k, const Plan& p, float* pv, int32_t* pi,
                  float* scores, float eps, uint32_t* row_floor, bool f16, cudaStream_t st) {
  auto kern = cosine_topk_bf16_kernel<KMAX, WS, FN>;
  constexpr int SMEM_ALLOC = SC::ALLOC;
  static std::atomic<uint64_t> attr_done{0};
  if (attr_needed(attr_done)) {
    IRR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
    attr_set(attr_done);
  }
  const int num_kb = (D + BLOCK_K - 1) / BLOCK_K;
  // a gallery streamed by a single query tile is read exactly once: do not let it displace the
  // query tiles in L2; with several query tiles the gallery tiles are the L2-shared operand
  const uint64_t g_policy = p.m_tiles == 1 ? kPolicyEvictFirst : kPolicyEvictNormal;
  // programmatic dependent launch: this kernel's set-up overlaps the workspace-zeroing kernel
  // before it, and the partial-list merge after it is scheduled while this one drains
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_ALLOC;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (!WS && knobs().pdl) ? 1 : 0;
  if (!WS) profile_mark_start(st);
  const cudaError_t e = cudaLaunchKernelEx(
      &cfg, kern, qm.full, qm.tail, tg, gin, qin, static_cast<int>(Q), static_cast<int>(N), num_kb,
      static_cast<int>(k), p.m_tiles, p.n_tiles, p.tiles_per_chunk, p.n_chunks, pv, pi, scores,
      g_policy, eps, qm.tail_tile, qm.tail_bytes, row_floor, f16 ? 1 : 0,
      knobs().m64 ? 64 : BLOCK_M);
  if (!WS) profile_mark_stop(st);
  if (e != cudaSuccess) return static_cast<irr_status>(static_cast<int>(e));
  return IRR_OK;
}

}  // namespace

// workspace: [g_inv_norm fp32 N][part_val fp32 chunks*Q*k][part_idx i32 chunks*Q*k]
//            [row_floor u32 Q | tile_rows_done u32 n_tiles]   (the last two are zeroed per call)
size_t bf16_topk_workspace_bytes(int64_t Q, int64_t N, int32_t D, int32_t k) {
  // the caller may or may not pass cached norms: size for the larger of the two plans
  size_t parts = 0;
  for (int cached = 0; cached < 2; ++cached) {
    const Plan p = use_pair(Q, cached) ? make_plan_pair(Q, N, D, k) : make_plan(Q, N);
    const size_t n = static_cast<size_t>(p.n_chunks) * Q * k;
    if (n > parts) parts = n;
  }
  const size_t n_tiles = static_cast<size_t>((N + BLOCK_N - 1) / BLOCK_N);
  return align_up(static_cast<size_t>(N) * 4, 256) + align_up(parts * 4, 256) * 2 +
         align_up(static_cast<size_t>(Q) * 4, 256) + align_up(n_tiles * 4, 256) + 256;
}

irr_status bf16_cosine_topk(const void* q, const void* g, const float* g_inv_norm, int64_t Q,
                            int64_t N, int32_t D, int32_t k, float eps, int64_t idx_offset,
                            float* out_val, int64_t* out_idx, void* ws, size_t ws_bytes,
                            cudaStream_t st, irr_dtype dt) {
  if (device_cc() / 10 != 10) return IRR_ERR_UNSUPPORTED_DEVICE;
  const bool f16 = dt == IRR_F16;
  if (N > 0x7fffff00ll || Q > 0x7fffff00ll) return IRR_ERR_INVALID_ARG;
  if (ws_bytes < bf16_topk_workspace_bytes(Q, N, D, k)) return IRR_ERR_WORKSPACE_TOO_SMALL;
  const bool cached = g_inv_norm != nullptr;
  const bool pair = use_pair(Q, cached);
  const Plan p = pair ? make_plan_pair(Q, N, D, k) : make_plan(Q, N);
  uint8_t* w = static_cast<uint8_t*>(ws);
  float* gin_ws = reinterpret_cast<float*>(w);
  w += align_up(static_cast<size_t>(N) * 4, 256);
  const size_t parts = static_cast<size_t>(p.n_chunks) * Q * k;
  float* pv = reinterpret_cast<float*>(w);
  w += align_up(parts * 4, 256);
  int32_t* pi = reinterpret_cast<int32_t*>(w);
  w += align_up(parts * 4, 256);
  uint32_t* row_floor = reinterpret_cast<uint32_t*>(w);
  w += align_up(static_cast<size_t>(Q) * 4, 256);
  uint32_t* tile_done = reinterpret_cast<uint32_t*>(w);
  int mode = pair ? pair_norm_mode(cached, p.m_tiles) : NORMS_CACHED;
  // zero the rows' shared floors and (if used) the per-tile norm counters: a memset, or — in front
  // of the single-CTA kernel, whose set-up then overlaps it — a kernel
  const size_t zero_bytes =
      mode == NORMS_PRODUCERS ? align_up(static_cast<size_t>(Q) * 4, 256) + static_cast<size_t>(p.n_tiles) * 4
                              : static_cast<size_t>(Q) * 4;
  if (!pair && knobs().pdl) {
    zero_words_kernel<<<static_cast<unsigned>((zero_bytes / 4 + 255) / 256), 256, 0, st>>>(
        row_floor, static_cast<int>(zero_bytes / 4));
    IRR_LAUNCH_CHECK();
  } else {
    IRR_CUDA_TRY(cudaMemsetAsync(row_floor, 0, zero_bytes, st));
  }

  QueryMaps qm;
  CUtensorMap tg;
  if (!encode_queries(&qm, q, Q, D, f16) ||
      !encode_bf16_rows(&tg, g, N, D, pair ? P_B_ROWS : BLOCK_N, f16))
    return IRR_ERR_UNSUPPORTED_DEVICE;
  const CUtensorMap& tq = qm.full;
  irr_status s = IRR_OK;
  if (pair) {
    const bool small_k = k <= 4;
    bool refused = false;
#define IRR_LAUNCH_PAIR(KM, NM, GIN) \
  s = launch_pair<KM, NM>(tq, tg, GIN, Q, N, D, k, p, pv, pi, row_floor, f16, g, eps, tile_done, st, &refused, qm)
    if (mode == NORMS_PRODUCERS) {   // gin_ws is written by the kernel's own norm producers
      if (small_k) IRR_LAUNCH_PAIR(4, NORMS_PRODUCERS, gin_ws); else IRR_LAUNCH_PAIR(16, NORMS_PRODUCERS, gin_ws);
      if (s == IRR_OK && refused) mode = NORMS_CACHED;   // no co-resident grid: pre-pass instead
    } else if (mode == NORMS_FUSED) {
      if (small_k) IRR_LAUNCH_PAIR(4, NORMS_FUSED, nullptr); else IRR_LAUNCH_PAIR(16, NORMS_FUSED, nullptr);
    }
    if (s == IRR_OK && mode == NORMS_CACHED) {
      const float* gin = g_inv_norm;
      if (!gin) {
        s = row_inv_norms(g, N, D, dt, eps, gin_ws, st);
        if (s != IRR_OK) return s;
        gin = gin_ws;
      }
      if (small_k) IRR_LAUNCH_PAIR(4, NORMS_CACHED, gin); else IRR_LAUNCH_PAIR(16, NORMS_CACHED, gin);
    }
#undef IRR_LAUNCH_PAIR
  } else {
    // fuse the gallery norms into the tile stream when every gallery tile has exactly one consumer;
    // otherwise they come from the caller's cache or from one streaming pre-pass
    const bool fuse = !cached && p.m_tiles == 1;
    const float* gin = g_inv_norm;
    if (!gin && !fuse) {
      s = row_inv_norms(g, N, D, dt, eps, gin_ws, st);
      if (s != IRR_OK) return s;
      gin = gin_ws;
    }
#define IRR_LAUNCH_SC(KM, FN)                                                                   \
  s = launch<KM, false, FN>(qm, tg, FN ? nullptr : gin, nullptr, Q, N, D, k, p, pv, pi, nullptr, \
                            eps, row_floor, f16, st)
    if (fuse) { if (k <= 4) IRR_LAUNCH_SC(4, true); else IRR_LAUNCH_SC(16, true); }
    else      { if (k <= 4) IRR_LAUNCH_SC(4, false); else IRR_LAUNCH_SC(16, false); }
#undef IRR_LAUNCH_SC
  }
  if (s != IRR_OK) return s;
  return merge_partials(pv, pi, p.n_chunks, Q, k, q, D, dt, eps, idx_offset, out_val, out_idx, st);
}

irr_status bf16_cosine_scores(const void* q, const void* g, int64_t Q, int64_t N, int32_t D,
                              float eps, float* out_scores, void* ws, size_t ws_bytes,
                              cudaStream_t st) {
  if (device_cc() / 10 != 10) return IRR_ERR_UNSUPPORTED_DEVICE;
  if (N > 0x7fffff00ll || Q > 0x7fffff00ll) return IRR_ERR_INVALID_ARG;
  const size_t need = align_up(static_cast<size_t>(N) * 4, 256) + align_up(static_cast<size_t>(Q) * 4, 256);
  if (ws_bytes < need) return IRR_ERR_WORKSPACE_TOO_SMALL;
  float* gin = static_cast<float*>(ws);
  float* qin = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + align_up(static_cast<size_t>(N) * 4, 256));
  irr_status s = row_inv_norms(g, N, D, IRR_BF16, eps, gin, st);
  if (s != IRR_OK) return s;
  s = row_inv_norms(q, Q, D, IRR_BF16, eps, qin, st);
  if (s != IRR_OK) return s;
  const Plan p = make_plan(Q, N);
  QueryMaps qm;
  CUtensorMap tg;
  if (!encode_queries(&qm, q, Q, D, false) || !encode_bf16_rows(&tg, g, N, D, BLOCK_N))
    return IRR_ERR_UNSUPPORTED_DEVICE;
  return launch<4, true, false>(qm, tg, gin, qin, Q, N, D, 1, p, nullptr, nullptr, out_scores, eps,
                                nullptr, false, st);
}

// dense [Q,N] cosine scores with both inverse norms supplied (a block of the large-k path)
irr_status bf16_scores_block(const void* q, const void* g, const float* g_inv_norm,
                             const float* q_inv_norm, int64_t Q, int64_t N, int32_t D, float eps,
                             float* out_scores, cudaStream_t st, irr_dtype dt) {
  if (device_cc() / 10 != 10) return IRR_ERR_UNSUPPORTED_DEVICE;
  const bool f16 = dt == IRR_F16;
  const Plan p = make_plan(Q, N);
  QueryMaps qm;
  CUtensorMap tg;
  if (!encode_queries(&qm, q, Q, D, f16) || !encode_bf16_rows(&tg, g, N, D, BLOCK_N, f16))
    return IRR_ERR_UNSUPPORTED_DEVICE;
  return launch<4, true, false>(qm, tg, g_inv_norm, q_inv_norm, Q, N, D, 1, p, nullptr, nullptr,
                                out_scores, eps, nullptr, f16, st);
}

}  // namespace irr
