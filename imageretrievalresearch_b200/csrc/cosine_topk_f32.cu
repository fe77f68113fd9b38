// cosine_topk_f32.cu — K1 exactness path: fp32 inputs, FFMA accumulation (tensor cores cannot hold
// the 1e-5-relative bar of the fp32 mode: kind::tf32 keeps 10 mantissa bits), same running top-k
// and tie rule as the bf16 tensor-core kernel.  Replaces cos(q[i][None], G) + torch.topk(sim, k)
// (train/train_efficient_cos_con_ce_loss.py:273-276,385-388; ipynb:238) for fp32 embeddings.
//
// CTA = 128 threads, score tile 64 queries x 64 gallery rows (three CTAs per SM: a 10k-row gallery
// split over K is 314 CTAs, all resident at once), K-slabs of 16 staged row-major in a
// four-deep cp.async ring (a slab is in flight for three slabs' worth of FMAs: with one slab of
// register prefetch the kernel waited on every load — long-scoreboard stalls were its largest
// stall reason and a slab took 3060 cycles against 1120 of instruction issue); each thread owns a
// 4 x 8 register micro-tile.  The scaled tile goes to shared memory once, and 64 threads (one per
// query row) fold it into their running sorted top-k in increasing column order.  Work unit =
// (query tile, chunk of gallery tiles); the per-unit partial lists are folded by topk_merge.cu.
#include <stdlib.h>

#include "irr_common.cuh"
#include "irr_kernels.h"

namespace irr {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int THREADS = 256;     // large-batch kernel
constexpr int S_THREADS = 128;   // small kernel: 16 x 8 threads, 4 x 8 outputs each
constexpr int S_LD = BN + 1;
// small kernel: ring of ST slabs, rows of BK floats padded to 80 bytes — an odd number of 16-byte
// units, so the LDS.128 of eight consecutive rows hit eight different bank groups
constexpr int ST = 4;
constexpr int ROW_LD = BK + 4;
constexpr int SMEM_A = ST * BM * ROW_LD * 4;
constexpr int SMEM_B = ST * BN * ROW_LD * 4;
constexpr int SMEM_S = BM * S_LD * 4;
constexpr int SMEM_L = BM * 16 * 8;   // running top-k lists between tiles (KMAX <= 16), kept out of
                                     // the registers the FMA loop needs
constexpr int SMEM_BYTES = SMEM_A + SMEM_B + SMEM_S + SMEM_L;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  // src_bytes = 0: the 16 destination bytes are zero-filled (rows / k past the end)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Plan {
  int m_tiles, n_tiles, tiles_per_chunk, n_chunks;
};

Plan make_plan(int64_t Q, int64_t N) {
  Plan p;
  p.m_tiles = static_cast<int>((Q + BM - 1) / BM);
  p.n_tiles = static_cast<int>((N + BN - 1) / BN);
  if (p.m_tiles < 1) p.m_tiles = 1;
  if (p.n_tiles < 1) p.n_tiles = 1;
  const int slots = num_sms() * 3;  // three resident CTAs per SM
  int best = 1;
  double best_cost = 1e300;
  const int max_tpc = p.n_tiles < 64 ? p.n_tiles : 64;
  for (int tpc = 1; tpc <= max_tpc; ++tpc) {
    const long long units = 1ll * ((p.n_tiles + tpc - 1) / tpc) * p.m_tiles;
    const long long waves = (units + slots - 1) / slots;
    const double cost = static_cast<double>(waves) * (tpc + 0.1);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = tpc; }
  }
  p.tiles_per_chunk = best;
  p.n_chunks = (p.n_tiles + best - 1) / best;
  return p;
}

// WRITE_SCORES: instead of selecting, publish the dense cosine tile (both norms applied) — the
// first stage of the large-k path (topk_select.cu).
//
// Every dot product is summed as (first half of K) + (second half of K), each half as (its even
// k, ascending) + (its odd k, ascending) — the two lanes of the packed FFMA2 accumulators.
// SPLIT = 1: one CTA runs both halves (the first half's partial tile waits in its score buffer).
// SPLIT = 2: a cluster of two CTAs runs one half each and the second CTA hands its partial tile to
// the first through distributed shared memory — for searches with too few gallery tiles to give
// every SM a CTA (configs[1]: 10k rows = 79 tiles on 148 SMs).  Same additions in the same order,
// so the two variants return identical bits and a gallery scanned in blocks (StreamedGallery)
// ranks exactly like the resident one whichever variant each block size selects.
template <int KMAX, bool WRITE_SCORES, int SPLIT>
__global__ void __launch_bounds__(S_THREADS, 3)
cosine_topk_f32_kernel(const float* __restrict__ q, const float* __restrict__ g,
                       const float* __restrict__ g_inv_norm, int Q, int N, int D, int k,
                       int m_tiles, int n_tiles, int tiles_per_chunk,
                       float* __restrict__ part_val, int32_t* __restrict__ part_idx,
                       const float* __restrict__ q_inv_norm, float* __restrict__ scores_out) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* As = reinterpret_cast<float*>(smem);                      // [ST][BM][ROW_LD]
  float* Bs = reinterpret_cast<float*>(smem + SMEM_A);             // [ST][BN][ROW_LD]
  float* Ss = reinterpret_cast<float*>(smem + SMEM_A + SMEM_B);    // [BM][S_LD]
  float* Lv = reinterpret_cast<float*>(smem + SMEM_A + SMEM_B + SMEM_S);   // [KMAX][BM]
  int32_t* Li = reinterpret_cast<int32_t*>(Lv + BM * KMAX);
  const uint32_t as_u32 = smem_u32(As), bs_u32 = smem_u32(Bs);

  const int t = threadIdx.x;
  const int ty = t >> 3, tx = t & 7;
  const int unit = blockIdx.x / SPLIT;
  const uint32_t rank = SPLIT == 2 ? cluster_ctarank() : 0u;   // which half of K this CTA sums
  const int chunk = unit / m_tiles, mt = unit - chunk * m_tiles;
  const int m0 = mt * BM;
  const int t0 = chunk * tiles_per_chunk, t1 = min(t0 + tiles_per_chunk, n_tiles);

  // global -> smem staging assignment: a slab row is four 16-byte pieces, two of A and two of B
  // per thread.  This thread's micro-tile: query rows ty*4 + i, gallery rows tx + 8*j.
  const float* a_src[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int f = t + i * S_THREADS;
    a_ok[i] = m0 + (f >> 2) < Q;
    a_src[i] = q + static_cast<size_t>(a_ok[i] ? m0 + (f >> 2) : 0) * D + (f & 3) * 4;
  }

  if (!WRITE_SCORES && t < BM) {   // (read again only by this thread: no barrier needed)
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      Lv[j * BM + t] = kNegInf;
      Li[j * BM + t] = -1;
    }
  }

  const int num_ks = (D + BK - 1) / BK;
  const int half_ks = (num_ks + 1) / 2;
  for (int tile = t0; tile < t1; ++tile) {
    const int n0 = tile * BN;
    // SPLIT == 2: the first CTA is done with the previous tile's score buffer before the second
    // one writes this tile's partial into it
    if (SPLIT == 2) cluster_sync_all();
    const float* b_src[2];
    bool b_ok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = t + i * S_THREADS;
      const int row = f >> 2;
      b_ok[i] = n0 + row < N;
      b_src[i] = g + static_cast<size_t>(b_ok[i] ? n0 + row : 0) * D + (f & 3) * 4;
    }
    // packed accumulators: .x sums the even k of a K-half, .y the odd k (one FFMA2 per two k)
    float2 acc2[4][8];
    float acc[4][8];
    auto zero_acc = [&]() {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc2[i][j] = make_float2(0.f, 0.f);
    };
    auto fold_acc = [&]() {   // a K-half's dot products: (even k) + (odd k)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = acc2[i][j].x + acc2[i][j].y;
    };
    zero_acc();

    // stage K-slab ks (16 k of the 64 query rows and of the tile's 64 gallery rows) into ring slot
    auto issue = [&](int ks, int slot) {
      const int kk = ks * BK;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int f = t + i * S_THREADS;
        const int row = f >> 2, kq = (f & 3) * 4;
        // (a zero-filled piece — row or k past the end — still names a valid address)
        const bool a_in = a_ok[i] && kk + kq < D, b_in = b_ok[i] && kk + kq < D;
        cp_async16(as_u32 + static_cast<uint32_t>((slot * BM + row) * ROW_LD + kq) * 4u,
                   a_in ? a_src[i] + kk : q, a_in ? 16u : 0u);
        cp_async16(bs_u32 + static_cast<uint32_t>((slot * BN + row) * ROW_LD + kq) * 4u,
                   b_in ? b_src[i] + kk : g, b_in ? 16u : 0u);
      }
    };
    // K-slabs [ks0, ks1) into acc (ascending k; slab ks sits in ring slot (ks - ks0) % ST)
    auto run_k = [&](int ks0, int ks1) {
      if (ks0 >= ks1) return;
      __syncthreads();  // previous readers of the ring (and of the previous tile's Ss) are done
#pragma unroll
      for (int s = 0; s < ST - 1; ++s) {
        if (ks0 + s < ks1) issue(ks0 + s, s);
        cp_async_commit();
      }
      for (int ks = ks0; ks < ks1; ++ks) {
        cp_async_wait<ST - 2>();   // this thread's pieces of slab ks have landed ...
        __syncthreads();           // ... everybody's have, and everybody is done with slab ks - 1,
        const int nxt = ks + ST - 1;   // whose slot is refilled now
        if (nxt < ks1) issue(nxt, (nxt - ks0) % ST);
        cp_async_commit();         // (an empty group keeps the count uniform)
        const int slot = (ks - ks0) % ST;
        const float* a = As + (slot * BM + ty * 4) * ROW_LD;
        const float* b = Bs + (slot * BN + tx) * ROW_LD;
#pragma unroll
        for (int kq = 0; kq < BK; kq += 4) {
          float4 av[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(a + i * ROW_LD + kq);
#pragma unroll
          for (int jh = 0; jh < 8; jh += 4) {   // four gallery rows at a time
            float4 bv[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              bv[jj] = *reinterpret_cast<const float4*>(b + (jh + jj) * 8 * ROW_LD + kq);
            // k0,k1 of all 16 outputs, then k2,k3: 16 independent FFMA2 between two that touch the
            // same accumulator
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 a01 = make_float2(av[i].x, av[i].y);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                acc2[i][jh + jj] = __ffma2_rn(a01, make_float2(bv[jj].x, bv[jj].y), acc2[i][jh + jj]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 a23 = make_float2(av[i].z, av[i].w);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                acc2[i][jh + jj] = __ffma2_rn(a23, make_float2(bv[jj].z, bv[jj].w), acc2[i][jh + jj]);
            }
          }
        }
      }
    };
    // this thread's 4 x 8 elements of the score buffer (private to the thread: no barrier needed
    // between its own store and load)
    auto s_at = [&](int i, int j) -> float* {
      return Ss + (ty * 4 + i) * S_LD + tx + 8 * j;
    };
    if (SPLIT == 1) {
      run_k(0, half_ks);
      fold_acc();
      zero_acc();
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) *s_at(i, j) = acc[i][j];
      run_k(half_ks, num_ks);
      fold_acc();
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = *s_at(i, j) + acc[i][j];      // first half + second half
    } else {
      if (rank == 0) run_k(0, half_ks); else run_k(half_ks, num_ks);
      fold_acc();
      if (rank == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) st_shared_cluster_f32(mapa_rank(smem_u32(s_at(i, j)), 0), acc[i][j]);
      }
      cluster_sync_all();   // the second half's partial tile has landed in the first CTA
      if (rank == 1) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = acc[i][j] + *s_at(i, j);      // first half + second half
    }
    // scale by the inverse gallery norms and publish the tile
    float gn[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = n0 + tx + 8 * j;
      gn[j] = c < N ? __ldg(g_inv_norm + c) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) *s_at(i, j) = acc[i][j] * gn[j];
    __syncthreads();
    if (WRITE_SCORES) {
      for (int e = t; e < BM * BN; e += S_THREADS) {
        const int r = e / BN, c = e - r * BN;
        if (m0 + r < Q && n0 + c < N)
          scores_out[static_cast<size_t>(m0 + r) * N + n0 + c] =
              Ss[r * S_LD + c] * __ldg(q_inv_norm + m0 + r);
      }
    } else if (t < BM) {
      TopKList<KMAX, int32_t> top;
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        top.v[j] = Lv[j * BM + t];
        top.i[j] = Li[j * BM + t];
      }
      const int n_valid = min(BN, N - n0);
      const float* s = Ss + t * S_LD;
      for (int c = 0; c < n_valid; ++c) top.push_ordered(s[c], n0 + c);
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        Lv[j * BM + t] = top.v[j];
        Li[j * BM + t] = top.i[j];
      }
    }
  }
  if (!WRITE_SCORES && rank == 0 && t < BM && m0 + t < Q) {
    const size_t o = (static_cast<size_t>(chunk) * Q + m0 + t) * k;
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
      if (j < k) {
        part_val[o + j] = Lv[j * BM + t];
        part_idx[o + j] = Li[j * BM + t];
      }
  }
  // (SPLIT == 2: every DSMEM store precedes the last cluster barrier, so either CTA may exit now)
}


// ---------------------------------------------------------------------------------------------
// Large-batch variant (more than 64 queries): score tile 128 x 128, 8 x 8 register micro-tile per
// thread, packed FFMA2 (two fp32 lanes per instruction: acc[i][2j], acc[i][2j+1] against the splat
// of a[i] — each lane is an IEEE fp32 FMA and every output element still accumulates over k in
// ascending order, so the scores are bit-identical to the 64 x 128 kernel's).  Per k step a thread
// issues 4 LDS.128 + 32 FFMA2 for 64 FMAs (the small kernel: 3 LDS.128 + 32 FFMA for 32), which
// takes instruction issue off the critical path.  The score tile aliases the operand buffers (they
// are dead once the last K-slab is consumed) and the running top-k lists live in shared memory
// between tiles, so the accumulators are the only large register-resident state (2 CTAs per SM).
// ---------------------------------------------------------------------------------------------
constexpr int GM = 128, GN = 128;
constexpr int GA_LD = GM + 4, GB_LD = GN + 4, GS_LD = GN + 1;
constexpr int G_SMEM_AB = 2 * BK * (GA_LD + GB_LD) * 4;        // double-buffered K-slabs
constexpr int G_SMEM_S = GM * GS_LD * 4;                       // score tile (aliases the slabs)
constexpr int G_SMEM_TILE = G_SMEM_AB > G_SMEM_S ? G_SMEM_AB : G_SMEM_S;

template <int KMAX>
constexpr int big_smem_bytes() { return G_SMEM_TILE + GM * KMAX * 8; }

Plan make_plan_big(int64_t Q, int64_t N) {
  Plan p;
  p.m_tiles = static_cast<int>((Q + GM - 1) / GM);
  p.n_tiles = static_cast<int>((N + GN - 1) / GN);
  if (p.m_tiles < 1) p.m_tiles = 1;
  if (p.n_tiles < 1) p.n_tiles = 1;
  const int slots = num_sms() * 2;
  int best = 1;
  double best_cost = 1e300;
  const int max_tpc = p.n_tiles < 64 ? p.n_tiles : 64;
  for (int tpc = 1; tpc <= max_tpc; ++tpc) {
    const long long units = 1ll * ((p.n_tiles + tpc - 1) / tpc) * p.m_tiles;
    const long long waves = (units + slots - 1) / slots;
    const double cost = static_cast<double>(waves) * (tpc + 0.1);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = tpc; }
  }
  p.tiles_per_chunk = best;
  p.n_chunks = (p.n_tiles + best - 1) / best;
  return p;
}

template <int KMAX, bool WRITE_SCORES>
__global__ void __launch_bounds__(THREADS, 2)
cosine_topk_f32_big_kernel(const float* __restrict__ q, const float* __restrict__ g,
                           const float* __restrict__ g_inv_norm, int Q, int N, int D, int k,
                           int m_tiles, int n_tiles, int tiles_per_chunk,
                           float* __restrict__ part_val, int32_t* __restrict__ part_idx,
                           const float* __restrict__ q_inv_norm, float* __restrict__ scores_out) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* As = reinterpret_cast<float*>(smem);                               // [2][BK][GA_LD]
  float* Bs = As + 2 * BK * GA_LD;                                          // [2][BK][GB_LD]
  float* Ss = reinterpret_cast<float*>(smem);                               // [GM][GS_LD], aliases As/Bs
  float* Lv = reinterpret_cast<float*>(smem + G_SMEM_TILE);                 // [KMAX][GM] running lists
  int32_t* Li = reinterpret_cast<int32_t*>(Lv + GM * KMAX);

  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const int chunk = blockIdx.x / m_tiles, mt = blockIdx.x - chunk * m_tiles;
  const int m0 = mt * GM;
  const int t0 = chunk * tiles_per_chunk, t1 = min(t0 + tiles_per_chunk, n_tiles);

  // global -> smem staging: A and B tiles are 128 rows x 16 k = 512 float4 each, two per thread
  const float* a_src[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int f = t + i * THREADS;
    const int row = f >> 2;
    a_ok[i] = m0 + row < Q;
    a_src[i] = q + static_cast<size_t>(a_ok[i] ? m0 + row : 0) * D + (f & 3) * 4;
  }
  if (!WRITE_SCORES && t < GM) {
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      Lv[j * GM + t] = kNegInf;
      Li[j * GM + t] = -1;
    }
  }

  const int num_ks = (D + BK - 1) / BK;
  for (int tile = t0; tile < t1; ++tile) {
    const int n0 = tile * GN;
    const float* b_src[2];
    bool b_ok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = t + i * THREADS;
      const int row = f >> 2;
      b_ok[i] = n0 + row < N;
      b_src[i] = g + static_cast<size_t>(b_ok[i] ? n0 + row : 0) * D + (f & 3) * 4;
    }
    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);

    float4 ra[2], rb[2];
    auto gload = [&](int ks) {
      const int kk = ks * BK;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int kq = ((t + i * THREADS) & 3) * 4;
        ra[i] = (a_ok[i] && kk + kq < D) ? __ldg(reinterpret_cast<const float4*>(a_src[i] + kk)) : z;
        rb[i] = (b_ok[i] && kk + kq < D) ? __ldg(reinterpret_cast<const float4*>(b_src[i] + kk)) : z;
      }
    };
    auto sstore = [&](int buf) {
      float* a = As + buf * BK * GA_LD;
      float* b = Bs + buf * BK * GB_LD;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int f = t + i * THREADS;
        // k rows 8..15 of a slab keep their columns XOR 8: the four k-quads of a warp's 32 scalar
        // stores (row stride 132 = 4 mod 32) then cover all 32 banks instead of colliding in pairs
        const int kq = (f & 3) * 4, row = (f >> 2) ^ (kq & 8);
        a[(kq + 0) * GA_LD + row] = ra[i].x;
        a[(kq + 1) * GA_LD + row] = ra[i].y;
        a[(kq + 2) * GA_LD + row] = ra[i].z;
        a[(kq + 3) * GA_LD + row] = ra[i].w;
        b[(kq + 0) * GB_LD + row] = rb[i].x;
        b[(kq + 1) * GB_LD + row] = rb[i].y;
        b[(kq + 2) * GB_LD + row] = rb[i].z;
        b[(kq + 3) * GB_LD + row] = rb[i].w;
      }
    };

    gload(0);
    __syncthreads();  // previous tile's readers of the score tile (which aliases As/Bs) are done
    sstore(0);
    __syncthreads();
    for (int ks = 0; ks < num_ks; ++ks) {
      const int buf = ks & 1;
      if (ks + 1 < num_ks) gload(ks + 1);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const int sw = kk & 8;   // (compile-time per unrolled step)
        const float4 a0 = *reinterpret_cast<const float4*>(As + buf * BK * GA_LD + kk * GA_LD + ((ty * 4) ^ sw));
        const float4 a1 = *reinterpret_cast<const float4*>(As + buf * BK * GA_LD + kk * GA_LD + 64 + ((ty * 4) ^ sw));
        const float4 b0 = *reinterpret_cast<const float4*>(Bs + buf * BK * GB_LD + kk * GB_LD + ((tx * 4) ^ sw));
        const float4 b1 = *reinterpret_cast<const float4*>(Bs + buf * BK * GB_LD + kk * GB_LD + 64 + ((tx * 4) ^ sw));
        const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float2 bp[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w),
                              make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 as2 = make_float2(ar[i], ar[i]);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(as2, bp[j], acc[i][j]);
        }
      }
      if (ks + 1 < num_ks) {
        sstore(buf ^ 1);
        __syncthreads();
      }
    }
    // scale by the inverse gallery norms and publish the tile (aliases the operand buffers: every
    // thread must be done reading the last K-slab first)
    float gn[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      gn[j] = c < N ? __ldg(g_inv_norm + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = j < 2 ? tx * 4 + 2 * j : 64 + tx * 4 + 2 * (j - 2);
        Ss[r * GS_LD + c] = acc[i][j].x * gn[2 * j];
        Ss[r * GS_LD + c + 1] = acc[i][j].y * gn[2 * j + 1];
      }
    }
    __syncthreads();
    if (WRITE_SCORES) {
      for (int e = t; e < GM * GN; e += THREADS) {
        const int r = e / GN, c = e - r * GN;
        if (m0 + r < Q && n0 + c < N)
          scores_out[static_cast<size_t>(m0 + r) * N + n0 + c] =
              Ss[r * GS_LD + c] * __ldg(q_inv_norm + m0 + r);
      }
    } else if (t < GM) {
      TopKList<KMAX, int32_t> top;
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        top.v[j] = Lv[j * GM + t];
        top.i[j] = Li[j * GM + t];
      }
      const int n_valid = min(GN, N - n0);
      const float* s = Ss + t * GS_LD;
      for (int c = 0; c < n_valid; ++c) top.push_ordered(s[c], n0 + c);
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        Lv[j * GM + t] = top.v[j];
        Li[j * GM + t] = top.i[j];
      }
    }
  }
  if (!WRITE_SCORES && t < GM && m0 + t < Q) {
    const size_t o = (static_cast<size_t>(chunk) * Q + m0 + t) * k;
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
      if (j < k) {
        part_val[o + j] = Lv[j * GM + t];
        part_idx[o + j] = Li[j * GM + t];
      }
  }
}

// IRR_F32_SMALL_TILES=1 forces the 64 x 128 kernel (measurement / bit-equality knob, not an API)
// IRR_F32_SPLITK=0: never split K over a CTA pair (measurement knob, not an API)
bool split_k_enabled() {
  static const bool on = []() {
    const char* e = getenv("IRR_F32_SPLITK");
    return !(e && e[0] == '0');
  }();
  return on;
}

bool use_big(int64_t Q) {
  static const bool forced_small = []() {
    const char* e = getenv("IRR_F32_SMALL_TILES");
    return e && e[0] == '1';
  }();
  return Q > BM && !forced_small;
}

}  // namespace

size_t f32_topk_workspace_bytes(int64_t Q, int64_t N, int32_t k) {
  const Plan p = use_big(Q) ? make_plan_big(Q, N) : make_plan(Q, N);
  const size_t parts = static_cast<size_t>(p.n_chunks) * Q * k;
  return align_up(static_cast<size_t>(N) * 4, 256) + align_up(parts * 4, 256) * 2 + 256;
}

irr_status f32_cosine_topk(const void* q, const void* g, const float* g_inv_norm, int64_t Q,
                           int64_t N, int32_t D, int32_t k, float eps, int64_t idx_offset,
                           float* out_val, int64_t* out_idx, void* ws, size_t ws_bytes,
                           cudaStream_t st) {
  if (N > 0x7fffff00ll || Q > 0x7fffff00ll) return IRR_ERR_INVALID_ARG;
  if (ws_bytes < f32_topk_workspace_bytes(Q, N, k)) return IRR_ERR_WORKSPACE_TOO_SMALL;
  const bool big = use_big(Q);
  const Plan p = big ? make_plan_big(Q, N) : make_plan(Q, N);
  uint8_t* w = static_cast<uint8_t*>(ws);
  float* gin_ws = reinterpret_cast<float*>(w);
  w += align_up(static_cast<size_t>(N) * 4, 256);
  const size_t parts = static_cast<size_t>(p.n_chunks) * Q * k;
  float* pv = reinterpret_cast<float*>(w);
  w += align_up(parts * 4, 256);
  int32_t* pi = reinterpret_cast<int32_t*>(w);

  const float* gin = g_inv_norm;
  if (!gin) {
    irr_status s = row_inv_norms(g, N, D, IRR_F32, eps, gin_ws, st);
    if (s != IRR_OK) return s;
    gin = gin_ws;
  }
  const int grid = p.m_tiles * p.n_chunks;
  profile_mark_start(st);
  if (big) {
#define IRR_LAUNCH_BIG(KM)                                                                        \
  do {                                                                                            \
    auto kern = cosine_topk_f32_big_kernel<KM, false>;                                            \
    static std::atomic<uint64_t> attr_done{0};                                                    \
    if (attr_needed(attr_done)) {                                                                 \
      IRR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                        big_smem_bytes<KM>()));                                   \
      attr_set(attr_done);                                                                        \
    }                                                                                             \
    kern<<<grid, THREADS, big_smem_bytes<KM>(), st>>>(                                            \
        static_cast<const float*>(q), static_cast<const float*>(g), gin, static_cast<int>(Q),     \
        static_cast<int>(N), D, k, p.m_tiles, p.n_tiles, p.tiles_per_chunk, pv, pi, nullptr,      \
        nullptr);                                                                                 \
  } while (0)
    if (k <= 4) IRR_LAUNCH_BIG(4); else IRR_LAUNCH_BIG(16);
#undef IRR_LAUNCH_BIG
  } else {
    // fewer CTAs than half the resident slots (three per SM): split K over a cluster of two CTAs
    // (same bits, see the kernel) — 10k rows: 157 tiles -> 314 CTAs on 444 slots
    const bool split = 2 * grid <= 3 * num_sms() && split_k_enabled();
#define IRR_LAUNCH_SMALL(KM, SP)                                                                  \
  do {                                                                                            \
    auto kern = cosine_topk_f32_kernel<KM, false, SP>;                                            \
    static std::atomic<uint64_t> attr_done{0};                                                    \
    if (attr_needed(attr_done)) {                                                                 \
      IRR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); \
      attr_set(attr_done);                                                                        \
    }                                                                                             \
    cudaLaunchConfig_t cfg = {};                                                                  \
    cfg.gridDim = dim3(grid * SP);                                                                \
    cfg.blockDim = dim3(S_THREADS);                                                               \
    cfg.dynamicSmemBytes = SMEM_BYTES;                                                            \
    cfg.stream = st;                                                                              \
    cudaLaunchAttribute at[1];                                                                    \
    at[0].id = cudaLaunchAttributeClusterDimension;                                               \
    at[0].val.clusterDim.x = SP;                                                                  \
    at[0].val.clusterDim.y = 1;                                                                   \
    at[0].val.clusterDim.z = 1;                                                                   \
    cfg.attrs = at;                                                                               \
    cfg.numAttrs = SP == 2 ? 1 : 0;                                                               \
    IRR_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, static_cast<const float*>(q),                     \
                                    static_cast<const float*>(g), gin, static_cast<int>(Q),       \
                                    static_cast<int>(N), static_cast<int>(D), static_cast<int>(k),\
                                    p.m_tiles, p.n_tiles, p.tiles_per_chunk, pv, pi,              \
                                    static_cast<const float*>(nullptr),                           \
                                    static_cast<float*>(nullptr)));                               \
  } while (0)
    if (split) { if (k <= 4) IRR_LAUNCH_SMALL(4, 2); else IRR_LAUNCH_SMALL(16, 2); }
    else       { if (k <= 4) IRR_LAUNCH_SMALL(4, 1); else IRR_LAUNCH_SMALL(16, 1); }
#undef IRR_LAUNCH_SMALL
  }
  profile_mark_stop(st);
  IRR_LAUNCH_CHECK();
  return merge_partials(pv, pi, p.n_chunks, Q, k, q, D, IRR_F32, eps, idx_offset, out_val, out_idx, st);
}

// dense [Q,N] cosine scores (both norms applied) for the large-k path; g_inv_norm / q_inv_norm given
irr_status f32_cosine_scores(const void* q, const void* g, const float* g_inv_norm,
                             const float* q_inv_norm, int64_t Q, int64_t N, int32_t D,
                             float* out_scores, cudaStream_t st) {
  if (use_big(Q)) {
    const Plan pb = make_plan_big(Q, N);
    auto kb = cosine_topk_f32_big_kernel<4, true>;
    IRR_CUDA_TRY(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      big_smem_bytes<4>()));
    kb<<<pb.m_tiles * pb.n_chunks, THREADS, big_smem_bytes<4>(), st>>>(
        static_cast<const float*>(q), static_cast<const float*>(g), g_inv_norm, static_cast<int>(Q),
        static_cast<int>(N), D, 1, pb.m_tiles, pb.n_tiles, pb.tiles_per_chunk, nullptr, nullptr,
        q_inv_norm, out_scores);
    IRR_LAUNCH_CHECK();
    return IRR_OK;
  }
  const Plan p = make_plan(Q, N);
  auto kern = cosine_topk_f32_kernel<4, true, 1>;
  IRR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  kern<<<p.m_tiles * p.n_chunks, S_THREADS, SMEM_BYTES, st>>>(
      static_cast<const float*>(q), static_cast<const float*>(g), g_inv_norm, static_cast<int>(Q),
      static_cast<int>(N), D, 1, p.m_tiles, p.n_tiles, p.tiles_per_chunk, nullptr, nullptr,
      q_inv_norm, out_scores);
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

}  // namespace irr
