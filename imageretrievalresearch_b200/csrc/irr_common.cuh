// irr_common.cuh — shared host/device helpers for the sm_100a retrieval-ranking kernels.
//
// Everything here is hand-written PTX for Blackwell (mbarrier, TMA bulk copies, tcgen05 / TMEM);
// there is no CUTLASS / CuTe dependency.  Compile with -gencode arch=compute_100a,code=sm_100a.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/irr_b200.h"

namespace irr {

constexpr int kWarp = 32;
constexpr float kNegInf = -__builtin_huge_valf();

// ATen's cosine_embedding_loss adds this to both squared norms
constexpr float kCosEmbEps = 1e-12f;
// utils/contrastive_loss.py:34 (self.eps) — added to the squared distance before the sqrt
constexpr float kContrastiveEps = 1e-9f;

#define IRR_CUDA_TRY(expr)                              \
  do {                                                  \
    cudaError_t _e = (expr);                            \
    if (_e != cudaSuccess) return (irr_status)(int)_e;  \
  } while (0)

#define IRR_LAUNCH_CHECK()                              \
  do {                                                  \
    cudaError_t _e = cudaGetLastError();                \
    if (_e != cudaSuccess) return (irr_status)(int)_e;  \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int dtype_bytes(irr_dtype dt) { return dt == IRR_F32 ? 4 : 2; }

// Number of SMs of the current device (148 on B200); 148 when no device is visible so that the
// workspace-size queries stay callable on a CPU-only build box.
int num_sms();
// compute capability major*10+minor of the current device, 0 if none
int device_cc();

// cudaFuncSetAttribute once per (kernel instantiation, device), not once per call: one bit per
// device in a function-local static
inline bool attr_needed(std::atomic<uint64_t>& done) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  return (done.load(std::memory_order_relaxed) >> dev & 1ull) == 0;
}
inline void attr_set(std::atomic<uint64_t>& done) {
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64)
    done.fetch_or(1ull << dev, std::memory_order_relaxed);
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// two packed bf16 (one 32-bit word) -> two fp32, exact
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// two packed fp16 (one 32-bit word) -> two fp32, exact
__device__ __forceinline__ float2 f16x2(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
// two packed 16-bit floats of either kind; `f16` is warp-uniform
__device__ __forceinline__ float2 unpack16x2(uint32_t w, bool f16) {
  return f16 ? f16x2(w) : make_float2(bf16lo(w), bf16hi(w));
}

// 128-bit streaming global load (read once: do not allocate in L1)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (it
// is woken by the completion, so the hint costs no latency) or the hint expires, instead of the
// thread spinning through issue slots — and joules — that the arithmetic needs.
__device__ __forceinline__ bool mbar_try_wait_suspend(uint32_t bar, uint32_t parity, uint32_t ns);
// Wait with a watchdog: a protocol bug must end in a trapped kernel (a reported launch failure),
// never in a hung GPU.  The slow path is only entered when the first probe fails; it polls through
// parked try_waits (an ncu source view of the Q=4096 search showed half of all executed warp
// instructions in spin loops before this: the epilogue warps idle ~80 % of a tile's time).
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int tag) {
  const uint64_t t0 = global_timer_ns();
  while (!mbar_try_wait_suspend(bar, parity, 4000u)) {
    if (global_timer_ns() - t0 > 4000000000ull) {
      printf("irr_b200: mbarrier watchdog: block %d thread %d tag %d parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, tag, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, tag);
}
// Same contract for waits made by MANY threads at once (whole warps waiting for bulk-copied data).
__device__ __forceinline__ bool mbar_try_wait_suspend(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
static __device__ __noinline__ void mbar_wait_parked_slow(uint32_t bar, uint32_t parity, int tag) {
  const uint64_t t0 = global_timer_ns();
  while (!mbar_try_wait_suspend(bar, parity, 20000u)) {
    if (global_timer_ns() - t0 > 4000000000ull) {
      printf("irr_b200: mbarrier watchdog: block %d thread %d tag %d parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, tag, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_parked_slow(bar, parity, tag);
}

// ---- TMA -------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 2-D tiled load global -> shared, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, int c0, int c1,
                                            uint32_t bar, uint64_t cache_policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "l"(cache_policy)
      : "memory");
}
// 1-D bulk copy global -> shared (SASS: UBLKCP); bytes % 16 == 0, both addresses 16-B aligned
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_dst),
      "l"(gsrc), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d_hint(uint32_t smem_dst, const void* gsrc, uint32_t bytes,
                                                  uint32_t bar, uint64_t cache_policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], "
      "%2, [%3], %4;" ::"r"(smem_dst),
      "l"(gsrc), "r"(bytes), "r"(bar), "l"(cache_policy)
      : "memory");
}
// L2 eviction-priority policies for the .L2::cache_hint operand (createpolicy encodings)
constexpr uint64_t kPolicyEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kPolicyEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kPolicyEvictLast = 0x14F0000000000000ull;

// ---- tcgen05 / TMEM ----------------------------------------------------------------------------
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// whole warp; writes the TMEM base address to *smem_slot
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; one thread issues for the CTA
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread l of the warp receives columns [c, c+32) of
// TMEM lane (lane field of taddr + l).  A warp may only touch lanes 32*(warp_id%4) .. +31.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// ---- CTA pairs (cluster of 2, tcgen05 cta_group::2) --------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
// wait on a LOCAL mbarrier whose arrivals come from both CTAs of the cluster and publish data
// the peer wrote into this CTA's shared memory (acquire at cluster scope)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster_suspend(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
static __device__ __noinline__ void mbar_wait_cluster_slow(uint32_t bar, uint32_t parity, int tag) {
  const uint64_t t0 = global_timer_ns();
  while (!mbar_try_wait_cluster_suspend(bar, parity, 4000u)) {
    if (global_timer_ns() - t0 > 4000000000ull) {
      printf("irr_b200: mbarrier watchdog: block %d thread %d tag %d parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, tag, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  mbar_wait_cluster_slow(bar, parity, tag);
}
// fp32 store into the shared memory of any CTA of the cluster (address from mapa_rank)
__device__ __forceinline__ void st_shared_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
// 2-D tiled load issued by either CTA of a pair; the bytes are accounted on the mbarrier at
// `bar_cluster_addr`, which may live in the peer (leader) CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const void* tmap, int c0, int c1,
                                                 uint32_t bar_cluster_addr, uint64_t cache_policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(bar_cluster_addr), "l"(cache_policy)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
// M=256 MMA across the two SMs of a pair: each CTA supplies 128 rows of A and half of B from the
// same shared-memory offsets; issued by one thread of the leader CTA only
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this shared-memory offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}

// UMMA shared-memory matrix descriptor for a K-major bf16 tile stored as rows of 128 bytes
// (64 elements) with the 128-byte swizzle TMA applies: 8-row groups are 1024 B apart (SBO),
// LBO unused, descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);       // start address, bits [0,14)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                 // stride byte offset, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                         // version = 1
  d |= static_cast<uint64_t>(2) << 61;                         // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::f16: D=f32, A=B=fp16 (format 0) or bf16 (format 1), both K-major
__host__ __device__ constexpr uint32_t umma_idesc_16(int m, int n, bool f16) {
  return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// monotone float <-> uint32 map (larger float -> larger unsigned); 0 is below every float
__device__ __forceinline__ uint32_t orderable(float v) {
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
// the same map with every NaN on top (torch.topk's order: NaN is the largest value);
// from_orderable(0xffffffff) is a NaN again
__device__ __forceinline__ uint32_t orderable_nan_top(float v) {
  return v != v ? 0xffffffffu : orderable(v);
}
__device__ __forceinline__ float from_orderable(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// ---- per-thread sorted top-k list in registers -----------------------------------------------
// "better" = larger score, ties -> lower index (the rule irr_b200.h promises).  NaN scores (a
// gallery or query row with a non-finite element: torch's cosine_similarity gives NaN there) order
// the way torch.topk orders them: above every number; among themselves by the lower index.
__device__ __forceinline__ bool cand_better(float va, long long ia, float vb, long long ib) {
  const bool na = va != va, nb = vb != vb;
  if (na || nb) return na && (!nb || ia < ib);
  return va > vb || (va == vb && ia < ib);
}

template <int KMAX, typename IdxT>
struct TopKList {
  float v[KMAX];
  IdxT i[KMAX];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      v[j] = kNegInf;
      i[j] = static_cast<IdxT>(-1);
    }
  }
  // candidates offered in increasing index order: strict '>' keeps the lower index first.
  // "a beats b" for a later candidate a: b is a number and a is larger or NaN — !(a <= b) is
  // true for a NaN a, and a NaN b is never displaced
  __device__ __forceinline__ void push_ordered(float s, IdxT idx) {
    if (v[KMAX - 1] == v[KMAX - 1] && !(s <= v[KMAX - 1])) {
      v[KMAX - 1] = s;
      i[KMAX - 1] = idx;
#pragma unroll
      for (int j = KMAX - 1; j > 0; --j) {
        if (v[j - 1] == v[j - 1] && !(v[j] <= v[j - 1])) {
          float tv = v[j]; v[j] = v[j - 1]; v[j - 1] = tv;
          IdxT ti = i[j]; i[j] = i[j - 1]; i[j - 1] = ti;
        }
      }
    }
  }
  // Same insert as push_ordered without a serial bubble chain: rank the candidate against all
  // entries at once (independent compares), then every slot picks {keep, take from above, take the
  // candidate} — high ILP, which matters with a single epilogue warp per scheduler.  `take` must
  // imply that s beats v[KMAX-1]; lanes with take == false leave their list untouched.
  __device__ __forceinline__ void insert_ranked(bool take, float s, IdxT idx) {
    int c = 0;  // entries that stay ahead of s: the earlier (lower) index wins ties
    const bool s_nan = s != s;   // a NaN goes behind the NaNs already listed, ahead of every number
#pragma unroll
    for (int j = 0; j < KMAX; ++j) c += (s_nan ? (v[j] != v[j]) : !(v[j] < s)) ? 1 : 0;
    if (!take) c = KMAX;
#pragma unroll
    for (int j = KMAX - 1; j > 0; --j) {
      const bool from_above = j > c;
      const bool here = j == c;
      v[j] = from_above ? v[j - 1] : (here ? s : v[j]);
      i[j] = from_above ? i[j - 1] : (here ? idx : i[j]);
    }
    if (c == 0) { v[0] = s; i[0] = idx; }
  }
  // candidates in arbitrary order: full (score, index) comparison; idx < 0 = padding
  __device__ __forceinline__ void push_any(float s, IdxT idx) {
    if (idx < 0) return;
    if (i[KMAX - 1] < 0 || cand_better(s, idx, v[KMAX - 1], i[KMAX - 1])) {
      v[KMAX - 1] = s;
      i[KMAX - 1] = idx;
#pragma unroll
      for (int j = KMAX - 1; j > 0; --j) {
        if (i[j - 1] < 0 || cand_better(v[j], i[j], v[j - 1], i[j - 1])) {
          float tv = v[j]; v[j] = v[j - 1]; v[j - 1] = tv;
          IdxT ti = i[j]; i[j] = i[j - 1]; i[j - 1] = ti;
        }
      }
    }
  }
  __device__ __forceinline__ void pop_front() {
#pragma unroll
    for (int j = 0; j < KMAX - 1; ++j) {
      v[j] = v[j + 1];
      i[j] = i[j + 1];
    }
    v[KMAX - 1] = kNegInf;
    i[KMAX - 1] = static_cast<IdxT>(-1);
  }
};

// One warp merges its 32 per-lane sorted lists into the global top-k: k rounds of
// "shuffle-argmax over the lane heads, winner pops".  Lane 0 receives round j's winner.
template <int KMAX, typename IdxT, typename Emit>
__device__ __forceinline__ void warp_merge_topk(TopKList<KMAX, IdxT>& L, int k, Emit emit) {
  const int lane = threadIdx.x & 31;
  for (int j = 0; j < k; ++j) {
    float bv = L.v[0];
    long long bi = static_cast<long long>(L.i[0]);
    int bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      int ol = __shfl_xor_sync(0xffffffffu, bl, o);
      // total order so that every lane of the butterfly agrees on the winner: real beats
      // padding, then (score, index), then the lower lane for identical keys
      bool take;
      if (bi < 0 || oi < 0) take = (bi < 0 && oi >= 0) || (bi < 0 && oi < 0 && ol < bl);
      else if (ov == bv && oi == bi) take = ol < bl;
      else take = cand_better(ov, oi, bv, bi);
      if (take) { bv = ov; bi = oi; bl = ol; }
    }
    if (lane == bl && bi >= 0) L.pop_front();
    emit(j, bv, bi);
  }
}

#endif  // __CUDACC__

}  // namespace irr
