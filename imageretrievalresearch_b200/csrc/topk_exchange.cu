// topk_exchange.cu — K3x: the sharded gallery's one exchange step fused with the candidate merge,
// over NVLink / NVSwitch peer memory.
//
// Every rank owns one exchange buffer that is mapped into all processes of the box (CUDA VMM /
// symmetric memory on the host side; this file only sees G plain device pointers).  One kernel per
// rank then does what `ncclAllGather` + the merge kernel did:
//   push   this rank's [Q,k] (score fp32, global index int64) list is stored straight into slot
//          `rank` of EVERY rank's buffer (remote stores travel over NVLink), the last CTA to finish
//          publishes the call's epoch into each peer's flag word (st.release.sys);
//   wait   each CTA spins (ld.acquire.sys) until all G flags of its OWN buffer reached the epoch;
//   merge  one warp per query folds the G*k candidates that now sit in local memory
//          (score descending, ties -> lower global index, idx < 0 = padding).
// No NCCL, no host round trip, no pack / unpack copies; the call's epoch lives in device memory, so
// the launch has no per-call arguments and can be captured in a CUDA graph.
//
// Buffer layout (irr_topk_exchange_bytes):
//   [0,   64)  uint32 flag[g]  = last epoch whose list rank g finished storing into THIS buffer
//   [256, 264) uint32 epoch (calls completed by the owner), uint32 done (CTAs that finished pushing)
//   [512, ..)  two parity halves; half (epoch & 1) holds G slots [scores Q*k fp32 | pad | indices Q*k i64]
// Why two halves are enough, for both call orders the library uses.  Fused calls (push n, merge n):
// a rank pushes n+2 (overwriting the half call n used) only after its merge n+1, i.e. after every
// peer pushed n+1, which every peer does after its own merge n.  Lagged calls (per search: merge
// n-1, THEN push n — the rendezvous of a search is with the peers' PREVIOUS push, so a rank is
// never held up by peers that are less than one search behind): a rank pushes n+1 (overwriting the
// half of n-1) only after its merge n, i.e. after every peer pushed n, which every peer does after
// its own merge n-1.
//
// Deadlock freedom: a CTA only ever waits for pushes, and a push waits for nothing — but every CTA
// of the fused kernel also waits for its OWN rank's flag, which is published by the last CTA of the
// same grid to finish pushing.  The whole grid therefore has to be resident at once: it is sized
// from cudaOccupancyMaxActiveBlocksPerMultiprocessor (not from an assumed register count) and
// launched COOPERATIVELY, so the driver either co-schedules all of it — also next to kernels of
// other streams, under MPS partitions or green contexts — or refuses the launch, in which case the
// call is issued as the push kernel followed by the wait+merge kernel (stream order then
// guarantees the push has finished before any CTA waits).  Every wait has a watchdog (trap, never a
// hung GPU).
#include <stdlib.h>

#include "irr_common.cuh"
#include "irr_kernels.h"

namespace irr {
namespace {

constexpr int XT = 256;          // threads per CTA
constexpr int XW = XT / 32;      // query rows merged per CTA iteration
constexpr size_t X_STATE = 256;  // byte offset of {epoch, done}
constexpr size_t X_DATA = 512;   // byte offset of the first parity half
constexpr uint32_t X_PARTS = 2;

struct Peers {
  uint8_t* p[IRR_MAX_PEERS];
};

struct XGeom {
  size_t n;           // Q*k
  size_t idx_off;     // bytes from slot start to the int64 indices
  size_t slot_bytes;
  size_t half_bytes;  // distance between the two parity halves
};

XGeom make_geom(int32_t G, int64_t Q, int32_t k, size_t buf_bytes) {
  XGeom x;
  x.n = static_cast<size_t>(Q) * k;
  x.idx_off = align_up(x.n * 4, 16);
  x.slot_bytes = x.idx_off + align_up(x.n * 8, 16);
  x.half_bytes = buf_bytes ? (buf_bytes - X_DATA) / X_PARTS / 256 * 256
                           : align_up(static_cast<size_t>(G) * x.slot_bytes, 256);
  return x;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- push: local list -> slot `rank` of every rank's buffer, then publish the epoch ----------
__device__ __forceinline__ void push_phase(const float* __restrict__ lv,
                                           const int64_t* __restrict__ li, const Peers& peers, int G,
                                           int rank, const XGeom& x, uint32_t epoch,
                                           uint32_t* state) {
  const size_t slot = X_DATA + (epoch % X_PARTS) * x.half_bytes + static_cast<size_t>(rank) * x.slot_bytes;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < x.n; e += stride) {
    const float v = lv[e];
    const int64_t ix = li[e];
    // start with the own buffer's neighbour so that the G ranks do not all hit rank 0 first
    for (int i = 0; i < G; ++i) {
      int g = rank + 1 + i;
      if (g >= G) g -= G;
      uint8_t* dst = peers.p[g] + slot;
      reinterpret_cast<float*>(dst)[e] = v;
      reinterpret_cast<int64_t*>(dst + x.idx_off)[e] = ix;
    }
  }
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // this CTA's remote stores are visible system-wide before it checks in
    const unsigned prev = atomicAdd(state + 1, 1u);
    s_last = prev == gridDim.x - 1;
    __threadfence_system();
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) {
      state[1] = 0;       // self-resetting for the next call on this stream
      state[0] = epoch;
    }
    if (threadIdx.x < G)
      st_release_sys(reinterpret_cast<uint32_t*>(peers.p[threadIdx.x]) + rank, epoch);
  }
}

// ---- wait: all G lists of this epoch have landed in the local buffer ---------------------------
static __device__ __noinline__ void wait_flag_slow(const uint32_t* f, uint32_t epoch,
                                                   unsigned long long timeout_ns) {
  const uint64_t t0 = global_timer_ns();
  while (static_cast<int32_t>(ld_acquire_sys(f) - epoch) < 0) {
    if (global_timer_ns() - t0 > timeout_ns) {
      printf("irr_b200: exchange watchdog: block %d waited %llu ms for the list of rank %d "
             "(epoch %u)\n", (int)blockIdx.x, timeout_ns / 1000000ull, (int)threadIdx.x, epoch);
      __trap();
    }
    __nanosleep(64);
  }
}
__device__ __forceinline__ void wait_phase(const uint8_t* mine, int G, uint32_t epoch,
                                           unsigned long long timeout_ns) {
  if (threadIdx.x < G) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(mine) + threadIdx.x;
    if (static_cast<int32_t>(ld_acquire_sys(f) - epoch) < 0) wait_flag_slow(f, epoch, timeout_ns);
  }
  __syncthreads();
}

// ---- merge: one warp per query over the G*k candidates in the local buffer ---------------------
template <int KMAX>
__device__ __forceinline__ void merge_phase(const uint8_t* mine, int G, int64_t Q, int k,
                                            const XGeom& x, uint32_t epoch,
                                            float* __restrict__ out_val,
                                            int64_t* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const uint8_t* half = mine + X_DATA + (epoch % X_PARTS) * x.half_bytes;
  const int total = G * k;
  for (int64_t qi = static_cast<int64_t>(blockIdx.x) * XW + (threadIdx.x >> 5); qi < Q;
       qi += static_cast<int64_t>(gridDim.x) * XW) {
    TopKList<KMAX, long long> L;
    L.reset();
    for (int c = lane; c < total; c += 32) {
      const int g = c / k, j = c - g * k;
      const uint8_t* slot = half + static_cast<size_t>(g) * x.slot_bytes;
      const size_t o = static_cast<size_t>(qi) * k + j;
      // .cg: the lists were written by other GPUs during this kernel's lifetime — read them at L2
      L.push_any(__ldcg(reinterpret_cast<const float*>(slot) + o),
                 static_cast<long long>(__ldcg(reinterpret_cast<const long long*>(slot + x.idx_off) + o)));
    }
    warp_merge_topk<KMAX, long long>(L, k, [&](int j, float v, long long i) {
      if (lane == 0) {
        out_val[qi * k + j] = i >= 0 ? v : kNegInf;
        out_idx[qi * k + j] = i >= 0 ? i : -1;
      }
    });
  }
}

// mode: IRR_XCHG_FUSED = push + wait + merge, IRR_XCHG_MERGE = wait + merge of the epoch already pushed
template <int KMAX>
__global__ void __launch_bounds__(XT, 2)
exchange_merge_kernel(const float* __restrict__ lv, const int64_t* __restrict__ li,
                      const __grid_constant__ Peers peers, int G, int rank, int64_t Q, int k,
                      const __grid_constant__ XGeom x, int mode, unsigned long long timeout_ns,
                      float* __restrict__ out_val, int64_t* __restrict__ out_idx) {
  uint8_t* mine = peers.p[rank];
  uint32_t* state = reinterpret_cast<uint32_t*>(mine + X_STATE);
  // read before this CTA checks in: the last CTA to check in is the one that advances it
  uint32_t epoch = __ldcg(state);
  if (mode == IRR_XCHG_FUSED) {
    ++epoch;
    push_phase(lv, li, peers, G, rank, x, epoch, state);
  }
  wait_phase(mine, G, epoch, timeout_ns);
  merge_phase<KMAX>(mine, G, Q, k, x, epoch, out_val, out_idx);
}

__global__ void __launch_bounds__(XT)
exchange_push_kernel(const float* __restrict__ lv, const int64_t* __restrict__ li,
                     const __grid_constant__ Peers peers, int G, int rank,
                     const __grid_constant__ XGeom x) {
  uint32_t* state = reinterpret_cast<uint32_t*>(peers.p[rank] + X_STATE);
  const uint32_t epoch = __ldcg(state) + 1u;
  push_phase(lv, li, peers, G, rank, x, epoch, state);
}

__global__ void __launch_bounds__(32)
exchange_wait_kernel(const uint8_t* __restrict__ mine, int G, unsigned long long timeout_ns) {
  const uint32_t epoch = __ldcg(reinterpret_cast<const uint32_t*>(mine + X_STATE));
  wait_phase(mine, G, epoch, timeout_ns);
}

// CTAs of `kern` (XT threads, no dynamic shared memory) the current device holds at once
template <typename K>
int resident_ctas(K kern) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, XT, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  return per_sm * num_sms();
}

unsigned long long exchange_timeout_ns() {
  static unsigned long long cached = []() -> unsigned long long {
    long ms = 600000;   // a late peer is waited for as long as a collective library would
    if (const char* e = getenv("IRR_EXCHANGE_TIMEOUT_MS")) {
      const long v = atol(e);
      if (v > 0) ms = v;
    }
    return static_cast<unsigned long long>(ms) * 1000000ull;
  }();
  return cached;
}

}  // namespace

size_t topk_exchange_bytes(int32_t G, int64_t Q, int32_t k) {
  const XGeom x = make_geom(G, Q, k, 0);
  return X_DATA + X_PARTS * x.half_bytes;
}

irr_status topk_exchange_merge(const float* local_val, const int64_t* local_idx,
                               void* const* peer_bufs, int32_t G, int32_t rank, int64_t Q, int32_t k,
                               size_t buf_bytes, int32_t mode, float* out_val, int64_t* out_idx,
                               cudaStream_t st) {
  if (buf_bytes < topk_exchange_bytes(G, Q, k)) return IRR_ERR_WORKSPACE_TOO_SMALL;
  Peers peers = {};
  for (int g = 0; g < G; ++g) {
    if (!peer_bufs[g] || !aligned16(peer_bufs[g])) return IRR_ERR_INVALID_ARG;
    peers.p[g] = static_cast<uint8_t*>(peer_bufs[g]);
  }
  const XGeom x = make_geom(G, Q, k, buf_bytes);
  const unsigned long long tmo = exchange_timeout_ns();
  // CTAs of the fused kernel that fit the device at once (queried once per instantiation)
  static const int resident4 = resident_ctas(exchange_merge_kernel<4>);
  static const int resident16 = resident_ctas(exchange_merge_kernel<16>);
  const int resident = k <= 4 ? resident4 : resident16;
  const int cap = resident < 2 * num_sms() ? resident : 2 * num_sms();
  const int64_t push_ctas = static_cast<int64_t>((x.n + XT - 1) / XT);
  const int64_t merge_ctas = (Q + XW - 1) / XW;
  auto clamp = [&](int64_t v) { return static_cast<int>(v < 1 ? 1 : (v > cap ? cap : v)); };

  if (mode == IRR_XCHG_PUSH || (mode == IRR_XCHG_FUSED && k > IRR_MAX_K_FUSED)) {
    exchange_push_kernel<<<clamp(push_ctas), XT, 0, st>>>(local_val, local_idx, peers, G, rank, x);
    IRR_LAUNCH_CHECK();
    if (mode == IRR_XCHG_PUSH) return IRR_OK;
  }
  if (k > IRR_MAX_K_FUSED) {
    // large k: the sorted-merge kernel runs one CTA per query (not co-resident), so the wait is a
    // separate one-warp kernel in front of it; the merge reads the lists in place
    exchange_wait_kernel<<<1, 32, 0, st>>>(peers.p[rank], G, tmo);
    IRR_LAUNCH_CHECK();
    // which part holds the lists depends on the epoch in device memory: the merge reads it itself
    return merge_candidates_large_exchange(peers.p[rank], X_STATE, X_DATA, x.half_bytes,
                                           x.slot_bytes, x.idx_off, G, Q, k, out_val, out_idx, st);
  }
  const int grid = clamp(mode == IRR_XCHG_FUSED && push_ctas > merge_ctas ? push_ctas : merge_ctas);
  if (mode == IRR_XCHG_FUSED) {
    // cooperative: all CTAs resident, or the launch is refused (then: push kernel + merge kernel)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(XT);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const long long q64 = Q;
    const cudaError_t e =
        k <= 4 ? cudaLaunchKernelEx(&cfg, exchange_merge_kernel<4>, local_val, local_idx, peers,
                                    static_cast<int>(G), static_cast<int>(rank), static_cast<int64_t>(q64),
                                    static_cast<int>(k), x, static_cast<int>(mode), tmo, out_val, out_idx)
               : cudaLaunchKernelEx(&cfg, exchange_merge_kernel<16>, local_val, local_idx, peers,
                                    static_cast<int>(G), static_cast<int>(rank), static_cast<int64_t>(q64),
                                    static_cast<int>(k), x, static_cast<int>(mode), tmo, out_val, out_idx);
    if (e == cudaSuccess) return IRR_OK;
    if (e != cudaErrorCooperativeLaunchTooLarge && e != cudaErrorLaunchOutOfResources &&
        e != cudaErrorNotSupported)
      return static_cast<irr_status>(static_cast<int>(e));
    cudaGetLastError();
    exchange_push_kernel<<<clamp(push_ctas), XT, 0, st>>>(local_val, local_idx, peers, G, rank, x);
    IRR_LAUNCH_CHECK();
    mode = IRR_XCHG_MERGE;
  }
  // wait + merge only: a CTA waits for pushes of OTHER grids (this rank's own push has completed in
  // stream order), so residency does not matter
  const int mgrid = clamp(merge_ctas);
  if (k <= 4)
    exchange_merge_kernel<4><<<mgrid, XT, 0, st>>>(local_val, local_idx, peers, G, rank, Q, k, x, mode,
                                                   tmo, out_val, out_idx);
  else
    exchange_merge_kernel<16><<<mgrid, XT, 0, st>>>(local_val, local_idx, peers, G, rank, Q, k, x,
                                                    mode, tmo, out_val, out_idx);
  IRR_LAUNCH_CHECK();
  return IRR_OK;
}

}  // namespace irr
