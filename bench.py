"""bench.py — headline benchmark of the retrieval-ranking hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU torch path

Metric (BASELINE.json): top-k cosine queries/s at a 1M x 1536 gallery.  Workload: cosine top-3
over a 1,000,000 x 1536 bf16 synthetic gallery with Q=4096 queries per step (the configuration
BASELINE.json's target is quoted on; it fits one GPU).  With N>1 the gallery is row-sharded over
the N ranks (N/G rows each, so total work is fixed: "strong" scaling) and every step ends with the
path's one exchange of the [Q,k] candidates: ONE kernel per rank that stores its list into every
peer over NVLink peer memory, flags, waits and merges (csrc/topk_exchange.cu); if peer memory
cannot be mapped the ranks agree on an NCCL all-gather + merge kernel instead (`config.exchange`
says which ran).

A step = one search of all Q queries: the tcgen05 top-k kernel — whose four norm-producer warps
per CTA recompute the inverse row norms of the gallery shard inside the same launch, every step:
nothing is cached across steps — the partial-list merge, and for
N>1 the candidate exchange + merge.  `value` times that with the queries resident in HBM; `e2e`
times the same search driven through `irr.SearchPipeline` — the package's serving loop — with every
step's queries coming from pinned host memory and its [Q,k] results going back to pinned host
memory inside the timed region, the copies of neighbouring steps overlapped with the search (the
gallery is the resident index, as in the reference where the embeddings live on the device:
inference/training_analysis.ipynb:222).
The gallery (3.07 GB, or 384 MB per rank at N=8) is larger than the 126 MB L2, so no flush is used.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

# a rank that dies must not leave the others spinning for the library's default ten minutes
os.environ.setdefault("IRR_EXCHANGE_TIMEOUT_MS", "60000")

METRIC = "top-k cosine queries/s at 1Mx1536 gallery"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=1536)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="also time Q=1 and Q=64 (extra keys)")
    return ap.parse_args()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained"), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback"}


def workload_name(a, world):
    return (f"cosine top-{a.k} over a {a.rows}x{a.dim} bf16 gallery, Q={a.queries} per step, "
            f"gallery row-sharded over {world} GPU(s)" + (" + NCCL all-gather merge" if world > 1 else ""))


# -------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks", 0x100: "display",
               0x10: "sync_boost"}

    def __init__(self, index: int, enabled: bool = True):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        if not enabled:            # only rank 0 polls NVML: eight pollers on one box perturb the
            self.nv, self.err = None, "not sampled on this rank"   # launches they are meant to observe
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # no NVML: report that, do not fake numbers
            self.nv, self.err = None, repr(e)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) \
                    if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.010)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.err}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# -------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU torch path (the oracle's literal loop), rank 0 only
# -------------------------------------------------------------------------------------------------
def host_gallery(rows, dim, seed=3):
    """fp32 gallery on the host (the reference keeps fp32 embeddings); generated in slabs."""
    import torch
    g = torch.empty(rows, dim, dtype=torch.float32)
    gen = torch.Generator().manual_seed(seed)
    for lo in range(0, rows, 65536):
        hi = min(rows, lo + 65536)
        g[lo:hi] = torch.randn(hi - lo, dim, generator=gen).bfloat16().float()
    return g


def run_reference(a):
    import torch
    from oracle import reference_path as ref
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = host_gallery(a.rows, a.dim)
    gen = torch.Generator().manual_seed(11)
    sample_q = 2                                      # queries per step (bounded sample)
    q = torch.randn(sample_q * (a.steps + a.warmup), a.dim, generator=gen).bfloat16().float()
    # one untimed query sizes the sample so that the whole run stays within a few minutes
    t0 = time.perf_counter()
    ref.cos_topk_loop(q[:1], g, a.k)
    per_query = time.perf_counter() - t0
    if per_query * sample_q * (a.steps + a.warmup) > 240:
        sample_q = 1
    for w in range(a.warmup):
        ref.cos_topk_loop(q[w * sample_q:(w + 1) * sample_q], g, a.k)
    t0 = time.perf_counter()
    for s in range(a.steps):
        lo = (a.warmup + s) * sample_q
        ref.cos_topk_loop(q[lo:lo + sample_q], g, a.k)
    dt = time.perf_counter() - t0
    value = sample_q * a.steps / dt
    sample = (f"{sample_q} queries per step over the full {a.rows}x{a.dim} fp32 gallery: the "
              f"reference's per-query CosineSimilarity(dim=1,eps=1e-6)+torch.topk loop "
              f"(train/train_efficient_cos_con_ce_loss.py:270-276), torch CPU, {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(a, a.gpus), "Q": a.queries, "N": a.rows, "D": a.dim,
                   "k": a.k, "sampled_queries_per_step": sample_q},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def cpu_baseline(a, gallery_dev):
    """Oracle loop on the host cores, ~10-30 s, on the same gallery (copied back as fp32)."""
    import torch
    from oracle import reference_path as ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.empty(gallery_dev.shape, dtype=torch.float32)
    for lo in range(0, g.shape[0], 131072):
        g[lo:lo + 131072] = gallery_dev[lo:lo + 131072].float().cpu()
    q = torch.randn(64, a.dim, generator=torch.Generator().manual_seed(11)).bfloat16().float()
    ref.cos_topk_loop(q[:1], g, a.k)                  # warm-up (page-in, thread pool)
    n, t0 = 0, time.perf_counter()
    while n < 64 and (time.perf_counter() - t0 < 15.0 or n < 2):
        ref.cos_topk_loop(q[n:n + 1], g, a.k)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} queries over the full {a.rows}x{a.dim} gallery (fp32 on the host), the "
                      f"reference's per-query cos+topk loop, torch CPU with {cores} threads, {dt:.1f} s"}


# -------------------------------------------------------------------------------------------------
# this repo's arm
# -------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    import imageretrievalresearch_b200 as irr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA: the product path has no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = irr.load_library()
    Q, N, D, k = a.queries, a.rows, a.dim, a.k

    # synthetic inputs: this rank's contiguous shard of the gallery, queries replicated
    lo, hi = irr.shard_bounds(N, world, rank)
    gen = torch.Generator(device=dev).manual_seed(3 + rank)
    shard = torch.randn(hi - lo, D, device=dev, dtype=torch.bfloat16, generator=gen)
    qgen = torch.Generator(device=dev).manual_seed(11)
    queries = torch.randn(Q, D, device=dev, dtype=torch.bfloat16, generator=qgen)
    if world > 1:
        gallery = irr.ShardedGallery(shard, N, cache_norms=False)
        search = lambda q: gallery.search(q, k)
    else:
        search = lambda q: irr.cosine_topk(q, shard, k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- device-resident timing (value) with per-launch timing of the dominant kernel ----
    for _ in range(a.warmup):
        search(queries)
    # N > 1 with the peer-memory exchange: the stream of searches runs lagged (each search's
    # rendezvous is with the peers' PREVIOUS push, ShardedGallery.search_lagged), so a step does
    # not cost the slowest of N kernels; every search's merged result is still produced inside the
    # timed region (the last one by flush()).  IRR_BENCH_LAGGED=0 times the plain search instead.
    lagged = (world > 1 and gallery.transport == "peer" and k <= 16
              and os.environ.get("IRR_BENCH_LAGGED", "1") != "0")
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(a.steps)]
    for s, e in kev:            # force creation of the raw cudaEvent_t handles
        s.record(); e.record()
    # NVML is initialised and the sampler thread is running BEFORE the ranks line up: anything a
    # rank does between the barrier and its first launch is waited for by every other rank in the
    # first exchange and would be charged to all of their timers
    sampler = ClockSampler(local, enabled=(rank == 0))
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(); stop.record()          # create the raw handles ahead of the barrier, too
    with sampler as clocks:
        barrier()
        start.record()
        prof_every = int(os.environ.get("IRR_BENCH_PROF_EVERY", "1"))
        for j, (s, e) in enumerate(kev):
            if prof_every > 0 and j % prof_every == 0:
                lib.irr_profile_next_topk(s.cuda_event, e.cuda_event)
            if lagged:
                gallery.search_lagged(queries, k)
            else:
                search(queries)
        if lagged:
            gallery.flush()
        stop.record()
        barrier()
    ms_total = max_over_ranks(start.elapsed_time(stop))
    ms_step = ms_total / a.steps
    timed = [(s, e) for j, (s, e) in enumerate(kev) if prof_every > 0 and j % prof_every == 0]
    kernel_ms = sum(s.elapsed_time(e) for s, e in timed) / len(timed) if timed else ms_step
    kernel_ms = max_over_ranks(kernel_ms)
    value = Q / (ms_step * 1e-3)

    # ---- end to end through the public API with host buffers (e2e) ----
    # irr.SearchPipeline: every step's queries are copied from pinned host memory and its [Q,k]
    # results are read back to pinned host memory inside the timed region; the copy of step i+1
    # and the read-back of step i-1 overlap the search of step i (two steps in flight), and the
    # host touches every result (the generator hands them out one by one).
    q_host = queries.cpu().pin_memory()
    pipe = irr.SearchPipeline(lambda q, kk: search(q), Q, D, k, torch.bfloat16, dev, depth=2)

    def e2e_run(n):
        seen = 0
        for v_host, i_host in pipe.run(q_host for _ in range(n)):
            seen += int(i_host[0, 0] >= 0)          # the caller holds the results
        return seen

    e2e_run(max(3, a.warmup))
    barrier()
    t0 = time.perf_counter()
    e2e_run(a.steps)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / a.steps
    e2e = {"value": Q / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": pipe.h2d_bytes_per_batch,
           "d2h_bytes_per_step": pipe.d2h_bytes_per_batch,
           "api": "SearchPipeline.run over pinned host batches (copy / search / read-back overlapped)"}

    # ---- roofline of the dominant kernel (cosine_topk_bf16_kernel) ----
    peaks = measured_peaks()
    n_local = hi - lo
    flops = 2.0 * Q * n_local * D
    bytes_alg = n_local * D * 2 + Q * D * 2 + Q * k * 12
    t_tensor = flops / (peaks["bf16_tflops"] * 1e12)
    t_hbm = bytes_alg / (peaks["hbm_gbs"] * 1e9)
    traffic = None
    tp = ROOT / "profiles" / "dram_traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text()).get(f"Q{Q}_N{n_local}_D{D}")
    if t_tensor >= t_hbm:
        ach = flops / (kernel_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops"], "traffic": traffic,
                "peak_sustained": peaks["bf16_tflops_sustained"],
                "frac_of_sustained": ach / peaks["bf16_tflops_sustained"] if peaks["bf16_tflops_sustained"] else None}
    else:
        ach = bytes_alg / (kernel_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic}
    kname = "cosine_topk_bf16_pair_kernel" if Q > 384 else "cosine_topk_bf16_kernel"
    roof.update({"kernel": kname, "kernel_ms": kernel_ms, "peak_source": peaks["source"],
                 "kernel_share_of_step": kernel_ms / ms_step,
                 "algorithmic": {"flops": flops, "bytes": bytes_alg}})

    extra = {}
    if a.sweep and world == 1:
        sweep = []
        for qs in (1, 64):
            qq = queries[:qs].contiguous()
            for _ in range(3):
                search(qq)
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(a.steps):
                search(qq)
            s1.record()
            torch.cuda.synchronize()
            ms = s0.elapsed_time(s1) / a.steps
            b = n_local * D * 2 + qs * D * 2 + qs * k * 12
            sweep.append({"Q": qs, "ms_per_step": ms, "queries_per_s": qs / (ms * 1e-3),
                          "hbm_frac_of_step": b / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]})
        extra["sweep"] = sweep

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline(a, shard)
    if world > 1:
        dist.barrier()

    if rank == 0:
        # top-k (gallery norms produced inside it), partial merge (+ the fused exchange/merge kernel, or
        # the lagged pair merge-of-previous + push); 257..512 queries
        # without cached norms still run the streaming norm pre-pass kernel
        launches_per_step = 2 + (1 if 256 < Q <= 512 else 0) + ((2 if lagged else 1) if world > 1 else 0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(a, world), "Q": Q, "N": N, "D": D, "k": k,
                       "parallelism": f"gallery rows sharded x{world}",
                       "l2": "inputs larger than L2 (gallery shard streamed every step); no flush",
                       "norms": "inverse gallery norms recomputed every step inside the top-k kernel (no cached state)",
                       "exchange": (("peer-memory exchange+merge kernel (" + gallery._peer.mapping + ")"
                                     + (", lagged by one search in the value loop" if lagged else ""))
                                    if world > 1 and gallery.transport == "peer" else
                                    ("nccl all-gather + merge kernel" if world > 1 else "none"))},
            "e2e": e2e, "gpu_launches": launches_per_step * a.steps, "clocks": clocks.summary(),
            "roofline": roof, "cpu_baseline": cpu,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
