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

After the timed regions the line is VERIFIED: the merged top-k of 16 sampled queries from a plain
search is compared with a chunked torch fp32 scan of every rank's shard (all-gathered and merged in
torch) — `"verified": true`, or the process exits non-zero.  The line also carries the other
configurations BASELINE.json names (`sweep`: Q=1 / Q=64 as fractions of the HBM roofline; `losses`:
the fused loss kernel at 4096 x 1536 fp32 and bf16; `fp32_10k`: configs[1]) and `torch_gpu_baseline`:
torch's own CUDA path on the same GPU (batched bf16 cuBLAS + torch.topk, and the reference's literal
per-query loop).

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
    ap.add_argument("--no-sweep", action="store_true",
                    help="skip the extra keys (Q=1/64, losses, fp32_10k, torch_gpu_baseline)")
    return ap.parse_args()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained"), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback"}


def workload_name(a, world, transport=None):
    how = ""
    if world > 1:
        how = {"peer": " + peer-memory exchange/merge kernel over NVLink",
               "collective": " + NCCL all-gather and merge kernel"}.get(transport, " + candidate exchange and merge")
    return (f"cosine top-{a.k} over a {a.rows}x{a.dim} bf16 gallery, Q={a.queries} per step, "
            f"gallery row-sharded over {world} GPU(s)" + how)


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
def graphed_us(fn, calls, replays=5):
    """`calls` invocations of fn(i) captured into one CUDA graph; microseconds per invocation on the
    device (the python wrapper costs more host time than the small kernels take)."""
    import torch
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for i in range(calls):
            fn(i)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for i in range(calls):
                fn(i)
        for _ in range(3):
            g.replay()
        st.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(st)
        for _ in range(replays):
            g.replay()
        e.record(st)
        st.synchronize()
    return s.elapsed_time(e) / (replays * calls) * 1e3


def event_ms(fn, iters, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def verify_against_torch(irr, dist, search_plain, queries, shard, lo, N, k, world, dev, n_check=16):
    """Outside the timed regions: the merged top-k of `n_check` sampled queries against a chunked
    torch fp32 scan of this rank's shard, all-gathered over the ranks and merged in torch
    (normalize + matmul + topk: torch's own restatement of the reference's cos + topk,
    train/train_efficient_cos_con_ce_loss.py:273-276).  Indices must be identical except where the
    torch score gap is below 1e-4 (the kernels are within ~3e-6 of fp64 on bf16 inputs; iid
    galleries have near-ties); values within the bf16-mode bar of 2e-2 absolute."""
    import torch
    import torch.nn.functional as F
    Q = queries.shape[0]
    sel = torch.linspace(0, Q - 1, steps=min(n_check, Q), device=dev).long()
    res = search_plain(queries)
    got_v, got_i = res.values[sel], res.indices[sel]
    qs = F.normalize(queries[sel].float(), dim=1, eps=1e-6)
    kk = min(k + 2, max(shard.shape[0], 1))
    best_v = torch.full((len(sel), kk), -float("inf"), device=dev)
    best_i = torch.full((len(sel), kk), -1, dtype=torch.int64, device=dev)
    for b0 in range(0, shard.shape[0], 131072):
        blk = F.normalize(shard[b0:b0 + 131072].float(), dim=1, eps=1e-6)
        v, i = torch.topk(qs @ blk.T, min(kk, blk.shape[0]), dim=1)
        allv, alli = torch.cat([best_v, v], 1), torch.cat([best_i, i + lo + b0], 1)
        o = torch.argsort(allv, dim=1, descending=True, stable=True)[:, :kk]
        best_v, best_i = allv.gather(1, o), alli.gather(1, o)
    if world > 1:
        gv = [torch.empty_like(best_v) for _ in range(world)]
        gi = [torch.empty_like(best_i) for _ in range(world)]
        dist.all_gather(gv, best_v)
        dist.all_gather(gi, best_i)
        allv, alli = torch.cat(gv, 1), torch.cat(gi, 1)
        o = torch.argsort(allv, dim=1, descending=True, stable=True)[:, :kk]
        best_v, best_i = allv.gather(1, o), alli.gather(1, o)
    want_v, want_i = best_v[:, :k], best_i[:, :k]
    val_err = (got_v - want_v).abs().max().item()
    same = got_i == want_i
    # a differing index is fine only where torch's own candidates are within 1e-4 of each other
    near = torch.zeros_like(same)
    for j in range(k):
        hit = (best_i == got_i[:, j:j + 1])
        picked = torch.where(hit, best_v, torch.full_like(best_v, -float("inf"))).max(dim=1).values
        near[:, j] = (picked - want_v[:, j]).abs() < 1e-4
    bad = int((~same & ~near).sum().item())
    ok = bad == 0 and val_err < 2e-2
    if world > 1:
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
    return {"verified": ok, "queries_checked": int(len(sel)), "max_abs_score_err": val_err,
            "indices_identical_frac": same.float().mean().item(), "bad_indices": bad,
            "against": "chunked torch fp32 normalize+matmul+topk over every rank's shard, merged in torch"}


def torch_gpu_baseline(a, queries, shard, ours_ms):
    """torch's own CUDA path on the same GPU: (1) the batched restatement in bf16 — normalise, cuBLAS
    GEMM per gallery chunk, torch.topk, running merge — with the gallery normalised every step like
    the timed search does; (2) the reference's literal per-query loop
    (inference/training_analysis.ipynb:238, train_efficient_cos_con_ce_loss.py:273-276) on fp32
    embeddings for 16 sampled queries."""
    import torch
    import torch.nn.functional as F
    Q, k = queries.shape[0], a.k
    chunk = 65536

    def batched():
        qn = F.normalize(queries.float(), dim=1, eps=1e-6).bfloat16()
        bv = bi = None
        for b0 in range(0, shard.shape[0], chunk):
            gn = F.normalize(shard[b0:b0 + chunk].float(), dim=1, eps=1e-6).bfloat16()
            v, i = torch.topk(qn @ gn.T, k, dim=1)
            i = i + b0
            if bv is None:
                bv, bi = v, i
            else:
                allv, alli = torch.cat([bv, v], 1), torch.cat([bi, i], 1)
                o = torch.topk(allv, k, dim=1).indices
                bv, bi = allv.gather(1, o), alli.gather(1, o)
        return bv, bi

    ms_b = event_ms(batched, 3, warm=1)
    g32 = shard.float()
    cos = torch.nn.CosineSimilarity(dim=1, eps=1e-6)
    q32 = queries[:16].float()

    def loop():
        for j in range(q32.shape[0]):
            torch.topk(cos(q32[j].unsqueeze(0), g32), k)

    ms_l = event_ms(loop, 2, warm=1) / q32.shape[0]
    del g32
    torch.cuda.empty_cache()
    return {"batched_bf16_cublas_topk": {"ms_per_step": ms_b, "queries_per_s": Q / (ms_b * 1e-3),
                                         "what": f"normalize + bf16 matmul + torch.topk per {chunk}-row chunk, gallery normalised every step"},
            "reference_loop_same_gpu": {"ms_per_query": ms_l, "queries_per_s": 1e3 / ms_l,
                                        "what": "per-query CosineSimilarity(dim=1,eps=1e-6)+torch.topk on fp32 embeddings, 16 sampled queries"},
            "x_over_torch_cuda_batched": ms_b / ours_ms,
            "x_over_reference_loop_same_gpu": (ms_l * Q) / ours_ms}


def extra_configs(irr, a, queries, shard, peaks, search):
    """The other configurations BASELINE.json names, each a few milliseconds of GPU time."""
    import torch
    D, k, n_local = a.dim, a.k, shard.shape[0]
    out = {}

    def settle():
        """The headline loop leaves the board at its power cap with the SM clock shed to ~1300 MHz,
        and the governor takes a moment to give it back: a 3 ms measurement started right after
        inherits that state (the loss figures came out 30 % slow on some boxes, while the same
        process-fresh measurement next to it did not).  Each configuration starts from idle."""
        torch.cuda.synchronize()
        time.sleep(0.5)

    settle()
    sweep = []
    for qs in (1, 64):
        qq = queries[:qs].contiguous()
        ms = event_ms(lambda: search(qq), a.steps)
        b = n_local * D * 2 + qs * D * 2 + qs * k * 12
        sweep.append({"Q": qs, "ms_per_step": ms, "queries_per_s": qs / (ms * 1e-3),
                      "hbm_frac_of_step": b / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                      "algorithmic_bytes": b})
    out["sweep"] = sweep
    # the same Q as the headline against a Gallery handle (inverse norms cached once, as a resident
    # index would hold them): what recomputing the norms inside the kernel every step costs
    settle()
    handle = irr.Gallery(shard)
    ms = event_ms(lambda: handle.search(queries, k), a.steps, warm=a.warmup)
    Q = queries.shape[0]
    out["cached_norms"] = {"Q": Q, "ms_per_step": ms, "queries_per_s": Q / (ms * 1e-3),
                           "tensor_frac_of_step": 2.0 * Q * n_local * D / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"],
                           "what": "headline search through irr.Gallery (1/max(|g|,eps) cached at construction)"}
    del handle
    # configs[2]: fused contrastive + cosine-embedding losses fwd+bwd, 4096 x 1536 triplets.
    # 151 MB (fp32) per call is about the size of L2: rotate over 6 input sets (> 2x L2).
    B, LD = 4096, 1536
    losses = []
    for dt in (torch.float32, torch.bfloat16):
        sets = [[torch.randn(B, LD, device=queries.device).to(dt) for _ in range(3)] for _ in range(6)]
        settle()
        us = min(graphed_us(lambda i: irr.triplet_losses_fwd_bwd(*sets[i % 6], 0.3), 24) for _ in range(3))
        by = 6 * B * LD * sets[0][0].element_size()
        losses.append({"dtype": str(dt).replace("torch.", ""), "B": B, "D": LD, "us": us,
                       "algorithmic_bytes": by, "GBps": by / us / 1e3,
                       "hbm_frac": by / us / 1e3 / peaks["hbm_gbs"], "triplets_per_s": B / (us * 1e-6),
                       "timing": "CUDA graph of 24 launches over 6 rotating input sets, best of 3, from idle"})
        del sets
    out["losses"] = losses
    # configs[1]: fp32 exactness path, 10k x 1536, Q=64, k=3 (gallery fits L2: report time)
    g32 = torch.randn(10_000, 1536, device=queries.device)
    q32 = torch.randn(64, 1536, device=queries.device)
    settle()
    us = min(graphed_us(lambda i: irr.cosine_topk(q32, g32, 3), 10) for _ in range(3))
    out["fp32_10k"] = {"Q": 64, "N": 10_000, "D": 1536, "k": 3, "us": us, "queries_per_s": 64 / (us * 1e-6),
                       "timing": "CUDA graph of 10 searches, best of 3, from idle"}
    return out


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
    gallery = None
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
    transport = gallery.transport if gallery is not None else None
    # N > 1 with the peer-memory exchange: the stream of searches runs lagged (each search's
    # rendezvous is with the peers' PREVIOUS push, ShardedGallery.search_lagged), so a step does
    # not cost the slowest of N kernels; every search's merged result is still produced inside the
    # timed region (the last one by flush()).  `value_plain` times the plain search (every step
    # ends in a rendezvous) in a second region; IRR_BENCH_LAGGED=0 makes that one the headline.
    can_lag = world > 1 and transport == "peer" and k <= 16
    lagged = can_lag and os.environ.get("IRR_BENCH_LAGGED", "1") != "0"
    prof_every = int(os.environ.get("IRR_BENCH_PROF_EVERY", "1"))
    # NVML is initialised and the sampler thread is running BEFORE the ranks line up: anything a
    # rank does between the barrier and its first launch is waited for by every other rank in the
    # first exchange and would be charged to all of their timers
    sampler = ClockSampler(local, enabled=(rank == 0))

    def timed_region(use_lagged, profile):
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(a.steps)]
        for s, e in kev:            # force creation of the raw cudaEvent_t handles
            s.record(); e.record()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(); stop.record()          # create the raw handles ahead of the barrier, too
        barrier()
        start.record()
        for j, (s, e) in enumerate(kev):
            if profile and prof_every > 0 and j % prof_every == 0:
                lib.irr_profile_next_topk(s.cuda_event, e.cuda_event)
            if use_lagged:
                gallery.search_lagged(queries, k)
            else:
                search(queries)
        if use_lagged:
            gallery.flush()
        stop.record()
        barrier()
        ms_step = max_over_ranks(start.elapsed_time(stop)) / a.steps
        timed = [(s, e) for j, (s, e) in enumerate(kev) if profile and prof_every > 0 and j % prof_every == 0]
        kernel_ms = sum(s.elapsed_time(e) for s, e in timed) / len(timed) if timed else ms_step
        return ms_step, max_over_ranks(kernel_ms)

    with sampler as clocks:
        ms_step, kernel_ms = timed_region(lagged, True)
        ms_other = timed_region(not lagged, False)[0] if can_lag else None
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

    # ---- the result itself, outside the timed regions, at every N ----
    verified = verify_against_torch(irr, dist, search, queries, shard, lo, N, k, world, dev)

    # ---- roofline of the dominant kernel ----
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
    kname = "cosine_topk_bf16_pair_kernel" if Q > 128 else "cosine_topk_bf16_kernel"
    roof.update({"kernel": kname, "kernel_ms": kernel_ms, "peak_source": peaks["source"],
                 "kernel_share_of_step": kernel_ms / ms_step,
                 "algorithmic": {"flops": flops, "bytes": bytes_alg}})

    extra = {}
    if not a.no_sweep and world == 1:
        extra = extra_configs(irr, a, queries, shard, peaks, search)
        extra["torch_gpu_baseline"] = torch_gpu_baseline(a, queries, shard, ms_step)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline(a, shard)
    if world > 1:
        dist.barrier()

    if rank == 0:
        # kernels per step: the top-k kernel (gallery norms produced inside it) and the partial-list
        # merge, + for N > 1 the fused exchange/merge kernel (lagged: merge-of-previous + push)
        launches_per_step = 2 + ((2 if lagged else 1) if world > 1 else 0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(a, world, transport), "Q": Q, "N": N, "D": D, "k": k,
                       "parallelism": f"gallery rows sharded x{world}",
                       "l2": "inputs larger than L2 (gallery shard streamed every step); no flush",
                       "norms": "inverse gallery norms recomputed every step inside the top-k kernel (no cached state)",
                       "exchange": (("peer-memory exchange+merge kernel (" + gallery._peer.mapping + ")"
                                     + (", lagged by one search in the value loop" if lagged else ""))
                                    if world > 1 and transport == "peer" else
                                    ("nccl all-gather + merge kernel" if world > 1 else "none"))},
            "e2e": e2e, "gpu_launches": launches_per_step * a.steps, "clocks": clocks.summary(),
            "roofline": roof, "cpu_baseline": cpu,
        }
        line.update(verified)
        if can_lag:
            ms_l, ms_p = (ms_step, ms_other) if lagged else (ms_other, ms_step)
            line["value_lagged"] = Q / (ms_l * 1e-3)
            line["value_plain"] = Q / (ms_p * 1e-3)
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not verified["verified"]:
        print(f"bench.py: rank {rank}: result verification FAILED: {verified}", file=sys.stderr, flush=True)
        sys.exit(3)


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
