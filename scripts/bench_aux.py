"""Secondary measurements (not the headline bench line): the fused loss kernel (BASELINE.json
configs[2]), the fp32 exactness config (configs[1]) and the small-Q searches, each as one JSON line.

    python scripts/bench_aux.py [losses] [f32] [smallq]
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import imageretrievalresearch_b200 as irr

PEAKS = {"hbm_gbs": 6541.1}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAKS = json.load(open(p))


def timed(fn, iters, warm=5):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def graphed(fn, calls):
    """Capture `calls` invocations into one CUDA graph (the python wrapper costs more host time
    than these kernels take on the device) and return ms per invocation."""
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
        for _ in range(5):
            g.replay()
        e.record(st)
        st.synchronize()
    return s.elapsed_time(e) / (5 * calls)


def losses():
    B, D = 4096, 1536
    for dt in (torch.float32, torch.bfloat16):
        # 151 MB per call at fp32 is about the size of L2: rotate over 6 input sets (> 2x L2)
        sets = [[torch.randn(B, D, device="cuda").to(dt) for _ in range(3)] for _ in range(6)]
        ms = graphed(lambda i: irr.triplet_losses_fwd_bwd(*sets[i % 6], 0.3), 24)
        by = 6 * B * D * sets[0][0].element_size()
        print(json.dumps({"what": "fused triplet losses fwd+bwd", "B": B, "D": D, "dtype": str(dt),
                          "us": ms * 1e3, "algorithmic_bytes": by, "GBps": by / ms / 1e6,
                          "hbm_frac": by / ms / 1e6 / PEAKS["hbm_gbs"],
                          "triplets_per_s": B / (ms * 1e-3), "l2": "6 rotating input sets", "timing": "CUDA graph of 24 launches"}), flush=True)
        q, p_, n = [t.clone().requires_grad_(True) for t in sets[0]]

        def ag(i):
            tl = irr.triplet_losses(q, p_, n, 0.3)
            (tl.loss_cos + tl.loss_con).backward()
            q.grad = p_.grad = n.grad = None
        ms = timed(ag, 30)
        print(json.dumps({"what": "autograd triplet losses fwd + bwd (2 kernels + torch glue)", "B": B,
                          "dtype": str(dt), "us": ms * 1e3}), flush=True)
    # the reference's op-by-op path on the same GPU (torch CUDA), for context
    q, p_, n = [torch.randn(B, D, device="cuda").requires_grad_(True) for _ in range(3)]
    cel = torch.nn.CosineEmbeddingLoss(0.3)
    one = torch.ones(1, device="cuda")

    def torch_path(i):
        l = cel(q, p_, one) + cel(q, n, -one)
        d1 = (p_ - q).pow(2).sum(1)
        d0 = (n - q).pow(2).sum(1)
        l = l + (0.5 * d1).mean() + (0.5 * torch.relu(0.3 - (d0 + 1e-9).sqrt()).pow(2)).mean()
        l.backward()
        q.grad = p_.grad = n.grad = None
    ms = timed(torch_path, 30)
    print(json.dumps({"what": "torch CUDA op-by-op losses fwd+bwd (reference's calls on the GPU)",
                      "B": B, "us": ms * 1e3}), flush=True)


def f32():
    g = torch.randn(10_000, 1536, device="cuda")
    q = torch.randn(64, 1536, device="cuda")
    ms = timed(lambda i: irr.cosine_topk(q, g, 3), 50)
    print(json.dumps({"what": "fp32 cosine top-3, 10k x 1536, Q=64 (configs[1])", "us": ms * 1e3,
                      "queries_per_s": 64 / (ms * 1e-3)}), flush=True)
    cos = torch.nn.CosineSimilarity(dim=1, eps=1e-6)

    def loop(i):
        for j in range(64):
            torch.topk(cos(q[j].unsqueeze(0), g), 3)
    ms = timed(loop, 5, warm=2)
    print(json.dumps({"what": "reference loop on the same GPU (torch CUDA), 10k x 1536, Q=64",
                      "us": ms * 1e3, "queries_per_s": 64 / (ms * 1e-3)}), flush=True)


def smallq():
    N, D = 1_000_000, 1536
    g = torch.randn(N, D, device="cuda", dtype=torch.bfloat16)
    gal = irr.Gallery(g)
    for Q in (1, 8, 64, 128, 256, 1024, 4096):
        q = torch.randn(Q, D, device="cuda", dtype=torch.bfloat16)
        ms = timed(lambda i: irr.cosine_topk(q, g, 3), 20)
        msc = timed(lambda i: gal.search(q, 3), 20)
        by = N * D * 2 + Q * D * 2 + Q * 36
        print(json.dumps({"what": "bf16 cosine top-3 1M x 1536", "Q": Q, "ms": ms, "ms_cached_norms": msc,
                          "queries_per_s": Q / (ms * 1e-3), "GBps": by / ms / 1e6,
                          "TFLOPs": 2.0 * Q * N * D / ms / 1e9}), flush=True)




def config5():
    """One GPU's share of BASELINE.json configs[4]: 10M x 2560 bf16 over 8 GPUs -> 1.25M rows per
    GPU, Q=8192, k=10."""
    N, D, Q, k = 1_250_000, 2560, 8192, 10
    g = torch.randn(N, D, device="cuda", dtype=torch.bfloat16)
    q = torch.randn(Q, D, device="cuda", dtype=torch.bfloat16)
    ms = timed(lambda i: irr.cosine_topk(q, g, k), 5, warm=2)
    fl = 2.0 * Q * N * D
    print(json.dumps({"what": "bf16 cosine top-10, 1.25M x 2560 shard, Q=8192 (configs[4] per GPU)",
                      "ms": ms, "TFLOPs": fl / ms / 1e9, "frac_burst": fl / ms / 1e9 / PEAKS["bf16_tflops"],
                      "queries_per_s_per_gpu_shard": Q / (ms * 1e-3)}), flush=True)


def notebook():
    """The reference notebook's evaluation size (inference/training_analysis.ipynb:277): 8736 queries
    against the 8736 positives, D=1920 (rexnet_150), top-150 + class de-dup + top1/top3."""
    Q = N = 8736
    D = 1920
    q = torch.randn(Q, D, device="cuda")
    g = torch.randn(N, D, device="cuda")
    cls = torch.arange(N, device="cuda") // 70
    for dt in (torch.float32, torch.bfloat16):
        qq, gg = q.to(dt), g.to(dt)
        ms = timed(lambda i: irr.top1_top3_dedup(qq, gg, cls, cls, k=150), 5, warm=2)
        print(json.dumps({"what": "notebook evaluation 8736 x 8736 x 1920, k=150 + class de-dup + top1/top3",
                          "dtype": str(dt), "ms": ms, "queries_per_s": Q / (ms * 1e-3)}), flush=True)
        ms3 = timed(lambda i: irr.top1_top3(qq, gg, cls, cls, k=3), 5, warm=2)
        print(json.dumps({"what": "same size, fused k=3 path", "dtype": str(dt), "ms": ms3}), flush=True)


def losses_big():
    """The loss kernel away from launch/ramp effects: 64k triplets (2.4 GB of traffic per launch)."""
    B, D = 65536, 1536
    q, p_, n = [torch.randn(B, D, device="cuda") for _ in range(3)]
    ms = graphed(lambda i: irr.triplet_losses_fwd_bwd(q, p_, n, 0.3), 4)
    by = 6 * B * D * 4
    print(json.dumps({"what": "fused triplet losses fwd+bwd", "B": B, "D": D, "dtype": "torch.float32",
                      "us": ms * 1e3, "GBps": by / ms / 1e6, "hbm_frac": by / ms / 1e6 / PEAKS["hbm_gbs"]}),
          flush=True)


def pool():
    """get_fm (global average pool) and the CE pair at the training batch and at a large batch."""
    for B in (64, 2048):
        fm = torch.randn(B, 1536, 7, 7, device="cuda")
        ms = graphed(lambda i: irr.get_fm(fm), 10)
        by = fm.numel() * 4 + B * 1536 * 4
        print(json.dumps({"what": "get_fm global average pool fwd", "shape": list(fm.shape), "us": ms * 1e3,
                          "GBps": by / ms / 1e6, "hbm_frac": by / ms / 1e6 / PEAKS["hbm_gbs"]}), flush=True)
        ms = timed(lambda i: torch.nn.functional.adaptive_avg_pool2d(fm, 1).reshape(B, 1536), 20)
        print(json.dumps({"what": "torch AvgPool on the same GPU", "shape": list(fm.shape), "us": ms * 1e3}),
              flush=True)


def train_step():
    """The path's share of ONE training / validation step at the reference's own batch size
    (train/train_efficient_cos_con_ce_loss.py:230-287,374-405; B=64 pooled 1536-d embeddings, fp32):
    the reference's calls on the same GPU (four loss modules + backward, then the per-row
    cos/topk loop with its host syncs, plus the validation step's paired-score loops) against
    triplet_losses + backward + top1_top3 (+ the pair scores the fused loss already returns)."""
    import torch.nn as nn
    B, D = 64, 1536
    torch.manual_seed(0)
    mk = lambda: torch.randn(B, D, device="cuda", requires_grad=True)
    fm_ims, fm_poss, fm_negs = mk(), mk(), mk()
    clss = (torch.arange(B, device="cuda") % 8)
    cos = nn.CosineSimilarity(dim=1, eps=1e-6)
    cel = nn.CosineEmbeddingLoss(margin=0.3)
    one, mone = torch.ones(1, device="cuda"), -torch.ones(1, device="cuda")

    def con(a, b, y, m=0.3):      # utils/contrastive_loss.py:56-61
        d = (b - a).pow(2).sum(1)
        return (0.5 * (y * d + (1 - y) * torch.relu(m - (d + 1e-9).sqrt()).pow(2))).mean()

    def ref_train(i):
        loss = cel(fm_ims, fm_poss, one) + cel(fm_ims, fm_negs, mone) + con(fm_ims, fm_poss, 1.0) + con(fm_ims, fm_negs, 0.0)
        loss.backward()
        top3 = top1 = 0
        for idx in range(B):
            sim = cos(fm_ims[idx].unsqueeze(0), fm_poss)
            vals, inds = torch.topk(sim, k=3)
            if clss[idx] == clss[inds[0]] or clss[idx] == clss[inds[1]] or clss[idx] == clss[inds[2]]:
                top3 += 1
            if clss[idx] in clss[inds[0]]:
                top1 += 1
        fm_ims.grad = fm_poss.grad = fm_negs.grad = None
        return top1, top3

    def ref_val(i):
        with torch.no_grad():
            cel(fm_ims, fm_poss, one); cel(fm_ims, fm_negs, mone); con(fm_ims, fm_poss, 1.0); con(fm_ims, fm_negs, 0.0)
            sims, unsims, top3, top1 = [], [], 0, 0
            for idx in range(B):
                sims.append(cos(fm_ims[idx].unsqueeze(0), fm_poss[idx].unsqueeze(0)))
                unsims.append(cos(fm_ims[idx].unsqueeze(0), fm_negs[idx].unsqueeze(0)))
                sim = cos(fm_ims[idx].unsqueeze(0), fm_poss)
                vals, inds = torch.topk(sim, k=3)
                if clss[idx] == clss[inds[0]] or clss[idx] == clss[inds[1]] or clss[idx] == clss[inds[2]]:
                    top3 += 1
                if clss[idx] in clss[inds[0]]:
                    top1 += 1
            return torch.mean(torch.FloatTensor(sims)).item(), torch.mean(torch.FloatTensor(unsims)).item()

    def ours_train(i):
        tl = irr.triplet_losses(fm_ims, fm_poss, fm_negs, 0.3)
        (tl.loss_cos + tl.loss_con).backward()
        t1, t3, _ = irr.top1_top3(fm_ims, fm_poss, clss, clss, k=3)
        fm_ims.grad = fm_poss.grad = fm_negs.grad = None
        return t1, t3

    def ours_val(i):
        with torch.no_grad():
            tl = irr.triplet_losses(fm_ims, fm_poss, fm_negs, 0.3, pair_scores=True)
            t1, t3, _ = irr.top1_top3(fm_ims, fm_poss, clss, clss, k=3)
            return tl.pair_cos_pos.mean().item(), tl.pair_cos_neg.mean().item()

    for name, fn, n in (("reference calls on the GPU, training step", ref_train, 10),
                        ("this package, training step", ours_train, 200),
                        ("reference calls on the GPU, validation step", ref_val, 10),
                        ("this package, validation step", ours_val, 200)):
        fn(0)
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        for i in range(n):
            fn(i)
        torch.cuda.synchronize()
        us = (time.perf_counter() - t0) / n * 1e6
        print(json.dumps({"what": "hot-path share of one step, B=64 x 1536 fp32 (wall clock, host syncs included)",
                          "who": name, "us": us}), flush=True)


def streamed():
    """Host-resident gallery scanned block by block (StreamedGallery): bound by the host link."""
    N, D = 1_000_000, 1536
    host = torch.randn(N, D).to(torch.bfloat16).pin_memory()
    for Q, blk in ((64, 1 << 17), (4096, 1 << 17), (4096, 1 << 15)):
        sg = irr.StreamedGallery(host, blk, "cuda", buffers=3)
        q = torch.randn(Q, D, device="cuda", dtype=torch.bfloat16)
        ms = timed(lambda i: sg.search(q, 3), 5, warm=2)
        by = N * D * 2
        print(json.dumps({"what": "streamed scan of a pinned host gallery 1M x 1536 bf16, top-3", "Q": Q,
                          "block_rows": blk, "ms": ms, "host_link_GBps": by / ms / 1e6,
                          "queries_per_s": Q / (ms * 1e-3)}), flush=True)
    # plain pinned->device copy of the same bytes: the link's own ceiling on this box
    dst = torch.empty(N, D, dtype=torch.bfloat16, device="cuda")
    ms = timed(lambda i: dst.copy_(host, non_blocking=True), 5, warm=2)
    print(json.dumps({"what": "cudaMemcpyAsync pinned->device of the same 3.07 GB", "ms": ms,
                      "host_link_GBps": N * D * 2 / ms / 1e6}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["losses", "f32", "smallq"]
    for w in which:
        globals()[w]()
