"""GPU bring-up: run every kernel once against plain torch on the same device and print diagnostics.

Each stage runs in its own subprocess under a timeout so that a trapped kernel in one stage
does not take the others down:   python scripts/bringup.py [stage ...]
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

STAGES = ["norms", "merge", "f32", "scores_small", "scores", "topk_bf16", "topk_bf16_big", "losses",
          "hits", "timing"]


def ref_topk(q, g, k, eps=1e-6):
    import torch
    qn = q.float() / q.float().norm(dim=1, keepdim=True).clamp_min(eps)
    gn = g.float() / g.float().norm(dim=1, keepdim=True).clamp_min(eps)
    s = qn.double() @ gn.double().T
    v, i = torch.sort(s, dim=1, descending=True, stable=True)
    return v[:, :k].float(), i[:, :k], s


def report(name, ok, msg=""):
    print(f"[{'PASS' if ok else 'FAIL'}] {name} {msg}", flush=True)
    return ok


def stage_norms():
    import torch
    import imageretrievalresearch_b200._ops as ops
    ok = True
    for dt in (torch.float32, torch.bfloat16):
        x = (torch.randn(1000, 1536, device="cuda") * torch.rand(1000, 1, device="cuda") * 3).to(dt)
        x[5] = 0
        got = ops.row_inv_norms(x, 1e-6)
        ref = 1.0 / x.float().norm(dim=1).clamp_min(1e-6)
        err = ((got - ref).abs() / ref).max().item()
        ok &= report(f"row_inv_norms {dt}", err < 1e-5, f"max rel err {err:.2e}")
    return ok


def stage_merge():
    import torch
    import imageretrievalresearch_b200._ops as ops
    ok = True
    for (G, Q, k) in [(8, 100, 3), (2, 7, 10), (4, 33, 1), (8, 64, 16)]:
        vals = torch.randn(G, Q, k, device="cuda").sort(dim=2, descending=True).values
        vals[0, :, 0] = vals[1, :, 0]  # force cross-rank ties
        idx = torch.stack([torch.randperm(1000, device="cuda")[:k].sort().values + g * 1000
                           for g in range(G) for _ in range(Q)]).view(G, Q, k)
        v, i = ops.topk_merge(vals, idx)
        flat_v = vals.permute(1, 0, 2).reshape(Q, G * k)
        flat_i = idx.permute(1, 0, 2).reshape(Q, G * k)
        # order: value desc, idx asc
        order = torch.argsort(flat_i, dim=1, stable=True)
        fv, fi = flat_v.gather(1, order), flat_i.gather(1, order)
        o2 = torch.argsort(fv, dim=1, descending=True, stable=True)
        rv, ri = fv.gather(1, o2)[:, :k], fi.gather(1, o2)[:, :k]
        ok &= report(f"topk_merge G={G} Q={Q} k={k}", torch.equal(v, rv) and torch.equal(i, ri),
                     f"val mismatches {(v != rv).sum().item()} idx mismatches {(i != ri).sum().item()}")
    return ok


def _check_topk(name, q, g, k, tol_abs, eps=1e-6):
    import torch
    import imageretrievalresearch_b200 as irr
    res = irr.cosine_topk(q, g, k, eps)
    torch.cuda.synchronize()
    rv, ri, s = ref_topk(q, g, k, eps)
    verr = (res.values - rv).abs().max().item()
    same = (res.indices == ri)
    # allowed index differences: where the oracle's score gap is below the tolerance
    got_scores = s.gather(1, res.indices.clamp_min(0)).float()
    bad = (~same) & ((got_scores - rv).abs() > tol_abs)
    return report(name, verr <= tol_abs and not bad.any().item(),
                  f"max |dv| {verr:.3e}, idx equal {same.float().mean().item():.4f}, bad {bad.sum().item()}")


def stage_f32():
    import torch
    torch.manual_seed(1)
    ok = True
    g = torch.randn(10000, 1536, device="cuda") * (0.5 + 1.5 * torch.rand(10000, 1, device="cuda"))
    q = torch.randn(64, 1536, device="cuda")
    ok &= _check_topk("f32 10k x 1536 Q=64 k=3", q, g, 3, 1e-5)
    ok &= _check_topk("f32 ragged N=1000 D=100 Q=5 k=10", torch.randn(5, 100, device="cuda"),
                      torch.randn(1000, 100, device="cuda"), 10, 1e-5)
    ok &= _check_topk("f32 Q=200 N=777 D=64 k=1", torch.randn(200, 64, device="cuda"),
                      torch.randn(777, 64, device="cuda"), 1, 1e-5)
    return ok


def _check_scores(Q, N, D):
    import torch
    import imageretrievalresearch_b200._ops as ops
    torch.manual_seed(Q * 7 + N)
    q = torch.randn(Q, D, device="cuda").bfloat16()
    g = (torch.randn(N, D, device="cuda") * (0.5 + torch.rand(N, 1, device="cuda"))).bfloat16()
    got = ops.cosine_scores_bf16(q, g, 1e-6)
    torch.cuda.synchronize()
    _, _, s = ref_topk(q, g, 1)
    err = (got.double() - s).abs()
    ok = err.max().item() < 1e-4
    report(f"bf16 scores Q={Q} N={N} D={D}", ok, f"max abs err {err.max().item():.3e}")
    if not ok:
        bad = (err > 1e-4)
        rows = bad.any(dim=1).nonzero().flatten()[:8].tolist()
        cols = bad.any(dim=0).nonzero().flatten()[:16].tolist()
        print("   bad rows (first 8):", rows, " bad cols (first 16):", cols,
              " frac bad:", bad.float().mean().item())
        print("   got[0,:8]", got[0, :8].tolist())
        print("   ref[0,:8]", s[0, :8].float().tolist())
    return ok


def stage_scores_small():
    ok = _check_scores(128, 256, 64)
    ok &= _check_scores(128, 256, 128)
    return ok


def stage_scores():
    ok = _check_scores(128, 512, 1536)
    ok &= _check_scores(64, 1000, 1536)
    ok &= _check_scores(300, 3000, 1536)
    ok &= _check_scores(1, 257, 200)
    return ok


def stage_topk_bf16():
    import torch
    ok = True
    for (Q, N, D, k) in [(64, 10000, 1536, 3), (1, 5000, 1536, 3), (300, 70000, 1536, 10),
                         (128, 256, 64, 1), (5, 100, 72, 16)]:
        torch.manual_seed(N)
        q = torch.randn(Q, D, device="cuda").bfloat16()
        g = torch.randn(N, D, device="cuda").bfloat16()
        ok &= _check_topk(f"bf16 topk Q={Q} N={N} D={D} k={k}", q, g, k, 1e-4)
    return ok


def stage_topk_bf16_big():
    import torch
    import imageretrievalresearch_b200 as irr
    ok = True
    N, D = 1_000_000, 1536
    torch.manual_seed(3)
    g = torch.randn(N, D, device="cuda", dtype=torch.bfloat16)
    for Q in (1, 64, 4096):
        base = torch.randn(Q, D, device="cuda")
        pos = torch.randperm(N, device="cuda")[: Q * 3].view(Q, 3)
        sig = torch.tensor([0.010, 0.018, 0.026], device="cuda")
        for j in range(3):
            g[pos[:, j]] = (base + sig[j] * (D ** 0.5) * torch.randn(Q, D, device="cuda")).bfloat16()
        q = (3.7 * base).bfloat16()
        res = irr.cosine_topk(q, g, 3)
        torch.cuda.synchronize()
        # planted rows must come back (as a set; their order follows the noise level)
        hit = (res.indices.sort(dim=1).values == pos.sort(dim=1).values).all(dim=1).float().mean().item()
        # exact check of the values against torch on the returned rows
        gv = torch.nn.functional.cosine_similarity(q.float().unsqueeze(1), g[res.indices].float(), dim=2, eps=1e-6)
        verr = (gv - res.values).abs().max().item()
        ok &= report(f"bf16 1M x 1536 Q={Q} planted", hit > 0.999 and verr < 1e-4,
                     f"planted recovered {hit:.4f}, value err {verr:.2e}, top vals {res.values[0].tolist()}")
    return ok


def stage_losses():
    import torch
    import imageretrievalresearch_b200 as irr
    ok = True
    for dt, B, D in [(torch.float32, 4096, 1536), (torch.float32, 64, 1536), (torch.bfloat16, 333, 1920),
                     (torch.float32, 3, 4)]:
        torch.manual_seed(2)
        q = torch.nn.functional.normalize(torch.randn(B, D, device="cuda"), dim=1)
        p = torch.nn.functional.normalize(q + 0.02 * torch.randn(B, D, device="cuda"), dim=1)
        sig = torch.exp(torch.empty(B, 1, device="cuda").uniform_(-6.2, -1.6))
        n = torch.nn.functional.normalize(q + sig * torch.randn(B, D, device="cuda"), dim=1)
        n[: B // 2] = torch.nn.functional.normalize(torch.randn(B // 2, D, device="cuda"), dim=1)
        q, p, n = [(t * 1.7).to(dt) for t in (q, p, n)]
        for margin in (0.2, 0.3, 0.5):
            qr, pr, nr = [t.float().clone().requires_grad_(True) for t in (q, p, n)]
            cel = torch.nn.CosineEmbeddingLoss(margin)
            one = torch.ones(1, device="cuda")
            dis_p = (pr - qr).pow(2).sum(1)
            dis_n = (nr - qr).pow(2).sum(1)
            ref = torch.stack([cel(qr, pr, one), cel(qr, nr, -one), (0.5 * dis_p).mean(),
                               (0.5 * torch.relu(margin - (dis_n + 1e-9).sqrt()).pow(2)).mean()])
            w = torch.tensor([1.0, 0.7, 1.3, 2.0], device="cuda")
            (ref * w).sum().backward()
            out = irr.triplet_losses_fwd_bwd(q, p, n, margin, grad_scale=w.tolist(), pair_scores=True)
            torch.cuda.synchronize()
            lerr = ((out.losses - ref.detach()).abs() / ref.detach().abs().clamp_min(1e-6)).max().item()
            tol_g = 1e-4 if dt == torch.float32 else 2e-2
            gerrs = []
            for got, r in ((out.grad_qry, qr.grad), (out.grad_pos, pr.grad), (out.grad_neg, nr.grad)):
                gerrs.append(((got.float() - r).norm() / r.norm().clamp_min(1e-12)).item())
            pc = torch.nn.functional.cosine_similarity(q.float(), p.float(), dim=1, eps=1e-6)
            pcerr = (out.pair_cos[0] - pc).abs().max().item()
            ok &= report(f"fused triplet {dt} B={B} D={D} m={margin}",
                         lerr < 1e-5 and max(gerrs) < tol_g and pcerr < 1e-5,
                         f"loss rel {lerr:.2e} grads rel {['%.2e' % e for e in gerrs]} paircos {pcerr:.2e} ref {ref.tolist()}")
            # autograd path
            qa, pa, na = [t.clone().requires_grad_(True) for t in (q, p, n)]
            tl = irr.triplet_losses(qa, pa, na, margin)
            tot = tl.cos_pos * w[0] + tl.cos_neg * w[1] + tl.con_pos * w[2] + tl.con_neg * w[3]
            tot.backward()
            gerrs = [((a.grad.float() - r.grad).norm() / r.grad.norm().clamp_min(1e-12)).item()
                     for a, r in ((qa, qr), (pa, pr), (na, nr))]
            ok &= report(f"autograd triplet {dt} B={B} m={margin}", max(gerrs) < tol_g,
                         f"grads rel {['%.2e' % e for e in gerrs]}")
        # drop-in modules
        for lab in (1.0, 0.0):
            a, b = q.float().clone().requires_grad_(True), n.float().clone().requires_grad_(True)
            a2, b2 = q.clone().requires_grad_(True), n.clone().requires_grad_(True)
            dis = (b - a).pow(2).sum(1)
            ref = (0.5 * (lab * dis + (1 - lab) * torch.relu(0.3 - (dis + 1e-9).sqrt()).pow(2))).mean()
            ref.backward()
            got = irr.ContrastiveLoss(0.3)(a2, b2, lab)
            got.backward()
            e = abs(got.item() - ref.item()) / max(abs(ref.item()), 1e-6)
            ge = ((b2.grad.float() - b.grad).norm() / b.grad.norm().clamp_min(1e-12)).item()
            ok &= report(f"ContrastiveLoss {dt} B={B} label={lab}", e < 1e-5 and ge < (1e-4 if dt == torch.float32 else 2e-2),
                         f"loss rel {e:.2e} grad rel {ge:.2e}")
        for tg in (1.0, -1.0):
            a, b = q.float().clone().requires_grad_(True), n.float().clone().requires_grad_(True)
            a2, b2 = q.clone().requires_grad_(True), n.clone().requires_grad_(True)
            t = torch.tensor([tg], device="cuda")
            ref = torch.nn.CosineEmbeddingLoss(0.3)(a, b, t)
            ref.backward()
            got = irr.CosineEmbeddingLoss(0.3)(a2, b2, t)
            got.backward()
            e = abs(got.item() - ref.item()) / max(abs(ref.item()), 1e-6)
            ge = ((a2.grad.float() - a.grad).norm() / a.grad.norm().clamp_min(1e-12)).item()
            ok &= report(f"CosineEmbeddingLoss {dt} B={B} target={tg}", e < 1e-5 and ge < (1e-4 if dt == torch.float32 else 2e-2),
                         f"loss rel {e:.2e} grad rel {ge:.2e}")
    return ok


def stage_hits():
    import torch
    import imageretrievalresearch_b200 as irr
    torch.manual_seed(0)
    Q, N, k = 500, 2000, 3
    idx = torch.randint(0, N, (Q, k), device="cuda")
    ql = torch.randint(0, 8, (Q,), device="cuda")
    gl = torch.randint(0, 8, (N,), device="cuda")
    h = irr.topk_hits(idx, ql, gl)
    m = gl[idx] == ql[:, None]
    ref = torch.stack([m[:, 0].sum(), m.any(dim=1).sum()])
    ok = report("topk_hits class", torch.equal(h, ref), f"{h.tolist()} vs {ref.tolist()}")
    h = irr.topk_hits(idx)
    m = idx == torch.arange(Q, device="cuda")[:, None]
    ref = torch.stack([m[:, 0].sum(), m.any(dim=1).sum()])
    ok &= report("topk_hits instance", torch.equal(h, ref), f"{h.tolist()} vs {ref.tolist()}")
    cs = irr.CosineSimilarity(dim=1, eps=1e-6)
    a, b = torch.randn(1, 1536, device="cuda"), torch.randn(999, 1536, device="cuda")
    e = (cs(a, b) - torch.nn.functional.cosine_similarity(a, b, dim=1, eps=1e-6)).abs().max().item()
    ok &= report("CosineSimilarity broadcast", e < 1e-6, f"{e:.2e}")
    return ok


def _time(fn, iters=10, warm=3):
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


def stage_timing():
    import torch
    import imageretrievalresearch_b200 as irr
    N, D = 1_000_000, 1536
    g = torch.randn(N, D, device="cuda", dtype=torch.bfloat16)
    for Q in (1, 64, 4096):
        q = torch.randn(Q, D, device="cuda", dtype=torch.bfloat16)
        ms = _time(lambda: irr.cosine_topk(q, g, 3))
        fl = 2.0 * Q * N * D / (ms * 1e-3) / 1e12
        bw = N * D * 2 / (ms * 1e-3) / 1e9
        print(f"[TIME] bf16 1Mx1536 Q={Q}: {ms:.3f} ms  {Q / ms * 1e3:.0f} q/s  {fl:.1f} TFLOP/s  {bw:.0f} GB/s", flush=True)
    gal = irr.Gallery(g)
    for Q in (64, 4096):
        q = torch.randn(Q, D, device="cuda", dtype=torch.bfloat16)
        ms = _time(lambda: gal.search(q, 3))
        print(f"[TIME] cached-norm gallery Q={Q}: {ms:.3f} ms  {Q / ms * 1e3:.0f} q/s", flush=True)
    ms = _time(lambda: torch.matmul(q, g[:131072].T), iters=5)
    print(f"[TIME] torch bf16 matmul 4096x131072x1536: {ms:.3f} ms {2.0*4096*131072*1536/(ms*1e-3)/1e12:.1f} TFLOP/s")
    g32 = torch.randn(10000, D, device="cuda")
    q32 = torch.randn(64, D, device="cuda")
    ms = _time(lambda: irr.cosine_topk(q32, g32, 3))
    print(f"[TIME] f32 10kx1536 Q=64: {ms:.3f} ms", flush=True)
    B = 4096
    q, p, n = [torch.randn(B, D, device="cuda") for _ in range(3)]
    ms = _time(lambda: irr.triplet_losses_fwd_bwd(q, p, n, 0.3), iters=20)
    print(f"[TIME] fused triplet fwd+bwd B=4096 fp32: {ms * 1e3:.1f} us  {6 * B * D * 4 / (ms * 1e-3) / 1e9:.0f} GB/s", flush=True)
    return True


def main():
    stages = sys.argv[1:] or STAGES
    if len(stages) == 1 and stages[0].startswith("_run:"):
        name = stages[0][5:]
        ok = globals()["stage_" + name]()
        sys.exit(0 if ok else 1)
    results = {}
    for s in stages:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "_run:" + s], timeout=420)
            results[s] = r.returncode
        except subprocess.TimeoutExpired:
            results[s] = "timeout"
        print(f"== stage {s}: rc={results[s]} ({time.time() - t0:.1f}s)", flush=True)
    print("SUMMARY", results)
    sys.exit(0 if all(v == 0 for v in results.values()) else 1)


if __name__ == "__main__":
    main()
