"""Randomised stress of the fused loss kernel: random (B, D, dtype, margin, reduction, weights)
against torch fp64 autograd of the reference's formulas on the same GPU, for a fixed number of
seconds; also checks that repeated launches return identical bits (deterministic reduction).

    python scripts/stress_losses.py [seconds] [seed]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import imageretrievalresearch_b200 as irr

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
g = torch.Generator(device="cuda").manual_seed(seed)
rnd = torch.Generator().manual_seed(seed)


def ri(lo, hi):
    return int(torch.randint(lo, hi + 1, (1,), generator=rnd).item())


def reference(q, p, n, m, mean, w):
    """utils/contrastive_loss.py:56-61 and torch.nn.CosineEmbeddingLoss (ATen's 1e-12 convention),
    train/train_efficient_cos_con_ce_loss.py:230-237, in fp64 with autograd."""
    q, p, n = [t.double().requires_grad_(True) for t in (q, p, n)]

    def cos(a, b):
        return (a * b).sum(1) / torch.sqrt(((a * a).sum(1) + 1e-12) * ((b * b).sum(1) + 1e-12))

    def con(a, b, y):
        d = ((b - a) ** 2).sum(1)
        s = torch.sqrt(d + 1e-9)
        return 0.5 * (y * d + (1 - y) * torch.clamp(m - s, min=0) ** 2)

    red = (lambda x: x.mean()) if mean else (lambda x: x.sum())
    l = [red(1 - cos(q, p)), red(torch.clamp(cos(q, n) - m, min=0)), red(con(q, p, 1.0)), red(con(q, n, 0.0))]
    (w[0] * l[0] + w[1] * l[1] + w[2] * l[2] + w[3] * l[3]).backward()
    return torch.stack(l).detach(), q.grad, p.grad, n.grad


t0 = time.time()
cases = 0
while time.time() - t0 < secs:
    dt = (torch.float32, torch.bfloat16)[ri(0, 1)]
    B = (ri(1, 40), ri(1, 700), ri(100, 6000))[ri(0, 2)]
    D = 8 * ri(1, 320) if ri(0, 1) else (64, 1536, 1920, 2560)[ri(0, 3)]
    m = (0.2, 0.3, 0.5)[ri(0, 2)]
    mean = bool(ri(0, 1))
    w = tuple(float(x) for x in (0.5 + torch.rand(4, generator=rnd)))
    q = torch.nn.functional.normalize(torch.randn(B, D, device="cuda", generator=g), dim=1)
    p = torch.nn.functional.normalize(q + 0.02 * torch.randn(B, D, device="cuda", generator=g), dim=1)
    sig = torch.exp(torch.empty(B, 1, device="cuda").uniform_(-6.2, -1.6, generator=g))
    n = torch.where(torch.rand(B, 1, device="cuda", generator=g) < 0.5, torch.randn(B, D, device="cuda", generator=g),
                    q + sig * torch.randn(B, D, device="cuda", generator=g))
    if ri(0, 1):
        sc = 0.5 + 3.5 * torch.rand(B, 1, device="cuda", generator=g)
        q, p, n = q * sc, p * sc, n * sc
    q, p, n = q.to(dt), p.to(dt), n.to(dt)
    out = irr.triplet_losses_fwd_bwd(q, p, n, m, mean=mean, grad_scale=w)
    out2 = irr.triplet_losses_fwd_bwd(q, p, n, m, mean=mean, grad_scale=w)
    want, dq, dp, dn = reference(q, p, n, m, mean, w)
    tol_l, tol_g = (1e-5, 1e-4) if dt == torch.float32 else (2e-5, 1.5e-2)
    bad = None
    # absolute slack: 1 - cos for near-duplicates is a cancellation in fp32 (the reference's too)
    if not bool(((out.losses.double() - want).abs() <= tol_l * want.abs() + 2e-7 * (1 if mean else B)).all()):
        bad = ("losses", out.losses.tolist(), want.tolist())
    for name, got, ref in (("dq", out.grad_qry, dq), ("dp", out.grad_pos, dp), ("dn", out.grad_neg, dn)):
        den = ref.abs().max().clamp_min(1e-30)
        if float((got.double() - ref).abs().max() / den) > tol_g:
            bad = (name, float((got.double() - ref).abs().max() / den))
    if not (torch.equal(out.losses, out2.losses) and torch.equal(out.grad_qry, out2.grad_qry)
            and torch.equal(out.grad_neg, out2.grad_neg)):
        bad = ("not deterministic",)
    if bad:
        print("MISMATCH", dict(dtype=str(dt), B=B, D=D, m=m, mean=mean), bad, flush=True)
        sys.exit(1)
    cases += 1
print(f"ok: {cases} random loss cases in {time.time() - t0:.0f} s")
