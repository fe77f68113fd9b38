"""Randomised parity stress: random (Q, N, D, k, dtype, cached / uncached) searches against a torch
fp64 scan on the same GPU, for a fixed number of seconds.  Exits non-zero on the first mismatch.

    python scripts/stress_parity.py [seconds] [seed]
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


def check(q, gal, k, res, tol):
    s = torch.nn.functional.normalize(q.double(), dim=1) @ torch.nn.functional.normalize(gal.double(), dim=1).T
    ov, oi = torch.sort(s, dim=1, descending=True, stable=True)
    kk = min(k, gal.shape[0])
    gv, gi = res.values[:, :kk].double(), res.indices[:, :kk]
    if (gv - ov[:, :kk]).abs().max().item() > tol:
        return "values", float((gv - ov[:, :kk]).abs().max())
    same = gi == oi[:, :kk]
    gap = (s.gather(1, gi.clamp_min(0)) - ov[:, :kk]).abs()
    bad = (~same) & ((gap > tol) | (gi < 0))
    if int(bad.sum()):
        return "indices", int(bad.sum())
    if kk > 1 and not bool((res.values[:, : kk - 1] >= res.values[:, 1:kk]).all()):
        return "order", 0
    return None


t0 = time.time()
n = 0
kinds = {}
while time.time() - t0 < secs:
    dt = (torch.bfloat16, torch.float32, torch.float16)[ri(0, 2)]
    Q = (ri(1, 130), ri(1, 700), ri(1, 64), ri(200, 1300))[ri(0, 3)]
    N = (ri(1, 3000), ri(1000, 60000), ri(256, 9000))[ri(0, 2)]
    D = 8 * ri(1, 40) if ri(0, 1) else (64, 256, 1536, 1920)[ri(0, 3)]
    if dt == torch.float32:
        D = (D + 3) // 4 * 4
        Q, N = min(Q, 300), min(N, 20000)
    k = (ri(1, 16), 3, ri(17, 256), 1)[ri(0, 3)]
    k = min(k, N)
    q = torch.randn(Q, D, device="cuda", generator=g) * (0.2 + 3 * torch.rand(Q, 1, device="cuda", generator=g))
    gal = torch.randn(N, D, device="cuda", generator=g) * (0.2 + 3 * torch.rand(N, 1, device="cuda", generator=g))
    if N >= 8:   # exact duplicates: ties must resolve to the lower index
        src = torch.randint(0, N, (N // 8,), device="cuda", generator=g)
        dst = torch.randint(0, N, (N // 8,), device="cuda", generator=g)
        gal[dst] = gal[src]
    q, gal = q.to(dt), gal.to(dt)
    fp16_tensor = dt == torch.float16 and ri(0, 1) == 1
    tol = 2e-6 if (dt == torch.float32 or (dt == torch.float16 and not fp16_tensor)) else 2e-4
    for cached in (False, True):
        if cached:
            res = irr.Gallery(gal, fp16_tensor_path=fp16_tensor).search(q, k)
        else:
            res = irr.cosine_topk(q, gal, k, fp16_tensor_path=fp16_tensor)
        bad = check(q.float() if dt != torch.float32 else q, gal.float() if dt != torch.float32 else gal, k, res, tol)
        if bad:
            print("MISMATCH", dict(dtype=str(dt), Q=Q, N=N, D=D, k=k, cached=cached, fp16_tensor=fp16_tensor), bad, flush=True)
            sys.exit(1)
    n += 1
    kinds[str(dt)] = kinds.get(str(dt), 0) + 1
print(f"ok: {n} random cases x (uncached, cached) in {time.time() - t0:.0f} s", kinds)
