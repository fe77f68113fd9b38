"""Seeded synthetic inputs shaped like SURVEY.md §8d's configurations (test infrastructure).

CPU generators are bit-reproducible across machines (torch CPU Philox/MT streams are seeded per
call); the large configurations are generated on the device they will be searched on.
"""
from __future__ import annotations

from typing import Tuple

import torch


def _gen(seed: int, device="cpu") -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def unit(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(dim=1, keepdim=True)


def triplets(B: int, D: int, seed: int = 2, scaled: bool = False, device="cpu"
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """C3 triplets: q unit; p = unit(q + 0.02 randn) (cos ~ 0.79 at D=1536, like the positives of
    utils/binary_preds.csv); n: first half unit(randn) (both negative branches inactive), second
    half unit(q + sigma randn) with sigma log-uniform [0.002, 0.2] (straddles d < margin and
    cos > margin for margins 0.2/0.3/0.5).  `scaled` multiplies rows by U(0.5, 4)."""
    g = _gen(seed, device)
    r = lambda *s: torch.randn(*s, generator=g, device=device)
    q = unit(r(B, D))
    p = unit(q + 0.02 * r(B, D))
    sig = torch.exp(torch.empty(B, 1, device=device).uniform_(-6.2146, -1.6094, generator=g))
    n = unit(q + sig * r(B, D))
    h = B // 2
    if h:
        n[:h] = unit(r(h, D))
    if scaled:
        s = lambda: torch.empty(B, 1, device=device).uniform_(0.5, 4.0, generator=g)
        q, p, n = q * s(), p * s(), n * s()
    return q, p, n


def planted_gallery(N: int, D: int, Q: int, k: int, seed: int = 3, dtype=torch.float32,
                    device="cpu", sigmas=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """C2/C4 scheme: iid randn gallery (rows rescaled by U(0.5,2): un-normalised on purpose); for
    each query k planted rows base + sigma_j*sqrt(D)*randn at distinct random positions
    (cos ~ 0.93 / 0.83 / 0.69 for the default sigmas, gaps far above the bf16 tolerance), and
    query = 3.7 * base.  Returns (queries, gallery, planted positions [Q,k] in rank order)."""
    g = _gen(seed, device)
    r = lambda *s: torch.randn(*s, generator=g, device=device)
    gal = r(N, D) * torch.empty(N, 1, device=device).uniform_(0.5, 2.0, generator=g)
    if sigmas is None:
        sigmas = [0.010 + 0.008 * j for j in range(k)]
    base = r(Q, D)
    pos = torch.randperm(N, generator=g, device=device)[: Q * k].view(Q, k)
    for j in range(k):
        gal[pos[:, j]] = base + sigmas[j] * (D ** 0.5) * r(Q, D)
    queries = 3.7 * base
    return queries.to(dtype), gal.to(dtype), pos


def iid_gallery(N: int, D: int, Q: int, seed: int, dtype=torch.float32, device="cpu"):
    g = _gen(seed, device)
    gal = torch.randn(N, D, generator=g, device=device)
    gal = gal * torch.empty(N, 1, device=device).uniform_(0.5, 2.0, generator=g)
    q = torch.randn(Q, D, generator=g, device=device)
    return q.to(dtype), gal.to(dtype)


def tied_gallery(N: int, D: int, Q: int, seed: int = 7, dtype=torch.float32):
    """Gallery with exact duplicate rows (score ties) to exercise the lower-index rule."""
    q, gal = iid_gallery(N, D, Q, seed)
    g = _gen(seed + 1)
    src = torch.randint(0, N, (max(N // 4, 1),), generator=g)
    dst = torch.randint(0, N, (max(N // 4, 1),), generator=g)
    gal[dst] = gal[src]
    # make the best match of every query an exactly duplicated row at two distinct positions
    pairs = min(Q, N // 2)
    where = torch.randperm(N, generator=g)[: 2 * pairs].view(pairs, 2)
    for i in range(pairs):
        gal[where[i, 0]] = q[i] * 0.5
        gal[where[i, 1]] = q[i] * 0.5
    return q.to(dtype), gal.to(dtype)
