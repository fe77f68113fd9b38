"""Probe: how exact is an fp32 dot product computed on the bf16 tensor-core path from three-way
bf16 splits of the operands (x = hi + mid + lo, six cross terms, fp32 accumulation in TMEM)?
Expands q [Q,D] and g [N,D] to [.,6D] bf16 with torch, runs the dense-score kernel on them and
compares the recovered dot products with fp64.

    python scripts/split_bf16_probe.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from imageretrievalresearch_b200 import _ops


def split3(x):
    hi = x.to(torch.bfloat16)
    r1 = x - hi.float()
    mid = r1.to(torch.bfloat16)
    r2 = r1 - mid.float()
    lo = r2.to(torch.bfloat16)
    return hi, mid, lo


def expand(x, side, small_first):
    hi, mid, lo = split3(x)
    # pairs (a_part, b_part) whose products are kept: everything down to 2^-16 relative
    pairs = [("hi", "hi"), ("hi", "mid"), ("mid", "hi"), ("hi", "lo"), ("lo", "hi"), ("mid", "mid")]
    if small_first:
        pairs = pairs[::-1]
    parts = {"hi": hi, "mid": mid, "lo": lo}
    return torch.cat([parts[p[side]] for p in pairs], dim=1).contiguous()


def main():
    torch.manual_seed(0)
    dev = "cuda"
    for D in (1536, 1920):
        N, Q = 4096, 128
        g = torch.randn(N, D, device=dev) * (torch.rand(N, 1, device=dev) * 1.5 + 0.5)
        q = torch.randn(Q, D, device=dev)
        # planted near-duplicates: cos ~ 0.99 / 0.9 (monotone running sums: worst case for a
        # truncating accumulator)
        g[:Q] = q * 2.3 + 0.1 * torch.randn(Q, D, device=dev)
        g[Q:2 * Q] = q.abs() * 1.7          # all-positive products with |q|
        ref = (q.double() @ g.double().T)
        scale = q.double().norm(dim=1)[:, None] * g.double().norm(dim=1)[None, :]
        f32 = (q @ g.T).double()            # torch's own fp32 GEMM (TF32 off by default)
        for small_first in (False, True):
            qe, ge = expand(q, 0, small_first), expand(g, 1, small_first)
            s = _ops.cosine_scores_bf16(qe, ge, 1e-12).double()
            qi = _ops.row_inv_norms(qe, 1e-12).double()
            gi = _ops.row_inv_norms(ge, 1e-12).double()
            dots = s / (qi[:, None] * gi[None, :])
            err = (dots - ref).abs() / scale
            errp = err[:, :2 * Q]
            print(f"D={D} small_first={small_first}: split-bf16 max {err.max().item():.3e} "
                  f"mean {err.mean().item():.3e} | planted max {errp.max().item():.3e} "
                  f"mean signed {((dots - ref) / scale)[:, :2*Q].diagonal().mean().item():.3e}")
        e32 = (f32 - ref).abs() / scale
        print(f"D={D} torch fp32 matmul: max {e32.max().item():.3e} mean {e32.mean().item():.3e}")
        # abs-vector pair: q.abs() vs g row = |q|*1.7: every product positive
        qa = q.abs()
        refa = (qa.double() * g[Q:2 * Q].double()).sum(1)
        sa = _ops.cosine_scores_bf16(expand(qa, 0, True), expand(g[Q:2 * Q].contiguous(), 1, True), 1e-12).double()
        qi = _ops.row_inv_norms(expand(qa, 0, True), 1e-12).double()
        gi = _ops.row_inv_norms(expand(g[Q:2 * Q].contiguous(), 1, True), 1e-12).double()
        da = (sa / (qi[:, None] * gi[None, :])).diagonal()
        rel = (da - refa) / refa
        print(f"D={D} all-positive products: signed rel err mean {rel.mean().item():.3e} "
              f"min {rel.min().item():.3e} max {rel.max().item():.3e}")


if __name__ == "__main__":
    main()
