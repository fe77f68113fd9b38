"""Profiling driver for the bf16 top-k kernels: three searches of Q queries over a 1M x 1536
gallery (run under ncu with -k regex:cosine_topk_bf16 -s 1 -c 1).

    python scripts/prof_topk.py Q [cached] [N] [D] [k]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import imageretrievalresearch_b200 as irr

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cached = len(sys.argv) > 2 and sys.argv[2] == "cached"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
D = int(sys.argv[4]) if len(sys.argv) > 4 else 1536
k = int(sys.argv[5]) if len(sys.argv) > 5 else 3
g = torch.randn(N, D, device="cuda", dtype=torch.bfloat16)
q = torch.randn(Q, D, device="cuda", dtype=torch.bfloat16)
gal = irr.Gallery(g) if cached else None
for _ in range(3):
    r = gal.search(q, k) if cached else irr.cosine_topk(q, g, k)
torch.cuda.synchronize()
print("ok", Q, cached, r.indices[0].tolist())
