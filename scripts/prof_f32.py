"""Profiling driver for the fp32 exactness kernel at the notebook's size (run under ncu with
-k regex:cosine_topk_f32)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import imageretrievalresearch_b200 as irr

Q = N = 8736
D = 1920
q = torch.randn(Q, D, device="cuda")
g = torch.randn(N, D, device="cuda")
for _ in range(3):
    r = irr.cosine_topk(q, g, 3)
torch.cuda.synchronize()
print("ok", r.indices[0].tolist())
