"""configs[1] (10k x 1536 fp32, Q=64, k=3): a few searches for an ncu launch list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import imageretrievalresearch_b200 as irr

g = torch.randn(10_000, 1536, device="cuda")
q = torch.randn(64, 1536, device="cuda")
for _ in range(5):
    r = irr.cosine_topk(q, g, 3)
torch.cuda.synchronize()
print("ok", r.indices[0].tolist())
