"""Profiling driver for the fused loss kernel: a few eager launches per dtype over rotating input
sets (run under ncu with -k regex:loss_fwd_bwd).

    python scripts/prof_loss.py [f32|bf16] [B]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import imageretrievalresearch_b200 as irr

which = sys.argv[1] if len(sys.argv) > 1 else "f32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dt = torch.float32 if which == "f32" else torch.bfloat16
sets = [[torch.randn(B, 1536, device="cuda").to(dt) for _ in range(3)] for _ in range(6)]
for i in range(12):
    out = irr.triplet_losses_fwd_bwd(*sets[i % 6], 0.3)
torch.cuda.synchronize()
print(which, B, [float(x) for x in out.losses] if hasattr(out, "losses") else "ok")
