"""The notebook's evaluation size (8736 x 8736 x 1920, k = 150 + class de-dup) for an ncu launch
list:  python scripts/prof_notebook.py [bf16|f32]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import imageretrievalresearch_b200 as irr

dt = torch.float32 if len(sys.argv) > 1 and sys.argv[1] == "f32" else torch.bfloat16
Q = N = 8736
D = 1920
q = torch.randn(Q, D, device="cuda").to(dt)
g = torch.randn(N, D, device="cuda").to(dt)
cls = torch.arange(N, device="cuda") // 70
for _ in range(2):
    r = irr.top1_top3_dedup(q, g, cls, cls, k=150)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    r = irr.top1_top3_dedup(q, g, cls, cls, k=150)
e.record()
torch.cuda.synchronize()
print("ms per evaluation", s.elapsed_time(e) / 5, r)
