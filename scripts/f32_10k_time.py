"""configs[1] (10k x 1536 fp32, Q=64, k=3): per-search time in a CUDA graph and the main kernel's own
time (events armed through irr_profile_next_topk), e.g. with IRR_F32_SPLITK=0 / 1."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import imageretrievalresearch_b200 as irr
from bench import graphed_us

lib = irr.load_library()
g = torch.randn(10_000, 1536, device="cuda")
q = torch.randn(64, 1536, device="cuda")
us = min(graphed_us(lambda i: irr.cosine_topk(q, g, 3), 10) for _ in range(3))
ks = []
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); e.record()
    torch.cuda.synchronize()
    lib.irr_profile_next_topk(s.cuda_event, e.cuda_event)
    irr.cosine_topk(q, g, 3)
    torch.cuda.synchronize()
    ks.append(s.elapsed_time(e) * 1e3)
print(json.dumps({"splitk": os.environ.get("IRR_F32_SPLITK", "1"), "graph_us_per_search": round(us, 1),
                  "main_kernel_us_min": round(min(ks), 1), "main_kernel_us_median": round(sorted(ks)[5], 1)}))
