"""Single-tile batches over shard-sized galleries (what one of 8 GPUs holds of 1M rows): per-search
time in a CUDA graph, e.g. with IRR_PDL=0 / 1 (programmatic dependent launches along the chain)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import imageretrievalresearch_b200 as irr
from bench import graphed_us

for N in (125_000, 250_000, 500_000):
    g = torch.randn(N, 1536, device="cuda", dtype=torch.bfloat16)
    gal = irr.Gallery(g)
    for Q in (1, 64):
        q = torch.randn(Q, 1536, device="cuda", dtype=torch.bfloat16)
        out = {"N": N, "Q": Q, "pdl": os.environ.get("IRR_PDL", "1")}
        out["cached_us"] = round(min(graphed_us(lambda i: gal.search(q, 3), 10) for _ in range(3)), 1)
        out["uncached_us"] = round(min(graphed_us(lambda i: irr.cosine_topk(q, g, 3), 10) for _ in range(3)), 1)
        out["hbm_us"] = round(N * 1536 * 2 / 6541.1e3, 1)
        print(json.dumps(out), flush=True)
