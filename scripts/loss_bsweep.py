"""Fused loss kernel time against the number of triplets (fixed cost vs streaming rate).

    python scripts/loss_bsweep.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import imageretrievalresearch_b200 as irr
from bench import graphed_us, measured_peaks

peaks = measured_peaks()
D = 1536
for dt, name in ((torch.bfloat16, "bf16"), (torch.float32, "f32")):
    for B in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
        nsets = max(2, min(6, (1 << 30) // (3 * B * D * 4)))
        sets = [[torch.randn(B, D, device="cuda").to(dt) for _ in range(3)] for _ in range(nsets)]
        us = min(graphed_us(lambda i: irr.triplet_losses_fwd_bwd(*sets[i % nsets], 0.3), 4 * nsets) for _ in range(3))
        by = 6 * B * D * sets[0][0].element_size()
        print(json.dumps({"dtype": name, "B": B, "us": round(us, 2), "hbm_frac": round(by / us / 1e3 / peaks["hbm_gbs"], 3),
                          "sets": nsets}), flush=True)
        del sets
