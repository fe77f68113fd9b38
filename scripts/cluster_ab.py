"""A/B of the pair kernel's cluster size (2 = one CTA pair, 4 = two pairs sharing gallery tiles by
TMA multicast) on one box: N=1M x 1536 bf16, cached and uncached norms, alternating."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import imageretrievalresearch_b200 as irr

lib = irr.load_library()
N, D = 1_000_000, 1536
g = torch.randn(N, D, device="cuda", dtype=torch.bfloat16)
gal = irr.Gallery(g)


def t(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return round(s.elapsed_time(e) / iters, 3)


for Q in (int(x) for x in sys.argv[1:]):
    q = torch.randn(Q, D, device="cuda", dtype=torch.bfloat16)
    out = {"Q": Q}
    iters = 10 if Q <= 1024 else 5
    for rep in range(2):
        for cl in (2, 4):
            lib.irr_debug_set_cluster_size(cl)
            out[f"cached_cl{cl}_{rep}"] = t(lambda: gal.search(q, 3), iters)
            out[f"uncached_cl{cl}_{rep}"] = t(lambda: irr.cosine_topk(q, g, 3), iters)
    lib.irr_debug_set_cluster_size(0)
    print(json.dumps(out), flush=True)
