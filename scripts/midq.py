"""Mid-size query batches (HBM/tensor crossover): single-CTA vs CTA-pair kernel, cached norms."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import imageretrievalresearch_b200 as irr

N, D = 1_000_000, 1536
g = torch.randn(N, D, device="cuda", dtype=torch.bfloat16)
gal = irr.Gallery(g)
for Q in (int(x) for x in sys.argv[1:]):
    q = torch.randn(Q, D, device="cuda", dtype=torch.bfloat16)
    out = {}
    for mode, fn in (("cached", lambda: gal.search(q, 3)), ("uncached", lambda: irr.cosine_topk(q, g, 3))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            fn()
        e.record()
        torch.cuda.synchronize()
        out[mode] = round(s.elapsed_time(e) / 10, 3)
    print(json.dumps({"Q": Q, "pair_forced": os.environ.get("IRR_FORCE_PAIR"), **out}), flush=True)
