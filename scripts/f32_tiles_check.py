"""fp32 exactness path: the 128 x 128 FFMA2 kernel (default for more than 64 queries) against the
64 x 128 kernel (IRR_F32_SMALL_TILES=1): run once per setting, then compare /tmp/f32_out_{0,1}.pt
bit for bit; also times the notebook-size problem (8736 x 8736 x 1920) at k=3 and k=150.

    IRR_F32_SMALL_TILES=1 python scripts/f32_tiles_check.py; python scripts/f32_tiles_check.py
"""
import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import imageretrievalresearch_b200 as irr
mode = os.environ.get("IRR_F32_SMALL_TILES", "0")
torch.manual_seed(0)
out = {}
for (Q, N, D, k) in [(200, 5000, 1920, 3), (129, 777, 72, 10), (1000, 20000, 1536, 3), (70, 300, 64, 150)]:
    q = torch.randn(Q, D, device="cuda"); g = torch.randn(N, D, device="cuda")
    r = irr.cosine_topk(q, g, k)
    out[f"{Q}_{N}_{D}_{k}"] = (r.values.cpu(), r.indices.cpu())
torch.save(out, f"/tmp/f32_out_{mode}.pt")
Q = N = 8736; D = 1920
q = torch.randn(Q, D, device="cuda"); g = torch.randn(N, D, device="cuda")
for k in (3, 150):
    for _ in range(2): irr.cosine_topk(q, g, k)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): irr.cosine_topk(q, g, k)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(json.dumps({"small_tiles": mode, "Q": Q, "N": N, "D": D, "k": k, "ms": round(ms, 3), "TFLOPs": round(2.0*Q*N*D/ms/1e9, 1)}), flush=True)
