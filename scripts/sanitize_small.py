"""Every kernel once on tiny shapes — meant to run under `compute-sanitizer --tool memcheck`
(one tool per gpurun call, see the profiling guide)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import imageretrievalresearch_b200 as irr
from imageretrievalresearch_b200 import _ops

torch.manual_seed(0)
dev = "cuda"
for dt in (torch.bfloat16, torch.float32):
    g = torch.randn(3000, 256, device=dev).to(dt)
    for Q, k in ((5, 3), (130, 3), (600, 10), (40, 40)):
        q = torch.randn(Q, 256, device=dev).to(dt)
        r = irr.cosine_topk(q, g, k)
        assert r.indices.shape == (Q, k)
    irr.Gallery(g).search(torch.randn(7, 256, device=dev).to(dt), 3)
q = torch.randn(9, 64, device=dev)
g = torch.randn(100, 64, device=dev)
lab = torch.randint(0, 5, (100,), device=dev)
t1, t3, d = irr.top1_top3_dedup(q, g, lab[:9], lab, k=50)
irr.top1_top3(q, g, lab[:9], lab)
irr.CosineSimilarity(dim=1, eps=1e-6)(q[:1], g)
_ops.cosine_scores_bf16(q.bfloat16(), g.bfloat16(), 1e-6)
v = torch.randn(4, 9, 3, device=dev).sort(dim=2, descending=True).values
i = torch.randint(0, 1000, (4, 9, 3), device=dev)
_ops.topk_merge(v, i)
a, b, c = [torch.randn(70, 1536, device=dev, requires_grad=True) for _ in range(3)]
tl = irr.triplet_losses(a, b, c, 0.3, pair_scores=True)
(tl.loss_cos + tl.loss_con).backward()
irr.triplet_losses_fwd_bwd(a.detach().bfloat16(), b.detach().bfloat16(), c.detach().bfloat16(), 0.3)
irr.ContrastiveLoss(0.5)(a, b, 1.0).backward()
irr.CosineEmbeddingLoss(0.2)(a, c, torch.tensor([-1.0], device=dev)).backward()
fm = torch.randn(3, 130, 7, 7, device=dev, requires_grad=True)
irr.get_fm(fm).sum().backward()
la, lb = torch.randn(20, 125, device=dev, requires_grad=True), torch.randn(20, 125, device=dev, requires_grad=True)
irr.cross_entropy_pair(la, lb, torch.randint(0, 125, (20,), device=dev)).loss.backward()
# fp16 rows on the tensor path, streamed host gallery, peer-exchange protocol (virtual ranks)
gh = torch.randn(3000, 256, device=dev).half()
irr.cosine_topk(torch.randn(600, 256, device=dev).half(), gh, 3, fp16_tensor_path=True)
irr.cosine_topk(torch.randn(5, 256, device=dev).half(), gh, 3, fp16_tensor_path=True)
irr.StreamedGallery(torch.randn(2000, 64).bfloat16().pin_memory(), 700, dev).search(
    torch.randn(9, 64, device=dev).bfloat16(), 3)
from imageretrievalresearch_b200 import _lib
G, Q, k = 3, 50, 3
nb = _ops.topk_exchange_bytes(G, Q, k)
bufs = [torch.zeros(nb, dtype=torch.uint8, device=dev) for _ in range(G)]
ptrs = [b.data_ptr() for b in bufs]
xv = torch.randn(G, Q, k, device=dev).sort(dim=2, descending=True).values
xi = torch.randint(0, 1000, (G, Q, k), device=dev)
for r in range(1, G):
    _ops.topk_exchange_merge(xv[r], xi[r], ptrs, r, Q, k, nb, _lib.IRR_XCHG_PUSH, torch.device(dev, 0))
_ops.topk_exchange_merge(xv[0], xi[0], ptrs, 0, Q, k, nb, _lib.IRR_XCHG_FUSED, torch.device(dev, 0))
_ops.topk_exchange_merge(None, None, ptrs, 1, Q, k, nb, _lib.IRR_XCHG_MERGE, torch.device(dev, 0))
torch.cuda.synchronize()
print("sanitize_small: all kernels ran")
