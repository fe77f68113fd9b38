"""Host-side logic of the product package that does not need a device: shard planning, candidate
packing, and the guarantee that nothing silently runs on the CPU."""
import pytest
import torch
from hypothesis import given, settings, strategies as st

import imageretrievalresearch_b200 as irr
from imageretrievalresearch_b200 import sharded


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 10_000_019), st.integers(1, 16))
def test_shard_bounds_partition(n, world):
    prev = 0
    sizes = []
    for r in range(world):
        lo, hi = irr.shard_bounds(n, world, r)
        assert lo == prev and hi >= lo
        sizes.append(hi - lo)
        prev = hi
    assert prev == n
    assert max(sizes) - min(sizes) <= 1


def test_shard_bounds_examples():
    assert [irr.shard_bounds(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert irr.shard_bounds(1_000_000, 8, 7) == (875_000, 1_000_000)
    with pytest.raises(ValueError):
        irr.shard_bounds(10, 4, 4)


@pytest.mark.parametrize("Q,k", [(1, 1), (5, 3), (64, 3), (7, 10)])
def test_pack_unpack_roundtrip(Q, k):
    G = 3
    vals = torch.randn(G, Q, k)
    idx = torch.randint(-1, 1 << 40, (G, Q, k))
    msgs = torch.cat([sharded.pack_candidates(vals[g], idx[g]) for g in range(G)])
    v, i = sharded.unpack_candidates(msgs, G, Q, k)
    assert torch.equal(v, vals) and torch.equal(i, idx)
    assert msgs.numel() // G <= Q * k * 12 + 8


def test_no_cpu_fallback():
    q, g = torch.randn(4, 64), torch.randn(32, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        irr.cosine_topk(q, g, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        irr.ContrastiveLoss(0.3)(q, q, 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        irr.CosineEmbeddingLoss(0.3)(q, q, torch.ones(1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        irr.triplet_losses(q, q, q, 0.3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        irr.CosineSimilarity(dim=1, eps=1e-6)(q[:1], g)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        irr.topk_hits(torch.zeros(4, 3, dtype=torch.int64))


def test_product_package_never_imports_the_oracle():
    import pathlib
    pkg = pathlib.Path(irr.__file__).parent
    for f in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        text = f.read_text()
        assert "import oracle" not in text and "from oracle" not in text, f
        assert "/root/reference" not in text, f


def test_constructor_surface_matches_reference():
    # utils/contrastive_loss.py:31-34: ContrastiveLoss(margin) stores margin and eps=1e-9
    m = irr.ContrastiveLoss(0.5)
    assert m.margin == 0.5 and m.eps == 1e-9
    c = irr.CosineEmbeddingLoss(margin=0.3)
    assert c.margin == 0.3 and c.reduction == "mean"
    with pytest.raises(ValueError):
        irr.CosineEmbeddingLoss(reduction="none")
    with pytest.raises(ValueError):
        irr.CosineSimilarity(dim=0)


def test_torch_ops_are_registered_with_fake_kernels_and_no_cpu_kernel():
    """torch.ops.irr_b200.*: schema + fake (meta) kernels give shapes / dtypes without a device, the
    loss operator is differentiable through its registered backward, and there is no CPU kernel."""
    m = lambda *s, dt=torch.float32: torch.empty(*s, device="meta", dtype=dt)
    v, i = torch.ops.irr_b200.cosine_topk(m(5, 64), m(100, 64), 3, 1e-6, None, 0)
    assert (v.shape, v.dtype, i.shape, i.dtype) == ((5, 3), torch.float32, (5, 3), torch.int64)
    assert torch.ops.irr_b200.topk_hits(m(5, 3, dt=torch.int64), None, None, 0).shape == (2,)
    assert torch.ops.irr_b200.pair_cosine(m(1, 64), m(100, 64), 1e-6).shape == (100,)
    q = m(8, 64).requires_grad_(True)
    losses = irr.torch_ops.triplet_losses(q, m(8, 64), m(8, 64), 0.3)
    assert losses.shape == (4,) and losses.requires_grad
    losses.sum().backward()
    assert q.grad.shape == (8, 64)
    with pytest.raises(NotImplementedError):
        torch.ops.irr_b200.cosine_topk(torch.randn(5, 64), torch.randn(100, 64), 3, 1e-6, None, 0)


def test_serving_helpers_refuse_the_cpu():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        irr.SearchPipeline(lambda q, k: None, 4, 64, 3, torch.float32, "cpu")
    with pytest.raises(ValueError, match="depth"):
        irr.SearchPipeline(lambda q, k: None, 4, 64, 3, torch.float32, "cuda", depth=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        irr.Gallery(torch.randn(8, 64))
    with pytest.raises(ValueError):
        irr.shard_bounds(-1, 2, 0)
