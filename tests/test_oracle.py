"""The oracle (oracle/reference_path.py) against the golden vectors produced by running the
reference itself (oracle/gen_golden.py) — the pin that lets the GPU parity tests trust it."""
import numpy as np
import pytest
import torch

from oracle import reference_path as ref
from oracle import synthetic

MARGINS = (0.2, 0.3, 0.5)


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("tag", ["unit", "scaled"])
@pytest.mark.parametrize("margin", MARGINS)
def test_four_losses_match_reference(golden_losses, tag, margin):
    g = golden_losses
    q, p, n = T(g[f"{tag}_q"]), T(g[f"{tag}_p"]), T(g[f"{tag}_n"])
    got = ref.four_losses(q, p, n, margin)
    want = T(g[f"{tag}_m{margin}_losses"])
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-8)
    # both negative branches are exercised by the fixture (raw randn would give 0)
    # (row scaling pushes |n-q| past every margin, so the scaled set only keeps the cosine branch)
    assert want[1] > 0 and (want[3] > 0 or tag == "scaled")


@pytest.mark.parametrize("tag", ["unit", "scaled"])
@pytest.mark.parametrize("margin", MARGINS)
def test_grads_match_reference(golden_losses, tag, margin):
    g = golden_losses
    q, p, n = T(g[f"{tag}_q"]), T(g[f"{tag}_p"]), T(g[f"{tag}_n"])
    _, dq, dp, dn = ref.four_losses_and_grads(q, p, n, margin)
    key = f"{tag}_m{margin}"
    for got, name in ((dq, "_dq"), (dp, "_dp"), (dn, "_dn")):
        want = T(g[key + name])
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-9), name


@pytest.mark.parametrize("tag", ["unit", "scaled"])
@pytest.mark.parametrize("margin", MARGINS)
def test_restated_losses_match_reference(golden_losses, tag, margin):
    """The written-out ATen formula and the restated ContrastiveLoss, not just the torch module."""
    g = golden_losses
    q, p, n = T(g[f"{tag}_q"]), T(g[f"{tag}_p"]), T(g[f"{tag}_n"])
    want = T(g[f"{tag}_m{margin}_losses"])
    got = torch.stack([
        ref.cosine_embedding_loss(q, p, torch.tensor([1.0]), margin),
        ref.cosine_embedding_loss(q, n, torch.tensor([-1.0]), margin),
        ref.contrastive_loss(q, p, 1.0, margin),
        ref.contrastive_loss(q, n, 0.0, margin),
    ])
    assert torch.allclose(got, want, rtol=2e-6, atol=1e-8)
    sums = T(g[f"{tag}_m{margin}_con_sum"])
    assert torch.allclose(ref.contrastive_loss(q, p, 1.0, margin, mean=False), sums[0], rtol=1e-6)
    assert torch.allclose(ref.contrastive_loss(q, n, 0.0, margin, mean=False), sums[1], rtol=1e-6)


@pytest.mark.parametrize("margin", MARGINS)
def test_autocast_fp16_restatement_matches_reference_module(golden_autocast, margin):
    """precision=16 semantics (fp16 `fm2 - fm1`, fp32 everywhere else): the oracle's written-out
    policy against vectors produced by the unmodified reference module (oracle/gen_golden.py)."""
    g = golden_autocast
    q, p, n = T(g["ac_q"]), T(g["ac_p"]), T(g["ac_n"])
    assert q.dtype == torch.float16
    key = f"ac_m{margin}"
    w = (1024.0,) * 4
    l, dq, dp, dn = ref.four_losses_and_grads_autocast_fp16(q, p, n, margin, weights=w)
    assert torch.allclose(l, T(g[key + "_losses"]), rtol=1e-6, atol=1e-9)
    for got, name in ((dq, "_dq"), (dp, "_dp"), (dn, "_dn")):
        want = T(g[key + name])
        assert got.dtype == torch.float16
        # fp16 gradients: the two graphs round in the same places except the weighted sum
        assert ((got.float() - want.float()).abs() <= 2e-3 * want.float().abs() + 1e-3).all(), name
    # p - q of nearly equal fp16 rows is EXACT in fp16 (Sterbenz), so for tight positives the
    # autocast value and the widened value agree far below the 1e-5 bar; the fixture records both
    assert abs(float(l[2]) - float(g[key + "_con_pos_widened"])) <= 1e-5 * float(l[2])


def test_contrastive_docstring_shape(golden_losses):
    g = golden_losses
    got = ref.contrastive_loss(T(g["doc_a"]), T(g["doc_b"]), 1, 0.5)
    assert got.dim() == 0 and abs(got.item() - float(g["doc_loss"])) < 1e-6


@pytest.mark.parametrize("tag", ["unit", "scaled"])
def test_paired_scores(golden_losses, tag):
    g = golden_losses
    q, p, n = T(g[f"{tag}_q"]), T(g[f"{tag}_p"]), T(g[f"{tag}_n"])
    sims, unsims = ref.paired_scores(q, p, n)
    assert torch.equal(sims, T(g[f"{tag}_cos_sims"]))
    assert torch.equal(unsims, T(g[f"{tag}_cos_unsims"]))


def test_topk_loop_and_stable_match_reference(golden_retrieval):
    g = golden_retrieval
    q, gal = T(g["planted_q"]), T(g["planted_g"])
    v, i = ref.cos_topk_loop(q, gal, 3)
    assert torch.equal(v, T(g["planted_vals"])) and torch.equal(i, T(g["planted_inds"]))
    sv, si, s = ref.cos_topk_stable(q, gal, 3)
    assert torch.equal(si, T(g["planted_inds"]))          # planted gaps >> rounding
    assert torch.equal(si, T(g["planted_pos"]))           # and they are the planted rows, in order
    assert (sv.float() - T(g["planted_vals"])).abs().max() < 2e-7
    m = ref.topk_matches(T(g["planted_vals"]), T(g["planted_inds"]), s, 3, 1e-5, relative=True)
    assert m["bad_idx"] == 0 and m["val_err"] < 1e-5


def test_top1_top3_counts(golden_retrieval):
    g = golden_retrieval
    q, gal = T(g["planted_q"]), T(g["planted_g"])
    cq, cg = T(g["planted_clss_q"]), T(g["planted_clss_g"])
    # class flavour needs query labels != gallery labels tensor: emulate the loop with hits_from_indices
    _, si, _ = ref.cos_topk_stable(q, gal, 3)
    top1, top3 = ref.hits_from_indices(si, cq, cg)
    assert (top1, top3) == (int(g["planted_top1"]), int(g["planted_top3"]))
    assert 0 < top1 <= top3 < q.shape[0]
    # training-step flavour: gallery = batch of positives, one label vector (:270-281)
    bq, bp, cl = T(g["batch_q"]), T(g["batch_p"]), T(g["batch_clss"])
    assert ref.top1_top3_class_loop(bq, bp, cl) == (int(g["batch_top1"]), int(g["batch_top3"]))
    _, bi, _ = ref.cos_topk_stable(bq, bp, 3)
    assert ref.hits_from_indices(bi, cl, cl) == (int(g["batch_top1"]), int(g["batch_top3"]))
    # instance flavour: positives are aligned with queries, so the own index should be found
    t1, t3 = ref.hits_from_indices(bi, None, None)
    assert t1 == 16 and t3 == 16


def test_iid_values_k10(golden_retrieval):
    g = golden_retrieval
    sv, _, _ = ref.cos_topk_stable(T(g["iid_q"]), T(g["iid_g"]), 10)
    assert (sv.float() - T(g["iid_vals10"])).abs().max() < 3e-7


def test_stable_tie_rule():
    q, gal = synthetic.tied_gallery(400, 32, 8)
    _, si, s = ref.cos_topk_stable(q, gal, 3)
    # the two duplicated best rows tie exactly; the lower index must come first
    for r in range(8):
        assert s[r, si[r, 0]] == s[r, si[r, 1]] and si[r, 0] < si[r, 1]


def test_merge_candidates_equals_global_topk():
    q, gal = synthetic.tied_gallery(600, 48, 10, seed=9)
    _, want_i, s = ref.cos_topk_stable(q, gal, 4)
    bounds = [(0, 150), (150, 151), (151, 600)]
    vals, idxs = [], []
    for lo, hi in bounds:
        kk = min(4, hi - lo)
        v, i, _ = ref.cos_topk_stable(q, gal[lo:hi], kk)
        pad = 4 - kk
        vals.append(torch.cat([v.float(), torch.full((10, pad), -float("inf"))], 1))
        idxs.append(torch.cat([i + lo, torch.full((10, pad), -1, dtype=torch.int64)], 1))
    mv, mi = ref.merge_candidates(torch.stack(vals), torch.stack(idxs), 4)
    assert torch.equal(mi, want_i)
    assert torch.equal(mv, s.gather(1, want_i).float())


def test_synthetic_triplets_exercise_all_branches():
    q, p, n = synthetic.triplets(256, 128, seed=2)
    for m in MARGINS:
        l = ref.four_losses(q, p, n, m)
        assert (l > 0).all()
    d = (n - q).pow(2).sum(1).sqrt()
    assert (d < 0.2).any() and (d > 0.5).any()


def test_notebook_class_dedup_matches_reference(golden_retrieval):
    """ipynb:231-251: top-150, first 3 distinct classes, top1/top3 by class."""
    g = golden_retrieval
    q, p, cls = T(g["nb_q"]), T(g["nb_p"]), T(g["nb_cls"])
    top1, top3, top_r, top_i, top_v = ref.class_dedup_loop(q, p, cls, k=150)
    assert (top1, top3) == (int(g["nb_top1"]), int(g["nb_top3"]))
    assert 0 < top1 < top3 < 420
    assert top_r == [[c for c in row if c >= 0] for row in g["nb_r"].tolist()]
    assert top_i == [[c for c in row if c >= 0] for row in g["nb_i"].tolist()]
    # the batched oracle (stable top-150 + dedup) agrees with the loop
    sv, si, _ = ref.cos_topk_stable(q, p, 150)
    assert (sv.float() - T(g["nb_vals150"])).abs().max() < 3e-7
    lab, ind, val = ref.class_dedup_from_ranked(si, sv.float(), cls, 3)
    assert torch.equal(lab, T(g["nb_r"])) and torch.equal(ind, T(g["nb_i"]))
    assert (val - T(g["nb_v"])).abs().max() < 3e-7


def test_get_fm_and_loss_ce_match_reference(golden_pc):
    g = golden_pc
    fm = T(g["pool_fm"]).requires_grad_(True)
    emb = ref.get_fm(fm)
    assert torch.equal(emb.detach(), T(g["pool_out"]))
    (emb * T(g["pool_up"])).sum().backward()
    assert torch.equal(fm.grad, T(g["pool_grad"]))
    a, b = T(g["ce_a"]).requires_grad_(True), T(g["ce_b"]).requires_grad_(True)
    tot, la, lb = ref.loss_ce(a, b, T(g["ce_t"]))
    assert torch.allclose(torch.stack([tot, la, lb]).detach(), T(g["ce_losses"]), rtol=1e-6)
    (1.7 * tot).backward()
    assert torch.allclose(a.grad, T(g["ce_da"]), rtol=1e-6, atol=1e-9)
    assert torch.allclose(b.grad, T(g["ce_db"]), rtol=1e-6, atol=1e-9)


# -------------------------------------------------------------------------------------------------
# the independent closed-form restatement (numpy float64, no torch) against the same golden vectors
# -------------------------------------------------------------------------------------------------
from oracle import closed_forms as cf


@pytest.mark.parametrize("tag", ["unit", "scaled"])
@pytest.mark.parametrize("margin", MARGINS)
def test_closed_forms_match_reference_losses_and_grads(golden_losses, tag, margin):
    g = golden_losses
    q, p, n = g[f"{tag}_q"], g[f"{tag}_p"], g[f"{tag}_n"]
    losses, dq, dp, dn = cf.four_losses_and_grads(q, p, n, margin)
    key = f"{tag}_m{margin}"
    assert np.allclose(losses, g[key + "_losses"], rtol=2e-6, atol=1e-8)
    for got, name in ((dq, "_dq"), (dp, "_dp"), (dn, "_dn")):
        want = g[key + name].astype(np.float64)
        assert np.linalg.norm(got - want) <= 2e-6 * np.linalg.norm(want), name
    assert np.allclose(cf.pair_cos(q, p), g[f"{tag}_cos_sims"], atol=1e-6)
    assert np.allclose(cf.pair_cos(q, n), g[f"{tag}_cos_unsims"], atol=1e-6)
    sums = g[key + "_con_sum"]
    assert np.isclose(cf.contrastive(q, p, 1.0, margin, mean=False)[0], sums[0], rtol=1e-6)
    assert np.isclose(cf.contrastive(q, n, 0.0, margin, mean=False)[0], sums[1], rtol=1e-6)


def test_closed_forms_match_reference_retrieval(golden_retrieval):
    g = golden_retrieval
    s = cf.cos_scores(g["planted_q"], g["planted_g"])
    v, i = cf.topk_stable(s, 3)
    assert np.array_equal(i, g["planted_inds"]) and np.array_equal(i, g["planted_pos"])
    assert np.allclose(v, g["planted_vals"], rtol=1e-5)
    v10, _ = cf.topk_stable(cf.cos_scores(g["iid_q"], g["iid_g"]), 10)
    assert np.allclose(v10, g["iid_vals10"], rtol=1e-5, atol=1e-7)
    vb, _ = cf.topk_stable(cf.cos_scores(g["batch_q"], g["batch_p"]), 3)
    assert np.allclose(vb, g["batch_vals"], rtol=1e-5)


def test_closed_forms_agree_with_the_torch_oracle_on_edge_cases():
    q, gal = synthetic.tied_gallery(300, 32, 10, seed=3)
    gal[5] = 0
    q[2] = 0
    gal[7] = gal[7] * 1e-9                      # below eps: clamped, not normalised
    s = cf.cos_scores(q.numpy(), gal.numpy())
    _, want_i, want_s = ref.cos_topk_stable(q, gal, 4)
    assert np.allclose(s, want_s.double().numpy(), atol=2e-6)
    # exact duplicate rows tie in float64 as well: lower index first, like the torch oracle
    assert np.array_equal(cf.topk_stable(s, 2)[1][:, 0] < cf.topk_stable(s, 2)[1][:, 1],
                          np.ones(10, dtype=bool))
    vals = torch.randn(4, 9, 3).sort(dim=2, descending=True).values
    idx = torch.stack([torch.randperm(100)[:3].sort().values + 100 * gg for gg in range(4) for _ in range(9)]).view(4, 9, 3)
    vals[0, :, 0] = vals[1, :, 0]
    idx[3, ::2, 2] = -1
    mv, mi = cf.merge_candidates(vals.numpy(), idx.numpy(), 3)
    wv, wi = ref.merge_candidates(vals, idx, 3)
    assert np.array_equal(mi, wi.numpy()) and np.allclose(mv, wv.double().numpy())
