"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's retrieval-ranking path.

Nothing in the product package (imageretrievalresearch_b200/) imports this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and only as the
checker or the timed CPU baseline.

What it restates.  The reference (vitasoftAI/ImageRetrievalResearch) has no operator for this path:
it is four torch calls repeated in Python loops.  The arithmetic therefore lives in a third-party
dependency, PyTorch (pinned torch==1.12.0 in the reference's requirements.txt:167; torch 2.11.0 is
what runs here and on the GPU box, so torch 2.11 semantics are "the reference's own torch path").
Each function below cites the reference file:line it follows (paths relative to the reference
tree, *.ipynb:N are raw-JSON line numbers).

Parity pinning.  The reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
§4), so parity is pinned on outputs of the reference itself run in the authoring container:
oracle/gen_golden.py imports the reference's own ``ContrastiveLoss`` from
``utils/contrastive_loss.py`` by file path and makes the same torch calls the reference scripts
make, on small seeded inputs, and commits the vectors under tests/golden/.  tests/test_oracle.py
checks every function here against those vectors.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

CONTRASTIVE_EPS = 1e-9        # utils/contrastive_loss.py:34
COS_EPS = 1e-6                # train/train_efficient_cos_con_ce_loss.py:89
COS_EMB_EPS = 1e-12           # ATen cosine_embedding_loss EPSILON


# -------------------------------------------------------------------------------------------------
# a1/a2  cosine similarity + top-k
# -------------------------------------------------------------------------------------------------
def cos_topk_loop(queries: torch.Tensor, gallery: torch.Tensor, k: int, eps: float = COS_EPS
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's hot loop verbatim (train/train_efficient_cos_con_ce_loss.py:270-276,
    384-388; inference/training_analysis.ipynb:238): one CosineSimilarity + one torch.topk per
    query row.  Tie order is whatever torch.topk does (unspecified)."""
    cos = torch.nn.CosineSimilarity(dim=1, eps=eps)                    # :89
    vals, inds = [], []
    for idx in range(queries.shape[0]):                                # :270
        sim = cos(queries[idx].unsqueeze(0), gallery)                  # :273
        v, i = torch.topk(sim, k=k)                                    # :276
        vals.append(v)
        inds.append(i)
    return torch.stack(vals), torch.stack(inds)


def cos_scores(queries: torch.Tensor, gallery: torch.Tensor, eps: float = COS_EPS,
               dtype: torch.dtype = torch.float64) -> torch.Tensor:
    """[Q,N] matrix whose row i is cos(queries[i][None], gallery) — CosineSimilarity(dim=1, eps)
    (:89) with torch>=2.0's per-operand clamp x/max(|x|,eps) — evaluated in `dtype` (fp64 = ground
    truth for the tolerance checks).  Inputs are used as given (bf16 inputs are upcast, not
    re-rounded), which isolates the kernels' accumulation error (SURVEY.md §8c rule 3)."""
    q, g = queries.to(dtype), gallery.to(dtype)
    qn = q / q.norm(dim=1, keepdim=True).clamp_min(eps)
    gn = g / g.norm(dim=1, keepdim=True).clamp_min(eps)
    return qn @ gn.T


def cos_topk_stable(queries: torch.Tensor, gallery: torch.Tensor, k: int, eps: float = COS_EPS,
                    dtype: torch.dtype = torch.float64
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Batched restatement of :273-276 with the tie rule the kernels promise imposed on it:
    descending stable sort, so equal scores come back in increasing gallery index.
    Returns (values[Q,k], indices[Q,k], full score matrix)."""
    s = cos_scores(queries, gallery, eps, dtype)
    v, i = torch.sort(s, dim=1, descending=True, stable=True)
    return v[:, :k], i[:, :k], s


def topk_matches(got_val: torch.Tensor, got_idx: torch.Tensor, oracle_scores: torch.Tensor, k: int,
                 tol: float, relative: bool) -> Dict[str, float]:
    """Tolerance-aware comparison north_star prescribes: scores within `tol` (relative or absolute)
    of the oracle's, indices identical except where the oracle's own score gap is below `tol`."""
    ov, oi = torch.sort(oracle_scores, dim=1, descending=True, stable=True)
    ov, oi = ov[:, :k], oi[:, :k]
    gv = got_val.to(oracle_scores.dtype).cpu()
    gi = got_idx.cpu()
    denom = ov.abs().clamp_min(1e-30) if relative else torch.ones_like(ov)
    val_err = ((gv - ov).abs() / denom).max().item() if gv.numel() else 0.0
    same = gi == oi
    # a differing index is acceptable only if the row it names scores within tol of the oracle's
    picked = oracle_scores.gather(1, gi.clamp_min(0))
    gap = (picked - ov).abs() / denom
    bad = (~same) & ((gap > tol) | (gi < 0))
    return {"val_err": val_err, "idx_equal_frac": same.float().mean().item() if same.numel() else 1.0,
            "bad_idx": int(bad.sum().item())}


# -------------------------------------------------------------------------------------------------
# a3  top-1 / top-3 accounting
# -------------------------------------------------------------------------------------------------
def top1_top3_class_loop(fm_ims: torch.Tensor, fm_poss: torch.Tensor, clss: torch.Tensor,
                         k: int = 3, eps: float = COS_EPS) -> Tuple[int, int]:
    """Class flavour, verbatim (train/train_efficient_cos_con_ce_loss.py:267-281): per query,
    cos vs the batch of positives, topk, hit if the query's class is among the classes of the
    returned rows.  Returns hit COUNTS (top1, topk); the reference logs count/len (:286-287)."""
    cos = torch.nn.CosineSimilarity(dim=1, eps=eps)
    top3, top1 = 0, 0
    for idx in range(fm_ims.shape[0]):
        sim = cos(fm_ims[idx].unsqueeze(0), fm_poss)
        vals, inds = torch.topk(sim, k=k)
        if any(bool(clss[idx] == clss[inds[j]]) for j in range(k)):    # :279
            top3 += 1
        if bool(clss[idx] == clss[inds[0]]):                           # :281
            top1 += 1
    return top1, top3


def hits_from_indices(indices: torch.Tensor, q_label: Optional[torch.Tensor],
                      g_label: Optional[torch.Tensor], instance_offset: int = 0) -> Tuple[int, int]:
    """Hit counts from already-selected indices: class flavour (:279-281) when labels are given,
    instance flavour (inference/inference.py:237,242 — the query's own index is among the returned
    ones) otherwise."""
    indices = indices.cpu()
    Q = indices.shape[0]
    valid = indices >= 0
    if q_label is not None:
        m = (g_label.cpu()[indices.clamp_min(0)] == q_label.cpu()[:, None]) & valid
    else:
        m = (indices == (torch.arange(Q)[:, None] + instance_offset)) & valid
    return int(m[:, 0].sum()), int(m.any(dim=1).sum())


def class_dedup_loop(fms_ims_all: torch.Tensor, fms_poss_all: torch.Tensor, classes_all: torch.Tensor,
                     k: int = 150, n_distinct: int = 3, eps: float = COS_EPS):
    """The notebook's working inference evaluation, verbatim (inference/training_analysis.ipynb:
    231-251): per query top-k (150) by cosine, walk the ranked rows, keep the first 3 distinct
    classes; top3 / top1 by class.  Returns (top1 count, top3 count, top_r lists, top_i lists,
    top_v lists).  Tie order inside torch.topk is torch's own."""
    cos = torch.nn.CosineSimilarity(dim=1, eps=eps)                               # ipynb:187
    top1 = top3 = 0
    top_r_list, top_inds, top_vals = [], [], []
    for idx, (gt_reg, fm) in enumerate(zip(classes_all, fms_ims_all)):            # :231
        vals, inds = torch.topk(cos(fm, fms_poss_all), k=k)                       # :238
        classes = [int(classes_all[int(ind)]) for ind in inds]                    # :240
        top_i, top_v, top_r = [], [], []
        for num, (i, v, r) in enumerate(zip(inds, vals, classes)):                # :243
            if r not in top_r:
                top_r.append(r)
                top_v.append(float(v))
                top_i.append(int(i))
            if len(top_r) == n_distinct:
                break
        top3 += 1 if int(gt_reg) in top_r else 0                                  # :250
        top1 += 1 if int(gt_reg) == top_r[0] else 0                               # :251
        top_inds.append(top_i)
        top_vals.append(top_v)
        top_r_list.append(top_r)
    return top1, top3, top_r_list, top_inds, top_vals


def class_dedup_from_ranked(indices: torch.Tensor, values: torch.Tensor, g_label: torch.Tensor,
                            n_distinct: int = 3):
    """ipynb:240-249 applied to already-ranked lists: [Q,n] labels / indices / values, -1 / -inf pad."""
    Q = indices.shape[0]
    lab = torch.full((Q, n_distinct), -1, dtype=torch.int64)
    ind = torch.full((Q, n_distinct), -1, dtype=torch.int64)
    val = torch.full((Q, n_distinct), -float("inf"), dtype=torch.float32)
    for q in range(Q):
        seen = []
        for j in range(indices.shape[1]):
            g = int(indices[q, j])
            if g < 0:
                continue
            r = int(g_label[g])
            if r not in seen:
                lab[q, len(seen)], ind[q, len(seen)], val[q, len(seen)] = r, g, float(values[q, j])
                seen.append(r)
            if len(seen) == n_distinct:
                break
    return lab, ind, val


# -------------------------------------------------------------------------------------------------
# a4  paired scores
# -------------------------------------------------------------------------------------------------
def paired_scores(q: torch.Tensor, p: torch.Tensor, n: torch.Tensor, eps: float = COS_EPS
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos_sims / cos_unsims (train/train_efficient_cos_con_ce_loss.py:377-382; ipynb:232-235):
    row-wise cos(q_i, p_i) and cos(q_i, n_i)."""
    cos = torch.nn.CosineSimilarity(dim=1, eps=eps)
    sims = torch.stack([cos(q[i].unsqueeze(0), p[i].unsqueeze(0))[0] for i in range(q.shape[0])])
    unsims = torch.stack([cos(q[i].unsqueeze(0), n[i].unsqueeze(0))[0] for i in range(q.shape[0])])
    return sims, unsims


# -------------------------------------------------------------------------------------------------
# a5/a6/a7  losses
# -------------------------------------------------------------------------------------------------
def contrastive_loss(fm1: torch.Tensor, fm2: torch.Tensor, label, margin: float, mean: bool = True
                     ) -> torch.Tensor:
    """utils/contrastive_loss.py:56-61, restated line for line."""
    dis = (fm2 - fm1).pow(2).sum(1)                                                     # :56
    losses = 0.5 * (label * dis + (1 + -1 * label) *
                    F.relu(margin - (dis + CONTRASTIVE_EPS).sqrt()).pow(2))             # :59
    return losses.mean() if mean else losses.sum()                                      # :61


def cosine_embedding_loss(x1: torch.Tensor, x2: torch.Tensor, target: torch.Tensor, margin: float,
                          mean: bool = True) -> torch.Tensor:
    """torch.nn.CosineEmbeddingLoss(margin) as the reference constructs and calls it
    (train/train_efficient_cos_con_ce_loss.py:158,230-231), written out from ATen's formula so
    that the eps convention is explicit: EPSILON=1e-12 is added to the SQUARED norms."""
    prod = (x1 * x2).sum(1)
    m1 = (x1 * x1).sum(1) + COS_EMB_EPS
    m2 = (x2 * x2).sum(1) + COS_EMB_EPS
    c = prod / (m1 * m2).sqrt()
    t = target.to(c.dtype).expand_as(c) if target.numel() == 1 else target.to(c.dtype)
    zeros = torch.zeros_like(c)
    out = torch.where(t == 1, 1 - c, zeros) + torch.where(t == -1, (c - margin).clamp_min(0), zeros)
    return out.mean() if mean else out.sum()


def four_losses(q: torch.Tensor, p: torch.Tensor, n: torch.Tensor, margin: float,
                margin_con: Optional[float] = None) -> torch.Tensor:
    """[cos_pos, cos_neg, con_pos, con_neg] exactly as training_step composes them
    (train/train_efficient_cos_con_ce_loss.py:230-237) with the [1]-shaped labels of :97-100."""
    dev = q.device
    labels = {"con_pos": torch.tensor(1., device=dev).unsqueeze(0), "con_neg": torch.tensor(0., device=dev).unsqueeze(0),
              "cos_pos": torch.tensor(1., device=dev).unsqueeze(0), "cos_neg": torch.tensor(-1., device=dev).unsqueeze(0)}
    mk = margin if margin_con is None else margin_con
    cos_loss = torch.nn.CosineEmbeddingLoss(margin=margin)                              # :158
    return torch.stack([
        cos_loss(q, p, labels["cos_pos"]),                                              # :230
        cos_loss(q, n, labels["cos_neg"]),                                              # :231
        contrastive_loss(q, p, labels["con_pos"], mk),                                  # :235
        contrastive_loss(q, n, labels["con_neg"], mk),                                  # :236
    ])


def four_losses_and_grads(q, p, n, margin: float, weights=(1.0, 1.0, 1.0, 1.0),
                          margin_con: Optional[float] = None, dtype=torch.float32):
    """Losses plus d(sum_j w_j loss_j)/d(q,p,n) by torch autograd — what loss.backward() gives the
    reference (train/train_efficient_cos_con_ce_loss.py:245 + Lightning's backward)."""
    qr, pr, nr = [t.detach().to(dtype).clone().requires_grad_(True) for t in (q, p, n)]
    losses = four_losses(qr, pr, nr, margin, margin_con)
    w = torch.tensor(weights, dtype=dtype)
    (losses * w).sum().backward()
    return losses.detach(), qr.grad, pr.grad, nr.grad


def four_losses_autocast_fp16(q: torch.Tensor, p: torch.Tensor, n: torch.Tensor, margin: float,
                              margin_con: Optional[float] = None) -> torch.Tensor:
    """four_losses as the reference evaluates it on fp16 embeddings under ``precision=16``
    (train/train_efficient_cos_con_ce_loss.py:465), with autocast's dtype policy written out so that
    it runs on the CPU: ``fm2 - fm1`` (utils/contrastive_loss.py:56) is not on any autocast list and
    runs in the inputs' dtype — an fp16 subtraction, rounded to fp16 — while ``pow``, ``sum`` and
    ``cosine_embedding_loss`` are on the fp32 list: fp32 arithmetic on the fp16
    values.  Differentiable w.r.t. fp16 leaves (the casts round the gradients to fp16 and fp16
    ``.grad`` accumulates in fp16, as under autocast)."""
    assert q.dtype == p.dtype == n.dtype == torch.float16
    mk = margin if margin_con is None else margin_con
    one, minus = torch.tensor(1.).unsqueeze(0), torch.tensor(-1.).unsqueeze(0)
    cos_loss = torch.nn.CosineEmbeddingLoss(margin=margin)                              # :158, fp32 list
    def con(a, b, label):                                                               # :56-61
        dis = (b - a).float().pow(2).sum(1)               # fp16 sub, then the fp32-list ops
        losses = 0.5 * (label * dis + (1 + -1 * label) * F.relu(mk - (dis + CONTRASTIVE_EPS).sqrt()).pow(2))
        return losses.mean()
    return torch.stack([cos_loss(q.float(), p.float(), one), cos_loss(q.float(), n.float(), minus),
                        con(q, p, 1.0), con(q, n, 0.0)])


def four_losses_and_grads_autocast_fp16(q, p, n, margin: float, weights=(1.0, 1.0, 1.0, 1.0),
                                        margin_con: Optional[float] = None):
    qr, pr, nr = [t.detach().clone().requires_grad_(True) for t in (q, p, n)]
    losses = four_losses_autocast_fp16(qr, pr, nr, margin, margin_con)
    (losses * torch.tensor(weights)).sum().backward()
    return losses.detach(), qr.grad, pr.grad, nr.grad


# -------------------------------------------------------------------------------------------------
# a8 / next rows: the producer (get_fm) and the consumer (cross-entropy) either side of the path
# -------------------------------------------------------------------------------------------------
def get_fm(fm: torch.Tensor) -> torch.Tensor:
    """train/train_efficient_cos_con_ce_loss.py:103-122, verbatim."""
    pool = torch.nn.AvgPool2d((fm.shape[2], fm.shape[3]))                               # :120
    return torch.reshape(pool(fm), (-1, fm.shape[1]))                                   # :122


def loss_ce(lbl_ims: torch.Tensor, lbl_poss: torch.Tensor, clss: torch.Tensor):
    """train/train_efficient_cos_con_ce_loss.py:160,240-242: CrossEntropyLoss() on the query and
    positive logits, summed.  Returns (loss_ce, loss_ce_ims, loss_ce_poss)."""
    ce_loss = torch.nn.CrossEntropyLoss()                                               # :160
    loss_ce_ims = ce_loss(lbl_ims, clss)                                                # :240
    loss_ce_poss = ce_loss(lbl_poss, clss)                                              # :241
    return loss_ce_ims + loss_ce_poss, loss_ce_ims, loss_ce_poss                        # :242


# -------------------------------------------------------------------------------------------------
# K3  candidate merge (new in the sharded design; oracle = sort by (score desc, index asc))
# -------------------------------------------------------------------------------------------------
def merge_candidates(cand_val: torch.Tensor, cand_idx: torch.Tensor, k: int
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """[G,Q,k] -> [Q,k]: what the single-GPU stable top-k would return over the union of the
    shards' candidates; idx < 0 is padding and sorts last."""
    G, Q, kk = cand_val.shape
    v = cand_val.permute(1, 0, 2).reshape(Q, G * kk).clone().double()
    i = cand_idx.permute(1, 0, 2).reshape(Q, G * kk).clone()
    v[i < 0] = -float("inf")
    big = torch.iinfo(torch.int64).max
    order = torch.argsort(torch.where(i < 0, torch.full_like(i, big), i), dim=1, stable=True)
    v, i = v.gather(1, order), i.gather(1, order)
    o2 = torch.argsort(v, dim=1, descending=True, stable=True)
    v, i = v.gather(1, o2)[:, :k], i.gather(1, o2)[:, :k]
    out_v = torch.where(i < 0, torch.full_like(v, -float("inf")), v)
    return out_v.to(cand_val.dtype), torch.where(i < 0, torch.full_like(i, -1), i)
