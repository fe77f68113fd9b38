"""Second, independent restatement of the path in numpy float64 — TEST INFRASTRUCTURE ONLY.

oracle/reference_path.py restates the reference by making the reference's own torch calls; this
file restates the same quantities from their closed forms, without torch, in
float64, so that the oracle is pinned from two sides: the golden vectors produced by the
reference itself (tests/golden/, oracle/gen_golden.py) must agree with BOTH.  Only tests/ may
import it.

  cosine scores     CosineSimilarity(dim=1, eps): x1.x2 / (max(|x1|,eps) * max(|x2|,eps))
                    train/train_efficient_cos_con_ce_loss.py:89,273 (torch >= 2.0 clamp, §A.2)
  stable top-k      torch.topk(sim, k) with ties resolved to the lower index (:276)
  contrastive       utils/contrastive_loss.py:56-61   l = 0.5*(y*d + (1-y)*relu(m - sqrt(d+1e-9))^2)
  cosine embedding  torch.nn.CosineEmbeddingLoss (:158,230-231): c = P/sqrt((A+1e-12)(B+1e-12));
                    y=1: 1-c, y=-1: max(0, c-m)
  gradients         the analytic forms of §A.1 (batch mean -> factor 1/B)
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

COS_EPS = 1e-6
CONTRASTIVE_EPS = 1e-9
COSEMB_EPS = 1e-12


def f64(a) -> np.ndarray:
    return np.asarray(a, dtype=np.float64)


def cos_scores(queries, gallery, eps: float = COS_EPS) -> np.ndarray:
    q, g = f64(queries), f64(gallery)
    qn = np.maximum(np.sqrt((q * q).sum(1)), eps)
    gn = np.maximum(np.sqrt((g * g).sum(1)), eps)
    return (q @ g.T) / (qn[:, None] * gn[None, :])


def topk_stable(scores: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """k largest per row, descending, ties -> lower index (stable argsort of the negated row)."""
    idx = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(scores, idx, axis=1), idx


def pair_cos(a, b, eps: float = COS_EPS) -> np.ndarray:
    a, b = f64(a), f64(b)
    return (a * b).sum(1) / (np.maximum(np.sqrt((a * a).sum(1)), eps) * np.maximum(np.sqrt((b * b).sum(1)), eps))


def contrastive(a, b, y: float, margin: float, mean: bool = True):
    """loss and its gradients w.r.t. a (fm1) and b (fm2)."""
    a, b = f64(a), f64(b)
    delta = b - a
    d = (delta * delta).sum(1)
    s = np.sqrt(d + CONTRASTIVE_EPS)
    gap = np.maximum(margin - s, 0.0)
    per_row = 0.5 * (y * d + (1.0 - y) * gap * gap)
    scale = 1.0 / a.shape[0] if mean else 1.0
    coef = (y - (1.0 - y) * gap / s) * scale          # d l / d b = coef * (b - a)
    db = coef[:, None] * delta
    return per_row.sum() * scale, -db, db


def cosine_embedding(a, b, y: float, margin: float):
    """mean-reduced loss and its gradients w.r.t. a and b (y = +1 or -1)."""
    a, b = f64(a), f64(b)
    P = (a * b).sum(1)
    A = (a * a).sum(1) + COSEMB_EPS
    Bq = (b * b).sum(1) + COSEMB_EPS
    den = np.sqrt(A * Bq)
    c = P / den
    dc_da = b / den[:, None] - (c / A)[:, None] * a
    dc_db = a / den[:, None] - (c / Bq)[:, None] * b
    n = a.shape[0]
    if y == 1.0:
        return (1.0 - c).mean(), -dc_da / n, -dc_db / n
    active = (c - margin >= 0.0).astype(np.float64)   # clamp_min passes the gradient at equality
    return np.maximum(c - margin, 0.0).mean(), active[:, None] * dc_da / n, active[:, None] * dc_db / n


def four_losses_and_grads(q, p, n, margin: float):
    """(losses[4] = cos_pos, cos_neg, con_pos, con_neg; d(sum)/dq, /dp, /dn) — the composition of
    train/train_efficient_cos_con_ce_loss.py:230-237 without the cross-entropy term."""
    l0, a0, b0 = cosine_embedding(q, p, 1.0, margin)
    l1, a1, b1 = cosine_embedding(q, n, -1.0, margin)
    l2, a2, b2 = contrastive(q, p, 1.0, margin)
    l3, a3, b3 = contrastive(q, n, 0.0, margin)
    return np.array([l0, l1, l2, l3]), a0 + a1 + a2 + a3, b0 + b2, b1 + b3


def merge_candidates(cand_val: np.ndarray, cand_idx: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """[G,Q,k] lists -> [Q,k]: score descending, ties -> lower global index, idx < 0 = padding."""
    G, Q, kk = cand_val.shape
    v = np.transpose(cand_val, (1, 0, 2)).reshape(Q, G * kk).astype(np.float64)
    i = np.transpose(cand_idx, (1, 0, 2)).reshape(Q, G * kk)
    out_v = np.full((Q, k), -np.inf)
    out_i = np.full((Q, k), -1, dtype=np.int64)
    for r in range(Q):
        keep = i[r] >= 0
        order = np.lexsort((i[r][keep], -v[r][keep]))[:k]
        out_v[r, : order.size] = v[r][keep][order]
        out_i[r, : order.size] = i[r][keep][order]
    return out_v, out_i
