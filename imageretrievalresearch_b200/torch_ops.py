"""``torch.ops.irr_b200.*`` — the path registered as PyTorch custom operators.

The Python surface in ``retrieval.py`` / ``losses.py`` calls the C ABI directly; this module
registers the same entry points with ``torch.library`` (schema, fake/meta kernels for shape
inference, autograd for the loss) so that a training or evaluation step that contains them can be
captured by ``torch.compile(fullgraph=True)`` / ``torch.export`` without graph breaks — the
reference's per-row Python loops (train/train_efficient_cos_con_ce_loss.py:270-281) cannot be
traced at all.

    torch.ops.irr_b200.cosine_topk(queries, gallery, k, eps, gallery_inv_norm, idx_offset)
        -> (values fp32 [Q,k], indices int64 [Q,k])           cos + topk, :89,273,276
    torch.ops.irr_b200.topk_hits(indices, query_labels, gallery_labels, instance_offset)
        -> int64[2]                                            top1 / topk accounting, :279-281
    torch.ops.irr_b200.pair_cosine(x1, x2, eps) -> fp32 [N]    CosineSimilarity(dim=1), :377,381
    torch.ops.irr_b200.triplet_losses(qry, pos, neg, margin_cos, margin_con, mean)
        -> fp32[4] (cos_pos, cos_neg, con_pos, con_neg), differentiable     :230-237

Every operator is CUDA-only (no CPU kernel is registered: the path has no CPU fallback).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib, _ops
from ._lib import IRR_ROW_STATS, check
from .losses import _triplet_rows
from .retrieval import cosine_topk as _cosine_topk

_NS = "irr_b200"


@torch.library.custom_op(f"{_NS}::cosine_topk", mutates_args=(), device_types="cuda")
def cosine_topk(queries: Tensor, gallery: Tensor, k: int, eps: float,
                gallery_inv_norm: Optional[Tensor], idx_offset: int) -> Tuple[Tensor, Tensor]:
    res = _cosine_topk(queries, gallery, k, eps, gallery_inv_norm=gallery_inv_norm,
                       idx_offset=idx_offset, allow_short=True)
    return res.values, res.indices


@cosine_topk.register_fake
def _(queries, gallery, k, eps, gallery_inv_norm, idx_offset):
    Q = queries.shape[0]
    return (queries.new_empty((Q, k), dtype=torch.float32),
            queries.new_empty((Q, k), dtype=torch.int64))


@torch.library.custom_op(f"{_NS}::topk_hits", mutates_args=(), device_types="cuda")
def topk_hits(indices: Tensor, query_labels: Optional[Tensor], gallery_labels: Optional[Tensor],
              instance_offset: int) -> Tensor:
    return _ops.topk_hits(indices, query_labels, gallery_labels, instance_offset)


@topk_hits.register_fake
def _(indices, query_labels, gallery_labels, instance_offset):
    return indices.new_empty((2,), dtype=torch.int64)


@torch.library.custom_op(f"{_NS}::pair_cosine", mutates_args=(), device_types="cuda")
def pair_cosine(x1: Tensor, x2: Tensor, eps: float) -> Tensor:
    a, b = _ops.same_kind(_ops.as_rows(x1, "x1", keep_f16=True), _ops.as_rows(x2, "x2", keep_f16=True))
    _ops.check_same(a, b, "x1", "x2")
    return _ops.pair_cosine(a, b, eps)


@pair_cosine.register_fake
def _(x1, x2, eps):
    return x2.new_empty((x2.shape[0],), dtype=torch.float32)


# ---- losses: forward saves the per-row statistics, backward is its own operator -----------------
@torch.library.custom_op(f"{_NS}::triplet_losses_fwd", mutates_args=(), device_types="cuda")
def triplet_losses_fwd(qry: Tensor, pos: Tensor, neg: Tensor, margin_cos: float, margin_con: float,
                       mean: bool) -> Tuple[Tensor, Tensor]:
    lib = _lib.load()
    qr, pr, nr = _triplet_rows(qry, pos, neg)
    B, D = qr.shape
    dev = qr.device
    losses = torch.empty(4, dtype=torch.float32, device=dev)
    stats = torch.empty((B, IRR_ROW_STATS), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = _ops.zeroed_scratch(dev, lib.irr_triplet_loss_workspace_bytes(B, D, _ops.dtype_code(qr)))
        check(lib.irr_triplet_loss_fwd_bwd(
            _ops.ptr(qr), _ops.ptr(pr), _ops.ptr(nr), B, D, _ops.dtype_code(qr), margin_cos,
            margin_con, int(mean), 1e-6, _ops.ptr(losses), None, _ops.ptr(stats), None, None, None,
            None, _ops.ptr(ws), ws.numel(), _ops.stream_ptr(dev)), "irr_triplet_loss_fwd_bwd")
    return losses, stats


@triplet_losses_fwd.register_fake
def _(qry, pos, neg, margin_cos, margin_con, mean):
    return (qry.new_empty((4,), dtype=torch.float32),
            qry.new_empty((qry.shape[0], IRR_ROW_STATS), dtype=torch.float32))


@torch.library.custom_op(f"{_NS}::triplet_losses_bwd", mutates_args=(), device_types="cuda")
def triplet_losses_bwd(qry: Tensor, pos: Tensor, neg: Tensor, row_stats: Tensor, grad_losses: Tensor,
                       margin_cos: float, margin_con: float, mean: bool
                       ) -> Tuple[Tensor, Tensor, Tensor]:
    lib = _lib.load()
    qr, pr, nr = _triplet_rows(qry, pos, neg)
    B, D = qr.shape
    dev = qr.device
    gout = grad_losses.detach().to(torch.float32).contiguous()
    dq, dp, dn = torch.empty_like(qr), torch.empty_like(pr), torch.empty_like(nr)
    with torch.cuda.device(dev):
        check(lib.irr_triplet_loss_bwd(_ops.ptr(qr), _ops.ptr(pr), _ops.ptr(nr), _ops.ptr(row_stats),
                                       _ops.ptr(gout), B, D, _ops.dtype_code(qr), margin_cos,
                                       margin_con, int(mean), _ops.ptr(dq), _ops.ptr(dp), _ops.ptr(dn),
                                       _ops.stream_ptr(dev)), "irr_triplet_loss_bwd")
    return dq.to(qry.dtype), dp.to(pos.dtype), dn.to(neg.dtype)


@triplet_losses_bwd.register_fake
def _(qry, pos, neg, row_stats, grad_losses, margin_cos, margin_con, mean):
    return torch.empty_like(qry), torch.empty_like(pos), torch.empty_like(neg)


def _setup(ctx, inputs, output):
    qry, pos, neg, margin_cos, margin_con, mean = inputs
    ctx.save_for_backward(qry, pos, neg, output[1])
    ctx.meta = (margin_cos, margin_con, mean)


def _backward(ctx, grad_losses, _grad_stats):
    qry, pos, neg, stats = ctx.saved_tensors
    dq, dp, dn = torch.ops.irr_b200.triplet_losses_bwd(qry, pos, neg, stats, grad_losses, *ctx.meta)
    return dq, dp, dn, None, None, None


triplet_losses_fwd.register_autograd(_backward, setup_context=_setup)


def triplet_losses(qry: Tensor, pos: Tensor, neg: Tensor, margin_cos: float = 0.3,
                   margin_con: Optional[float] = None, mean: bool = True) -> Tensor:
    """fp32[4] = (cos_pos, cos_neg, con_pos, con_neg) through the registered operators
    (differentiable, traceable)."""
    mk = margin_cos if margin_con is None else margin_con
    return torch.ops.irr_b200.triplet_losses_fwd(qry, pos, neg, float(margin_cos), float(mk), bool(mean))[0]


__all__ = ["cosine_topk", "topk_hits", "pair_cosine", "triplet_losses", "triplet_losses_fwd",
           "triplet_losses_bwd"]
