"""Embedding losses: drop-ins for the reference's two loss modules plus the fused triplet form.

Reference (paths under the reference tree):
  ContrastiveLoss(margin)(fm1, fm2, label, mean=True)       utils/contrastive_loss.py:31-61
  torch.nn.CosineEmbeddingLoss(margin)(x1, x2, target)      train/train_efficient_cos_con_ce_loss.py:158,230-231
  loss_cos = pos + neg ; loss_con = pos + neg               train/train_efficient_cos_con_ce_loss.py:230-237
  labels 1. / 0. / 1. / -1. as shape-[1] tensors            train/train_efficient_cos_con_ce_loss.py:97-100
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Sequence, Tuple, Union

import torch

from . import _lib, _ops
from ._lib import IRR_LOSS_CONTRASTIVE, IRR_LOSS_COSINE_EMBEDDING, IRR_ROW_STATS, check

Label = Union[float, int, torch.Tensor]

_const_cache: dict = {}


def _label_tensor(label: Label, B: int, device: torch.device, what: str) -> torch.Tensor:
    """Python number or tensor of 1 / B elements -> contiguous fp32 device tensor."""
    if isinstance(label, torch.Tensor):
        t = label.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    else:
        key = (device.index, float(label))
        t = _const_cache.get(key)
        if t is None:
            t = torch.full((1,), float(label), dtype=torch.float32, device=device)
            _const_cache[key] = t
    if t.numel() not in (1, B):
        raise RuntimeError(f"{what} must have 1 or {B} elements, got {t.numel()}")
    return t


def _autocast_rows(tensors, names, autocast_exact: bool):
    """Rows for the loss kernels.  fp16 embeddings (what ``precision=16`` training produces,
    train/train_efficient_cos_con_ce_loss.py:465) stay fp16 when EVERY operand is fp16 and
    ``autocast_exact`` is set: the kernel then evaluates ``fm2 - fm1`` as the fp16 subtraction the
    reference's autocast performs (utils/contrastive_loss.py:56; everything after it in fp32) and
    emits fp16 gradients — the reference's numbers, not more exact ones.  Mixed fp16 / fp32 operands
    promote to fp32 in the reference, too, and are widened here (as is everything when
    ``autocast_exact=False``)."""
    keep = autocast_exact and all(t.dtype == torch.float16 for t in tensors)
    return [_ops.as_rows(t, n, keep_f16=keep) for t, n in zip(tensors, names)]


class _PairLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, label_t, kind, margin, mean, autocast_exact=True):
        lib = _lib.load()
        ar, br = _autocast_rows((a, b), ("input1", "input2"), autocast_exact)
        ar, br = _ops.same_kind(ar, br)
        _ops.check_same(ar, br, "input1", "input2")
        if ar.shape != br.shape:
            raise RuntimeError(f"shape mismatch: {tuple(ar.shape)} vs {tuple(br.shape)}")
        B, D = ar.shape
        dev = ar.device
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        stats = torch.empty((B, IRR_ROW_STATS), dtype=torch.float32, device=dev) if need_grad else None
        with torch.cuda.device(dev):
            ws = _ops.zeroed_scratch(dev, lib.irr_pair_loss_workspace_bytes(B, D, _ops.dtype_code(ar)))
            check(lib.irr_pair_loss_fwd_bwd(_ops.ptr(ar), _ops.ptr(br), _ops.ptr(label_t),
                                            label_t.numel(), B, D, _ops.dtype_code(ar), kind,
                                            margin, int(mean), _ops.ptr(loss), _ops.ptr(stats),
                                            None, None, 1.0, _ops.ptr(ws), ws.numel(),
                                            _ops.stream_ptr(dev)), "irr_pair_loss_fwd_bwd")
        if need_grad:
            ctx.save_for_backward(ar, br, label_t, stats)
            ctx.meta = (kind, margin, int(mean), a.dtype, b.dtype)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        ar, br, label_t, stats = ctx.saved_tensors
        kind, margin, mean, a_dtype, b_dtype = ctx.meta
        B, D = ar.shape
        dev = ar.device
        g = grad_out.detach().to(dtype=torch.float32).reshape(1).contiguous()
        da, db = torch.empty_like(ar), torch.empty_like(br)
        with torch.cuda.device(dev):
            check(lib.irr_pair_loss_bwd(_ops.ptr(ar), _ops.ptr(br), _ops.ptr(label_t),
                                        label_t.numel(), _ops.ptr(stats), _ops.ptr(g), B, D,
                                        _ops.dtype_code(ar), kind, margin, mean, _ops.ptr(da),
                                        _ops.ptr(db), _ops.stream_ptr(dev)), "irr_pair_loss_bwd")
        return (da.to(a_dtype) if ctx.needs_input_grad[0] else None,
                db.to(b_dtype) if ctx.needs_input_grad[1] else None, None, None, None, None, None)


class ContrastiveLoss(torch.nn.Module):
    """Same constructor, call signature and value as the reference's ``ContrastiveLoss``
    (utils/contrastive_loss.py:31-61): ``0.5*(label*d + (1-label)*relu(margin - sqrt(d+1e-9))^2)``
    with ``d = sum((fm2-fm1)^2, dim=1)``, reduced by mean (default) or sum.  One fused kernel per
    direction instead of ~12 elementwise launches; differentiable w.r.t. fm1 and fm2."""

    def __init__(self, margin: float, autocast_exact: bool = True) -> None:
        super().__init__()
        self.margin, self.eps = margin, 1e-9
        self.autocast_exact = autocast_exact   # fp16 inputs: reproduce autocast's fp16 `fm2 - fm1`

    def forward(self, fm1: torch.Tensor, fm2: torch.Tensor, label: Label, mean: bool = True
                ) -> torch.Tensor:
        _ops._require_cuda(fm1, "fm1")
        lab = _label_tensor(label, fm1.shape[0], fm1.device, "label")
        return _PairLossFn.apply(fm1, fm2, lab, IRR_LOSS_CONTRASTIVE, float(self.margin), bool(mean),
                                 self.autocast_exact)


class CosineEmbeddingLoss(torch.nn.Module):
    """Drop-in for ``torch.nn.CosineEmbeddingLoss(margin, reduction)`` on ``[B,D]`` inputs with a
    target of shape ``[1]`` or ``[B]`` holding 1 / -1
    (train/train_efficient_cos_con_ce_loss.py:97-100,158,230-231)."""

    def __init__(self, margin: float = 0.0, reduction: str = "mean") -> None:
        super().__init__()
        if reduction not in ("mean", "sum"):
            raise ValueError("reduction must be 'mean' or 'sum'")
        self.margin, self.reduction = margin, reduction

    def forward(self, input1: torch.Tensor, input2: torch.Tensor, target: Label) -> torch.Tensor:
        _ops._require_cuda(input1, "input1")
        tgt = _label_tensor(target, input1.shape[0], input1.device, "target")
        return _PairLossFn.apply(input1, input2, tgt, IRR_LOSS_COSINE_EMBEDDING, float(self.margin),
                                 self.reduction == "mean")


# -------------------------------------------------------------------------------------------------
# fused triplet form
# -------------------------------------------------------------------------------------------------
class TripletLosses(NamedTuple):
    cos_pos: torch.Tensor
    cos_neg: torch.Tensor
    con_pos: torch.Tensor
    con_neg: torch.Tensor
    pair_cos_pos: Optional[torch.Tensor]  # [B] cos(q_i, p_i): the reference's cos_sims
    pair_cos_neg: Optional[torch.Tensor]  # [B] cos(q_i, n_i): the reference's cos_unsims

    @property
    def loss_cos(self) -> torch.Tensor:   # train/train_efficient_cos_con_ce_loss.py:232
        return self.cos_pos + self.cos_neg

    @property
    def loss_con(self) -> torch.Tensor:   # train/train_efficient_cos_con_ce_loss.py:237
        return self.con_pos + self.con_neg


def _triplet_rows(q, p, n, autocast_exact: bool = True):
    qr, pr, nr = _autocast_rows((q, p, n), ("qry", "pos", "neg"), autocast_exact)
    if not (qr.dtype == pr.dtype == nr.dtype):     # mixed fp16 / fp32: the reference promotes to fp32
        qr, pr, nr = [t.float() if t.dtype == torch.float16 else t for t in (qr, pr, nr)]
    _ops.check_same(qr, pr, "qry", "pos")
    _ops.check_same(qr, nr, "qry", "neg")
    if not (qr.shape == pr.shape == nr.shape):
        raise RuntimeError("qry / pos / neg must have the same shape")
    return qr, pr, nr


class _TripletLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, p, n, margin_cos, margin_con, mean, want_pairs, pair_eps, autocast_exact=True):
        lib = _lib.load()
        qr, pr, nr = _triplet_rows(q, p, n, autocast_exact)
        B, D = qr.shape
        dev = qr.device
        need_grad = any(ctx.needs_input_grad[:3])
        losses = torch.empty(4, dtype=torch.float32, device=dev)
        pairs = torch.empty((2, B), dtype=torch.float32, device=dev) if want_pairs else None
        stats = torch.empty((B, IRR_ROW_STATS), dtype=torch.float32, device=dev) if need_grad else None
        with torch.cuda.device(dev):
            ws = _ops.zeroed_scratch(dev, lib.irr_triplet_loss_workspace_bytes(B, D, _ops.dtype_code(qr)))
            check(lib.irr_triplet_loss_fwd_bwd(
                _ops.ptr(qr), _ops.ptr(pr), _ops.ptr(nr), B, D, _ops.dtype_code(qr), margin_cos,
                margin_con, int(mean), pair_eps, _ops.ptr(losses), _ops.ptr(pairs), _ops.ptr(stats),
                None, None, None, None, _ops.ptr(ws), ws.numel(), _ops.stream_ptr(dev)),
                "irr_triplet_loss_fwd_bwd")
        if need_grad:
            ctx.save_for_backward(qr, pr, nr, stats)
            ctx.meta = (margin_cos, margin_con, int(mean), q.dtype, p.dtype, n.dtype)
        if want_pairs:
            ctx.mark_non_differentiable(pairs)
            return losses[0], losses[1], losses[2], losses[3], pairs
        return losses[0], losses[1], losses[2], losses[3]

    @staticmethod
    def backward(ctx, g0, g1, g2, g3, *unused):
        lib = _lib.load()
        qr, pr, nr, stats = ctx.saved_tensors
        margin_cos, margin_con, mean, qd, pd, nd = ctx.meta
        B, D = qr.shape
        dev = qr.device
        zero = None
        gs = []
        for g in (g0, g1, g2, g3):
            if g is None:
                if zero is None:
                    zero = torch.zeros((), dtype=torch.float32, device=dev)
                g = zero
            gs.append(g.detach().to(torch.float32).reshape(()))
        gout = torch.stack(gs).contiguous()
        dq, dp, dn = torch.empty_like(qr), torch.empty_like(pr), torch.empty_like(nr)
        with torch.cuda.device(dev):
            check(lib.irr_triplet_loss_bwd(_ops.ptr(qr), _ops.ptr(pr), _ops.ptr(nr), _ops.ptr(stats),
                                           _ops.ptr(gout), B, D, _ops.dtype_code(qr), margin_cos,
                                           margin_con, mean, _ops.ptr(dq), _ops.ptr(dp), _ops.ptr(dn),
                                           _ops.stream_ptr(dev)), "irr_triplet_loss_bwd")
        ng = ctx.needs_input_grad
        return (dq.to(qd) if ng[0] else None, dp.to(pd) if ng[1] else None,
                dn.to(nd) if ng[2] else None, None, None, None, None, None, None)


def triplet_losses(qry: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, margin: float = 0.3, *,
                   margin_con: Optional[float] = None, mean: bool = True,
                   pair_scores: bool = False, pair_eps: float = 1e-6,
                   autocast_exact: bool = True) -> TripletLosses:
    """The four embedding-loss scalars of a training / validation step from ONE pass over the
    triplets, each differentiable:

        cos_pos = CosineEmbeddingLoss(margin)(qry, pos,  1)     cos_neg = ...(qry, neg, -1)
        con_pos = ContrastiveLoss(margin_con)(qry, pos, 1.)     con_neg = ...(qry, neg, 0.)

    ``pair_scores=True`` also returns the row-wise ``cos(q_i,p_i)`` / ``cos(q_i,n_i)`` the
    reference logs as cos_sims / cos_unsims (train/train_efficient_cos_con_ce_loss.py:377-382).
    fp16 triplets (``precision=16``) are read as fp16 and ``pos - qry`` / ``neg - qry`` are rounded
    to fp16 like the reference's autocast does (``autocast_exact=False``: widened to fp32 first,
    which is more exact than — so not identical to — the reference).
    """
    mc = float(margin)
    mk = float(margin if margin_con is None else margin_con)
    out = _TripletLossFn.apply(qry, pos, neg, mc, mk, bool(mean), bool(pair_scores), float(pair_eps),
                               bool(autocast_exact))
    if pair_scores:
        return TripletLosses(out[0], out[1], out[2], out[3], out[4][0], out[4][1])
    return TripletLosses(out[0], out[1], out[2], out[3], None, None)


class TripletFwdBwd(NamedTuple):
    losses: torch.Tensor      # fp32[4]: cos_pos, cos_neg, con_pos, con_neg
    grad_qry: torch.Tensor
    grad_pos: torch.Tensor
    grad_neg: torch.Tensor
    pair_cos: Optional[torch.Tensor]  # [2,B]


def triplet_losses_fwd_bwd(qry: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor,
                           margin: float = 0.3, *, margin_con: Optional[float] = None,
                           mean: bool = True, grad_scale: Sequence[float] = (1.0, 1.0, 1.0, 1.0),
                           pair_scores: bool = False, pair_eps: float = 1e-6,
                           autocast_exact: bool = True) -> TripletFwdBwd:
    """Single-launch forward + backward: the four losses and the gradients of
    ``sum_j grad_scale[j] * loss_j`` w.r.t. qry / pos / neg, written in the same pass that reads the
    rows (2 x 3BD bytes of HBM traffic in total).  For training loops that own their backward; the
    autograd-integrated form is :func:`triplet_losses`."""
    lib = _lib.load()
    qr, pr, nr = _triplet_rows(qry, pos, neg, autocast_exact)
    B, D = qr.shape
    dev = qr.device
    losses = torch.empty(4, dtype=torch.float32, device=dev)
    pairs = torch.empty((2, B), dtype=torch.float32, device=dev) if pair_scores else None
    dq, dp, dn = torch.empty_like(qr), torch.empty_like(pr), torch.empty_like(nr)
    mc = float(margin)
    mk = float(margin if margin_con is None else margin_con)
    with torch.cuda.device(dev):
        ws = _ops.zeroed_scratch(dev, lib.irr_triplet_loss_workspace_bytes(B, D, _ops.dtype_code(qr)))
        check(lib.irr_triplet_loss_fwd_bwd(
            _ops.ptr(qr), _ops.ptr(pr), _ops.ptr(nr), B, D, _ops.dtype_code(qr), mc, mk, int(mean),
            float(pair_eps), _ops.ptr(losses), _ops.ptr(pairs), None, _ops.ptr(dq), _ops.ptr(dp),
            _ops.ptr(dn), _ops.f32x4(grad_scale), _ops.ptr(ws), ws.numel(), _ops.stream_ptr(dev)),
            "irr_triplet_loss_fwd_bwd")
    return TripletFwdBwd(losses, dq, dp, dn, pairs)
