"""The steps immediately before and after the retrieval-ranking path.

Reference (paths under the reference tree):
  get_fm(fm): AvgPool2d((H,W)) + reshape -> [B,C]             train/train_efficient_cos_con_ce_loss.py:103-122
  loss_ce = ce_loss(lbl_ims, clss) + ce_loss(lbl_poss, clss)  train/train_efficient_cos_con_ce_loss.py:160,240-242
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from . import _lib, _ops
from ._lib import IRR_BF16, IRR_F16, IRR_F32, check

_DT = {torch.float32: IRR_F32, torch.bfloat16: IRR_BF16, torch.float16: IRR_F16}


def _code(t: torch.Tensor, what: str) -> int:
    if t.dtype not in _DT:
        raise TypeError(f"{what}: unsupported dtype {t.dtype}")
    return _DT[t.dtype]


class _GetFmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fm, out_dtype):
        lib = _lib.load()
        _ops._require_cuda(fm, "fm")
        if fm.dim() != 4:
            raise ValueError(f"fm must be [B,C,H,W], got shape {tuple(fm.shape)}")
        x = fm.detach()
        if not x.is_contiguous():
            x = x.contiguous()           # channels_last feature maps are re-laid out once
        B, C, H, W = x.shape
        out = torch.empty((B, C), dtype=out_dtype, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.irr_avgpool_fwd(_ops.ptr(x), _code(x, "fm"), B * C, H * W, _ops.ptr(out),
                                      _code(out, "out"), _ops.stream_ptr(x.device)), "irr_avgpool_fwd")
        ctx.shape, ctx.in_dtype = (B, C, H, W), fm.dtype
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        B, C, H, W = ctx.shape
        g = grad_out.detach()
        if g.dtype not in (torch.float32, torch.bfloat16):
            g = g.float()
        g = g.contiguous()
        gfm = torch.empty((B, C, H, W), dtype=ctx.in_dtype, device=g.device)
        with torch.cuda.device(g.device):
            check(lib.irr_avgpool_bwd(_ops.ptr(g), _code(g, "grad"), B * C, H * W, _ops.ptr(gfm),
                                      _code(gfm, "grad_fm"), _ops.stream_ptr(g.device)), "irr_avgpool_bwd")
        return gfm, None


def get_fm(fm: torch.Tensor, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Drop-in for the reference's ``get_fm`` (train/train_efficient_cos_con_ce_loss.py:103-122):
    ``[B,C,H,W] -> [B,C]`` global average pool, differentiable.  ``out_dtype`` (fp32 / bf16) lets the
    embedding row be written directly in the dtype the gallery is stored in; default: the input's
    dtype (fp16 inputs produce fp32 rows, which is what the path's kernels consume)."""
    if out_dtype is None:
        out_dtype = torch.float32 if fm.dtype == torch.float16 else fm.dtype
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("out_dtype must be torch.float32 or torch.bfloat16")
    return _GetFmFn.apply(fm, out_dtype)


class CEPair(NamedTuple):
    loss: torch.Tensor      # 0-d fp32: ce(logits_a, target) + ce(logits_b, target)  (differentiable)
    loss_a: torch.Tensor    # 0-d fp32, detached
    loss_b: torch.Tensor    # 0-d fp32, detached


class _CEPairFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, target, ignore_index):
        lib = _lib.load()
        _ops._require_cuda(a, "logits")
        if a.dim() != 2 or a.shape != b.shape:
            raise ValueError("logits must both be [B, C]")
        if a.dtype != b.dtype:
            raise TypeError("both logits tensors must have the same dtype")
        xa, xb = a.detach().contiguous(), b.detach().contiguous()
        tgt = target.detach().to(device=xa.device, dtype=torch.int64).contiguous()
        B, C = xa.shape
        if tgt.numel() != B:
            raise ValueError(f"target has {tgt.numel()} entries for {B} rows")
        out = torch.empty(3, dtype=torch.float32, device=xa.device)
        with torch.cuda.device(xa.device):
            need = lib.irr_ce_pair_workspace_bytes(B)
            ws = torch.empty(need, dtype=torch.uint8, device=xa.device)   # owned by this call's graph node
            check(lib.irr_ce_pair_fwd(_ops.ptr(xa), _ops.ptr(xb), _ops.ptr(tgt), B, C, _code(xa, "logits"),
                                      ignore_index, _ops.ptr(out), _ops.ptr(ws), ws.numel(),
                                      _ops.stream_ptr(xa.device)), "irr_ce_pair_fwd")
        ctx.save_for_backward(xa, xb, tgt, ws)
        ctx.ignore_index = ignore_index
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, g, _unused):
        lib = _lib.load()
        xa, xb, tgt, ws = ctx.saved_tensors
        B, C = xa.shape
        gg = g.detach().to(torch.float32).reshape(1).contiguous()
        da, db = torch.empty_like(xa), torch.empty_like(xb)
        with torch.cuda.device(xa.device):
            check(lib.irr_ce_pair_bwd(_ops.ptr(xa), _ops.ptr(xb), _ops.ptr(tgt), B, C, _code(xa, "logits"),
                                      ctx.ignore_index, _ops.ptr(gg), _ops.ptr(ws), _ops.ptr(da),
                                      _ops.ptr(db), _ops.stream_ptr(xa.device)), "irr_ce_pair_bwd")
        return da, db, None, None


def cross_entropy_pair(logits_a: torch.Tensor, logits_b: torch.Tensor, target: torch.Tensor,
                       ignore_index: int = -100) -> CEPair:
    """``CrossEntropyLoss()(logits_a, target) + CrossEntropyLoss()(logits_b, target)`` — the
    reference's ``loss_ce`` (train/train_efficient_cos_con_ce_loss.py:240-242) — forward in one
    launch (+ a one-CTA deterministic finish) and backward in one launch for both tensors."""
    total, parts = _CEPairFn.apply(logits_a, logits_b, target, int(ignore_index))
    return CEPair(total, parts[1], parts[2])
