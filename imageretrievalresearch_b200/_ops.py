"""Tensor-level plumbing between torch and the C ABI: dtype/layout checks, stream and workspace
handling.  PyTorch is used for device memory and streams only; all arithmetic is in csrc/."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import IRR_BF16, IRR_F32, IRR_MAX_K, IRR_ROW_STATS, check


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: the B200 retrieval-ranking path runs on CUDA only "
            "(there is no CPU fallback)")


def as_rows(t: torch.Tensor, name: str, keep_f16: bool = False) -> torch.Tensor:
    """[rows, D] contiguous fp32/bf16 CUDA tensor.  fp16 (autocast embeddings) is widened to fp32,
    which is what the reference's autocast does for every reduction on this path —
    except with ``keep_f16``: the row-norm / row-wise cosine kernels read fp16 directly (exact fp32
    arithmetic on the fp16 values, no widening copy), and the search sends fp16 rows to the
    tensor-core path when the caller opted in (``fp16_tensor_path``)."""
    _require_cuda(t, name)
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D [rows, D], got shape {tuple(t.shape)}")
    if t.dtype == torch.float64 or (t.dtype == torch.float16 and not (keep_f16 and t.shape[1] % 8 == 0)):
        t = t.float()
    if t.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        raise TypeError(f"{name}: unsupported dtype {t.dtype}")
    t = t.detach()
    if not t.is_contiguous():
        t = t.contiguous()
    return t


def dtype_code(t: torch.Tensor) -> int:
    return {torch.bfloat16: IRR_BF16, torch.float16: _lib.IRR_F16}.get(t.dtype, IRR_F32)


def same_kind(a: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp16 rows meeting fp32 rows: widen the fp16 side (what the reference's autocast does)."""
    if a.dtype != b.dtype:
        if a.dtype == torch.float16 and b.dtype == torch.float32:
            a = a.float()
        elif b.dtype == torch.float16 and a.dtype == torch.float32:
            b = b.float()
    return a, b


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_scratch: dict = {}
_zeroed: dict = {}


def scratch(device: torch.device, nbytes: int) -> torch.Tensor:
    """Reusable uninitialised workspace for the current (device, stream).  While a CUDA graph is
    being captured the workspace comes from the graph's own memory pool instead (it must live and
    die with the graph, not with this cache)."""
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def zeroed_scratch(device: torch.device, nbytes: int) -> torch.Tensor:
    """Workspace that is zero-filled once (self-resetting sync words, see irr_b200.h).  While a
    CUDA graph is being captured the buffer (and its zero-fill, recorded as a graph node) comes
    from the graph's own memory pool and is never cached: memory of that pool is recycled once the
    capture ends, so a cached sync word could be non-zero — or somebody else's tensor — later."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _zeroed.get(key)
    if buf is not None and buf.numel() >= nbytes:
        return buf            # allocated eagerly (ordinary pool, kept alive here): fine inside a capture too
    if torch.cuda.is_current_stream_capturing():
        return torch.zeros(max(nbytes, 256), dtype=torch.uint8, device=device)
    buf = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
    _zeroed[key] = buf
    return buf


def check_same(a: torch.Tensor, b: torch.Tensor, an: str, bn: str) -> None:
    if a.device != b.device:
        raise RuntimeError(f"{an} and {bn} are on different devices ({a.device} vs {b.device})")
    if a.dtype != b.dtype:
        raise TypeError(f"{an} and {bn} have different dtypes ({a.dtype} vs {b.dtype})")
    if a.shape[1] != b.shape[1]:
        raise RuntimeError(
            f"embedding widths differ: {an} has D={a.shape[1]}, {bn} has D={b.shape[1]}")


# -------------------------------------------------------------------------------------------------
# thin typed calls
# -------------------------------------------------------------------------------------------------
def row_inv_norms(x: torch.Tensor, eps: float) -> torch.Tensor:
    lib = _lib.load()
    x = as_rows(x, "x", keep_f16=True)
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.irr_row_inv_norms(ptr(x), x.shape[0], x.shape[1], dtype_code(x), eps, ptr(out),
                                    stream_ptr(x.device)), "irr_row_inv_norms")
    return out


def cosine_topk_raw(q: torch.Tensor, g: torch.Tensor, k: int, eps: float,
                    g_inv_norm: Optional[torch.Tensor], idx_offset: int,
                    out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    Q, D = q.shape
    N = g.shape[0]
    dt = dtype_code(q)
    if out is None:
        vals = torch.empty((Q, k), dtype=torch.float32, device=q.device)
        idx = torch.empty((Q, k), dtype=torch.int64, device=q.device)
    else:
        vals, idx = out
    with torch.cuda.device(q.device):
        need = lib.irr_cosine_topk_workspace_bytes(Q, N, D, k, dt)
        ws = scratch(q.device, need)
        check(lib.irr_cosine_topk(ptr(q), ptr(g), ptr(g_inv_norm), Q, N, D, k, dt, eps, idx_offset,
                                  ptr(vals), ptr(idx), ptr(ws), ws.numel(), stream_ptr(q.device)),
              "irr_cosine_topk")
    return vals, idx


def cosine_scores_bf16(q: torch.Tensor, g: torch.Tensor, eps: float) -> torch.Tensor:
    """Dense [Q,N] scores from the tensor-core kernel (test / bring-up aid)."""
    lib = _lib.load()
    q, g = as_rows(q, "q"), as_rows(g, "g")
    check_same(q, g, "q", "g")
    if q.dtype != torch.bfloat16:
        raise TypeError("cosine_scores_bf16 needs bf16 inputs")
    Q, D = q.shape
    N = g.shape[0]
    out = torch.empty((Q, N), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        ws = scratch(q.device, (N + Q) * 4 + 1024)
        check(lib.irr_cosine_scores_bf16(ptr(q), ptr(g), Q, N, D, eps, ptr(out), ptr(ws),
                                         ws.numel(), stream_ptr(q.device)),
              "irr_cosine_scores_bf16")
    return out


def topk_merge(cand_val: torch.Tensor, cand_idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[G,Q,k] candidate lists -> merged [Q,k]."""
    lib = _lib.load()
    _require_cuda(cand_val, "cand_val")
    _require_cuda(cand_idx, "cand_idx")
    if cand_val.dim() != 3 or cand_val.shape != cand_idx.shape:
        raise ValueError("cand_val / cand_idx must both be [G, Q, k]")
    cand_val = cand_val.contiguous().float()
    cand_idx = cand_idx.contiguous().long()
    G, Q, k = cand_val.shape
    vals = torch.empty((Q, k), dtype=torch.float32, device=cand_val.device)
    idx = torch.empty((Q, k), dtype=torch.int64, device=cand_val.device)
    with torch.cuda.device(cand_val.device):
        check(lib.irr_topk_merge(ptr(cand_val), ptr(cand_idx), G, Q, k, ptr(vals), ptr(idx),
                                 stream_ptr(cand_val.device)), "irr_topk_merge")
    return vals, idx


def topk_merge_packed(gathered: torch.Tensor, G: int, Q: int, k: int, idx_byte_offset: int,
                      rank_bytes: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge straight out of an all-gather receive buffer of G packed messages
    [fp32 scores Q*k | pad | int64 indices Q*k] (no unpack copies)."""
    lib = _lib.load()
    vals = torch.empty((Q, k), dtype=torch.float32, device=gathered.device)
    idx = torch.empty((Q, k), dtype=torch.int64, device=gathered.device)
    base = gathered.data_ptr()
    with torch.cuda.device(gathered.device):
        check(lib.irr_topk_merge_strided(base, rank_bytes // 4, base + idx_byte_offset,
                                         rank_bytes // 8, G, Q, k, ptr(vals), ptr(idx),
                                         stream_ptr(gathered.device)), "irr_topk_merge_strided")
    return vals, idx


def topk_exchange_bytes(G: int, Q: int, k: int) -> int:
    return int(_lib.load().irr_topk_exchange_bytes(G, Q, k))


def topk_exchange_merge(vals: Optional[torch.Tensor], idx: Optional[torch.Tensor], peer_ptrs,
                        rank: int, Q: int, k: int, buf_bytes: int, mode: int,
                        device: torch.device,
                        out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
                        ) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """irr_topk_exchange_merge: push this rank's [Q,k] lists into every rank's exchange buffer over
    peer memory, wait for all G lists, merge (one kernel for k <= 16).  peer_ptrs: the G device
    pointers of the exchange buffers as mapped into this process."""
    lib = _lib.load()
    G = len(peer_ptrs)
    arr = (C.c_void_p * G)(*[int(p) for p in peer_ptrs])
    ov = oi = None
    if mode != _lib.IRR_XCHG_PUSH:
        if out is None:
            ov = torch.empty((Q, k), dtype=torch.float32, device=device)
            oi = torch.empty((Q, k), dtype=torch.int64, device=device)
        else:
            ov, oi = out
    with torch.cuda.device(device):
        check(lib.irr_topk_exchange_merge(ptr(vals), ptr(idx), arr, G, rank, Q, k, buf_bytes, mode,
                                          ptr(ov), ptr(oi), stream_ptr(device)),
              "irr_topk_exchange_merge")
    return None if ov is None else (ov, oi)


def topk_hits(idx: torch.Tensor, q_label: Optional[torch.Tensor], g_label: Optional[torch.Tensor],
              instance_offset: int) -> torch.Tensor:
    lib = _lib.load()
    _require_cuda(idx, "indices")
    idx = idx.contiguous().long()
    Q, k = idx.shape
    N = 0
    if q_label is not None:
        q_label = q_label.to(device=idx.device, dtype=torch.int64).contiguous()
        g_label = g_label.to(device=idx.device, dtype=torch.int64).contiguous()
        if q_label.numel() != Q:
            raise ValueError(f"query_labels has {q_label.numel()} entries for {Q} queries")
        N = g_label.numel()
    out = torch.empty(2, dtype=torch.int64, device=idx.device)
    with torch.cuda.device(idx.device):
        check(lib.irr_topk_hits(ptr(idx), Q, k, ptr(q_label), ptr(g_label), N, instance_offset,
                                ptr(out), stream_ptr(idx.device)), "irr_topk_hits")
    return out


def class_dedup(vals: torch.Tensor, idx: torch.Tensor, g_label: torch.Tensor, n_distinct: int,
                q_label: Optional[torch.Tensor]):
    lib = _lib.load()
    _require_cuda(idx, "indices")
    dev = idx.device
    vals = vals.to(device=dev, dtype=torch.float32).contiguous()
    idx = idx.contiguous().long()
    g_label = g_label.to(device=dev, dtype=torch.int64).contiguous()
    Q, k = idx.shape
    if vals.shape != idx.shape:
        raise ValueError("values / indices must have the same [Q, k] shape")
    hits = None
    if q_label is not None:
        q_label = q_label.to(device=dev, dtype=torch.int64).contiguous()
        if q_label.numel() != Q:
            raise ValueError(f"query_labels has {q_label.numel()} entries for {Q} queries")
        hits = torch.empty(2, dtype=torch.int64, device=dev)
    out_l = torch.empty((Q, n_distinct), dtype=torch.int64, device=dev)
    out_i = torch.empty((Q, n_distinct), dtype=torch.int64, device=dev)
    out_v = torch.empty((Q, n_distinct), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.irr_topk_class_dedup(ptr(vals), ptr(idx), Q, k, ptr(g_label), g_label.numel(),
                                       n_distinct, ptr(q_label), ptr(out_l), ptr(out_i), ptr(out_v),
                                       ptr(hits), stream_ptr(dev)), "irr_topk_class_dedup")
    return out_l, out_i, out_v, hits


def pair_cosine(x1: torch.Tensor, x2: torch.Tensor, eps: float) -> torch.Tensor:
    lib = _lib.load()
    N, D = x2.shape
    out = torch.empty(N, dtype=torch.float32, device=x2.device)
    with torch.cuda.device(x2.device):
        check(lib.irr_pair_cosine(ptr(x1), x1.shape[0], ptr(x2), N, D, dtype_code(x2), eps,
                                  ptr(out), stream_ptr(x2.device)), "irr_pair_cosine")
    return out


def f32x4(vals) -> "C.Array":
    return (C.c_float * 4)(*[float(v) for v in vals])


__all__ = ["IRR_MAX_K", "IRR_ROW_STATS"]
