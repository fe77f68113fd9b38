"""Row-sharded gallery across the GPUs of one box.

Queries are replicated, the gallery is split into contiguous row ranges (rank r owns
``shard_bounds(N, G, r)``).  Every rank runs the top-k kernel on its shard with its first row as
index offset, ONE all-gather moves the ``[Q,k]`` (score, global index) lists (``Q*k*12`` bytes per
rank, over NCCL / NVLink on GPUs), and the merge kernel folds the ``[G,Q,k]`` candidates; ties
resolve to the lower global index, which with contiguous shards equals the single-GPU answer.
The reference has no sharded retrieval (its gallery is the rank-local batch,
train/train_efficient_cos_con_ce_loss.py:385) — this is the scale-out of that same loop.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _ops
from .retrieval import Gallery, TopK


def shard_bounds(total_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row range [lo, hi) of `rank`; the first N % G ranks get one extra row."""
    if world_size < 1 or not (0 <= rank < world_size) or total_rows < 0:
        raise ValueError(f"bad shard request: N={total_rows}, world={world_size}, rank={rank}")
    base, rem = divmod(total_rows, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _packed_layout(Q: int, k: int) -> Tuple[int, int]:
    """(byte offset of the int64 indices, total bytes) of one rank's packed candidate message:
    [fp32 scores Q*k | pad to 8 B | int64 indices Q*k], total a multiple of 8."""
    val_bytes = Q * k * 4
    off = (val_bytes + 7) // 8 * 8
    return off, off + Q * k * 8


def pack_candidates(vals: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    Q, k = vals.shape
    off, total = _packed_layout(Q, k)
    buf = torch.zeros(total, dtype=torch.uint8, device=vals.device)
    buf[: Q * k * 4].view(torch.float32).copy_(vals.reshape(-1))
    buf[off:].view(torch.int64).copy_(idx.reshape(-1))
    return buf


def unpack_candidates(gathered: torch.Tensor, G: int, Q: int, k: int
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    off, total = _packed_layout(Q, k)
    g = gathered.view(G, total)
    vals = g[:, : Q * k * 4].contiguous().view(torch.float32).view(G, Q, k)
    idx = g[:, off:].contiguous().view(torch.int64).view(G, Q, k)
    return vals, idx


def exchange_candidates(vals: torch.Tensor, idx: torch.Tensor,
                        group: Optional[dist.ProcessGroup] = None
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
    """The path's one exchange step: all-gather every rank's [Q,k] lists -> [G,Q,k] on every rank.
    One collective of Q*k*12 bytes per rank (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
    G = dist.get_world_size(group)
    Q, k = vals.shape
    msg = pack_candidates(vals, idx)
    out = torch.empty(G * msg.numel(), dtype=torch.uint8, device=msg.device)
    dist.all_gather_into_tensor(out, msg, group=group)
    return unpack_candidates(out, G, Q, k)


class ShardedGallery:
    """This rank's shard of a row-sharded gallery plus the process group it is sharded over."""

    def __init__(self, local_embeddings: torch.Tensor, total_rows: int,
                 group: Optional[dist.ProcessGroup] = None, eps: float = 1e-6,
                 cache_norms: bool = True) -> None:
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        lo, hi = shard_bounds(total_rows, self.world, self.rank)
        if local_embeddings.shape[0] != hi - lo:
            raise ValueError(
                f"rank {self.rank} must hold rows [{lo},{hi}) = {hi - lo} rows, "
                f"got {local_embeddings.shape[0]}")
        self.total_rows = total_rows
        self.local = Gallery(local_embeddings, eps=eps, first_row=lo, cache_norms=cache_norms)

    @classmethod
    def from_full(cls, full_gallery: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                  **kw) -> "ShardedGallery":
        lo, hi = shard_bounds(full_gallery.shape[0], dist.get_world_size(group), dist.get_rank(group))
        return cls(full_gallery[lo:hi], full_gallery.shape[0], group, **kw)

    def search(self, queries: torch.Tensor, k: int) -> TopK:
        if k > self.total_rows:
            raise RuntimeError("selected index k out of range")
        if self.world == 1:
            return self.local.search(queries, k, allow_short=True)
        # the local top-k writes straight into this rank's packed message, ONE all-gather moves
        # the G messages, and the merge kernel reads them in place (irr_topk_merge_strided)
        Q = queries.shape[0]
        off, total = _packed_layout(Q, k)
        dev = self.local.embeddings.device
        msg = torch.empty(total, dtype=torch.uint8, device=dev)
        vals = msg[: Q * k * 4].view(torch.float32).view(Q, k)
        idx = msg[off:].view(torch.int64).view(Q, k)
        self.local.search(queries, k, allow_short=True, out=(vals, idx))
        gathered = torch.empty(self.world * total, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered, msg, group=self.group)
        mv, mi = _ops.topk_merge_packed(gathered, self.world, Q, k, off, total)
        return TopK(mv, mi)
