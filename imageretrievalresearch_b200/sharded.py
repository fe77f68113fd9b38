"""Row-sharded gallery across the GPUs of one box.

Queries are replicated, the gallery is split into contiguous row ranges (rank r owns
``shard_bounds(N, G, r)``).  Every rank runs the top-k kernel on its shard with its first row as
index offset, ONE all-gather moves the ``[Q,k]`` (score, global index) lists (``Q*k*12`` bytes per
rank, over NCCL / NVLink on GPUs), and the merge kernel folds the ``[G,Q,k]`` candidates; ties
resolve to the lower global index, which with contiguous shards equals the single-GPU answer.
The reference has no sharded retrieval (its gallery is the rank-local batch,
train/train_efficient_cos_con_ce_loss.py:385) — this is the scale-out of that same loop.

Two transports carry the exchange step:
  "peer"        (GPUs of one box) every rank's exchange buffer is mapped into all processes
                (torch symmetric memory = CUDA VMM + handle exchange) and ONE kernel per rank stores
                its lists into every peer over NVLink, publishes an epoch flag, waits for the G
                lists and merges them (csrc/topk_exchange.cu) — no collective library on the path;
  "collective"  one all-gather (NCCL on GPUs, gloo in the CPU tests) + the merge kernel.
"auto" (default) takes "peer" when all ranks can map each other's memory and agrees on the
answer across the group, so that no rank ever waits on a transport the others did not choose.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib, _ops
from .retrieval import CapturedSearch, Gallery, TopK


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


class PeerExchange:
    """The G exchange buffers of a process group, one per rank, each mapped into every process.

    Allocation and mapping are host plumbing; the protocol — remote stores, epoch flags, wait,
    merge — is the kernel in csrc/topk_exchange.cu behind ``irr_topk_exchange_merge``.
    mapping = "symm": torch symmetric memory (CUDA VMM allocation, handles passed between the
    processes by torch);  mapping = "ipc": an ordinary torch allocation exported with CUDA IPC
    (``irr_peer_export`` / ``irr_peer_import``), handles all-gathered over the process group.
    All ranks must make the same sequence of calls on one stream.
    """

    def __init__(self, group: Optional[dist.ProcessGroup], device: torch.device, nbytes: int,
                 mapping: str = "symm") -> None:
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > _lib.IRR_MAX_PEERS:
            raise RuntimeError(f"peer exchange supports up to {_lib.IRR_MAX_PEERS} ranks")
        self.device = device
        self.nbytes = int(nbytes)
        self.mapping = mapping
        self._imported: List[int] = []
        self.handle = None
        if mapping not in ("symm", "ipc"):
            raise ValueError("mapping must be 'symm' or 'ipc'")
        # step 1, local: allocate.  Every rank reports the outcome BEFORE anybody enters the
        # collective mapping step, so an allocation failure on one rank cannot leave the others
        # waiting in a rendezvous it never joins.
        err = None
        try:
            if mapping == "symm":
                import torch.distributed._symmetric_memory as symm_mem
                self.buf = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=device)
            else:
                self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        except Exception as e:  # noqa: BLE001
            err = e
        if not _agree(err is None, self.group, device):
            raise RuntimeError(f"peer exchange buffer ({mapping}) could not be allocated on every rank"
                               + (f": {type(err).__name__}: {err}" if err is not None else ""))
        # step 2, collective: map every rank's buffer into this process
        if mapping == "symm":
            self.handle = symm_mem.rendezvous(self.buf, self.group)
            self.ptrs: List[int] = [int(p) for p in self.handle.buffer_ptrs]
        else:
            self.ptrs = self._map_ipc()
        if len(self.ptrs) != self.world or self.ptrs[self.rank] != self.buf.data_ptr():
            raise RuntimeError("peer mapping returned an unexpected pointer table")

    def activate(self) -> None:
        """Collective, called once every rank reported a successful mapping: flags / epoch start
        at zero on every rank before anyone pushes."""
        self.buf.zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)

    def _map_ipc(self) -> List[int]:
        import ctypes as C
        lib = _lib.load()
        handle = (C.c_uint8 * 64)()
        off = C.c_uint64(0)
        with torch.cuda.device(self.device):
            st = lib.irr_peer_export(self.buf.data_ptr(), handle, C.byref(off))
        # every rank takes part in the gather even if its export failed (no mismatched collectives)
        mine = (bytes(handle), int(off.value)) if st == 0 else None
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        if any(e is None for e in everyone):
            bad = [r for r, e in enumerate(everyone) if e is None]
            raise RuntimeError(f"irr_peer_export failed on rank(s) {bad} (status here: {st})")
        ptrs: List[int] = []
        for r, (h, o) in enumerate(everyone):
            if r == self.rank:
                ptrs.append(self.buf.data_ptr())
                continue
            base = C.c_void_p()
            hb = (C.c_uint8 * 64).from_buffer_copy(h)
            with torch.cuda.device(self.device):
                _lib.check(lib.irr_peer_import(hb, C.byref(base)), "irr_peer_import")
            self._imported.append(int(base.value))
            ptrs.append(int(base.value) + o)
        return ptrs

    def fits(self, Q: int, k: int) -> bool:
        return _ops.topk_exchange_bytes(self.world, Q, k) <= self.nbytes

    def exchange_merge(self, vals: torch.Tensor, idx: torch.Tensor,
                       out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
        Q, k = vals.shape
        return _ops.topk_exchange_merge(vals, idx, self.ptrs, self.rank, Q, k, self.nbytes,
                                        _lib.IRR_XCHG_FUSED, self.device, out)

    # lagged exchange: per search "merge_pushed() (the previous search) then push()"; the last
    # search of a stream is collected with one more merge_pushed()
    def push(self, vals: torch.Tensor, idx: torch.Tensor) -> None:
        Q, k = vals.shape
        _ops.topk_exchange_merge(vals, idx, self.ptrs, self.rank, Q, k, self.nbytes,
                                 _lib.IRR_XCHG_PUSH, self.device)

    def merge_pushed(self, Q: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Wait for and merge the G lists of the epoch this rank pushed last."""
        return _ops.topk_exchange_merge(None, None, self.ptrs, self.rank, Q, k, self.nbytes,
                                        _lib.IRR_XCHG_MERGE, self.device)

    def close(self) -> None:
        """Collective: no rank may still be storing into a buffer that is about to be unmapped."""
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
        lib = _lib.load()
        with torch.cuda.device(self.device):
            for base in self._imported:
                lib.irr_peer_close(base)
        self._imported = []
        dist.barrier(self.group)
        self.handle = None
        self.buf = None


def _agree(ok: bool, group: Optional[dist.ProcessGroup], device: torch.device) -> bool:
    """True only if every rank of the group says True."""
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())


class ShardedGallery:
    """This rank's shard of a row-sharded gallery plus the process group it is sharded over."""

    def __init__(self, local_embeddings: torch.Tensor, total_rows: int,
                 group: Optional[dist.ProcessGroup] = None, eps: float = 1e-6,
                 cache_norms: bool = True, exchange: str = "auto") -> None:
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        exchange = os.environ.get("IRR_EXCHANGE", exchange)
        if exchange not in ("auto", "peer", "collective"):
            raise ValueError("exchange must be 'auto', 'peer' or 'collective'")
        self._exchange_mode = exchange
        self._peer: Optional[PeerExchange] = None
        self._peer_failed = False
        self.exchange_error: Optional[str] = None
        self._lagged: Optional[Tuple[int, int]] = None   # (Q, k) of the search whose result is pending
        self._live_captures = 0      # CUDA graphs holding this exchange buffer's mapped pointers
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

    def close(self) -> None:
        """Collective: release the peer-exchange buffers (no-op for the collective transport)."""
        if self._peer is not None:
            self._peer.close()
            self._peer = None

    @property
    def transport(self) -> str:
        """'peer' once a search went over peer memory, else 'collective' (what the next search
        will use is decided on first use when exchange='auto')."""
        return "peer" if self._peer is not None else "collective"

    def _peer_exchange(self, Q: int, k: int) -> Optional[PeerExchange]:
        """The peer transport sized for (Q, k), or None if this group uses the collective.  Growing
        or creating the buffers is collective; every rank takes the same branch because Q, k and
        the agreed outcome are the same everywhere."""
        dev = self.local.embeddings.device
        if (self._exchange_mode == "collective" or self._peer_failed or dev.type != "cuda"
                or self.world > _lib.IRR_MAX_PEERS):
            return None
        if self._peer is not None and self._peer.fits(Q, k):
            return self._peer
        need = max(2 * _ops.topk_exchange_bytes(self.world, Q, k), 1 << 20)
        if self._peer is not None:
            if self._live_captures > 0:
                # a captured graph has the current buffers' peer pointers baked into its kernel
                # arguments: replaying it after the buffers were unmapped would store into freed
                # memory on every rank
                raise RuntimeError(
                    f"a search with Q={Q}, k={k} needs a larger peer-exchange buffer, but "
                    f"{self._live_captures} captured search(es) still use the current one: "
                    "release() them first, or capture the largest (Q, k) first")
            self._peer.close()
            self._peer = None
        errs = []
        for mapping in os.environ.get("IRR_PEER_MAPPING", "symm,ipc").split(","):
            peer, err = None, None
            try:
                peer = PeerExchange(self.group, dev, need, mapping.strip())
            except Exception as e:  # noqa: BLE001 - any mapping failure selects the next option
                err = f"{mapping}: {type(e).__name__}: {e}"
            if _agree(peer is not None, self.group, dev):
                peer.activate()
                self._peer = peer
                return peer
            errs.append(err or f"{mapping}: another rank could not map peer memory")
        self._peer_failed = True
        self.exchange_error = "; ".join(errs)
        if self._exchange_mode == "peer":
            raise RuntimeError(f"exchange='peer' requested but unavailable: {self.exchange_error}")
        return None

    def capture(self, num_queries: int, k: int) -> CapturedSearch:
        """Collective: record the sharded search for a fixed batch shape as a CUDA graph on every
        rank (local top-k + the peer-memory exchange kernel; replay = one launch per rank).  Needs
        the peer transport for world > 1 — a collective library call is not recorded here."""
        emb = self.local.embeddings
        if self.world > 1 and self._peer_exchange(num_queries, k) is None:
            raise RuntimeError("capture() needs the peer-memory exchange (exchange='peer'/'auto' on "
                               f"GPUs of one box); unavailable: {self.exchange_error}")
        def released():
            self._live_captures -= 1
        cap = CapturedSearch(lambda q, kk: self.search(q, kk), num_queries, emb.shape[1], k,
                             emb.dtype, emb.device, on_release=released)
        self._live_captures += 1
        return cap

    def search_lagged(self, queries: torch.Tensor, k: int) -> Optional[TopK]:
        """A stream of searches with one search of slack between the ranks.  Enqueues this
        search (local top-k + push of its lists to every peer) and returns the merged result of
        the PREVIOUS ``search_lagged`` call (None on the first call); :meth:`flush` collects the
        last one.  ``search`` ends every call in a rendezvous, so every step costs the slowest of
        G kernels; here the rendezvous is with the peers' previous push, which has long happened
        unless a peer is more than one whole search behind.  Same results as ``search``, one call
        later.  Needs the peer-memory exchange, k <= 16, and the same (Q, k) from call to call."""
        if k > self.total_rows:
            raise RuntimeError("selected index k out of range")
        Q = queries.shape[0]
        if self.world == 1:
            prev, self._lagged_single = getattr(self, "_lagged_single", None), \
                self.local.search(queries, k, allow_short=True)
            return prev
        if k > _lib.IRR_MAX_K_FUSED:
            raise ValueError("search_lagged supports k <= 16")
        if self._lagged is not None and self._lagged != (Q, k):
            raise ValueError("search_lagged needs the same (Q, k) from call to call; flush() first")
        peer = self._peer_exchange(Q, k)
        if peer is None:
            raise RuntimeError("search_lagged needs the peer-memory exchange; unavailable: "
                               f"{self.exchange_error}")
        lv, li = self.local.search(queries, k, allow_short=True)
        prev = None
        if self._lagged is not None:
            prev = TopK(*peer.merge_pushed(Q, k))    # merge n-1 BEFORE push n (see topk_exchange.cu)
        peer.push(lv, li)
        self._lagged = (Q, k)
        return prev

    def flush(self) -> Optional[TopK]:
        """The result of the last :meth:`search_lagged` call (None if there is none pending)."""
        if self.world == 1:
            prev, self._lagged_single = getattr(self, "_lagged_single", None), None
            return prev
        if self._lagged is None:
            return None
        Q, k = self._lagged
        self._lagged = None
        return TopK(*self._peer.merge_pushed(Q, k))

    def search(self, queries: torch.Tensor, k: int) -> TopK:
        if k > self.total_rows:
            raise RuntimeError("selected index k out of range")
        if self._lagged is not None:
            raise RuntimeError("a lagged search is pending: flush() before a plain search")
        if self.world == 1:
            return self.local.search(queries, k, allow_short=True)
        peer = self._peer_exchange(queries.shape[0], k)
        if peer is not None:
            # local top-k, then ONE kernel: store into every peer over NVLink, flag, wait, merge
            lv, li = self.local.search(queries, k, allow_short=True)
            mv, mi = peer.exchange_merge(lv, li)
            return TopK(mv, mi)
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
