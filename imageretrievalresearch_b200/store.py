"""Gallery store: the data format on either side of the search path.

The reference keeps its gallery only in RAM: the notebook appends every pooled embedding to Python
lists and concatenates them (inference/training_analysis.ipynb:222-231,257), and inference.py
returns them in a dict (inference/inference.py:245, 'normalized_embeddings').  This module gives
that gallery a form that survives the process and scales past one GPU's memory:

* ``GalleryWriter`` / ``write_gallery`` — append ``get_fm`` outputs batch by batch into ONE file:
  a 4 KiB header, then row-major ``[rows, D]`` embeddings (fp32, bf16 or fp16, exactly the layout the
  kernels' TMA descriptors read — no transpose, no padding), then optional int64 labels and the
  fp32 inverse row norms ``1/max(|g|, eps)`` the search kernels consume.
* ``GalleryStore`` — memory-maps the file; ``load`` a row range to a resident :class:`Gallery`,
  ``load_shard`` this rank's contiguous rows as a :class:`ShardedGallery`, or ``stream`` it.
* ``StreamedGallery`` — searches a gallery that stays in host memory (pinned tensor or a mapped
  file) or is simply larger than HBM: row blocks are copied host->device on a copy stream into a
  ring of device buffers while the top-k kernel runs on the previous block; each block's ``[Q,k]``
  list (global indices via ``idx_offset``) is folded into the running list with the merge kernel, so
  the result is bit-identical to the resident search (ties -> lower global index).
* ``gather_embeddings`` — DDP-wide in-batch gallery for the training-loop evaluation
  (train/train_efficient_cos_con_ce_loss.py:270-281 ranks only the rank-local batch).

All arithmetic stays in the CUDA library; this file is host-side IO and stream plumbing.
"""
from __future__ import annotations

import os
import struct
from typing import Iterator, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.distributed as dist

from . import _ops
from .retrieval import Gallery, TopK, cosine_topk
from .sharded import ShardedGallery, shard_bounds

MAGIC = b"IRRGAL01"
VERSION = 1
HEADER_BYTES = 4096
ALIGN = 4096
# magic, version, dtype code, rows, dim, eps, emb_off, label_off, norm_off, file_bytes
_HEADER = struct.Struct("<8sIIQIfQQQQ")
# codes = irr_dtype (include/irr_b200.h); fp16 = the embeddings precision=16 training produces
_DTYPES = {0: (torch.float32, np.dtype("<f4"), 4), 1: (torch.bfloat16, np.dtype("<u2"), 2),
           2: (torch.float16, np.dtype("<f2"), 2)}
_CODES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def _align(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN


def block_ranges(rows: int, block_rows: int) -> List[Tuple[int, int]]:
    """Row blocks [lo, hi) a streamed scan visits, in order."""
    if rows < 0 or block_rows < 1:
        raise ValueError(f"bad block plan: rows={rows}, block_rows={block_rows}")
    return [(lo, min(lo + block_rows, rows)) for lo in range(0, rows, block_rows)]


def _as_numpy_rows(t: torch.Tensor) -> np.ndarray:
    """CPU view of a [rows, D] fp32 / bf16 tensor in the file's element type (bf16 -> raw uint16)."""
    t = t.detach().contiguous().cpu()
    if t.dtype == torch.bfloat16:
        return t.view(torch.uint16).numpy()
    return t.numpy()


class GalleryWriter:
    """Incremental writer: ``append(embeddings[, labels])`` per batch, ``close()`` seals the header.

    Rows are written in call order, so row ``i`` of the file is the ``i``-th embedding produced —
    the index the search returns.  Inverse norms are computed with the library's kernel when the
    batch lives on a CUDA device (``irr_row_inv_norms``); CPU batches are stored without norms and
    the reader computes them on load.
    """

    def __init__(self, path: Union[str, os.PathLike], dim: int, dtype: torch.dtype = torch.bfloat16,
                 eps: float = 1e-6, with_labels: bool = False) -> None:
        if dtype not in _CODES:
            raise TypeError(f"unsupported gallery dtype {dtype} (fp32, bf16 or fp16)")
        if dim < 1 or (dim * _DTYPES[_CODES[dtype]][2]) % 16 != 0:
            raise ValueError(f"D={dim}: rows must be a multiple of 16 bytes (irr_b200.h alignment contract)")
        self.path = os.fspath(path)
        self.dim, self.dtype, self.eps, self.with_labels = dim, dtype, float(eps), with_labels
        self.rows = 0
        self._labels: List[np.ndarray] = []
        self._norms: List[np.ndarray] = []
        self._norms_complete = True
        self._f = open(self.path, "wb")
        self._f.write(b"\0" * HEADER_BYTES)      # sealed by close()
        self._closed = False

    def append(self, embeddings: torch.Tensor, labels: Optional[torch.Tensor] = None) -> None:
        if self._closed:
            raise RuntimeError("writer is closed")
        if embeddings.dim() != 2 or embeddings.shape[1] != self.dim:
            raise ValueError(f"expected [rows, {self.dim}] embeddings, got {tuple(embeddings.shape)}")
        if self.with_labels != (labels is not None):
            raise ValueError("labels must be given for every batch or for none")
        e = embeddings.detach().to(self.dtype)
        if e.is_cuda:
            self._norms.append(_ops.row_inv_norms(e, self.eps).cpu().numpy())
        else:
            self._norms_complete = False
        self._f.write(_as_numpy_rows(e).tobytes())
        if labels is not None:
            lab = labels.detach().to(torch.int64).cpu().numpy().reshape(-1)
            if lab.shape[0] != e.shape[0]:
                raise ValueError("one label per embedding row")
            self._labels.append(lab)
        self.rows += e.shape[0]

    def close(self) -> None:
        if self._closed:
            return
        f = self._f
        esz = _DTYPES[_CODES[self.dtype]][2]
        emb_off = HEADER_BYTES
        end = emb_off + self.rows * self.dim * esz
        label_off = norm_off = 0
        if self.with_labels:
            label_off = _align(end)
            f.write(b"\0" * (label_off - end))
            f.write(np.concatenate(self._labels).astype("<i8").tobytes() if self._labels else b"")
            end = label_off + self.rows * 8
        if self._norms_complete and self.rows > 0:
            norm_off = _align(end)
            f.write(b"\0" * (norm_off - end))
            f.write(np.concatenate(self._norms).astype("<f4").tobytes())
            end = norm_off + self.rows * 4
        f.seek(0)
        f.write(_HEADER.pack(MAGIC, VERSION, _CODES[self.dtype], self.rows, self.dim, self.eps,
                             emb_off, label_off, norm_off, end))
        f.close()
        self._closed = True

    def __enter__(self) -> "GalleryWriter":
        return self

    def abort(self) -> None:
        """Give up on a partially written gallery: the file is removed (its header was never
        sealed — the magic is still zero — so even a copy taken now is rejected by the reader)."""
        if self._closed:
            return
        self._f.close()
        self._closed = True
        try:
            os.unlink(self.path)
        except OSError:
            pass

    def __exit__(self, exc_type, exc, tb) -> None:
        # an append that raised part-way leaves fewer rows than the caller meant to write: sealing
        # that would hand later searches a gallery that silently misses rows
        if exc_type is not None:
            self.abort()
        else:
            self.close()


class GalleryBuilder:
    """A resident gallery that grows batch by batch on the device — the GPU-side form of the
    notebook's ``fms_poss_all.append(...)`` / ``torch.cat`` accumulation
    (inference/training_analysis.ipynb:222-231,257).  Rows land in a preallocated ``[capacity, D]``
    buffer (doubled when full), their inverse norms are computed once as they arrive, and
    ``gallery()`` is a zero-copy :class:`Gallery` over the rows appended so far."""

    def __init__(self, dim: int, dtype: torch.dtype = torch.bfloat16, device="cuda",
                 capacity: int = 1 << 16, eps: float = 1e-6) -> None:
        if dtype not in _CODES:
            raise TypeError(f"unsupported gallery dtype {dtype} (fp32, bf16 or fp16)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GalleryBuilder lives on a CUDA device (there is no CPU fallback)")
        self.dim, self.dtype, self.eps = dim, dtype, float(eps)
        self.rows = 0
        self._emb = torch.empty((max(capacity, 1), dim), dtype=dtype, device=self.device)
        self._norm = torch.empty(max(capacity, 1), dtype=torch.float32, device=self.device)
        self._lab: Optional[torch.Tensor] = None

    def _reserve(self, need: int) -> None:
        cap = self._emb.shape[0]
        if need <= cap:
            return
        while cap < need:
            cap *= 2
        emb = torch.empty((cap, self.dim), dtype=self.dtype, device=self.device)
        norm = torch.empty(cap, dtype=torch.float32, device=self.device)
        emb[: self.rows].copy_(self._emb[: self.rows])
        norm[: self.rows].copy_(self._norm[: self.rows])
        self._emb, self._norm = emb, norm
        if self._lab is not None:
            lab = torch.empty(cap, dtype=torch.int64, device=self.device)
            lab[: self.rows].copy_(self._lab[: self.rows])
            self._lab = lab

    def append(self, embeddings: torch.Tensor, labels: Optional[torch.Tensor] = None) -> None:
        if embeddings.dim() != 2 or embeddings.shape[1] != self.dim:
            raise ValueError(f"expected [rows, {self.dim}] embeddings, got {tuple(embeddings.shape)}")
        if (labels is not None) != (self._lab is not None) and self.rows > 0:
            raise ValueError("labels must be given for every batch or for none")
        n = embeddings.shape[0]
        self._reserve(self.rows + n)
        dst = self._emb[self.rows: self.rows + n]
        dst.copy_(embeddings.detach())                      # converts dtype / device as needed
        self._norm[self.rows: self.rows + n].copy_(_ops.row_inv_norms(dst, self.eps))
        if labels is not None:
            if self._lab is None:
                self._lab = torch.empty(self._emb.shape[0], dtype=torch.int64, device=self.device)
            self._lab[self.rows: self.rows + n].copy_(labels.detach().reshape(-1))
        self.rows += n

    @property
    def labels(self) -> Optional[torch.Tensor]:
        return None if self._lab is None else self._lab[: self.rows]

    def gallery(self, first_row: int = 0) -> Gallery:
        g = Gallery(self._emb[: self.rows], eps=self.eps, first_row=first_row, cache_norms=False)
        g.inv_norm = self._norm[: self.rows]
        return g


def write_gallery(path: Union[str, os.PathLike], embeddings: torch.Tensor,
                  labels: Optional[torch.Tensor] = None, eps: float = 1e-6,
                  chunk_rows: int = 1 << 16) -> None:
    """Write a whole ``[rows, D]`` tensor (CPU or CUDA, fp32 / bf16) as one gallery file."""
    dt = torch.float32 if embeddings.dtype == torch.float64 else embeddings.dtype
    with GalleryWriter(path, embeddings.shape[1], dt, eps, labels is not None) as w:
        for lo, hi in block_ranges(embeddings.shape[0], chunk_rows):
            w.append(embeddings[lo:hi], None if labels is None else labels[lo:hi])


class GalleryStore:
    """Read side: a memory-mapped gallery file."""

    def __init__(self, path: Union[str, os.PathLike]) -> None:
        self.path = os.fspath(path)
        size = os.path.getsize(self.path)
        if size < HEADER_BYTES:
            raise ValueError(f"{self.path}: not a gallery file (shorter than its header)")
        with open(self.path, "rb") as f:
            head = f.read(_HEADER.size)
        (magic, version, code, rows, dim, eps, emb_off, label_off, norm_off, file_bytes) = _HEADER.unpack(head)
        if magic != MAGIC:
            raise ValueError(f"{self.path}: bad magic {magic!r} (unsealed or foreign file)")
        if version != VERSION:
            raise ValueError(f"{self.path}: unsupported version {version}")
        if code not in _DTYPES:
            raise ValueError(f"{self.path}: unknown dtype code {code}")
        self.dtype, self._np_dtype, esz = _DTYPES[code]
        self.rows, self.dim, self.eps = int(rows), int(dim), float(eps)
        if file_bytes != size or emb_off + self.rows * self.dim * esz > size:
            raise ValueError(f"{self.path}: truncated (header says {file_bytes} bytes, file has {size})")
        self._emb = np.memmap(self.path, dtype=self._np_dtype, mode="r", offset=emb_off,
                              shape=(self.rows, self.dim)) if self.rows else None
        self._lab = (np.memmap(self.path, dtype="<i8", mode="r", offset=label_off, shape=(self.rows,))
                     if label_off and self.rows else None)
        self._norm = (np.memmap(self.path, dtype="<f4", mode="r", offset=norm_off, shape=(self.rows,))
                      if norm_off and self.rows else None)

    @property
    def has_labels(self) -> bool:
        return self._lab is not None

    @property
    def has_inv_norm(self) -> bool:
        return self._norm is not None

    def _rows(self, lo: int, hi: Optional[int]) -> Tuple[int, int]:
        hi = self.rows if hi is None else hi
        if not (0 <= lo <= hi <= self.rows):
            raise IndexError(f"rows [{lo},{hi}) outside a gallery of {self.rows} rows")
        return lo, hi

    def embeddings_np(self, lo: int = 0, hi: Optional[int] = None) -> np.ndarray:
        lo, hi = self._rows(lo, hi)
        return self._emb[lo:hi] if self.rows else np.zeros((0, self.dim), self._np_dtype)

    def embeddings(self, lo: int = 0, hi: Optional[int] = None) -> torch.Tensor:
        """CPU tensor (a copy) of rows [lo, hi) in the stored dtype."""
        a = np.array(self.embeddings_np(lo, hi), copy=True, order="C")
        t = torch.from_numpy(a)
        return t.view(torch.bfloat16) if self.dtype == torch.bfloat16 else t

    def labels(self, lo: int = 0, hi: Optional[int] = None) -> Optional[torch.Tensor]:
        if self._lab is None:
            return None
        lo, hi = self._rows(lo, hi)
        return torch.from_numpy(np.array(self._lab[lo:hi], copy=True))

    def inv_norm(self, lo: int = 0, hi: Optional[int] = None) -> Optional[torch.Tensor]:
        if self._norm is None:
            return None
        lo, hi = self._rows(lo, hi)
        return torch.from_numpy(np.array(self._norm[lo:hi], copy=True))

    def load(self, device: Union[str, torch.device], lo: int = 0, hi: Optional[int] = None,
             chunk_rows: int = 1 << 16) -> Gallery:
        """Rows [lo, hi) as a resident :class:`Gallery` (indices are global: first_row = lo); the
        stored inverse norms are used when the file has them, else computed on the device."""
        lo, hi = self._rows(lo, hi)
        device = torch.device(device)
        dev = torch.empty((hi - lo, self.dim), dtype=self.dtype, device=device)
        for a, b in block_ranges(hi - lo, chunk_rows):
            dev[a:b].copy_(self.embeddings(lo + a, lo + b))
        g = Gallery(dev, eps=self.eps, first_row=lo, cache_norms=not self.has_inv_norm)
        if self.has_inv_norm:
            g.inv_norm = self.inv_norm(lo, hi).to(device)
        return g

    def load_shard(self, device: Union[str, torch.device],
                   group: Optional[dist.ProcessGroup] = None, **kw) -> ShardedGallery:
        """This rank's contiguous rows (``shard_bounds``) as a :class:`ShardedGallery`."""
        lo, hi = shard_bounds(self.rows, dist.get_world_size(group), dist.get_rank(group))
        local = self.load(device, lo, hi)
        sg = ShardedGallery(local.embeddings, self.rows, group, eps=self.eps, cache_norms=False, **kw)
        sg.local.inv_norm = local.inv_norm
        return sg

    def stream(self, device: Union[str, torch.device], block_rows: int = 1 << 18,
               buffers: int = 2) -> "StreamedGallery":
        return StreamedGallery(self, block_rows, device, eps=self.eps, buffers=buffers)


class StreamedGallery:
    """Search a gallery that lives in host memory (or in a mapped file) block by block.

    source: a CPU ``[N, D]`` fp32 / bf16 tensor (pinned = copied straight from it; pageable = staged
    through pinned bounce buffers) or a :class:`GalleryStore`.  Per block: host->device copy on a
    dedicated copy stream into one of ``buffers`` device blocks, ``irr_cosine_topk`` on the compute
    stream with ``idx_offset`` = the block's first row, merge into the running ``[Q,k]`` list.  The
    copy of block b+1 overlaps the kernel of block b.  The scan is bound by the host link
    (PCIe / C2C), not by HBM: bytes per search = N*D*s over the link.
    """

    def __init__(self, source: Union[torch.Tensor, GalleryStore], block_rows: int,
                 device: Union[str, torch.device], eps: float = 1e-6, buffers: int = 2,
                 first_row: int = 0, inv_norm: Optional[torch.Tensor] = None) -> None:
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("StreamedGallery streams to a CUDA device (there is no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if buffers < 2:
            raise ValueError("need at least two device buffers to overlap copy and search")
        self.store = source if isinstance(source, GalleryStore) else None
        if self.store is not None:
            self.rows, self.dim, self.dtype = self.store.rows, self.store.dim, self.store.dtype
            self._host_norm = self.store.inv_norm() if self.store.has_inv_norm else None
            self._src = None
        else:
            if source.is_cuda or source.dim() != 2:
                raise ValueError("source must be a CPU [N, D] tensor or a GalleryStore")
            if source.dtype not in _CODES:
                raise TypeError(f"unsupported gallery dtype {source.dtype}")
            self._src = source.contiguous()
            self.rows, self.dim, self.dtype = source.shape[0], source.shape[1], source.dtype
            self._host_norm = None if inv_norm is None else inv_norm.detach().float().cpu().contiguous()
        self.eps, self.first_row = float(eps), int(first_row)
        self.block_rows = int(min(max(block_rows, 1), max(self.rows, 1)))
        self.blocks = block_ranges(self.rows, self.block_rows)
        self._direct = self._src is not None and self._src.is_pinned()
        self._dev = [torch.empty((self.block_rows, self.dim), dtype=self.dtype, device=self.device)
                     for _ in range(buffers)]
        self._dev_norm = ([torch.empty(self.block_rows, dtype=torch.float32, device=self.device)
                           for _ in range(buffers)] if self._host_norm is not None else None)
        if self._host_norm is not None and not self._host_norm.is_pinned():
            self._host_norm = self._host_norm.pin_memory()
        # pageable / mapped sources: pinned bounce buffers, one per device buffer
        self._bounce = (None if self._direct else
                        [torch.empty((self.block_rows, self.dim), dtype=self.dtype).pin_memory()
                         for _ in range(buffers)])
        self._copy_stream = torch.cuda.Stream(self.device)
        self._ready = [torch.cuda.Event() for _ in range(buffers)]     # block landed in buffer i
        self._free = [torch.cuda.Event() for _ in range(buffers)]      # search done with buffer i
        self._h2d_done = [torch.cuda.Event() for _ in range(buffers)]  # bounce buffer i reusable
        self._used = [False] * buffers
        self.bytes_per_search = self.rows * self.dim * _DTYPES[_CODES[self.dtype]][2]

    def _host_block(self, lo: int, hi: int, i: int) -> torch.Tensor:
        if self._direct:
            return self._src[lo:hi]
        b = self._bounce[i][: hi - lo]
        if self._used[i]:
            self._h2d_done[i].synchronize()          # previous copy out of this bounce buffer
        if self.store is not None:
            dst = b.view(torch.uint16).numpy() if self.dtype == torch.bfloat16 else b.numpy()
            np.copyto(dst, self.store.embeddings_np(lo, hi))
        else:
            b.copy_(self._src[lo:hi])
        return b

    def _enqueue_copy(self, blk: int) -> None:
        lo, hi = self.blocks[blk]
        i = blk % len(self._dev)
        host = self._host_block(lo, hi, i)
        cs = self._copy_stream
        if self._used[i]:
            cs.wait_event(self._free[i])             # the search that last read buffer i is done
        with torch.cuda.stream(cs):
            self._dev[i][: hi - lo].copy_(host, non_blocking=True)
            if self._dev_norm is not None:
                self._dev_norm[i][: hi - lo].copy_(self._host_norm[lo:hi], non_blocking=True)
            self._h2d_done[i].record(cs)
            self._ready[i].record(cs)
        self._used[i] = True

    def search(self, queries: torch.Tensor, k: int) -> TopK:
        if k > self.rows:
            raise RuntimeError("selected index k out of range")
        q = _ops.as_rows(queries, "queries")
        if q.device != self.device:
            raise RuntimeError(f"queries are on {q.device}, the stream target is {self.device}")
        Q = q.shape[0]
        cur = torch.cuda.current_stream(self.device)
        self._copy_stream.wait_stream(cur)           # buffers may still be read by earlier work
        pair_v = torch.empty((2, Q, k), dtype=torch.float32, device=self.device)
        pair_i = torch.empty((2, Q, k), dtype=torch.int64, device=self.device)
        nb = len(self._dev)
        for b in range(min(nb - 1, len(self.blocks))):
            self._enqueue_copy(b)
        for b, (lo, hi) in enumerate(self.blocks):
            if b + nb - 1 < len(self.blocks):
                self._enqueue_copy(b + nb - 1)       # keep nb-1 blocks in flight ahead of the search
            i = b % nb
            cur.wait_event(self._ready[i])
            slot = 0 if b == 0 else 1
            cosine_topk(q, self._dev[i][: hi - lo], k, self.eps,
                        gallery_inv_norm=None if self._dev_norm is None else self._dev_norm[i][: hi - lo],
                        idx_offset=self.first_row + lo, allow_short=True,
                        out=(pair_v[slot], pair_i[slot]))
            self._free[i].record(cur)
            if b > 0:
                mv, mi = _ops.topk_merge(pair_v, pair_i)
                pair_v[0].copy_(mv)
                pair_i[0].copy_(mi)
        return TopK(pair_v[0], pair_i[0])


def gather_embeddings(local: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                      labels: Optional[torch.Tensor] = None
                      ) -> Tuple[torch.Tensor, Optional[torch.Tensor], int]:
    """DDP-wide in-batch gallery: all-gather every rank's ``[B, D]`` embeddings (and labels) in rank
    order.  Returns (gathered ``[G*B, D]``, gathered labels or None, first row of this rank's
    block).  Every rank must contribute the same B.  Evaluation only: no autograd through the
    collective (the reference's ranking metrics carry no gradient either)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    local = local.detach().contiguous()
    out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    lab = None
    if labels is not None:
        labels = labels.detach().to(torch.int64).contiguous()
        lab = torch.empty(world * labels.shape[0], dtype=torch.int64, device=labels.device)
        dist.all_gather_into_tensor(lab, labels, group=group)
    return out, lab, rank * local.shape[0]


def iter_blocks(store: GalleryStore, block_rows: int) -> Iterator[Tuple[int, torch.Tensor]]:
    """(first row, CPU tensor) for each block of a store, in order."""
    for lo, hi in block_ranges(store.rows, block_rows):
        yield lo, store.embeddings(lo, hi)
