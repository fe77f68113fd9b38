"""Retrieval-ranking surface: the same calls the reference makes, fused and batched.

Reference call sites this module stands in for (paths under the reference tree):
  cos = CosineSimilarity(dim=1, eps=1e-6)            train/train_efficient_cos_con_ce_loss.py:89
  sim = cos(fm_ims[idx].unsqueeze(0), fm_poss)       :273, :385       ipynb:238
  vals, inds = torch.topk(sim, k=3)                  :276, :388       inference/inference.py:235,240
  top3 / top1 class tests                            :279-281, :390-392
  instance tests  len(inds[idx == inds])             inference/inference.py:237,242
  paired scores cos(q[i][None], p[i][None])          :377, :381       ipynb:232,234
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Tuple

import torch

from . import _ops
from ._lib import IRR_MAX_K


class TopK(NamedTuple):
    values: torch.Tensor   # [Q, k] fp32, descending
    indices: torch.Tensor  # [Q, k] int64, ties -> lower gallery index


class CosineSimilarity(torch.nn.Module):
    """Drop-in for ``torch.nn.CosineSimilarity(dim=1, eps)`` on 2-D embeddings.

    ``forward(x1, x2)`` accepts the two shapes the reference uses: row-wise pairs ``[N,D] x [N,D]``
    and one query against a gallery ``[1,D]`` / ``[D]`` x ``[N,D]``; returns ``[N]`` fp32.
    Evaluation-only (the reference uses it for metrics): the result carries no autograd graph.
    """

    def __init__(self, dim: int = 1, eps: float = 1e-8) -> None:
        super().__init__()
        if dim not in (1, -1):
            raise ValueError("only dim=1 (the embedding axis of [rows, D]) is supported")
        self.dim, self.eps = dim, eps

    def forward(self, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        if x1.dim() == 1:
            x1 = x1.unsqueeze(0)
        if x2.dim() == 1:
            x2 = x2.unsqueeze(0)
        if x2.shape[0] == 1 and x1.shape[0] != 1:
            x1, x2 = x2, x1  # cosine is symmetric
        a, b = _ops.same_kind(_ops.as_rows(x1, "x1", keep_f16=True),
                              _ops.as_rows(x2, "x2", keep_f16=True))
        _ops.check_same(a, b, "x1", "x2")
        if a.shape[0] not in (1, b.shape[0]):
            raise RuntimeError(
                f"The size of tensor a ({a.shape[0]}) must match the size of tensor b "
                f"({b.shape[0]}) at non-singleton dimension 0")
        return _ops.pair_cosine(a, b, self.eps)


def cosine_topk(queries: torch.Tensor, gallery: torch.Tensor, k: int, eps: float = 1e-6, *,
                gallery_inv_norm: Optional[torch.Tensor] = None, idx_offset: int = 0,
                allow_short: bool = False,
                out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                fp16_tensor_path: bool = False) -> TopK:
    """Fused ``topk(cos(q_i[None], gallery), k)`` for every query row.

    Row ``i`` of the result is what the reference's per-query loop yields for query ``i``:
    values sorted descending, int64 indices; ties resolve to the lower gallery index (torch.topk
    leaves tie order unspecified).  fp32 inputs take the exact FFMA path, bf16 inputs the tcgen05
    tensor-core path; both accumulate and emit fp32.  fp16 inputs (what the reference's
    ``precision=16`` training produces) are widened to fp32 like the reference's autocast does —
    unless ``fp16_tensor_path=True``, which sends them to the tensor cores as fp16 (kind::f16):
    same ranking rule, scores within ~2e-5 of the fp32 result (tensor-core accumulation truncates;
    measured 1.0e-5 at D=1920), i.e. outside the fp32-mode bar of 1e-5 relative, hence opt-in.  The Q x N score matrix is never
    materialised.

    gallery_inv_norm: cached ``1/max(|g|, eps)`` per gallery row (see :class:`Gallery`).
    idx_offset: added to the returned indices (first row of this shard).
    allow_short: a shard with fewer than ``k`` rows pads with (-inf, -1) instead of raising.
    out: optional preallocated contiguous (values fp32 [Q,k], indices int64 [Q,k]) to write into.
    """
    q, g = _ops.same_kind(_ops.as_rows(queries, "queries", keep_f16=fp16_tensor_path),
                          _ops.as_rows(gallery, "gallery", keep_f16=fp16_tensor_path))
    _ops.check_same(q, g, "queries", "gallery")
    if not isinstance(k, int) or k < 1:
        raise ValueError(f"k must be a positive int, got {k!r}")
    if k > IRR_MAX_K:
        raise ValueError(f"k={k} exceeds the supported maximum of {IRR_MAX_K}")
    if k > g.shape[0] and not allow_short:
        raise RuntimeError("selected index k out of range")  # torch.topk's message
    if gallery_inv_norm is not None:
        gallery_inv_norm = gallery_inv_norm.to(device=g.device, dtype=torch.float32).contiguous()
        if gallery_inv_norm.numel() != g.shape[0]:
            raise ValueError("gallery_inv_norm must have one entry per gallery row")
    if out is not None:
        ov, oi = out
        if (ov.shape != (q.shape[0], k) or oi.shape != (q.shape[0], k) or ov.dtype != torch.float32
                or oi.dtype != torch.int64 or not ov.is_contiguous() or not oi.is_contiguous()
                or ov.device != q.device or oi.device != q.device):
            raise ValueError("out must be contiguous (fp32 [Q,k], int64 [Q,k]) on the query device")
    vals, idx = _ops.cosine_topk_raw(q, g, k, eps, gallery_inv_norm, idx_offset, out)
    return TopK(vals, idx)


def topk_hits(indices: torch.Tensor, query_labels: Optional[torch.Tensor] = None,
              gallery_labels: Optional[torch.Tensor] = None, instance_offset: int = 0
              ) -> torch.Tensor:
    """int64[2] = (#queries hit at rank 1, #queries hit anywhere in the k columns), on the device.

    class flavour (labels given): hit when ``query_labels[i] == gallery_labels[indices[i, j]]``
    (train/train_efficient_cos_con_ce_loss.py:279-281); instance flavour (no labels): hit when
    ``indices[i, j] == i + instance_offset`` (inference/inference.py:237,242).
    """
    if (query_labels is None) != (gallery_labels is None):
        raise ValueError("give both query_labels and gallery_labels, or neither")
    return _ops.topk_hits(indices, query_labels, gallery_labels, instance_offset)


def top1_top3(queries: torch.Tensor, gallery: torch.Tensor,
              query_labels: Optional[torch.Tensor] = None,
              gallery_labels: Optional[torch.Tensor] = None, *, k: int = 3, eps: float = 1e-6
              ) -> Tuple[torch.Tensor, torch.Tensor, TopK]:
    """The whole evaluation loop of training_step / validation_step in three launches.

    Returns (top1, topk) as 0-d fp32 device tensors holding hits / Q — the quantities the reference
    logs as ``train_top1`` / ``train_top3`` — plus the TopK lists.  No host synchronisation.
    """
    res = cosine_topk(queries, gallery, k, eps)
    hits = topk_hits(res.indices, query_labels, gallery_labels)
    frac = hits.to(torch.float32) / float(res.indices.shape[0])
    return frac[0], frac[1], res


class DedupTopK(NamedTuple):
    labels: torch.Tensor            # [Q, n] int64: first n distinct labels in rank order (-1 pad)
    indices: torch.Tensor           # [Q, n] int64: gallery row each label was first seen at
    values: torch.Tensor            # [Q, n] fp32: its cosine score
    hits: Optional[torch.Tensor]    # int64[2] (top1, topn) when query labels were given


def class_dedup_topk(topk: TopK, gallery_labels: torch.Tensor, n_distinct: int = 3,
                     query_labels: Optional[torch.Tensor] = None) -> DedupTopK:
    """The notebook's class de-duplication (inference/training_analysis.ipynb:240-251): walk each
    query's ranked list and keep the first ``n_distinct`` DISTINCT gallery labels; with
    ``query_labels`` also count top1 (label equals the first distinct label) and topn (label among
    them), on the device."""
    if not 1 <= n_distinct <= 8:
        raise ValueError("n_distinct must be in 1..8")
    l, i, v, h = _ops.class_dedup(topk.values, topk.indices, gallery_labels, n_distinct, query_labels)
    return DedupTopK(l, i, v, h)


def top1_top3_dedup(queries: torch.Tensor, gallery: torch.Tensor, query_labels: torch.Tensor,
                    gallery_labels: torch.Tensor, *, k: int = 150, n_distinct: int = 3,
                    eps: float = 1e-6) -> Tuple[torch.Tensor, torch.Tensor, DedupTopK]:
    """The working inference evaluation of the reference (ipynb:231-257) without the per-query
    Python loop: top-``k`` (150) cosine rows per query, first 3 distinct classes, top1 / top3 as
    hits / Q (0-d fp32 device tensors)."""
    res = cosine_topk(queries, gallery, min(k, gallery.shape[0]), eps)
    d = class_dedup_topk(res, gallery_labels, n_distinct, query_labels)
    frac = d.hits.to(torch.float32) / float(res.indices.shape[0])
    return frac[0], frac[1], d


class Gallery:
    """A resident gallery shard: embeddings plus cached inverse row norms.

    The reference re-normalises the whole gallery for every query
    (train/train_efficient_cos_con_ce_loss.py:273 inside the ``for idx`` loop); a gallery that is
    searched more than once caches ``1/max(|g|, eps)`` instead.
    """

    def __init__(self, embeddings: torch.Tensor, eps: float = 1e-6, first_row: int = 0,
                 cache_norms: bool = True, fp16_tensor_path: bool = False) -> None:
        self.fp16_tensor_path = fp16_tensor_path
        self.embeddings = _ops.as_rows(embeddings, "embeddings", keep_f16=fp16_tensor_path)
        self.eps = eps
        self.first_row = int(first_row)
        self.inv_norm = _ops.row_inv_norms(self.embeddings, eps) if cache_norms else None

    @property
    def num_rows(self) -> int:
        return self.embeddings.shape[0]

    def search(self, queries: torch.Tensor, k: int, allow_short: bool = False,
               out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> TopK:
        return cosine_topk(queries, self.embeddings, k, self.eps, gallery_inv_norm=self.inv_norm,
                           idx_offset=self.first_row, allow_short=allow_short, out=out,
                           fp16_tensor_path=self.fp16_tensor_path)

    def capture(self, num_queries: int, k: int) -> "CapturedSearch":
        """Record ``search`` for a fixed batch shape as a CUDA graph (see CapturedSearch)."""
        return CapturedSearch(lambda q, kk: self.search(q, kk), num_queries,
                              self.embeddings.shape[1], k, self.embeddings.dtype,
                              self.embeddings.device)


class CapturedSearch:
    """A search over a resident gallery recorded once as a CUDA graph and replayed with one launch.

    The reference pays ~10 kernel launches and several host syncs per query; the fused path is
    down to 3-5 launches per *batch*, and for small batches (Q <= 64 against a sharded gallery the
    kernels take tens of microseconds) those launches are what is left.  ``search_fn(queries, k)``
    must enqueue only graph-capturable work on the current stream (the single-GPU search and the
    peer-memory exchange both qualify: no host syncs, no allocation outside torch's graph pool,
    the exchange epoch lives in device memory).  Collective for a sharded gallery: every rank
    captures and replays in lockstep.
    """

    def __init__(self, search_fn, num_queries: int, dim: int, k: int, dtype: torch.dtype,
                 device: torch.device, warmup: int = 2, on_release=None) -> None:
        self.queries = torch.zeros((num_queries, dim), dtype=dtype, device=device)
        self.k = k
        self._on_release = on_release
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(warmup):          # allocate workspaces / exchange buffers before capture
                search_fn(self.queries, k)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = search_fn(self.queries, k)

    def __call__(self, queries: Optional[torch.Tensor] = None, k: Optional[int] = None) -> TopK:
        """Replay.  `queries` (optional) is copied into the static input first; the returned TopK
        tensors are the graph's static outputs, overwritten by the next replay.  `k`, if given,
        must be the captured k (lets a replay stand in wherever a ``search(queries, k)`` is taken)."""
        if k is not None and k != self.k:
            raise ValueError(f"this search was captured for k={self.k}, not k={k}")
        if queries is not None:
            self.queries.copy_(queries)
        self.graph.replay()
        return self.result

    def release(self) -> None:
        """Drop the graph (and tell the owner that its baked-in pointers are no longer live)."""
        if self.graph is not None:
            self.graph = None
            self.result = None
            if self._on_release is not None:
                self._on_release()
                self._on_release = None

    def __del__(self):
        try:
            self.release()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    # the TopK this returns is overwritten by the next replay (SearchPipeline orders its read-back
    # of batch n before the replay for batch n+1 when it sees this)
    static_outputs = True


class SearchPipeline:
    """Serving loop over host-resident query batches with copy / search / read-back overlap.

    The reference's inference loop moves one batch at a time to the GPU, runs its per-query loop
    and pulls every index back with ``.item()`` (inference/inference.py:225-245).  Here batch i+1 is
    copied host->device on a copy stream while batch i is searched, and batch i's ``[Q,k]`` lists
    travel back to pinned host memory on a third stream while batch i+1 is searched; the host only
    waits for the batch it is about to hand out.  ``search_fn(queries, k)`` is any of the searches
    of this package (``Gallery.search``, ``ShardedGallery.search``, a ``CapturedSearch``...).
    """

    def __init__(self, search_fn, num_queries: int, dim: int, k: int, dtype: torch.dtype,
                 device: torch.device, depth: int = 2, static_outputs: Optional[bool] = None) -> None:
        """static_outputs: the search returns the SAME output tensors on every call (a
        CapturedSearch replay) — the next search may then only start once the previous batch's
        read-back has finished.  Default: detected from ``search_fn.static_outputs``."""
        if static_outputs is None:
            static_outputs = bool(getattr(search_fn, "static_outputs", False))
        self.static_outputs = static_outputs
        self._last_r: Optional[int] = None
        if depth < 2:
            raise ValueError("depth >= 2: one batch in flight while the next is being copied")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("SearchPipeline feeds a CUDA device (there is no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.search_fn, self.k, self.depth, self.device = search_fn, k, depth, device
        self.shape, self.dtype = (num_queries, dim), dtype
        self._q = [torch.empty(self.shape, dtype=dtype, device=device) for _ in range(depth)]
        # one more result slot than batches in flight: what was handed out stays untouched until
        # the caller asks for the next batch
        self._v = [torch.empty((num_queries, k), dtype=torch.float32).pin_memory() for _ in range(depth + 1)]
        self._i = [torch.empty((num_queries, k), dtype=torch.int64).pin_memory() for _ in range(depth + 1)]
        self._h2d, self._d2h = torch.cuda.Stream(device), torch.cuda.Stream(device)
        self._landed = [torch.cuda.Event() for _ in range(depth)]    # queries of slot s on the device
        self._searched = [torch.cuda.Event() for _ in range(depth)]  # search of slot s enqueued & done
        self._back = [torch.cuda.Event() for _ in range(depth + 1)]  # results of slot r on the host
        self.h2d_bytes_per_batch = num_queries * dim * torch.empty((), dtype=dtype).element_size()
        self.d2h_bytes_per_batch = num_queries * k * 12

    def _enqueue(self, s: int, r: int, q_host: torch.Tensor, used: bool) -> None:
        if tuple(q_host.shape) != self.shape or q_host.dtype != self.dtype or q_host.is_cuda:
            raise ValueError(f"expected a host batch {self.shape} of {self.dtype}")
        cur = torch.cuda.current_stream(self.device)
        if used:
            self._h2d.wait_event(self._searched[s])      # slot's previous search has read its queries
        with torch.cuda.stream(self._h2d):
            self._q[s].copy_(q_host, non_blocking=True)
            self._landed[s].record(self._h2d)
        cur.wait_event(self._landed[s])
        if self.static_outputs and self._last_r is not None:
            cur.wait_event(self._back[self._last_r])     # the outputs about to be overwritten are on the host
        self._last_r = r
        res = self.search_fn(self._q[s], self.k)
        self._searched[s].record(cur)
        self._d2h.wait_event(self._searched[s])
        with torch.cuda.stream(self._d2h):
            if not self.static_outputs:
                res.values.record_stream(self._d2h)
                res.indices.record_stream(self._d2h)
            self._v[r].copy_(res.values, non_blocking=True)
            self._i[r].copy_(res.indices, non_blocking=True)
            self._back[r].record(self._d2h)

    def run(self, host_batches):
        """Yield (values, indices) — pinned host tensors, valid until the next batch is requested —
        for every batch of ``host_batches`` (pinned host ``[Q, D]`` tensors), in order."""
        pending = []          # result slots in flight, oldest first
        n = 0
        self._last_r = None
        for q_host in host_batches:
            if len(pending) == self.depth:               # hand out the oldest batch first
                r = pending.pop(0)
                self._back[r].synchronize()
                yield self._v[r], self._i[r]
            r = n % (self.depth + 1)
            self._enqueue(n % self.depth, r, q_host, used=n >= self.depth)
            pending.append(r)
            n += 1
        for r in pending:
            self._back[r].synchronize()
            yield self._v[r], self._i[r]
