"""imageretrievalresearch_b200 — B200-native (sm_100a) retrieval-ranking hot path of
vitasoftAI/ImageRetrievalResearch: cosine similarity -> top-k -> top1/top3 accounting and the
contrastive / cosine-embedding losses, as hand-written CUDA kernels behind a C ABI
(include/irr_b200.h, libirr_b200.so) with a torch-facing surface that mirrors the reference's calls.

There is no CPU fallback: every function raises if the CUDA library is missing or the tensors are
not on a CUDA device.
"""
from ._lib import IRR_MAX_K, IrrError, LIB_PATH, load as load_library
from .losses import (ContrastiveLoss, CosineEmbeddingLoss, TripletFwdBwd, TripletLosses,
                     triplet_losses, triplet_losses_fwd_bwd)
from .producer_consumer import CEPair, cross_entropy_pair, get_fm
from .retrieval import (CapturedSearch, CosineSimilarity, DedupTopK, Gallery, SearchPipeline, TopK, class_dedup_topk, cosine_topk,
                        top1_top3, top1_top3_dedup, topk_hits)
from .store import (GalleryBuilder, GalleryStore, GalleryWriter, StreamedGallery, block_ranges, gather_embeddings,
                    write_gallery)
from . import torch_ops  # registers torch.ops.irr_b200.*
from .sharded import PeerExchange, ShardedGallery, exchange_candidates, shard_bounds

__all__ = [
    "IRR_MAX_K", "IrrError", "LIB_PATH", "load_library",
    "ContrastiveLoss", "CosineEmbeddingLoss", "TripletLosses", "TripletFwdBwd",
    "triplet_losses", "triplet_losses_fwd_bwd",
    "CosineSimilarity", "Gallery", "TopK", "DedupTopK", "cosine_topk", "top1_top3", "topk_hits",
    "class_dedup_topk", "top1_top3_dedup", "get_fm", "cross_entropy_pair", "CEPair",
    "GalleryBuilder", "GalleryStore", "GalleryWriter", "StreamedGallery", "write_gallery", "gather_embeddings",
    "block_ranges", "torch_ops", "ShardedGallery", "PeerExchange", "CapturedSearch", "SearchPipeline", "exchange_candidates", "shard_bounds",
]
