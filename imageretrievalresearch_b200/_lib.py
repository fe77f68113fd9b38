"""ctypes binding of libirr_b200.so — the C ABI declared in include/irr_b200.h.

The product path has no CPU fallback: if the library is missing this module raises, and every
wrapper raises on a non-zero irr_status.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# IRR_B200_LIB: load another build of the same library (A/B measurements of two builds in one
# session — profiles/); not an API
LIB_PATH = Path(os.environ["IRR_B200_LIB"]) if os.environ.get("IRR_B200_LIB") else _PKG / "libirr_b200.so"

IRR_F32, IRR_BF16, IRR_F16 = 0, 1, 2
IRR_MAX_K = 256
IRR_MAX_K_FUSED = 16
IRR_ROW_STATS = 8
IRR_LOSS_COSINE_EMBEDDING, IRR_LOSS_CONTRASTIVE = 1, 2
IRR_MAX_PEERS = 16
IRR_XCHG_FUSED, IRR_XCHG_PUSH, IRR_XCHG_MERGE = 0, 1, 2

_i32, _i64, _f32, _sz, _vp = C.c_int32, C.c_int64, C.c_float, C.c_size_t, C.c_void_p

# name -> (restype, argtypes); kept in one table so that tests can check it against the header
SIGNATURES = {
    "irr_version": (_i32, []),
    "irr_status_string": (C.c_char_p, [_i32]),
    "irr_profile_next_topk": (None, [_vp, _vp]),
    "irr_debug_occupy_sms": (_i32, [_i32, _i32, _i64, _vp]),
    "irr_cosine_topk_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32, _i32]),
    "irr_cosine_topk": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _f32, _i64, _vp, _vp,
                               _vp, _sz, _vp]),
    "irr_cosine_scores_bf16": (_i32, [_vp, _vp, _i64, _i64, _i32, _f32, _vp, _vp, _sz, _vp]),
    "irr_row_inv_norms": (_i32, [_vp, _i64, _i32, _i32, _f32, _vp, _vp]),
    "irr_topk_merge": (_i32, [_vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp]),
    "irr_topk_merge_strided": (_i32, [_vp, _i64, _vp, _i64, _i32, _i64, _i32, _vp, _vp, _vp]),
    "irr_topk_exchange_bytes": (_sz, [_i32, _i64, _i32]),
    "irr_topk_exchange_merge": (_i32, [_vp, _vp, C.POINTER(_vp), _i32, _i32, _i64, _i32, _sz, _i32,
                                       _vp, _vp, _vp]),
    "irr_peer_export": (_i32, [_vp, _vp, C.POINTER(C.c_uint64)]),
    "irr_peer_import": (_i32, [_vp, C.POINTER(_vp)]),
    "irr_peer_close": (_i32, [_vp]),
    "irr_cosine_topk_sharded_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32, _i32]),
    "irr_cosine_topk_sharded": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _f32, _i64,
                                       C.POINTER(_vp), _i32, _i32, _sz, _vp, _vp, _vp, _sz, _vp]),
    "irr_topk_hits": (_i32, [_vp, _i64, _i32, _vp, _vp, _i64, _i64, _vp, _vp]),
    "irr_topk_class_dedup": (_i32, [_vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp,
                                    _vp]),
    "irr_pair_cosine": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _f32, _vp, _vp]),
    "irr_triplet_loss_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "irr_triplet_loss_fwd_bwd": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _f32, _f32, _i32, _f32,
                                        _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_f32), _vp, _sz,
                                        _vp]),
    "irr_triplet_loss_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _f32, _i32,
                                    _vp, _vp, _vp, _vp]),
    "irr_pair_loss_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "irr_pair_loss_fwd_bwd": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _f32, _i32, _vp,
                                     _vp, _vp, _vp, _f32, _vp, _sz, _vp]),
    "irr_pair_loss_bwd": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _f32, _i32,
                                 _vp, _vp, _vp]),
    "irr_avgpool_fwd": (_i32, [_vp, _i32, _i64, _i32, _vp, _i32, _vp]),
    "irr_avgpool_bwd": (_i32, [_vp, _i32, _i64, _i32, _vp, _i32, _vp]),
    "irr_ce_pair_workspace_bytes": (_sz, [_i64]),
    "irr_ce_pair_fwd": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i64, _vp, _vp, _sz, _vp]),
    "irr_ce_pair_bwd": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m imageretrievalresearch_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = symbol missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class IrrError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        msg = load().irr_status_string(status).decode()
        super().__init__(f"{where}: irr_status {status} ({msg})")


def check(status: int, where: str) -> None:
    if status != 0:
        raise IrrError(status, where)
