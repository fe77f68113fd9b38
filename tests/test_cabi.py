"""The C-ABI library loads on a CPU-only box and exports exactly what include/irr_b200.h declares.
No compute is launched here; argument validation returns before anything touches a device."""
import ctypes as C
import re
from pathlib import Path

import pytest

from imageretrievalresearch_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "irr_b200.h").read_text()


def declared_symbols():
    names = re.findall(r"IRR_API\s+[\w\s\*]+?\b(irr_\w+)\s*\(", HEADER)
    assert len(names) >= 15
    return sorted(set(names))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in irr_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_header_cites_reference_for_each_entry_point():
    for needle in ("train/train_efficient_cos_con_ce_loss.py", "utils/contrastive_loss.py",
                   "inference/inference.py"):
        assert needle in HEADER


def test_no_torch_types_in_abi():
    # prose may mention torch; no declaration may
    assert "at::" not in HEADER and "#include <torch" not in HEADER and "c10::" not in HEADER
    sigs = re.findall(r"IRR_API[^;]+;", HEADER)
    assert len(sigs) >= 25
    for sig in sigs:
        assert "Tensor" not in sig and "std::" not in sig and "torch" not in sig


def test_version_and_status_strings():
    lib = _lib.load()
    assert lib.irr_version() >= 100
    assert lib.irr_status_string(0) == b"ok"
    for code in range(-7, 0):
        assert len(lib.irr_status_string(code)) > 3


def test_constants_match_header():
    assert int(re.search(r"#define IRR_MAX_K (\d+)", HEADER).group(1)) == _lib.IRR_MAX_K
    assert int(re.search(r"#define IRR_ROW_STATS (\d+)", HEADER).group(1)) == _lib.IRR_ROW_STATS


def test_workspace_queries():
    lib = _lib.load()
    small = lib.irr_cosine_topk_workspace_bytes(64, 10_000, 1536, 3, _lib.IRR_F32)
    big = lib.irr_cosine_topk_workspace_bytes(4096, 1_000_000, 1536, 3, _lib.IRR_BF16)
    assert 0 < small < big < 64 << 20
    assert lib.irr_cosine_topk_workspace_bytes(-1, 10, 8, 3, 0) == 0
    assert lib.irr_triplet_loss_workspace_bytes(4096, 1536, 0) >= 256
    assert lib.irr_pair_loss_workspace_bytes(64, 1536, 1) >= 256


def test_argument_validation_without_a_device():
    lib = _lib.load()
    buf = C.create_string_buffer(4096)
    p = C.addressof(buf)
    p16 = (p + 15) // 16 * 16
    # null pointers
    assert lib.irr_cosine_topk(None, None, None, 4, 4, 8, 1, 0, 1e-6, 0, None, None, None, 0, None) == -1
    # k too large
    assert lib.irr_cosine_topk(p16, p16, None, 4, 400, 8, 257, 1, 1e-6, 0, p16, p16, p16, 4096, None) == -5
    # D violates the 16-byte row contract (bf16 needs D % 8 == 0)
    assert lib.irr_cosine_topk(p16, p16, None, 4, 40, 12, 3, 1, 1e-6, 0, p16, p16, p16, 4096, None) == -3
    # misaligned base pointer
    assert lib.irr_row_inv_norms(p16 + 4, 4, 8, 0, 1e-6, p16, None) == -3
    # bad dtype
    assert lib.irr_row_inv_norms(p16, 4, 8, 7, 1e-6, p16, None) == -2
    # merge / hits / losses
    assert lib.irr_topk_merge(None, None, 2, 4, 3, None, None, None) == -1
    assert lib.irr_topk_hits(p16, 4, 3, p16, None, 0, 0, p16, None) == -1
    assert lib.irr_pair_loss_fwd_bwd(p16, p16, p16, 3, 4, 8, 0, 2, 0.3, 1, p16, None, None, None, 1.0,
                                     p16, 4096, None) == -1  # label_count not in {1, B}
    assert lib.irr_pair_loss_fwd_bwd(p16, p16, p16, 1, 4, 8, 0, 9, 0.3, 1, p16, None, None, None, 1.0,
                                     p16, 4096, None) == -1  # unknown loss kind
    assert lib.irr_triplet_loss_fwd_bwd(p16, p16, p16, 4, 8, 0, 0.3, 0.3, 1, 1e-6, p16, None, None,
                                        p16, None, None, None, p16, 4096, None) == -1  # partial grads
    # Q == 0 is a no-op
    assert lib.irr_cosine_topk(None, None, None, 0, 4, 8, 1, 0, 1e-6, 0, p16, p16, None, 0, None) == 0


def test_exchange_entry_points_validate_without_a_device():
    lib = _lib.load()
    buf = C.create_string_buffer(4096)
    p16 = (C.addressof(buf) + 15) // 16 * 16
    # layout: header + two halves of G slots [fp32 Q*k | pad16 | int64 Q*k]
    assert lib.irr_topk_exchange_bytes(8, 4096, 3) == 512 + 2 * 8 * (4096 * 3 * 12)
    assert lib.irr_topk_exchange_bytes(2, 1, 1) >= 512 + 2 * 2 * 32
    assert lib.irr_topk_exchange_bytes(17, 1, 1) == 0          # > IRR_MAX_PEERS
    assert lib.irr_topk_exchange_bytes(0, 1, 1) == 0
    assert int(re.search(r"#define IRR_MAX_PEERS (\d+)", HEADER).group(1)) == _lib.IRR_MAX_PEERS
    ptrs = (C.c_void_p * 2)(p16, p16)
    call = lib.irr_topk_exchange_merge
    assert call(p16, p16, None, 2, 0, 4, 3, 1 << 20, 0, p16, p16, None) == -1      # no pointer table
    assert call(p16, p16, ptrs, 2, 2, 4, 3, 1 << 20, 0, p16, p16, None) == -1      # rank out of range
    assert call(p16, p16, ptrs, 2, 0, 4, 3, 1 << 20, 7, p16, p16, None) == -1      # unknown mode
    assert call(p16, p16, ptrs, 2, 0, 4, 257, 1 << 20, 0, p16, p16, None) == -5    # k too large
    assert call(None, None, ptrs, 2, 0, 4, 3, 1 << 20, 0, p16, p16, None) == -1    # fused needs lists
    assert call(p16, p16, ptrs, 2, 0, 4, 3, 600, 0, p16, p16, None) == -4          # buffer too small
    bad = (C.c_void_p * 2)(p16, None)
    assert call(p16, p16, bad, 2, 0, 4, 3, 1 << 20, 0, p16, p16, None) == -1       # unmapped peer
    assert lib.irr_cosine_topk_sharded_workspace_bytes(64, 1000, 1536, 3, 1) > \
        lib.irr_cosine_topk_workspace_bytes(64, 1000, 1536, 3, 1)
    assert lib.irr_peer_export(None, None, None) == -1
    assert lib.irr_peer_import(None, None) == -1
    assert lib.irr_peer_close(None) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libirr_b200.so")
    with pytest.raises(ImportError, match="no CPU"):
        _lib.load()
