"""Build recipe of libirr_b200.so (the C-ABI shared library, include/irr_b200.h).

Plain nvcc, in-tree: objects under build/, the library next to this file so that it travels to
the GPU box with the repo snapshot.  sm_100a only; -lineinfo so ncu's source page maps to csrc/.

    python -m imageretrievalresearch_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = REPO_ROOT / "build" / "irr_b200"
LIB_PATH = PKG_DIR / "libirr_b200.so"

SOURCES = [
    "irr_cabi.cu",
    "row_norms.cu",
    "topk_merge.cu",
    "topk_exchange.cu",
    "topk_select.cu",
    "cosine_topk_f32.cu",
    "cosine_topk_bf16.cu",
    "triplet_loss.cu",
    "producer_consumer.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libirr_b200.so cannot be built (there is no CPU fallback)")


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sorted(CSRC.glob("*")) + [REPO_ROOT / "include" / "irr_b200.h", Path(__file__)]
    for f in files:
        if f.is_file():
            h.update(f.name.encode())
            h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = BUILD_DIR / "fingerprint"
    return LIB_PATH.exists() and stamp.exists() and stamp.read_text() == _fingerprint()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link libirr_b200.so.  Idempotent."""
    if not force and is_current():
        return LIB_PATH
    nvcc = _nvcc()
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    logs: dict[str, str] = {}

    def compile_one(src: str) -> Path:
        obj = BUILD_DIR / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src] = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
            "-o", str(LIB_PATH), *map(str, objs)]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    (BUILD_DIR / "ptxas.log").write_text("\n".join(f"==== {k}\n{v}" for k, v in logs.items()))
    (BUILD_DIR / "fingerprint").write_text(_fingerprint())
    if verbose:
        print((BUILD_DIR / "ptxas.log").read_text())
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
