"""Build libb200vs.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m b200vs.build            # from mlx-vector-db_b200/
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
ROOT = PKG_DIR.parent                      # mlx-vector-db_b200/
CSRC = ROOT / "csrc"
BUILD = ROOT / "build"
LIB = PKG_DIR / "libb200vs.so"
SOURCES = ["gemm_inst_plain.cu", "gemm_inst_general.cu", "store.cu", "scan_topk.cu", "merge.cu", "search.cu", "gemm_topk.cu", "exchange.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v",
]
# tuning knobs (diagnostic): epilogue warp groups of K3 (4 warps each), STREAMING / RESIDENT variants
if os.environ.get("B200VS_EPI_GROUPS"):
    NVCC_FLAGS.append("-DVS_EPI_GROUPS=" + os.environ["B200VS_EPI_GROUPS"])
if os.environ.get("B200VS_EPI_GROUPS_RES"):
    NVCC_FLAGS.append("-DVS_EPI_GROUPS_RES=" + os.environ["B200VS_EPI_GROUPS_RES"])
# diagnostic: B200VS_DEBUG_BUILD=1 adds K3's timing-experiment epilogues (B200VS_GEMM_DBGMODE)
if os.environ.get("B200VS_RES_TN"):          # diagnostic: K3 RESIDENT tile width (128 | 256)
    NVCC_FLAGS.append("-DVS_RES_TN=" + os.environ["B200VS_RES_TN"])
if os.environ.get("B200VS_RES_ISSUERS"):     # diagnostic: MMA-issuing warps of K3 RESIDENT (1 | 2)
    NVCC_FLAGS.append("-DVS_RES_ISSUERS=" + os.environ["B200VS_RES_ISSUERS"])
if os.environ.get("B200VS_DEBUG_BUILD") == "1":
    NVCC_FLAGS.append("-DVS_GEMM_DEBUG_MODES")


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: libb200vs cannot be built (no CPU fallback exists)")
    return exe


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) +
                    [ROOT.parent / "include" / "b200vs.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())      # diagnostic -D knobs change the binary too
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile (if the sources changed since the last build) and return the library path.
    Safe to call from several processes at once (one rank per GPU under torchrun): an exclusive
    file lock serialises the builders, later ones find the stamp up to date, and the library is
    moved into place atomically so a reader never sees a half-written file."""
    import fcntl
    # B200VS_SKIP_BUILD=1: use the library as shipped (GPU boxes get the .so built here with the
    # snapshot; a source edit racing the snapshot must not trigger a rebuild there)
    if os.environ.get("B200VS_SKIP_BUILD") == "1" and LIB.exists() and not force:
        return LIB
    BUILD.mkdir(exist_ok=True)
    with open(BUILD / "lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    stamp = BUILD / "stamp.txt"
    dig = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    nvcc = _nvcc()

    def compile_one(src: str):
        obj = BUILD / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD / (Path(src).stem + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr[-4000:]}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB.with_suffix(f".so.tmp{os.getpid()}")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp),
           *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    os.replace(tmp, LIB)
    stamp.write_text(dig)
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True)
