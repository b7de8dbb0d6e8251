"""ctypes binding of libb200vs.so (the C-ABI declared in include/b200vs.h).

There is no CPU fallback: if the shared library is missing this module raises, and every
compute entry point fails with RuntimeError when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path
from typing import List

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libb200vs.so"
HEADER = PKG_DIR.parent.parent / "include" / "b200vs.h"

VS_OK, VS_ERR_INVALID, VS_ERR_CUDA, VS_ERR_OOM, VS_ERR_STATE = 0, 1, 2, 3, 4
METRIC_COSINE, METRIC_EUCLIDEAN, METRIC_DOT = 0, 1, 2
METRICS = {"cosine": METRIC_COSINE, "euclidean": METRIC_EUCLIDEAN, "dot_product": METRIC_DOT}
SHADOW_NONE, SHADOW_BF16, SHADOW_FP8 = 0, 1, 2
SEARCH_AUTO, SEARCH_SCAN_FP32, SEARCH_SCAN_BF16, SEARCH_GEMM, SEARCH_GEMM_NOCERT, SEARCH_GEMM_FP8 = 0, 1, 2, 3, 4, 5
SEARCH_TMA, SEARCH_LDG = 0x100, 0x200
SEARCH_MODES = {"auto": SEARCH_AUTO, "scan_fp32": SEARCH_SCAN_FP32, "scan_bf16": SEARCH_SCAN_BF16,
                "gemm": SEARCH_GEMM, "gemm_nocert": SEARCH_GEMM_NOCERT, "gemm_fp8": SEARCH_GEMM_FP8}

_lib = None


def declared_symbols() -> List[str]:
    """Every function include/b200vs.h declares with VS_API."""
    text = HEADER.read_text()
    return re.findall(r"VS_API\s+[\w\s\*]+?\b(vs_\w+)\s*\(", text)


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    import os
    path = Path(os.environ.get("B200VS_LIB") or LIB_PATH)     # diagnostic: A/B-test another build of the library
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python -m b200vs.build` "
            "(this engine has no CPU fallback)")
    L = C.CDLL(str(path))
    p, i32, i64, f32p, i32p, u32p = C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p
    sig = {
        "vs_last_error": (C.c_char_p, []),
        "vs_version": (C.c_char_p, []),
        "vs_launch_count": (i64, []),
        "vs_profile": (i32, [i32]),
        "vs_profile_read": (i32, [i32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
        "vs_create": (i32, [i32, i32, i32, i32, i64, C.POINTER(p)]),
        "vs_destroy": (i32, [p]),
        "vs_append": (i32, [p, f32p, i64, i32, p]),
        "vs_append_ids": (i32, [p, f32p, i64, i32, i64, p]),
        "vs_count": (i64, [p]),
        "vs_reset": (i32, [p]),
        "vs_memory_bytes": (i64, [p]),
        "vs_read_rows": (i32, [p, i64, i64, f32p, i32, p]),
        "vs_search": (i32, [p, f32p, i32, i32, i32, u32p, i64, f32p, i32p, p]),
        "vs_search_submit": (i32, [p, f32p, i32, i32, i32, u32p, i64, f32p, i32p, p, C.POINTER(p)]),
        "vs_search_complete": (i32, [p, p]),
        "vs_search_submit_on": (i32, [p, f32p, i32, i32, i32, u32p, i64, f32p, i32p, p, p, C.POINTER(p)]),
        "vs_exchange_result": (i32, [p, p, p, i64, p, p, i32, C.c_uint32, p, p, p, i32, i32, f32p, i32p, p, p]),
        "vs_search_host": (i32, [p, f32p, i32, i32, i32, u32p, i64, f32p, i32p]),
        "vs_fallback_count": (i64, [p]),
        "vs_retry_count": (i64, [p]),
        "vs_merge": (i32, [i32, i32, f32p, i32p, i32, i32, i32, i64, f32p, i32p, p]),
        "vs_rescore": (i32, [p, f32p, i32, i32p, i32, i32, f32p, i32p, p]),
        "vs_exchange_push": (i32, [i32, p, i64, p, p, i32, C.c_uint32, p, p]),
        "vs_exchange_wait": (i32, [i32, p, i32, C.c_uint32, p]),
        "vs_exchange_wait_merge": (i32, [i32, i32, p, i32, C.c_uint32, p, i64, i32, i32, f32p, i32p, p]),
        "vs_debug_gemm_scores": (i32, [p, f32p, i32, f32p, p]),
        "vs_normalize_rows": (i32, [i32, f32p, i64, i32, f32p, p]),
        "vs_score_matrix": (i32, [i32, i32, f32p, i32, f32p, i64, i32, f32p, p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int) -> None:
    """Map a C status code to the Python exception the reference's callers expect."""
    if rc == VS_OK:
        return
    msg = lib().vs_last_error().decode("utf-8", "replace")
    if rc == VS_ERR_INVALID:
        raise ValueError(msg)
    if rc == VS_ERR_OOM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
