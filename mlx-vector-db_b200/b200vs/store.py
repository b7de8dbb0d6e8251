"""B200-native drop-in for the reference's flat vector store.

Mirrors `MLXVectorStore`, `MLXVectorStoreConfig` and `create_optimized_vector_store` of the
reference (service/optimized_vector_store.py:51-246): same method names, positional order,
defaults, return shapes and error behaviour, with the arithmetic done by hand-written
sm_100a kernels behind the C-ABI of include/b200vs.h.  Methods the reference's callers use
but the reference never defined (`batch_query`, `optimize`, `health_check`,
`get_stats()['memory_usage_mb']`) are defined here from the callers'
expectations.  No CPU fallback: every search runs on the GPU or raises.
"""
from __future__ import annotations

import ctypes as C
import json
import logging
import shutil
import threading
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import _cabi

logger = logging.getLogger("b200vs.store")

try:  # torch is plumbing (device tensors, streams); numpy inputs do not need it
    import torch
except Exception:  # pragma: no cover
    torch = None


@dataclass
class MLXVectorStoreConfig:
    """service/optimized_vector_store.py:51-56, plus engine knobs with neutral defaults."""
    dimension: int = 384
    metric: str = "cosine"
    enable_hnsw: bool = False     # accepted for compatibility; this engine is exact-only
    jit_compile: bool = True
    # --- engine extensions (not in the reference) ---
    device: int = 0
    shadow_bf16: bool = True      # keep a 16-bit copy for the tensor-core candidate kernels
    shadow_fp8: bool = False      # additionally keep an e4m3 copy (cosine; search_mode "gemm_fp8")
    max_vectors: int = 0          # address-space reservation; 0 = derive from device memory
    search_mode: str = "auto"     # auto | scan_fp32 | scan_bf16 | gemm | gemm_nocert
    persist: bool = True          # write appended rows to disk on every add (reference does)


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def _to_host_f32(x) -> np.ndarray:
    """`mx.array(vectors, dtype=mx.float32)` (service/optimized_vector_store.py:215-216)."""
    if _is_torch(x):
        x = x.detach().to("cpu").numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


class MLXVectorStore:
    """Flat (N, D) fp32 store resident in HBM; exact brute-force top-k search."""

    def __init__(self, store_path: str, config: Optional[MLXVectorStoreConfig] = None):
        self.store_path = Path(store_path).expanduser()
        self.config = config or MLXVectorStoreConfig()
        self._lock = threading.RLock()
        self.store_path.mkdir(parents=True, exist_ok=True)
        self._metadata: List[Dict] = []
        self._vector_count = 0
        self._handle = C.c_void_p()
        self._segments = 0
        if self.config.enable_hnsw:
            logger.warning("enable_hnsw=True ignored: this engine serves exact search only")
        # reference :211-213 -- only cosine / euclidean get a score function, and only with
        # jit_compile; dot_product is an extension here (service/models.py:23-27 advertises it)
        self._metric_id = _cabi.METRICS.get(self.config.metric) if self.config.jit_compile else None
        self._create_handle()
        self._load_store()
        logger.info("B200 store initialised: %s", self.store_path)

    # ------------------------------------------------------------------ native handle
    def _create_handle(self):
        metric = self._metric_id if self._metric_id is not None else _cabi.METRIC_COSINE
        _cabi.check(_cabi.lib().vs_create(
            int(self.config.device), int(self.config.dimension), metric,
            (_cabi.SHADOW_BF16 if self.config.shadow_bf16 else 0) |
            (_cabi.SHADOW_FP8 if self.config.shadow_fp8 else 0),
            int(self.config.max_vectors), C.byref(self._handle)))

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            _cabi.lib().vs_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _flags(self) -> int:
        return _cabi.SEARCH_MODES[self.config.search_mode]

    # ------------------------------------------------------------------ add
    def add_vectors(self, vectors, metadata: List[Dict]):
        """service/optimized_vector_store.py:96-114.  Appends in place (K1); no O(N) copy."""
        with self._lock:
            if _is_torch(vectors) and vectors.is_cuda:
                v = vectors.detach().to(torch.float32).contiguous()
                if v.ndim == 1:
                    v = v.reshape(1, -1)
                self._check_dim(v.shape)
                stream = torch.cuda.current_stream(v.device).cuda_stream
                _cabi.check(_cabi.lib().vs_append(self._handle, C.c_void_p(v.data_ptr()),
                                                  v.shape[0], 1, C.c_void_p(stream)))
                host_rows = None
                m = v.shape[0]
            else:
                host_rows = _to_host_f32(vectors)
                if host_rows.ndim == 1:
                    host_rows = host_rows.reshape(1, -1)
                self._check_dim(host_rows.shape)
                m = host_rows.shape[0]
                _cabi.check(_cabi.lib().vs_append(self._handle, host_rows.ctypes.data_as(C.c_void_p),
                                                  m, 0, None))
            self._metadata.extend(metadata)
            self._vector_count = int(_cabi.lib().vs_count(self._handle))
            if self.config.persist:
                if host_rows is None:
                    host_rows = self._read_rows(self._vector_count - m, m)
                self._persist_append(host_rows, list(metadata))
            return {"vectors_added": len(metadata), "total_vectors": self._vector_count}

    def _check_dim(self, shape):
        if len(shape) != 2 or shape[1] != self.config.dimension:
            raise ValueError(f"vectors must have shape (m, {self.config.dimension}), got {tuple(shape)}")

    # ------------------------------------------------------------------ query
    def query(self, query_vector, k: int = 10, filter_metadata: Optional[Dict] = None,
              use_hnsw: bool = True) -> Tuple:
        """service/optimized_vector_store.py:116-192 -> (indices, scores, metadata), best first."""
        if self._vector_count == 0:
            return [], [], []
        if self._metric_id is None:
            raise RuntimeError("no compiled similarity function available")  # reference :153-154
        q = _to_host_f32(query_vector).reshape(-1)
        if q.shape[0] != self.config.dimension:
            raise ValueError(f"query must have {self.config.dimension} components, got {q.shape[0]}")
        res = self._search_host(q.reshape(1, -1), k, filter_metadata)
        return res[0] if res else ([], [], [])

    def batch_query(self, queries, k: int = 10, filter_metadata: Optional[Dict] = None):
        """Missing in the reference (api/routes/vectors.py:291, tests/demo.py:134 call it).
        Returns a list of per-query `(indices, scores, metadata)` tuples in `query()`'s raw
        convention (cosine -> similarity, descending), computed in one batched GPU pass."""
        q = _to_host_f32(queries)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.ndim != 2 or q.shape[1] != self.config.dimension:
            raise ValueError(f"queries must have shape (B, {self.config.dimension}), got {q.shape}")
        if self._vector_count == 0:
            return [([], [], []) for _ in range(q.shape[0])]
        if self._metric_id is None:
            raise RuntimeError("no compiled similarity function available")
        return self._search_host(q, k, filter_metadata)

    def search_arrays(self, queries: np.ndarray, k: int = 10, flags: Optional[int] = None,
                      row_mask=None) -> Tuple[np.ndarray, np.ndarray]:
        """(ids (B, k) int32, scores (B, k) fp32) as arrays; unused slots id -1."""
        q = _to_host_f32(queries)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        B = q.shape[0]
        k = int(k)
        ids = np.full((B, max(k, 0)), -1, dtype=np.int32)
        scores = np.zeros((B, max(k, 0)), dtype=np.float32)
        if B == 0 or k <= 0:
            return ids, scores
        mask_ptr = None
        if row_mask is not None:
            mask_ptr = C.c_void_p(row_mask.data_ptr())
        _cabi.check(_cabi.lib().vs_search_host(
            self._handle, q.ctypes.data_as(C.c_void_p), B, k,
            self._flags() if flags is None else flags, mask_ptr,
            scores.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p)))
        return ids, scores

    def _search_host(self, q: np.ndarray, k: int, filter_metadata: Optional[Dict]):
        B = q.shape[0]
        k = int(k)
        if k == 0:
            return [([], [], []) for _ in range(B)]
        row_mask = None
        n_live = self._vector_count
        if filter_metadata:
            # reference :159-167: exact-match AND over keys; here the predicate becomes a
            # device bitmap consumed by the scan instead of a row gather.
            hit = np.fromiter((all(m.get(key) == val for key, val in filter_metadata.items())
                               for m in self._metadata), dtype=np.bool_, count=len(self._metadata))
            n_live = int(hit.sum())
            if n_live == 0:
                return [([], [], []) for _ in range(B)]
            if torch is None:
                raise RuntimeError("metadata filters need torch for the device bitmap")
            bits = np.packbits(hit, bitorder="little")
            pad = (-bits.size) % 4
            if pad:
                bits = np.concatenate([bits, np.zeros(pad, np.uint8)])
            row_mask = torch.from_numpy(bits.view(np.int32).copy()).to(f"cuda:{self.config.device}")
            torch.cuda.current_stream(row_mask.device).synchronize()
        # the reference slices `argsort(...)[:k]` (:178,:181): k > N returns N results, a negative k
        # drops the last |k| of the (filtered) rows
        kk = min(k, n_live) if k > 0 else max(0, n_live + k)
        if kk == 0:
            return [([], [], []) for _ in range(B)]
        ids, scores = self.search_arrays(q, kk, row_mask=row_mask)
        out = []
        for b in range(B):
            idx = ids[b].tolist()
            out.append((idx, scores[b].tolist(), [self._metadata[i] for i in idx]))
        return out

    # ------------------------------------------------------------------ misc surface
    def clear(self):
        """service/optimized_vector_store.py:198-209."""
        with self._lock:
            try:
                if self.store_path.exists():
                    shutil.rmtree(self.store_path)
                self.store_path.mkdir(parents=True, exist_ok=True)
                _cabi.check(_cabi.lib().vs_reset(self._handle))
                self._metadata, self._vector_count, self._segments = [], 0, 0
            except Exception as e:  # reference logs and carries on
                logger.error("clear failed for %s: %s", self.store_path, e)

    def get_stats(self) -> Dict[str, Any]:
        """service/optimized_vector_store.py:241-242 + the `memory_usage_mb` key its callers
        read (api/routes/vectors.py:131, api/routes/monitoring.py:153)."""
        return {"vector_count": self._vector_count, "dimension": self.config.dimension,
                "metric": self.config.metric, "index_type": "flat",
                "memory_usage_mb": _cabi.lib().vs_memory_bytes(self._handle) / 2**20}

    def optimize(self):
        """Called by api/routes/vectors.py:425 and admin.py:230; compacts the on-disk segment
        log into the reference's single `vectors.npz` (key `vectors`)."""
        with self._lock:
            if self._vector_count:
                self._write_snapshot(self._read_rows(0, self._vector_count))

    def health_check(self) -> Dict[str, Any]:
        """tests/demo.py:254 expects {'healthy': bool, 'issues': list}."""
        issues = []
        n = int(_cabi.lib().vs_count(self._handle))
        if n != len(self._metadata):
            issues.append(f"{n} vectors but {len(self._metadata)} metadata entries")
        if self._metric_id is None:
            issues.append("no similarity function (jit_compile=False or unknown metric)")
        return {"healthy": not issues, "issues": issues}

    # ------------------------------------------------------------------ persistence
    # On-disk format is the reference's (service/optimized_vector_store.py:218-239):
    # `vectors.npz` with key `vectors` (N, D) fp32 and `metadata.jsonl`, one JSON object per
    # row.  Appends additionally write `segments/seg_XXXXXX.npy` so an add costs O(m), not
    # O(N); `optimize()` folds the segments back into `vectors.npz`.
    def _read_rows(self, first: int, m: int) -> np.ndarray:
        out = np.empty((m, self.config.dimension), np.float32)
        if m:
            _cabi.check(_cabi.lib().vs_read_rows(self._handle, first, m,
                                                 out.ctypes.data_as(C.c_void_p), 0, None))
        return out

    def _persist_append(self, rows: np.ndarray, metadata: List[Dict]):
        seg_dir = self.store_path / "segments"
        seg_dir.mkdir(exist_ok=True)
        np.save(seg_dir / f"seg_{self._segments:06d}.npy", rows)
        self._segments += 1
        with open(self.store_path / "metadata.jsonl", "a") as f:
            for m in metadata:
                f.write(json.dumps(m) + "\n")

    def _write_snapshot(self, rows: np.ndarray):
        tmp = self.store_path / "vectors.tmp.npz"
        np.savez(str(tmp), vectors=rows)
        tmp.replace(self.store_path / "vectors.npz")
        seg_dir = self.store_path / "segments"
        if seg_dir.exists():
            shutil.rmtree(seg_dir)
        self._segments = 0
        with open(self.store_path / "metadata.jsonl", "w") as f:
            for m in self._metadata:
                f.write(json.dumps(m) + "\n")

    def _save_store(self):
        self.optimize()

    def _load_store(self):
        try:
            parts = []
            vp = self.store_path / "vectors.npz"
            if vp.exists():
                parts.append(np.load(str(vp))["vectors"].astype(np.float32, copy=False))
            seg_dir = self.store_path / "segments"
            segs = sorted(seg_dir.glob("seg_*.npy")) if seg_dir.exists() else []
            parts.extend(np.load(str(s)) for s in segs)
            self._segments = len(segs)
            meta = []
            mp = self.store_path / "metadata.jsonl"
            if mp.exists():
                with open(mp) as f:
                    meta = [json.loads(line) for line in f if line.strip()]
            for p in parts:
                p = np.ascontiguousarray(p, dtype=np.float32)
                if p.size:
                    self._check_dim(p.shape)
                    _cabi.check(_cabi.lib().vs_append(self._handle, p.ctypes.data_as(C.c_void_p),
                                                      p.shape[0], 0, None))
            self._metadata = meta
            self._vector_count = int(_cabi.lib().vs_count(self._handle))
        except Exception as e:  # reference :237-239: log and start empty
            logger.error("loading %s failed, starting empty: %s", self.store_path, e)
            _cabi.lib().vs_reset(self._handle)
            self._metadata, self._vector_count, self._segments = [], 0, 0


def create_optimized_vector_store(store_path: str, dimension: int = 384, jit_compile: bool = True,
                                  enable_hnsw: bool = False, **kwargs) -> MLXVectorStore:
    """service/optimized_vector_store.py:244-246."""
    config = MLXVectorStoreConfig(dimension=dimension, jit_compile=jit_compile,
                                  enable_hnsw=enable_hnsw, **kwargs)
    return MLXVectorStore(store_path, config)
