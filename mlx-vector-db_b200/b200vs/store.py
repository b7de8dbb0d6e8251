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
from array import array
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


class _MetadataIndex:
    """Exact-match index over the metadata list: (key, value) -> row ids, built incrementally at
    `add_vectors`, so that a filter becomes a few vectorised set operations instead of the
    reference's Python predicate over all N dicts per query
    (service/optimized_vector_store.py:159-167).  Semantics are the reference's
    `all(meta.get(key) == val ...)`: a row without the key has the value None.  Values that
    cannot be hashed (or NaN, which never equals itself) are left to the generic predicate."""

    def __init__(self):
        self._rows: Dict[Any, Dict[Any, array]] = {}
        self._generic_keys = set()          # keys that hold an unhashable value somewhere

    def extend(self, first_row: int, metadata: List[Dict]):
        for r, meta in enumerate(metadata, start=first_row):
            for key, val in meta.items():
                try:
                    if val != val:              # NaN
                        raise TypeError
                    self._rows.setdefault(key, {}).setdefault(val, array("q")).append(r)
                except TypeError:
                    self._generic_keys.add(key)

    def lookup(self, n_rows: int, key, val) -> Optional[np.ndarray]:
        """bool (n_rows,) hit vector for `meta.get(key) == val`, or None when this pair needs
        the generic predicate."""
        if key in self._generic_keys:
            return None
        try:
            if val is None or val != val:
                return None
            ids = self._rows.get(key, {}).get(val)
        except TypeError:
            return None
        hit = np.zeros(n_rows, dtype=np.bool_)
        if ids is not None and len(ids):
            idx = np.frombuffer(ids, dtype=np.int64)
            hit[idx[idx < n_rows]] = True
        return hit


class MLXVectorStore:
    """Flat (N, D) fp32 store resident in HBM; exact brute-force top-k search.

    Limits (the reference has none because it sorts all N scores): k <= 1024 per query
    (larger k raises ValueError), N < 2**31 rows."""

    def __init__(self, store_path: str, config: Optional[MLXVectorStoreConfig] = None):
        self.store_path = Path(store_path).expanduser()
        self.config = config or MLXVectorStoreConfig()
        self._lock = threading.RLock()
        self.store_path.mkdir(parents=True, exist_ok=True)
        self._metadata: List[Dict] = []
        self._index = _MetadataIndex()
        self._vector_count = 0
        self._version = 0                 # bumped by every add / clear: invalidates cached bitmaps
        self._mask_cache: Dict[Any, Tuple[int, Any, int]] = {}
        self._handle = C.c_void_p()
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
            on_device = _is_torch(vectors) and vectors.is_cuda
            if on_device:
                v = vectors.detach().to(torch.float32).contiguous()
                if v.ndim == 1:
                    v = v.reshape(1, -1)
                host_rows = None
            else:
                v = host_rows = _to_host_f32(vectors)
                if v.ndim == 1:
                    v = host_rows = v.reshape(1, -1)
            self._check_dim(v.shape)
            m = int(v.shape[0])
            metadata = list(metadata)
            if len(metadata) != m:
                # the reference extends its list with whatever it is given (:104) and then serves
                # wrong or missing metadata; keep one entry per row instead
                logger.warning("%d vectors but %d metadata entries: %s", m, len(metadata),
                               "padding with {}" if len(metadata) < m else "dropping the surplus")
                metadata = metadata[:m] + [{} for _ in range(m - len(metadata))]
            # metadata first: a concurrent (lock-free, like the reference's) query never sees a
            # row id without its metadata entry
            first = len(self._metadata)
            self._metadata.extend(metadata)
            try:
                if on_device:
                    stream = torch.cuda.current_stream(v.device).cuda_stream
                    _cabi.check(_cabi.lib().vs_append(self._handle, C.c_void_p(v.data_ptr()), m, 1,
                                                      C.c_void_p(stream)))
                else:
                    _cabi.check(_cabi.lib().vs_append(self._handle, host_rows.ctypes.data_as(C.c_void_p),
                                                      m, 0, None))
            except Exception:
                del self._metadata[first:]
                raise
            self._index.extend(first, metadata)
            self._vector_count = int(_cabi.lib().vs_count(self._handle))
            self._version += 1
            self._mask_cache.clear()
            if self.config.persist:
                if host_rows is None:
                    host_rows = self._read_rows(self._vector_count - m, m)
                self._persist_append(first, host_rows, metadata)
            return {"vectors_added": m, "total_vectors": self._vector_count}

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
                      row_mask=None, mask_live: int = -1) -> Tuple[np.ndarray, np.ndarray]:
        """(ids (B, k) int32, scores (B, k) fp32) as arrays; unused slots id -1.
        row_mask: int32 device tensor, bit i of word i/32 = row i takes part; mask_live = its
        popcount (lets AUTO use the tensor-core path for the filtered search)."""
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
            self._flags() if flags is None else flags, mask_ptr, int(mask_live),
            scores.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p)))
        return ids, scores

    # -- metadata filter -> device bitmap (reference :159-167: exact-match AND over keys) --
    def _filter_hits(self, filter_metadata: Dict, n_rows: int) -> np.ndarray:
        hit = None
        generic = {}
        for key, val in filter_metadata.items():
            h = self._index.lookup(n_rows, key, val)
            if h is None:
                generic[key] = val
            else:
                hit = h if hit is None else (hit & h)
        if generic:   # unhashable / None / NaN values: the reference's predicate, row by row
            g = np.fromiter((all(m.get(key) == val for key, val in generic.items())
                             for m in self._metadata[:n_rows]), dtype=np.bool_, count=n_rows)
            hit = g if hit is None else (hit & g)
        return hit

    def _row_mask(self, filter_metadata: Dict):
        """(device bitmap, rows set) for this filter; cached until the next add / clear.
        Caller holds the lock."""
        if torch is None:
            raise RuntimeError("metadata filters need torch for the device bitmap")
        try:
            ckey = tuple(sorted(filter_metadata.items(), key=lambda kv: repr(kv[0])))
            hash(ckey)
        except TypeError:
            ckey = None
        if ckey is not None:
            ent = self._mask_cache.get(ckey)
            if ent is not None and ent[0] == self._version:
                return ent[1], ent[2]
        n_rows = self._vector_count
        hit = self._filter_hits(filter_metadata, n_rows)
        n_live = int(hit.sum())
        mask = None
        if n_live:
            bits = np.packbits(hit, bitorder="little")
            # whole 32-bit words, plus one spare word: kernels read the word of their 32-row chunk
            words = (n_rows + 31) // 32 + 1
            buf = np.zeros(words * 4, np.uint8)
            buf[:bits.size] = bits
            dev = torch.device("cuda", int(self.config.device))
            mask = torch.from_numpy(buf.view(np.int32)).to(dev)
            torch.cuda.current_stream(dev).synchronize()     # vs_search_host runs on the store's own stream
        if ckey is not None:
            if len(self._mask_cache) >= 16:
                self._mask_cache.clear()
            self._mask_cache[ckey] = (self._version, mask, n_live)
        return mask, n_live

    def _search_host(self, q: np.ndarray, k: int, filter_metadata: Optional[Dict]):
        B = q.shape[0]
        k = int(k)
        if k == 0:
            return [([], [], []) for _ in range(B)]
        if filter_metadata:
            # The bitmap is built for the rows present now; the lock keeps an append from
            # growing the store between building it and the search that indexes it.
            with self._lock:
                row_mask, n_live = self._row_mask(filter_metadata)
                if n_live == 0:
                    return [([], [], []) for _ in range(B)]
                kk = min(k, n_live) if k > 0 else max(0, n_live + k)
                if kk == 0:
                    return [([], [], []) for _ in range(B)]
                ids, scores = self.search_arrays(q, kk, row_mask=row_mask, mask_live=n_live)
        else:
            n_live = self._vector_count
            # the reference slices `argsort(...)[:k]` (:178,:181): k > N returns N results, a negative
            # k drops the last |k| of the rows
            kk = min(k, n_live) if k > 0 else max(0, n_live + k)
            if kk == 0:
                return [([], [], []) for _ in range(B)]
            ids, scores = self.search_arrays(q, kk)
        meta = self._metadata
        out = []
        for b in range(B):
            idx = ids[b].tolist()
            out.append((idx, scores[b].tolist(), [meta[i] for i in idx]))
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
                self._metadata, self._vector_count = [], 0
                self._index = _MetadataIndex()
                self._version += 1
                self._mask_cache.clear()
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
    # row.  Appends additionally write `segments/seg_<first row>_<rows>.npy` so an add costs
    # O(m), not O(N); `optimize()` folds the segments back into `vectors.npz`.  Every file is
    # written to a temporary name and renamed; a segment names the rows it holds, so after a
    # crash between the snapshot's rename and the removal of the segments it covers, loading
    # skips those segments instead of ingesting their rows twice.
    def _read_rows(self, first: int, m: int) -> np.ndarray:
        out = np.empty((m, self.config.dimension), np.float32)
        if m:
            _cabi.check(_cabi.lib().vs_read_rows(self._handle, first, m,
                                                 out.ctypes.data_as(C.c_void_p), 0, None))
        return out

    def _persist_append(self, first: int, rows: np.ndarray, metadata: List[Dict]):
        seg_dir = self.store_path / "segments"
        seg_dir.mkdir(exist_ok=True)
        tmp = seg_dir / f"tmp_{first:012d}.npy"
        np.save(tmp, rows)
        tmp.replace(seg_dir / f"seg_{first:012d}_{rows.shape[0]}.npy")
        with open(self.store_path / "metadata.jsonl", "a") as f:
            f.write("".join(json.dumps(m) + "\n" for m in metadata))

    def _write_snapshot(self, rows: np.ndarray):
        mtmp = self.store_path / "metadata.tmp.jsonl"
        with open(mtmp, "w") as f:
            for m in self._metadata[:rows.shape[0]]:
                f.write(json.dumps(m) + "\n")
        tmp = self.store_path / "vectors.tmp.npz"
        np.savez(str(tmp), vectors=rows)
        tmp.replace(self.store_path / "vectors.npz")
        mtmp.replace(self.store_path / "metadata.jsonl")
        seg_dir = self.store_path / "segments"
        if seg_dir.exists():
            shutil.rmtree(seg_dir)

    def _save_store(self):
        self.optimize()

    def _load_store(self):
        try:
            parts = []
            covered = 0
            vp = self.store_path / "vectors.npz"
            if vp.exists():
                snap = np.load(str(vp))["vectors"].astype(np.float32, copy=False)
                parts.append(snap)
                covered = snap.shape[0]
            seg_dir = self.store_path / "segments"
            for seg in (sorted(seg_dir.glob("seg_*.npy")) if seg_dir.exists() else []):
                first, m = (int(x) for x in seg.stem.split("_")[1:3])
                if first + m <= covered:
                    continue                 # already inside the snapshot (interrupted optimize())
                if first != covered:
                    raise ValueError(f"segment {seg.name} does not continue the store at row {covered}")
                parts.append(np.load(str(seg)))
                covered += m
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
            count = int(_cabi.lib().vs_count(self._handle))
            if len(meta) != count:
                # e.g. a crash while metadata.jsonl was being appended to: keep the rows usable
                logger.error("%s: %d vectors but %d metadata entries; %s", self.store_path, count, len(meta),
                             "padding with {}" if len(meta) < count else "dropping the surplus")
                meta = meta[:count] + [{} for _ in range(count - len(meta))]
            self._metadata = meta
            self._index = _MetadataIndex()
            self._index.extend(0, meta)
            self._vector_count = count
            self._version += 1
        except Exception as e:  # reference :237-239: log and start empty
            logger.error("loading %s failed, starting empty: %s", self.store_path, e)
            _cabi.lib().vs_reset(self._handle)
            self._metadata, self._vector_count = [], 0
            self._index = _MetadataIndex()


def create_optimized_vector_store(store_path: str, dimension: int = 384, jit_compile: bool = True,
                                  enable_hnsw: bool = False, **kwargs) -> MLXVectorStore:
    """service/optimized_vector_store.py:244-246."""
    config = MLXVectorStoreConfig(dimension=dimension, jit_compile=jit_compile,
                                  enable_hnsw=enable_hnsw, **kwargs)
    return MLXVectorStore(store_path, config)
