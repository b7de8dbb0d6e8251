"""CPU oracle for the exact vector-search hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in NumPy fp32, the arithmetic the reference
(Theseus-AT/mlx-vector-db) executes on its `matmul -> argsort[:k]` path.  It is
the checker for the CUDA engine: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product
path (`mlx-vector-db_b200/`) never imports it and has no CPU fallback.

PARITY PIN: the reference's arithmetic lives in the un-vendored third-party library `mlx`
(pinned `mlx>=0.25.2`, reference `requirements.txt:7`), which is not importable in this image,
and the reference ships no golden vectors, seeds or known-answer files for this path
(SURVEY.md section 8c).  What pins this restatement instead:
  * `tests/golden/ref_*.npz|json` -- outputs of the reference's OWN modules
    (`service/optimized_vector_store.py`, `performance/mlx_optimized.py`, imported unmodified
    from /root/reference by `tests/golden/make_reference_golden.py`) executed over a NumPy
    stand-in for the `mlx.core` primitives (`tests/golden/mlx_standin/`).  The restatement
    agrees with them BIT FOR BIT (`tests/test_reference_golden.py`): op order, clamps, slicing,
    id mapping, filter semantics, degenerate cases and exception types are the reference's.
  * NOT pinned: the arithmetic inside the mlx binary itself (accumulation order of its matmul,
    the tie order of its argsort -- assumed stable).  BASELINE.json's tolerance
    (scores 1e-5, ids exact outside 1e-6 ties) is orders of magnitude wider than such effects.
  * an independent float64 restatement and the behavioural assertions of the reference's own
    tests (`tests/test_integration.py:110,133-136,158-160`, `tests/demo.py:232,238,243`),
    `tests/test_oracle.py`.

All paths below are relative to the reference root.
"""
from __future__ import annotations

import json
import shutil
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

EPS = np.float32(1e-8)  # service/optimized_vector_store.py:36, performance/mlx_optimized.py:45


# --------------------------------------------------------------------------- #
# conversions
# --------------------------------------------------------------------------- #
def to_f32(x) -> np.ndarray:
    """`mx.array(vectors, dtype=mx.float32)` -- service/optimized_vector_store.py:215-216."""
    if hasattr(x, "detach"):  # torch tensor stand-in for mx.array
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def _row_norms(v: np.ndarray) -> np.ndarray:
    """`sqrt(sum(square(v), axis=1, keepdims=True))` -- performance/mlx_optimized.py:41-42
    (the store spells it `mx.linalg.norm(axis=1, keepdims=True)`,
    service/optimized_vector_store.py:34-35: same maths)."""
    return np.sqrt(np.sum(np.square(v), axis=1, keepdims=True, dtype=np.float32))


# --------------------------------------------------------------------------- #
# scoring functions
# --------------------------------------------------------------------------- #
def normalize_vectors(vectors) -> np.ndarray:
    """performance/mlx_optimized.py:110-125 -- row normalise with the 1e-8 clamp."""
    v = to_f32(vectors)
    if v.ndim != 2:
        raise ValueError(f"vectors must be 2-D for normalisation, got shape {v.shape}")
    if v.shape[0] == 0:
        return v.copy()
    n = np.maximum(_row_norms(v), EPS)
    return v / n


def cosine_similarity_single(query, db) -> np.ndarray:
    """service/optimized_vector_store.py:31-41 == performance/mlx_optimized.py:26-57.

    Op order: norms of both operands, clamp at 1e-8, divide *each operand* by its
    norm, then one fp32 matmul of the normalised operands, flattened to (N,)."""
    q = to_f32(query)
    v = to_f32(db)
    if q.ndim == 1:
        q = q.reshape(1, -1)
    elif not (q.ndim == 2 and q.shape[0] == 1):
        raise ValueError(f"query_vector must be 1-D or 2-D with one row, got shape {q.shape}")
    qn = q / np.maximum(_row_norms(q), EPS)
    vn = v / np.maximum(_row_norms(v), EPS)
    return (vn @ qn.T).reshape(-1)


def cosine_similarity_batch(queries, db) -> np.ndarray:
    """performance/mlx_optimized.py:59-88 -> (B, N) fp32, materialised."""
    q = to_f32(queries)
    v = to_f32(db)
    if q.ndim != 2:
        raise ValueError(f"query_vectors must be 2-D, got shape {q.shape}")
    if v.ndim != 2:
        raise ValueError(f"db_vectors must be 2-D, got shape {v.shape}")
    if q.shape[1] != v.shape[1]:
        raise ValueError(f"dimension mismatch: queries {q.shape[1]}, db {v.shape[1]}")
    qn = q / np.maximum(_row_norms(q), EPS)
    vn = v / np.maximum(_row_norms(v), EPS)
    return qn @ vn.T


def euclidean_distance(query, db) -> np.ndarray:
    """service/optimized_vector_store.py:43-48 == performance/mlx_optimized.py:139-148.
    Direct-difference form, true (not squared) distance."""
    q = to_f32(query)
    v = to_f32(db)
    if q.ndim == 1:
        q = q.reshape(1, -1)
    diff = v - q
    return np.sqrt(np.sum(diff * diff, axis=1, dtype=np.float32))


def dot_product(query, db) -> np.ndarray:
    """performance/mlx_optimized.py:150-156."""
    q = to_f32(query)
    v = to_f32(db)
    if q.ndim == 1:
        return v @ q
    return (v @ q.T).reshape(-1)


# --------------------------------------------------------------------------- #
# top-k
# --------------------------------------------------------------------------- #
def top_k_indices(scores, k: int) -> np.ndarray:
    """performance/mlx_optimized.py:90-108: `argsort(-scores)[:min(k, N)]`.

    Tie rule: stable sort -> equal keys keep ascending row order (assumption about
    mlx's CPU argsort)."""
    s = np.asarray(scores)
    if s.ndim != 1:
        raise ValueError("scores must be a 1-D array")
    if k <= 0:
        return np.zeros((0,), dtype=np.int32)
    kk = min(int(k), s.shape[0])
    if kk == 0:
        return np.zeros((0,), dtype=np.int32)
    return np.argsort(-s, kind="stable")[:kk].astype(np.int32)


def similarity_search(query, db, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
    """performance/mlx_optimized.py:199-215 -> (idx (k,), scores (k,))."""
    q = to_f32(query)
    if q.ndim == 2 and q.shape[0] == 1:
        q = q.reshape(-1)
    elif q.ndim != 1:
        raise ValueError(f"query_vector must be 1-D or (1, D), got shape {q.shape}")
    s = cosine_similarity_single(q, db)
    idx = top_k_indices(s, k)
    return idx, s[idx]


def batch_similarity_search(queries, db, k: int = 10, chunk: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """performance/mlx_optimized.py:217-248 -> (idx (B,k) , scores (B,k)).

    `chunk` > 0 processes that many queries at a time so the (B, N) score,
    negation and index temporaries fit in host RAM (scores can differ from the unchunked
    call in the last ulp: BLAS picks its sgemm kernel by shape)."""
    q = to_f32(queries)
    v = to_f32(db)
    if q.ndim != 2 or v.ndim != 2:
        raise ValueError("queries and db must be 2-D")
    B = q.shape[0]
    if v.shape[0] == 0:
        return np.zeros((B, 0), np.int32), np.zeros((B, 0), np.float32)
    kk = min(int(k), v.shape[0])
    if kk <= 0:
        return np.zeros((B, 0), np.int32), np.zeros((B, 0), np.float32)
    step = chunk if chunk and chunk > 0 else max(B, 1)
    out_i = np.empty((B, kk), np.int32)
    out_s = np.empty((B, kk), np.float32)
    for b0 in range(0, B, step):
        s = cosine_similarity_batch(q[b0:b0 + step], v)          # :221
        order = np.argsort(-s, axis=1, kind="stable")[:, :kk]    # :235-236
        out_i[b0:b0 + step] = order
        out_s[b0:b0 + step] = np.take_along_axis(s, order, axis=1)  # :239-244
    return out_i, out_s


def vector_concatenation(existing, new) -> np.ndarray:
    """performance/mlx_optimized.py:127-137."""
    a = to_f32(existing)
    b = to_f32(new)
    if a.shape[0] == 0:
        return b
    if b.shape[0] == 0:
        return a
    if a.shape[1] != b.shape[1]:
        raise ValueError("dimensions of the vectors to concatenate differ")
    return np.concatenate([a, b], axis=0)


def vector_addition(existing, new, normalize: bool = False) -> np.ndarray:
    """performance/mlx_optimized.py:250-255."""
    out = vector_concatenation(existing, new)
    return normalize_vectors(out) if normalize else out


# --------------------------------------------------------------------------- #
# generic scored search (all three metrics), used by the parity tests
# --------------------------------------------------------------------------- #
def score_matrix(queries, db, metric: str) -> np.ndarray:
    """(B, N) scores for `metric` in reference op order, one query row at a time for
    euclidean (the reference has no batched L2; it is the per-query function)."""
    q = to_f32(queries)
    if q.ndim == 1:
        q = q.reshape(1, -1)
    if metric == "cosine":
        return cosine_similarity_batch(q, db)
    if metric == "euclidean":
        return np.stack([euclidean_distance(qi, db) for qi in q], axis=0)
    if metric == "dot_product":
        return q @ to_f32(db).T
    raise ValueError(f"unknown metric {metric!r}")


def search(queries, db, k: int, metric: str = "cosine") -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Full-sort top-k per query.  Returns (ids (B,kk), scores (B,kk), score_matrix (B,N)).
    Order: cosine/dot descending, euclidean ascending; ties -> lower id first
    (service/optimized_vector_store.py:176-184)."""
    s = score_matrix(queries, db, metric)
    kk = max(0, min(int(k), s.shape[1]))
    key = s if metric == "euclidean" else -s
    order = np.argsort(key, axis=1, kind="stable")[:, :kk].astype(np.int32)
    return order, np.take_along_axis(s, order, axis=1), s


# --------------------------------------------------------------------------- #
# the store (service/optimized_vector_store.py:59-246)
# --------------------------------------------------------------------------- #
class OracleVectorStore:
    """Restatement of `MLXVectorStore` without HNSW (out of scope)."""

    def __init__(self, store_path: Optional[str] = None, dimension: int = 384,
                 metric: str = "cosine", jit_compile: bool = True):
        self.store_path = Path(store_path).expanduser() if store_path else None
        self.dimension = dimension
        self.metric = metric
        self._vectors: Optional[np.ndarray] = None
        self._metadata: List[Dict] = []
        self._vector_count = 0
        # :211-213 -- only cosine / euclidean get a score function, and only if jit_compile
        self._fn = None
        if jit_compile:
            if metric == "cosine":
                self._fn = cosine_similarity_single
            elif metric == "euclidean":
                self._fn = euclidean_distance
        if self.store_path is not None:
            self.store_path.mkdir(parents=True, exist_ok=True)
            self._load_store()

    # :96-114
    def add_vectors(self, vectors, metadata: Sequence[Dict]) -> Dict:
        new = to_f32(vectors)
        if self._vectors is None:
            self._vectors = new
        else:
            self._vectors = np.concatenate([self._vectors, new], axis=0)
        self._metadata.extend(metadata)
        self._vector_count = self._vectors.shape[0]
        self._save_store()
        return {"vectors_added": len(metadata), "total_vectors": self._vector_count}

    # :116-145 (HNSW branch dropped) + :149-192
    def query(self, query_vector, k: int = 10, filter_metadata: Optional[Dict] = None,
              use_hnsw: bool = True):
        if self._vector_count == 0:
            return [], [], []
        q = to_f32(query_vector)
        if self._fn is None:
            raise RuntimeError("no compiled similarity function available")
        target = self._vectors
        original = None
        if filter_metadata:
            original = [i for i, m in enumerate(self._metadata)
                        if all(m.get(key) == val for key, val in filter_metadata.items())]
            if not original:
                return [], [], []
            target = self._vectors[original]
        if target.shape[0] == 0:
            return [], [], []
        s = self._fn(q, target)
        # `[:k]` is a Python slice in the reference (:178,:181): k = 0 -> empty, k < 0 drops the last |k|
        if self.metric == "euclidean":
            order = np.argsort(s, kind="stable")[:k]
        else:
            order = np.argsort(-s, kind="stable")[:k]
        top = s[order]
        idx = [original[i] for i in order.tolist()] if original is not None else order.tolist()
        return idx, top.tolist(), [self._metadata[i] for i in idx]

    def batch_query(self, queries, k: int = 10):
        """Missing in the reference (api/routes/vectors.py:291 calls it); defined as a
        loop over `query` -- the same convention the engine adopts."""
        return [self.query(q, k) for q in to_f32(queries)]

    # :198-209
    def clear(self):
        if self.store_path is not None and self.store_path.exists():
            shutil.rmtree(self.store_path)
            self.store_path.mkdir(parents=True, exist_ok=True)
        self._vectors, self._metadata, self._vector_count = None, [], 0

    # :241-242
    def get_stats(self) -> Dict:
        return {"vector_count": self._vector_count, "dimension": self.dimension,
                "metric": self.metric, "index_type": "flat"}

    # :218-223 -- `mx.savez(vectors.npz, vectors=...)` + one JSON object per line
    def _save_store(self):
        if self.store_path is None or self._vectors is None:
            return
        np.savez(str(self.store_path / "vectors.npz"), vectors=self._vectors)
        with open(self.store_path / "metadata.jsonl", "w") as f:
            for m in self._metadata:
                f.write(json.dumps(m) + "\n")

    # :225-239
    def _load_store(self):
        p = self.store_path / "vectors.npz"
        if not p.exists():
            return
        try:
            self._vectors = np.load(str(p))["vectors"].astype(np.float32)
            self._vector_count = self._vectors.shape[0]
            mp = self.store_path / "metadata.jsonl"
            if mp.exists():
                with open(mp) as f:
                    self._metadata = [json.loads(line) for line in f]
        except Exception:
            self._vectors, self._metadata, self._vector_count = None, [], 0


# --------------------------------------------------------------------------- #
# timed variant for bench.py's CPU baseline / reference arm
# --------------------------------------------------------------------------- #
def batch_similarity_search_timed(queries, db, k: int = 10, threads: int = 1, return_scores: bool = False):
    """`batch_similarity_search` (performance/mlx_optimized.py:217-248) with its three phases
    timed separately and spread over `threads` host threads where NumPy itself is serial:
    (1) normalise queries and the WHOLE database (:69-83, done on every call by the
    reference), (2) one fp32 GEMM (:86, BLAS threads), (3) per-row full stable argsort of the
    negated scores + gather (:235-244).  Same arithmetic and tie rule as the untimed function.
    Returns (ids, scores, {"normalize_s", "matmul_s", "argsort_s"}); with `return_scores` the
    dict also carries the full (B, N) score matrix under "score_matrix" (for compare_topk)."""
    import time
    from concurrent.futures import ThreadPoolExecutor

    q = to_f32(queries)
    v = to_f32(db)
    if q.ndim == 1:
        q = q.reshape(1, -1)
    B, n = q.shape[0], v.shape[0]
    kk = max(0, min(int(k), n))
    threads = max(1, int(threads))
    t0 = time.perf_counter()
    qn = q / np.maximum(_row_norms(q), EPS)
    vn = np.empty_like(v)

    def _norm_block(lo_hi):
        lo, hi = lo_hi
        blk = v[lo:hi]
        np.divide(blk, np.maximum(_row_norms(blk), EPS), out=vn[lo:hi])

    step = max(1, (n + threads - 1) // threads)
    blocks = [(lo, min(n, lo + step)) for lo in range(0, n, step)]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(_norm_block, blocks))
        t1 = time.perf_counter()
        s = qn @ vn.T
        t2 = time.perf_counter()
        ids = np.empty((B, kk), np.int32)
        sc = np.empty((B, kk), np.float32)

        def _sort_row(b):
            order = np.argsort(-s[b], kind="stable")[:kk]
            ids[b] = order
            sc[b] = s[b][order]

        list(ex.map(_sort_row, range(B)))
        t3 = time.perf_counter()
    t = {"normalize_s": t1 - t0, "matmul_s": t2 - t1, "argsort_s": t3 - t2}
    if return_scores:
        t["score_matrix"] = s
    return ids, sc, t


def batch_similarity_search_optimized_timed(queries, db_normalized, k: int = 10, threads: int = 1):
    """The "optimised CPU" line of BASELINE.md section 2: what a careful CPU implementation of the
    same search would do -- database normalised ONCE (passed in, untimed), one fp32 GEMM, then
    `argpartition` to the best k and a sort of those k instead of a full sort of N.  Same results
    as `batch_similarity_search` up to the tie order inside the k-th score.  Separates the
    reference's algorithmic waste (per-call re-normalisation, full argsort) from the hardware.
    Returns (ids, scores, {"matmul_s", "select_s"})."""
    import time
    from concurrent.futures import ThreadPoolExecutor

    q = to_f32(queries)
    if q.ndim == 1:
        q = q.reshape(1, -1)
    B, n = q.shape[0], db_normalized.shape[0]
    kk = max(0, min(int(k), n))
    t0 = time.perf_counter()
    qn = q / np.maximum(_row_norms(q), EPS)
    s = qn @ db_normalized.T
    t1 = time.perf_counter()
    ids = np.empty((B, kk), np.int32)
    sc = np.empty((B, kk), np.float32)

    def _select_row(b):
        part = np.argpartition(-s[b], kk - 1)[:kk] if kk < n else np.arange(n)
        order = part[np.lexsort((part, -s[b][part]))]
        ids[b] = order
        sc[b] = s[b][order]

    with ThreadPoolExecutor(max_workers=max(1, int(threads))) as ex:
        list(ex.map(_select_row, range(B)))
    t2 = time.perf_counter()
    return ids, sc, {"matmul_s": t1 - t0, "select_s": t2 - t1}
