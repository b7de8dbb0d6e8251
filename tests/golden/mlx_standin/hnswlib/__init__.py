"""Import stub: the reference's service/optimized_vector_store.py imports
performance/hnsw_index.py, which imports hnswlib at module level (performance/hnsw_index.py:14).
The exact-search path never touches it (enable_hnsw defaults to False)."""


class Index:  # pragma: no cover - never constructed by the exact path
    def __init__(self, *a, **k):
        raise RuntimeError("hnswlib is not available; the golden generator only runs the exact path")
