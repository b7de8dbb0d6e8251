"""GPU: `ShardedMLXVectorStore` (the reference's store surface over row shards,
service/optimized_vector_store.py:96-192) on the real kernels at world size 1 -- the multi-rank host
logic is covered under gloo in tests/test_sharded_store_cpu.py, the multi-shard kernels in
tests/test_sharded_gpu.py and tests/test_baseline_sizes_gpu.py.  Includes a store large enough
(>= 65 536 rows) for the filtered batch to take the masked tensor-core path."""
import numpy as np
import pytest
import torch

from oracle import datasets, vs_oracle

pytestmark = pytest.mark.gpu


class _Close:
    """(ids, scores, metadata) with scores compared at the contract's 1e-5."""

    def __init__(self, t):
        self.t = t

    def __eq__(self, other):
        a, b = self.t, other.t
        return (list(a[0]) == list(b[0]) and list(a[2]) == list(b[2]) and len(a[1]) == len(b[1]) and
                bool(np.allclose(np.asarray(a[1], np.float64), np.asarray(b[1], np.float64), atol=1e-5)))

    def __repr__(self):
        return repr(self.t)[:400]


def _norm(t):
    return _Close(t)


def test_sharded_store_surface_world1(tmp_path, native_lib):
    from b200vs.sharded_store import ShardedMLXVectorStore
    from b200vs.store import MLXVectorStoreConfig
    n, d = 70_000, 64
    db = datasets.make_db(n, d)
    meta = [{"id": i, "category": "A" if i % 3 else "B", "shard": i % 7} for i in range(n)]
    q = datasets.make_queries(20, d)
    cfg = MLXVectorStoreConfig(dimension=d, metric="cosine", persist=True, max_vectors=100_000)
    st = ShardedMLXVectorStore(str(tmp_path / "s"), cfg)
    for lo, hi in ((0, 1), (1, 30_000), (30_000, n)):
        r = st.add_vectors(torch.from_numpy(db[lo:hi]).cuda() if lo else db[lo:hi], meta[lo:hi])
        assert r == {"vectors_added": hi - lo, "total_vectors": hi}
    ora = vs_oracle.OracleVectorStore(None, dimension=d, metric="cosine")
    ora.add_vectors(db, meta)
    assert _norm(st.query(q[0], k=10)) == _norm(ora.query(q[0], k=10))
    got = st.batch_query(q, k=10, filter_metadata={"category": "A"})          # 2/3 of the rows: masked K3
    for b in range(20):
        assert _norm(got[b]) == _norm(ora.query(q[b], k=10, filter_metadata={"category": "A"})), b
    got = st.batch_query(q[:4], k=5, filter_metadata={"shard": 3, "category": "B"})   # 1/21 of the rows: masked scan
    for b in range(4):
        assert _norm(got[b]) == _norm(ora.query(q[b], k=5, filter_metadata={"shard": 3, "category": "B"})), b
    assert st.query(q[0], k=3, filter_metadata={"category": "nope"}) == ([], [], [])
    assert st.health_check()["healthy"]
    st.optimize()
    st.close()
    snap = np.load(tmp_path / "s" / "vectors.npz")["vectors"]
    np.testing.assert_array_equal(snap, db)
    st2 = ShardedMLXVectorStore(str(tmp_path / "s"), cfg)                     # reload from the reference format
    assert st2.get_stats()["vector_count"] == n
    assert _norm(st2.query(q[1], k=10, filter_metadata={"category": "B"})) == \
        _norm(ora.query(q[1], k=10, filter_metadata={"category": "B"}))
    st2.clear()
    assert st2.query(q[0], k=3) == ([], [], []) and st2.get_stats()["vector_count"] == 0
    st2.close()
