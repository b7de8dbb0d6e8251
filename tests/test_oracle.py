"""CPU tests of the oracle: golden fixtures, an independent float64 restatement, and the
behavioural pins of the reference's own tests (tests/test_integration.py:110,133-136,158-160;
tests/demo.py:232,238,243)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import compare, datasets, vs_oracle

GOLDEN = sorted(p for p in (Path(__file__).parent / "golden").glob("*.npz") if not p.name.startswith("ref_"))
METRICS = ("cosine", "euclidean", "dot_product")


def f64_scores(q, db, metric):
    """Independent restatement in float64 with the maths written differently."""
    q = np.asarray(q, np.float64)
    db = np.asarray(db, np.float64)
    if metric == "cosine":
        qn = np.maximum(np.linalg.norm(q, axis=1), 1e-8)
        dn = np.maximum(np.linalg.norm(db, axis=1), 1e-8)
        return np.einsum("bd,nd->bn", q, db) / qn[:, None] / dn[None, :]
    if metric == "euclidean":
        return np.sqrt(((db[None, :, :] - q[:, None, :]) ** 2).sum(-1))
    return np.einsum("bd,nd->bn", q, db)


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
@pytest.mark.parametrize("metric", METRICS)
def test_oracle_matches_golden(path, metric):
    g = np.load(path)
    ids, scores, _ = vs_oracle.search(g["q"], g["db"], int(g["k"]), metric)
    np.testing.assert_array_equal(ids, g[f"ids_{metric}"])
    np.testing.assert_allclose(scores, g[f"scores_{metric}"], rtol=0, atol=0)


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dist", ["normal", "uniform"])
def test_oracle_vs_float64(metric, dist):
    db = datasets.make_db(3000, 96, dist)
    q = datasets.make_queries(6, 96, dist)
    ids, scores, S = vs_oracle.search(q, db, 10, metric)
    S64 = f64_scores(q, db, metric)
    # fp32 rounding scales with sum|q_i x_i| (un-normalised for dot_product)
    np.testing.assert_allclose(S, S64, rtol=2e-5, atol=2e-5 if metric == "dot_product" else 2e-6)
    key = S64 if metric == "euclidean" else -S64
    ids64 = np.argsort(key, axis=1, kind="stable")[:, :10]
    rep = compare.compare_topk(ids64, np.take_along_axis(S64, ids64, 1), ids, scores, S64,
                               tie_rtol=2e-6)
    assert rep.ok, rep.first_failure


def test_single_and_batch_agree():
    db = datasets.make_db(2000, 64)
    q = datasets.make_queries(5, 64)
    bi, bs = vs_oracle.batch_similarity_search(q, db, 10)
    ci, cs = vs_oracle.batch_similarity_search(q, db, 10, chunk=2)
    # BLAS picks different sgemm kernels per batch shape: scores agree to ~1 ulp, not bitwise
    np.testing.assert_array_equal(bi, ci)
    np.testing.assert_allclose(bs, cs, atol=2e-7)
    for b in range(5):
        i1, s1 = vs_oracle.similarity_search(q[b], db, 10)
        np.testing.assert_array_equal(i1, bi[b])
        np.testing.assert_allclose(s1, bs[b], atol=2e-7)


def test_edge_cases():
    db = datasets.make_db(7, 16)
    q = datasets.make_queries(2, 16)
    ids, scores = vs_oracle.batch_similarity_search(q, db, 12)      # k > N -> N results
    assert ids.shape == (2, 7)
    assert vs_oracle.top_k_indices(np.arange(5.0), 0).shape == (0,)  # k <= 0 -> empty
    assert vs_oracle.top_k_indices(np.arange(5.0), -3).shape == (0,)
    e_i, e_s = vs_oracle.batch_similarity_search(q, np.zeros((0, 16), np.float32), 5)
    assert e_i.shape == (2, 0) and e_s.shape == (2, 0)
    with pytest.raises(ValueError):
        vs_oracle.cosine_similarity_batch(q, datasets.make_db(4, 8))
    with pytest.raises(ValueError):
        vs_oracle.cosine_similarity_single(q, db)                    # (2, D) query is rejected
    with pytest.raises(ValueError):
        vs_oracle.normalize_vectors(np.zeros(4, np.float32))
    # stable ties: equal scores -> ascending index
    np.testing.assert_array_equal(vs_oracle.top_k_indices(np.array([1., 3., 3., 2., 3.]), 4), [1, 2, 4, 3])
    # zero row scores 0 through the clamp, not NaN
    z = np.zeros((3, 16), np.float32)
    z[1] = 1
    s = vs_oracle.cosine_similarity_single(np.ones(16, np.float32), z)
    assert np.isfinite(s).all() and s[0] == 0 and abs(s[1] - 1) < 1e-6


def test_reference_behavioural_pins(tmp_path):
    # tests/test_integration.py: 100 x 384 rand; count; self-query rank-1 > 0.999; filter -> doc
    rng = np.random.default_rng(0)
    vecs = rng.random((100, 384), dtype=np.float32)
    meta = [{"id": f"doc_{i}", "content_hash": f"hash_{i}"} for i in range(100)]
    st = vs_oracle.OracleVectorStore(str(tmp_path / "s"), dimension=384)
    assert st.add_vectors(vecs, meta) == {"vectors_added": 100, "total_vectors": 100}
    assert st.get_stats()["vector_count"] == 100
    idx, sc, md = st.query(vecs[0], k=5)
    assert len(idx) == 5 and md[0]["id"] == "doc_0" and sc[0] > 0.999
    idx, sc, md = st.query(vecs[10], k=1, filter_metadata={"content_hash": "hash_10"})
    assert len(md) == 1 and md[0]["id"] == "doc_10"
    # tests/demo.py: AND semantics, empty on no match
    meta2 = [{"id": i, "category": "A" if i < 10 else "B", "priority": i % 3,
              "lang": "de" if i % 2 == 0 else "en"} for i in range(20)]
    st2 = vs_oracle.OracleVectorStore(None, dimension=32)
    v2 = rng.standard_normal((20, 32)).astype(np.float32)
    st2.add_vectors(v2, meta2)
    _, _, r = st2.query(v2[0], k=10, filter_metadata={"category": "A"})
    assert r and all(m["category"] == "A" for m in r)
    _, _, r = st2.query(v2[0], k=10, filter_metadata={"priority": 1, "lang": "en"})
    assert r and all(m["priority"] == 1 and m["lang"] == "en" for m in r)
    assert st2.query(v2[0], k=10, filter_metadata={"category": "C"}) == ([], [], [])
    # persistence round trip in the reference's on-disk format
    st3 = vs_oracle.OracleVectorStore(str(tmp_path / "s"), dimension=384)
    assert st3.get_stats()["vector_count"] == 100
    assert st3.query(vecs[3], k=1)[0] == [3]
    # unsupported metric / jit_compile=False -> RuntimeError at query time
    st4 = vs_oracle.OracleVectorStore(None, dimension=32, metric="dot_product")
    st4.add_vectors(v2, meta2)
    with pytest.raises(RuntimeError):
        st4.query(v2[0])


def test_comparator_tie_rules():
    S = np.array([[0.9, 0.5, 0.5, 0.1]], np.float32)
    ref_ids = np.array([[0, 1, 2]]); ref_s = np.array([[0.9, 0.5, 0.5]], np.float32)
    ok = compare.compare_topk(ref_ids, ref_s, np.array([[0, 2, 1]]), ref_s, S)
    assert ok.ok and ok.id_tie_ok == 2
    bad = compare.compare_topk(ref_ids, ref_s, np.array([[0, 1, 3]]), np.array([[0.9, 0.5, 0.1]]), S)
    assert not bad.ok
    bad_score = compare.compare_topk(ref_ids, ref_s, ref_ids, ref_s + 1e-3, S)
    assert not bad_score.ok
    assert compare.recall_at_k(ref_ids, np.array([[0, 1, 3]])) == pytest.approx(2 / 3)
