"""GPU: the drop-in surface -- `MLXVectorStore` / `create_optimized_vector_store`
(service/optimized_vector_store.py:59-246) and the performance/mlx_optimized.py functions --
checked on the reference's own behavioural assertions and against the oracle store."""
import sys

import numpy as np
import pytest
import torch

from oracle import compare, datasets, vs_oracle

pytestmark = pytest.mark.gpu


def test_reference_integration_scenario(tmp_path, native_lib):
    """tests/test_integration.py:83-160 of the reference, minus HTTP."""
    from b200vs import create_optimized_vector_store
    rng = np.random.default_rng(0)
    vecs = rng.random((100, 384), dtype=np.float32)
    meta = [{"id": f"doc_{i}", "content_hash": f"hash_{i}"} for i in range(100)]
    st = create_optimized_vector_store(str(tmp_path / "s"), dimension=384)
    assert st.query(vecs[0]) == ([], [], [])                       # empty store (:117)
    assert st.add_vectors(vecs.tolist(), meta) == {"vectors_added": 100, "total_vectors": 100}
    assert st.get_stats()["vector_count"] == 100                   # :110
    idx, sc, md = st.query(vecs[0].tolist(), k=5)                  # :133-136
    assert len(idx) == 5 and md[0]["id"] == "doc_0" and sc[0] > 0.999
    assert all(isinstance(i, int) for i in idx) and all(isinstance(s, float) for s in sc)
    idx, sc, md = st.query(vecs[10], k=1, filter_metadata={"content_hash": "hash_10"})  # :158-160
    assert len(md) == 1 and md[0]["id"] == "doc_10" and idx == [10]
    stats = st.get_stats()
    assert stats["index_type"] == "flat" and stats["metric"] == "cosine" and stats["memory_usage_mb"] > 0
    assert st.health_check() == {"healthy": True, "issues": []}
    # same answers as the oracle store, including the filter path
    ora = vs_oracle.OracleVectorStore(None, dimension=384)
    ora.add_vectors(vecs, meta)
    for qi in (0, 7, 99):
        a = st.query(vecs[qi], k=10)
        b = ora.query(vecs[qi], k=10)
        assert a[0] == b[0] and a[2] == b[2]
        np.testing.assert_allclose(a[1], b[1], atol=1e-5)
    st.close()
    # persistence round trip; the files are the reference's format after optimize()
    st2 = create_optimized_vector_store(str(tmp_path / "s"), dimension=384)
    assert st2.get_stats()["vector_count"] == 100
    assert st2.query(vecs[3], k=1)[0] == [3]
    st2.optimize()
    z = np.load(tmp_path / "s" / "vectors.npz")
    np.testing.assert_array_equal(z["vectors"], vecs)
    assert len((tmp_path / "s" / "metadata.jsonl").read_text().splitlines()) == 100
    ora2 = vs_oracle.OracleVectorStore(str(tmp_path / "s"), dimension=384)   # reference-format reader
    assert ora2.get_stats()["vector_count"] == 100
    st2.clear()
    assert st2.get_stats()["vector_count"] == 0 and st2.query(vecs[0]) == ([], [], [])
    st2.close()


def test_demo_filter_semantics(make_store):
    """tests/demo.py:229-243: AND over keys, empty result on no match."""
    rng = np.random.default_rng(1)
    v = rng.standard_normal((20, 128)).astype(np.float32)
    meta = [{"id": f"doc_{i}", "category": "A" if i < 10 else "B", "priority": i % 3,
             "lang": "de" if i % 2 == 0 else "en"} for i in range(20)]
    st = make_store(128)
    st.add_vectors(v, meta)
    ora = vs_oracle.OracleVectorStore(None, dimension=128)
    ora.add_vectors(v, meta)
    for flt in ({"category": "A"}, {"priority": 1, "lang": "en"}, {"category": "C"}):
        got = st.query(v[0], k=10, filter_metadata=flt)
        want = ora.query(v[0], k=10, filter_metadata=flt)
        assert got[0] == want[0] and got[2] == want[2]
        np.testing.assert_allclose(got[1], want[1], atol=1e-5)
    assert st.query(v[0], k=10, filter_metadata={"category": "C"}) == ([], [], [])


def test_batch_query_and_errors(make_store):
    d = 96
    db = datasets.make_db(5000, d)
    q = datasets.make_queries(20, d)
    for metric in ("cosine", "euclidean"):
        st = make_store(d, metric)
        st.add_vectors(db, [{"i": i} for i in range(5000)])
        res = st.batch_query(q, k=7)
        assert len(res) == 20
        ref_ids, ref_scores, S = vs_oracle.search(q, db, 7, metric)
        ids = np.array([r[0] for r in res]); sc = np.array([r[1] for r in res], np.float32)
        rep = compare.compare_topk(ref_ids, ref_scores, ids, sc, S)
        assert rep.ok, f"{rep}"
        assert res[3][2] == [{"i": i} for i in res[3][0]]
        single = st.query(q[3], k=7)
        assert single[0] == res[3][0]
    st = make_store(d)
    with pytest.raises(ValueError):
        st.add_vectors(np.zeros((3, d + 1), np.float32), [{}] * 3)
    st.add_vectors(db[:10], [{}] * 10)
    with pytest.raises(ValueError):
        st.query(np.zeros(d + 1, np.float32))
    assert st.query(q[0], k=0) == ([], [], [])
    assert len(st.query(q[0], k=50)[0]) == 10                      # k > N -> N results
    # reference :153-154: no score function -> RuntimeError
    nojit = make_store(d, jit_compile=False)
    nojit.add_vectors(db[:10], [{}] * 10)
    with pytest.raises(RuntimeError):
        nojit.query(q[0])
    # dot_product: advertised by service/models.py:23-27, served here
    dot = make_store(d, "dot_product")
    dot.add_vectors(db, [{}] * 5000)
    ref_ids, ref_scores, S = vs_oracle.search(q[:2], db, 5, "dot_product")
    ids, sc = dot.search_arrays(q[:2], 5)
    assert compare.compare_topk(ref_ids, ref_scores, ids, sc, S).ok


def test_ops_module_matches_oracle(native_lib):
    from b200vs import ops
    d = 64
    db = datasets.make_db(3000, d)
    q = datasets.make_queries(5, d)
    tdb, tq = torch.from_numpy(db).cuda(), torch.from_numpy(q).cuda()
    np.testing.assert_allclose(ops.compute_cosine_similarity_single(tq[0], tdb).cpu().numpy(),
                               vs_oracle.cosine_similarity_single(q[0], db), atol=2e-6)
    np.testing.assert_allclose(ops.compute_cosine_similarity_batch(tq, tdb).cpu().numpy(),
                               vs_oracle.cosine_similarity_batch(q, db), atol=2e-6)
    np.testing.assert_allclose(ops.compute_euclidean_distance(tq[1], tdb).cpu().numpy(),
                               vs_oracle.euclidean_distance(q[1], db), rtol=1e-5)
    np.testing.assert_allclose(ops.compute_dot_product(tq[2], tdb).cpu().numpy(),
                               vs_oracle.dot_product(q[2], db), atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(ops.normalize_vectors(tdb).cpu().numpy(),
                               vs_oracle.normalize_vectors(db), atol=1e-6)
    s = vs_oracle.cosine_similarity_single(q[0], db)
    np.testing.assert_array_equal(ops.fast_top_k_indices(torch.from_numpy(s).cuda(), 10).cpu().numpy(),
                                  vs_oracle.top_k_indices(s, 10))
    assert ops.fast_top_k_indices(torch.from_numpy(s).cuda(), 0).shape == (0,)
    i1, s1 = ops.optimized_similarity_search(tq[0], tdb, 10)
    ri, rs = vs_oracle.similarity_search(q[0], db, 10)
    S = vs_oracle.cosine_similarity_batch(q, db)
    assert compare.compare_topk(ri[None], rs[None], i1.cpu().numpy()[None], s1.cpu().numpy()[None], S[:1]).ok
    ib, sb = ops.optimized_batch_similarity_search(tq, tdb, 10)
    rbi, rbs = vs_oracle.batch_similarity_search(q, db, 10)
    assert compare.compare_topk(rbi, rbs, ib.cpu().numpy(), sb.cpu().numpy(), S).ok
    ei, es = ops.optimized_batch_similarity_search(tq, tdb[:0], 10)
    assert ei.shape == (5, 0) and es.shape == (5, 0)
    cat = ops.optimized_vector_addition(tdb[:10], tdb[10:30], normalize=True)
    np.testing.assert_allclose(cat.cpu().numpy(), vs_oracle.vector_addition(db[:10], db[10:30], True), atol=1e-6)
    with pytest.raises(ValueError):
        ops.compute_cosine_similarity_batch(tq, tdb[:, :32])
    with pytest.raises(ValueError):
        ops.compute_cosine_similarity_single(tq, tdb)
    with pytest.raises(ValueError):
        ops.normalize_vectors(tq[0])
    with pytest.raises(ValueError):
        ops.fast_vector_concatenation(tdb, tdb[:, :32])
    ops.warmup_compiled_functions(64, 50)
    ops._db_cache.clear()


def test_import_path_shims(native_lib):
    """`from service.optimized_vector_store import MLXVectorStore` resolves to the engine."""
    for m in ("service", "service.optimized_vector_store", "performance", "performance.mlx_optimized"):
        sys.modules.pop(m, None)
    from service.optimized_vector_store import MLXVectorStore as A
    from performance.mlx_optimized import optimized_batch_similarity_search as f
    import b200vs
    assert A is b200vs.MLXVectorStore and f is b200vs.ops.optimized_batch_similarity_search


def test_http_shim_replays_reference_integration_test(tmp_path, native_lib):
    """The reference's tests/test_integration.py:46-160 flow against the /vectors shim:
    add 100 x 384, count == 100, query with row 0 -> 5 results, rank-1 is doc_0 with
    similarity > 0.999, filtered query finds exactly the matching document; plus the batched
    route the reference cannot serve."""
    from fastapi.testclient import TestClient
    from b200vs.api_shim import create_app
    app = create_app(str(tmp_path / "stores"), dimension=384, persist=False, max_vectors=10000)
    client = TestClient(app)
    rng = np.random.default_rng(0)
    vecs = rng.random((100, 384), dtype=np.float32)
    meta = [{"id": f"doc_{i}", "content_hash": f"hash_{i}"} for i in range(100)]
    body = {"user_id": "u", "model_id": "m"}
    r = client.post("/vectors/add", json={**body, "vectors": vecs.tolist(), "metadata": meta})
    assert r.status_code == 200 and r.json()["vectors_added"] == 100 and r.json()["total_vectors"] == 100
    r = client.get("/vectors/count", params=body)
    assert r.json()["count"] == 100                                             # :110
    r = client.post("/vectors/query", json={**body, "query": vecs[0].tolist(), "k": 5})
    res = r.json()["results"]
    assert r.status_code == 200 and len(res) == 5                               # :133
    assert res[0]["metadata"]["id"] == "doc_0" and res[0]["similarity_score"] > 0.999   # :134-136
    assert [x["rank"] for x in res] == [1, 2, 3, 4, 5]
    r = client.post("/vectors/query", json={**body, "query": vecs[10].tolist(), "k": 1,
                                            "filter_metadata": {"content_hash": "hash_10"}})
    res = r.json()["results"]
    assert len(res) == 1 and res[0]["metadata"]["id"] == "doc_10"               # :158-160
    r = client.post("/vectors/batch_query", json={**body, "queries": vecs[:7].tolist(), "k": 3})
    out = r.json()
    assert r.status_code == 200 and out["total_queries"] == 7
    assert [q[0]["metadata"]["id"] for q in out["results"]] == [f"doc_{i}" for i in range(7)]
    assert all(len(q) == 3 and q[0]["similarity_score"] > 0.999 for q in out["results"])
    r = client.post("/vectors/add", json={**body, "vectors": vecs[:2].tolist(), "metadata": meta[:1]})
    assert r.status_code == 422                                                  # schema validation
    # request strings never leave base_path, unknown stores are not created by reads
    for bad in ("../../x", "a/b", "..", "a_b", ""):
        r = client.post("/vectors/add", json={"user_id": bad, "model_id": "m", "vectors": vecs[:1].tolist(),
                                              "metadata": meta[:1]})
        assert r.status_code == 400, (bad, r.status_code)
    assert client.get("/vectors/count", params={"user_id": "nobody", "model_id": "m"}).status_code == 404
    assert not (tmp_path / "stores" / "nobody").exists()
    assert sorted(p.name for p in (tmp_path / "stores").iterdir()) == ["u"]
    app.state.store_manager.close()


def test_concurrent_queries_during_appends(make_store):
    """Threading contract of the reference: `query` takes no lock and may run while
    one `add_vectors` is in flight (api/routes/vectors.py:43 serves both from a 4-thread pool).
    Every result must be consistent with SOME prefix of the appended rows: ids below the row
    count seen after the call, sorted scores, and the planted best match once it is visible."""
    import threading
    d, chunk, chunks = 128, 5000, 24
    rng = np.random.default_rng(3)
    base = rng.standard_normal((chunk * chunks, d)).astype(np.float32)
    probe = rng.standard_normal(d).astype(np.float32)
    planted_at = chunk * 10 + 17
    base[planted_at] = probe * 3.0                       # cosine 1.0 with the probe
    st = make_store(d)
    st.add_vectors(base[:chunk], [{} for _ in range(chunk)])
    errors, seen_planted = [], []
    stop = threading.Event()

    def reader(batch):
        q = np.tile(probe, (batch, 1)) if batch > 1 else probe
        while not stop.is_set():
            try:
                if batch == 1:
                    ids, scores, _ = st.query(q, k=5)
                    res = [(ids, scores)]
                else:
                    res = [(i, s) for i, s, _ in st.batch_query(q, k=5)]
                n_after = st.get_stats()["vector_count"]
                for ids, scores in res:
                    assert len(ids) == 5 and all(0 <= i < n_after for i in ids), (ids, n_after)
                    assert all(a >= b for a, b in zip(scores, scores[1:]))
                    if ids[0] == planted_at:
                        assert scores[0] > 0.9999
                        seen_planted.append(True)
                    else:
                        assert scores[0] < 0.9          # random rows are far from the probe
            except Exception as e:                       # noqa: BLE001
                errors.append(repr(e))
                stop.set()

    threads = [threading.Thread(target=reader, args=(b,)) for b in (1, 1, 8)]
    for t in threads:
        t.start()
    for c in range(1, chunks):
        st.add_vectors(base[c * chunk:(c + 1) * chunk], [{} for _ in range(chunk)])
    import time as _t
    _t.sleep(0.2)
    stop.set()
    for t in threads:
        t.join(timeout=30)
    assert not errors, errors[:3]
    assert seen_planted, "the planted row was never returned after it became visible"
    ids, _, _ = st.query(probe, k=1)
    assert ids == [planted_at]


def test_ops_search_never_serves_a_stale_database(native_lib):
    """The functional API caches the ingested database per tensor OBJECT.  Two databases of equal
    shape searched back to back -- as numpy arrays (temporaries on the device each call) and as
    CUDA tensors allocated one after the other (the allocator reuses the address) -- must each be
    searched for real; an in-place update of a cached tensor must be seen too."""
    from b200vs import ops
    from oracle import vs_oracle
    n, d, k = 3000, 64, 5
    rng = np.random.default_rng(21)
    a = rng.standard_normal((n, d), dtype=np.float32)
    b = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((3, d), dtype=np.float32)
    ref_a, _ = vs_oracle.batch_similarity_search(q, a, k)
    ref_b, _ = vs_oracle.batch_similarity_search(q, b, k)
    assert not np.array_equal(ref_a, ref_b)
    for _ in range(2):
        ia, _ = ops.optimized_batch_similarity_search(q, a, k)
        ib, _ = ops.optimized_batch_similarity_search(q, b, k)
        np.testing.assert_array_equal(ia.cpu().numpy(), ref_a)
        np.testing.assert_array_equal(ib.cpu().numpy(), ref_b)
    ta = torch.from_numpy(a).cuda()
    ia, _ = ops.optimized_batch_similarity_search(q, ta, k)
    np.testing.assert_array_equal(ia.cpu().numpy(), ref_a)
    ptr = ta.data_ptr()
    del ta
    tb = torch.from_numpy(b).cuda()              # very likely the same address, same shape, _version 0
    ib, _ = ops.optimized_batch_similarity_search(q, tb, k)
    np.testing.assert_array_equal(ib.cpu().numpy(), ref_b)
    tb.copy_(torch.from_numpy(a))                # in-place update of a cached tensor
    ia, _ = ops.optimized_batch_similarity_search(q, tb, k)
    np.testing.assert_array_equal(ia.cpu().numpy(), ref_a)
    assert ptr  # (the address reuse itself is allocator behaviour, not asserted)
    ops._db_cache.clear()
