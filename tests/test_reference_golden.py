"""The oracle against fixtures produced by RUNNING THE REFERENCE'S OWN MODULES
(tests/golden/make_reference_golden.py: service/optimized_vector_store.py and
performance/mlx_optimized.py imported unmodified from /root/reference over a NumPy stand-in
for `mlx.core`).  This is what pins the oracle: op order, clamps, slicing, id mapping, filter
semantics, degenerate cases and exception types are the reference's; only the arithmetic inside
the mlx primitives is NumPy's.  CPU only.  The committed fixtures are all these tests need; the one
test that re-runs the generator is skipped wherever /root/reference is not mounted.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import vs_oracle

GOLD = Path(__file__).parent / "golden"
OPS = sorted(GOLD.glob("ref_ops_*.npz"))
STORES = sorted(GOLD.glob("ref_store_*.json"))


def test_fixtures_present():
    assert len(OPS) >= 4 and len(STORES) >= 4 and (GOLD / "ref_errors.json").exists()


@pytest.mark.parametrize("path", OPS, ids=[p.stem for p in OPS])
def test_ops_bit_exact(path):
    """Same NumPy primitives underneath, so the restatement must agree BIT FOR BIT with what the
    reference's functions returned -- any difference is a difference in op order or logic."""
    g = np.load(path)
    db, q, k = g["db"], g["q"], int(g["k"])
    for b, row in enumerate(q):
        np.testing.assert_array_equal(vs_oracle.cosine_similarity_single(row, db), g["cosine_single"][b])
        np.testing.assert_array_equal(vs_oracle.euclidean_distance(row, db), g["euclidean"][b])
        np.testing.assert_array_equal(vs_oracle.dot_product(row, db), g["dot"][b])
        idx, sc = vs_oracle.similarity_search(row, db, k)
        np.testing.assert_array_equal(idx, g["search_ids"][b])
        np.testing.assert_array_equal(sc, g["search_scores"][b])
    np.testing.assert_array_equal(vs_oracle.cosine_similarity_batch(q, db), g["cosine_batch"])
    bi, bs = vs_oracle.batch_similarity_search(q, db, k)
    np.testing.assert_array_equal(bi, g["batch_ids"])
    np.testing.assert_array_equal(bs, g["batch_scores"])
    assert bi.shape[1] == min(k, db.shape[0])
    np.testing.assert_array_equal(vs_oracle.normalize_vectors(db), g["normalized"])
    np.testing.assert_array_equal(vs_oracle.top_k_indices(g["cosine_single"][0], k), g["topk_of_first"])
    half = db.shape[0] // 2
    np.testing.assert_array_equal(vs_oracle.vector_addition(db[:half], db[half:], normalize=True),
                                  g["added_normalized"])
    # the generic search the GPU parity tests use is the same thing
    ids, scores, _ = vs_oracle.search(q, db, k, "cosine")
    np.testing.assert_array_equal(ids, g["batch_ids"])
    np.testing.assert_array_equal(scores, g["batch_scores"])


def replay_store(rec, db, q, store):
    """Re-run the recorded add/query sequence on `store` (oracle or engine)."""
    out = {"empty_before_add": [list(x) for x in store.query(q[0], k=rec["k"])], "adds": [], "queries": []}
    meta = [{"id": f"doc_{i}", "group": int(i % 5), "parity": "even" if i % 2 == 0 else "odd"}
            for i in range(db.shape[0])]
    lo = 0
    for m in rec["batches"]:
        out["adds"].append(store.add_vectors(db[lo:lo + m], meta[lo:lo + m]))
        lo += m
    for item in rec["queries"]:
        ids, scores, metas = store.query(q[item["q"]], k=item["k"], filter_metadata=item["filter"])
        out["queries"].append({"ids": [int(i) for i in ids], "scores": [float(np.float32(s)) for s in scores],
                               "meta_ids": [m_["id"] for m_ in metas]})
    return out


@pytest.mark.parametrize("path", STORES, ids=[p.stem for p in STORES])
def test_store_bit_exact(path, tmp_path):
    rec = json.loads(path.read_text())
    g = np.load(path.with_suffix(".npz"))
    db, q = g["db"], g["q"]
    st = vs_oracle.OracleVectorStore(str(tmp_path / "s"), dimension=db.shape[1], metric=rec["metric"])
    got = replay_store(rec, db, q, st)
    assert got["empty_before_add"] == rec["empty_before_add"] == [[], [], []]
    assert got["adds"] == rec["adds"]
    for want, have in zip(rec["queries"], got["queries"]):
        assert have["ids"] == want["ids"], want
        assert have["meta_ids"] == want["meta_ids"]
        assert have["scores"] == want["scores"]
    stats = st.get_stats()
    assert stats == rec["stats"]
    ids2, sc2, _ = st.query(q[0][None, :], k=rec["k"])
    assert [int(i) for i in ids2] == rec["query_2d_first"]["ids"]
    # on-disk layout and reload
    assert rec["disk_keys"] == ["vectors"] and rec["disk_metadata_lines"] == db.shape[0]
    with np.load(tmp_path / "s" / "vectors.npz") as z:
        assert list(z.files) == ["vectors"]
        np.testing.assert_array_equal(z["vectors"], db)
    st2 = vs_oracle.OracleVectorStore(str(tmp_path / "s"), dimension=db.shape[1], metric=rec["metric"])
    ids3, sc3, _ = st2.query(q[0], k=rec["k"])
    assert [int(i) for i in ids3] == rec["reloaded_first"]["ids"]
    assert [float(np.float32(s)) for s in sc3] == rec["reloaded_first"]["scores"]


def test_behavioural_pins_hold_in_the_reference_fixtures():
    """The reference's own test assertions (tests/test_integration.py:133-136,158-160;
    tests/demo.py:232,238,243) are true of what the reference returned."""
    rec = json.loads((GOLD / "ref_store_cosine_adversarial_130x32.json").read_text())
    first = rec["queries"][0]                     # query = stored row 3 (duplicated at 20, 21)
    assert first["ids"][:3] == [3, 20, 21] and first["scores"][0] > 0.999      # self-match, ties -> lower id
    for item in rec["queries"]:
        n_match = 130 if not item["filter"] else sum(
            all((i % 5 if key == "group" else ("even" if i % 2 == 0 else "odd")) == val
                for key, val in item["filter"].items()) for i in range(130))
        # `argsort(...)[:k]` is a Python slice: k > N -> N results, k = 0 -> none, k < 0 drops the last |k|
        assert len(item["ids"]) == (min(item["k"], n_match) if item["k"] >= 0 else max(0, n_match + item["k"]))
        if item["filter"] == {"group": 99}:
            assert item["ids"] == []               # no match -> empty, not an error
        elif item["filter"]:
            for i in item["ids"]:
                assert i % 5 == item["filter"]["group"]
                if "parity" in item["filter"]:
                    assert (i % 2 == 0) == (item["filter"]["parity"] == "even")
    big = [it for it in rec["queries"] if it["k"] > 130]
    assert big and len(big[0]["ids"]) == 130       # k > N returns N results


def test_error_behaviour():
    want = json.loads((GOLD / "ref_errors.json").read_text())
    z4 = np.zeros((2, 4), np.float32)
    db = np.arange(24, dtype=np.float32).reshape(6, 4)
    calls = {
        "cosine_single_2row_query": lambda: vs_oracle.cosine_similarity_single(z4, db),
        "cosine_batch_1d_query": lambda: vs_oracle.cosine_similarity_batch(np.zeros(4, np.float32), db),
        "cosine_batch_dim_mismatch": lambda: vs_oracle.cosine_similarity_batch(np.zeros((2, 5), np.float32), db),
        "topk_2d_scores": lambda: vs_oracle.top_k_indices(np.zeros((2, 3), np.float32), 2),
        "topk_k0": lambda: vs_oracle.top_k_indices(np.arange(5, dtype=np.float32), 0),
        "topk_k_gt_n": lambda: vs_oracle.top_k_indices(np.arange(5, dtype=np.float32), 9),
        "normalize_1d": lambda: vs_oracle.normalize_vectors(np.zeros(4, np.float32)),
        "normalize_empty": lambda: vs_oracle.normalize_vectors(np.zeros((0, 4), np.float32)),
        "concat_dim_mismatch": lambda: vs_oracle.vector_concatenation(db, np.zeros((2, 5), np.float32)),
        "concat_empty_left": lambda: vs_oracle.vector_concatenation(np.zeros((0, 4), np.float32), db),
        "search_2row_query": lambda: vs_oracle.similarity_search(z4, db, 3),
        "batch_search_empty_db": lambda: vs_oracle.batch_similarity_search(np.zeros((3, 4), np.float32),
                                                                          np.zeros((0, 4), np.float32), 3),
        "batch_search_k_gt_n": lambda: vs_oracle.batch_similarity_search(np.ones((3, 4), np.float32), db, 50),
    }
    for name, fn in calls.items():
        exp = want[name]
        if "raises" in exp:
            with pytest.raises(ValueError):
                fn()
            assert exp["raises"] == "ValueError"
        else:
            r = fn()
            shapes = [list(np.asarray(x).shape) for x in (r if isinstance(r, tuple) else (r,))]
            assert shapes == exp["ok"], name
    for label, kw in (("store_dot_product", {"metric": "dot_product"}), ("store_no_jit", {"jit_compile": False})):
        st = vs_oracle.OracleVectorStore(None, dimension=4, **kw)
        st.add_vectors(db, [{} for _ in range(6)])
        assert want[label + "_query"] == {"raises": "RuntimeError"}
        with pytest.raises(RuntimeError):
            st.query(np.zeros(4, np.float32), k=2)


@pytest.mark.skipif(not Path("/root/reference/service/optimized_vector_store.py").exists(),
                    reason="the reference tree is only mounted in the build container")
def test_fixtures_are_what_the_reference_produces_today(tmp_path):
    """Re-run the generator against the mounted reference and compare with the committed files, so a
    stale or hand-edited fixture cannot go unnoticed.  (CPU suite only; the GPU box has no reference.)"""
    import shutil
    import subprocess
    import sys
    work = tmp_path / "golden"
    shutil.copytree(GOLD / "mlx_standin", work / "mlx_standin")
    shutil.copy(GOLD / "make_reference_golden.py", work / "make_reference_golden.py")
    subprocess.run([sys.executable, str(work / "make_reference_golden.py")], check=True, capture_output=True)
    fresh = sorted(p.name for p in work.glob("ref_*"))
    assert fresh == sorted(p.name for p in GOLD.glob("ref_*"))
    for name in fresh:
        if name.endswith(".json"):
            assert json.loads((work / name).read_text()) == json.loads((GOLD / name).read_text()), name
        else:
            a, b = np.load(work / name), np.load(GOLD / name)
            assert sorted(a.files) == sorted(b.files), name
            for key in a.files:
                np.testing.assert_array_equal(a[key], b[key], err_msg=f"{name}:{key}")
