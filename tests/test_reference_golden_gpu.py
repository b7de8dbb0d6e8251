"""GPU: the CUDA engine against fixtures produced by RUNNING THE REFERENCE'S OWN MODULES
(tests/golden/make_reference_golden.py; see tests/golden/mlx_standin/mlx/core.py for what the
NumPy stand-in for `mlx.core` does and does not pin).  The engine is driven through the drop-in
surface -- `MLXVectorStore` and the `performance/mlx_optimized.py` function names -- so these
read like the reference's own calls.

Tolerance (BASELINE.json north_star): ids identical to the reference's except inside tie groups
(reference scores within 1e-6 relative), scores within 1e-5.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import compare, vs_oracle
from test_reference_golden import replay_store

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).parent / "golden"
OPS = sorted(GOLD.glob("ref_ops_*.npz"))
STORES = sorted(GOLD.glob("ref_store_*.json"))
ATOL = 1e-5


def ties_ok(ref_ids, ref_scores, got_ids, got_scores, full_scores, larger_is_better=True):
    """ids equal position by position, or differing only where the REFERENCE's scores of the two
    ids are within 1e-6 relative; engine scores within 1e-5 of the reference's for that id."""
    ref_ids, got_ids = np.asarray(ref_ids).reshape(1, -1), np.asarray(got_ids).reshape(1, -1)
    rep = compare.compare_topk(ref_ids, np.asarray(ref_scores).reshape(1, -1), got_ids,
                               np.asarray(got_scores).reshape(1, -1), np.asarray(full_scores).reshape(1, -1),
                               tie_rtol=1e-6, score_atol=ATOL, score_rtol=1e-5)
    return rep


@pytest.mark.parametrize("path", OPS, ids=[p.stem for p in OPS])
def test_ops_functions(path, native_lib):
    import torch
    from b200vs import ops
    g = np.load(path)
    db, q, k = g["db"], g["q"], int(g["k"])
    dbt = torch.from_numpy(db).cuda()
    for b, row in enumerate(q):
        np.testing.assert_allclose(ops.compute_cosine_similarity_single(row, dbt).cpu().numpy(),
                                   g["cosine_single"][b], atol=ATOL, rtol=0)
        np.testing.assert_allclose(ops.compute_euclidean_distance(row, dbt).cpu().numpy(),
                                   g["euclidean"][b], atol=ATOL, rtol=1e-5)   # distances of O(10)
        np.testing.assert_allclose(ops.compute_dot_product(row, dbt).cpu().numpy(),
                                   g["dot"][b], atol=5e-5, rtol=1e-5)         # un-normalised: sums of O(10) terms
        idx, sc = ops.optimized_similarity_search(row, dbt, k)
        assert idx.shape == g["search_ids"][b].shape
        rep = ties_ok(g["search_ids"][b], g["search_scores"][b], idx.cpu().numpy(), sc.cpu().numpy(),
                      g["cosine_single"][b])
        assert rep.ok, f"{path.stem} query {b}: {rep}"
    np.testing.assert_allclose(ops.compute_cosine_similarity_batch(q, dbt).cpu().numpy(), g["cosine_batch"],
                               atol=ATOL, rtol=0)
    bi, bs = ops.optimized_batch_similarity_search(q, dbt, k)
    assert tuple(bi.shape) == g["batch_ids"].shape == (q.shape[0], min(k, db.shape[0]))
    rep = compare.compare_topk(g["batch_ids"], g["batch_scores"], bi.cpu().numpy(), bs.cpu().numpy(),
                               g["cosine_batch"], tie_rtol=1e-6, score_atol=ATOL, score_rtol=1e-5)
    assert rep.ok, f"{path.stem}: {rep}"
    np.testing.assert_allclose(ops.normalize_vectors(dbt).cpu().numpy(), g["normalized"], atol=1e-6, rtol=0)
    half = db.shape[0] // 2
    np.testing.assert_allclose(ops.optimized_vector_addition(dbt[:half], dbt[half:], normalize=True).cpu().numpy(),
                               g["added_normalized"], atol=1e-6, rtol=0)
    ti = ops.fast_top_k_indices(torch.from_numpy(g["cosine_single"][0]).cuda(), k).cpu().numpy()
    rep = ties_ok(g["topk_of_first"], g["cosine_single"][0][g["topk_of_first"]], ti,
                  g["cosine_single"][0][ti], g["cosine_single"][0])
    assert rep.ok, f"{path.stem} fast_top_k_indices: {rep}"


@pytest.mark.parametrize("path", STORES, ids=[p.stem for p in STORES])
def test_store_sequences(path, make_store):
    """The recorded add / query / filtered-query sequence, replayed on the engine store."""
    rec = json.loads(path.read_text())
    g = np.load(path.with_suffix(".npz"))
    db, q = g["db"], g["q"]
    st = make_store(db.shape[1], rec["metric"])
    got = replay_store(rec, db, q, st)
    assert got["empty_before_add"] == rec["empty_before_add"] == [[], [], []]
    assert got["adds"] == rec["adds"]
    # Per-row scores do not depend on the filter, so the reference's full score vector (needed to
    # judge ties) is the oracle's -- which test_reference_golden.py shows is bit-identical to what
    # the reference computed.
    fn = vs_oracle.cosine_similarity_single if rec["metric"] == "cosine" else vs_oracle.euclidean_distance
    full = {qi: fn(q[qi], db) for qi in range(q.shape[0])}
    for want, have in zip(rec["queries"], got["queries"]):
        assert len(have["ids"]) == len(want["ids"]), want
        if not want["ids"]:
            assert have["ids"] == [] and have["meta_ids"] == []
            continue
        assert have["meta_ids"] == [f"doc_{i}" for i in have["ids"]]
        if want["filter"]:
            for i in have["ids"]:                                    # the filter really was applied
                assert all((i % 5 if key == "group" else ("even" if i % 2 == 0 else "odd")) == val
                           for key, val in want["filter"].items())
        rep = compare.compare_topk(np.asarray([want["ids"]]), np.asarray([want["scores"]]),
                                   np.asarray([have["ids"]]), np.asarray([have["scores"]]),
                                   full[want["q"]][None, :], tie_rtol=1e-6, score_atol=ATOL, score_rtol=1e-5)
        assert rep.ok, f"{path.stem} {want['q']} {want['filter']} k={want['k']}: {rep}"
    stats = st.get_stats()
    for key in ("vector_count", "dimension", "metric", "index_type"):
        assert stats[key] == rec["stats"][key]
    ids2, _, _ = st.query(q[0][None, :], k=rec["k"])
    assert len(ids2) == len(rec["query_2d_first"]["ids"])


def test_error_behaviour(native_lib):
    """Exception types and degenerate shapes equal the reference's (recorded in ref_errors.json)."""
    import torch
    from b200vs import ops
    want = json.loads((GOLD / "ref_errors.json").read_text())
    z4 = np.zeros((2, 4), np.float32)
    db = torch.arange(24, dtype=torch.float32).reshape(6, 4).cuda()
    calls = {
        "cosine_single_2row_query": lambda: ops.compute_cosine_similarity_single(z4, db),
        "cosine_batch_1d_query": lambda: ops.compute_cosine_similarity_batch(np.zeros(4, np.float32), db),
        "cosine_batch_dim_mismatch": lambda: ops.compute_cosine_similarity_batch(np.zeros((2, 5), np.float32), db),
        "topk_2d_scores": lambda: ops.fast_top_k_indices(torch.zeros((2, 3)).cuda(), 2),
        "topk_k0": lambda: ops.fast_top_k_indices(torch.arange(5, dtype=torch.float32).cuda(), 0),
        "topk_k_gt_n": lambda: ops.fast_top_k_indices(torch.arange(5, dtype=torch.float32).cuda(), 9),
        "normalize_1d": lambda: ops.normalize_vectors(np.zeros(4, np.float32)),
        "normalize_empty": lambda: ops.normalize_vectors(torch.zeros((0, 4)).cuda()),
        "concat_dim_mismatch": lambda: ops.fast_vector_concatenation(db, torch.zeros((2, 5)).cuda()),
        "concat_empty_left": lambda: ops.fast_vector_concatenation(torch.zeros((0, 4)).cuda(), db),
        "search_2row_query": lambda: ops.optimized_similarity_search(z4, db, 3),
        "batch_search_empty_db": lambda: ops.optimized_batch_similarity_search(torch.zeros((3, 4)).cuda(),
                                                                               torch.zeros((0, 4)).cuda(), 3),
        "batch_search_k_gt_n": lambda: ops.optimized_batch_similarity_search(torch.ones((3, 4)).cuda(), db, 50),
    }
    for name, fn in calls.items():
        exp = want[name]
        if "raises" in exp:
            with pytest.raises(ValueError):
                fn()
        else:
            r = fn()
            shapes = [list(x.shape) for x in (r if isinstance(r, tuple) else (r,))]
            assert shapes == exp["ok"], name
    assert want["store_no_jit_query"] == {"raises": "RuntimeError"}
