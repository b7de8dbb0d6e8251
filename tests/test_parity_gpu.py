"""GPU parity tests: the CUDA engine, called through the C-ABI (ctypes -> libb200vs.so),
against the CPU oracle on the same seeded inputs.

Tolerance (BASELINE.json north_star): fp32 ids identical to the oracle except inside tie
groups (oracle scores within 1e-6 relative, unit scale for cosine), scores within 1e-5
(+1e-5 relative for un-normalised metrics).  bf16-database modes report recall@k."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import compare, datasets, vs_oracle

pytestmark = pytest.mark.gpu

GOLDEN = sorted(p for p in (Path(__file__).parent / "golden").glob("*.npz") if not p.name.startswith("ref_"))
METRICS = ("cosine", "euclidean", "dot_product")


def flags_of(mode="scan_fp32", variant=None):
    from b200vs import _cabi
    f = _cabi.SEARCH_MODES[mode]
    if variant == "tma":
        f |= _cabi.SEARCH_TMA
    elif variant == "ldg":
        f |= _cabi.SEARCH_LDG
    return f


def check_parity(store, db, q, k, metric, flags, score_rtol=1e-5):
    ref_ids, ref_scores, S = vs_oracle.search(q, db, k, metric)
    ids, scores = store.search_arrays(q, k, flags=flags)
    kk = ref_ids.shape[1]
    assert (ids[:, kk:] == -1).all()
    rep = compare.compare_topk(ref_ids, ref_scores, ids[:, :kk], scores[:, :kk], S,
                               tie_rtol=1e-6, score_atol=1e-5, score_rtol=score_rtol)
    assert rep.ok, f"{rep}"
    return rep


# --------------------------------------------------------------------------- golden
@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("variant", ["ldg", "tma"])
def test_golden_fixtures(make_store, path, metric, variant):
    g = np.load(path)
    db, q, k = g["db"], g["q"], int(g["k"])
    st = make_store(db.shape[1], metric)
    st.add_vectors(db, [{} for _ in range(db.shape[0])])
    ids, scores = st.search_arrays(q, k, flags=flags_of("scan_fp32", variant))
    S = vs_oracle.score_matrix(q, db, metric)
    rep = compare.compare_topk(g[f"ids_{metric}"], g[f"scores_{metric}"], ids, scores, S)
    assert rep.ok, f"{rep}"


# --------------------------------------------------------------------------- shapes
SHAPES = [
    # N, D, B, k
    (1, 8, 1, 10),            # single row, k > N
    (37, 5, 3, 4),            # D not a multiple of 4, tiny
    (1000, 32, 1, 10),
    (4097, 100, 2, 10),       # ragged N and D
    (10007, 128, 5, 10),
    (30011, 384, 8, 10),
    (20000, 768, 1, 10),
    (9001, 1536, 4, 100),
    (5000, 64, 13, 1),        # B > 8 -> several passes, odd tail
    (3000, 96, 1, 1000),      # large k
    (600, 48, 2, 605),        # k > N
]


@pytest.mark.parametrize("shape", SHAPES, ids=[f"N{n}_D{d}_B{b}_k{k}" for n, d, b, k in SHAPES])
@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("variant", ["ldg", "tma"])
def test_scan_fp32_matches_oracle(make_store, shape, metric, variant):
    n, d, b, k = shape
    db = datasets.make_db(n, d, "normal")
    q = datasets.make_queries(b, d, "normal")
    st = make_store(d, metric)
    st.add_vectors(db, [{} for _ in range(n)])
    check_parity(st, db, q, k, metric, flags_of("scan_fp32", variant))


@pytest.mark.parametrize("metric", METRICS)
def test_uniform_distribution_near_ties(make_store, metric):
    """np.random.rand-style data (the reference's own test distribution): all-positive
    vectors, cosine scores cluster around 0.75 -> many near-ties."""
    n, d, b, k = 50000, 384, 8, 10
    db = datasets.make_db(n, d, "uniform")
    q = datasets.make_queries(b, d, "uniform")
    st = make_store(d, metric)
    st.add_vectors(db, [{} for _ in range(n)])
    check_parity(st, db, q, k, metric, flags_of("scan_fp32"))


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("variant", ["ldg", "tma"])
def test_adversarial_rows(make_store, metric, variant):
    """Exact duplicates (ties -> lower id first), zero row, sub-clamp norm, scaled copy,
    queries equal to stored rows (self-match pin of tests/test_integration.py:133-136)."""
    db, q = datasets.make_adversarial(1025, 64)
    st = make_store(64, metric)
    st.add_vectors(db, [{} for _ in range(db.shape[0])])
    ref_ids, ref_scores, S = vs_oracle.search(q, db, 12, metric)
    ids, scores = st.search_arrays(q, 12, flags=flags_of("scan_fp32", variant))
    rep = compare.compare_topk(ref_ids, ref_scores, ids, scores, S)
    assert rep.ok, f"{rep}"
    assert np.isfinite(scores).all()
    if metric == "cosine":
        # rows 2, 5 and 1024 are identical: exact tie, ascending id order
        assert ids[0, :3].tolist() == [2, 5, 1024]
        assert scores[0, 0] > 0.999


def test_all_rows_identical_mass_ties(make_store):
    """Every score equal: result must be ids 0..k-1 (stable order), via the merge kernel's
    overflow path."""
    n, d = 6000, 32
    db = np.tile(datasets.make_db(1, d), (n, 1))
    st = make_store(d, "cosine")
    st.add_vectors(db, [{} for _ in range(n)])
    for variant in ("ldg", "tma"):
        ids, scores = st.search_arrays(db[:1], 10, flags=flags_of("scan_fp32", variant))
        assert ids[0].tolist() == list(range(10))
        assert np.allclose(scores, 1.0, atol=1e-6)


def test_incremental_appends_equal_bulk(make_store):
    """Appending in ragged pieces (K1, in-place arena growth) gives the same store."""
    n, d = 20000, 96
    db = datasets.make_db(n, d)
    q = datasets.make_queries(4, d)
    a = make_store(d)
    a.add_vectors(db, [{} for _ in range(n)])
    b = make_store(d)
    cuts = [0, 1, 8, 1000, 1001, 7777, 19999, n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        r = b.add_vectors(db[lo:hi], [{} for _ in range(hi - lo)])
        assert r == {"vectors_added": hi - lo, "total_vectors": hi}
    ia, sa = a.search_arrays(q, 10)
    ib, sb = b.search_arrays(q, 10)
    np.testing.assert_array_equal(ia, ib)
    np.testing.assert_array_equal(sa, sb)
    np.testing.assert_array_equal(b._read_rows(0, n), db)
    # device-resident input (torch CUDA tensor stands in for mx.array)
    c = make_store(d)
    c.add_vectors(torch.from_numpy(db).cuda(), [{} for _ in range(n)])
    ic, sc = c.search_arrays(q, 10)
    np.testing.assert_array_equal(ia, ic)
    np.testing.assert_array_equal(sa, sc)


def test_query_while_appending_sees_prefix(make_store):
    """Searches see exactly the rows appended before them."""
    d = 64
    db = datasets.make_db(9000, d)
    q = datasets.make_queries(2, d)
    st = make_store(d)
    for hi in (1000, 4000, 9000):
        lo = st.get_stats()["vector_count"]
        st.add_vectors(db[lo:hi], [{} for _ in range(hi - lo)])
        check_parity(st, db[:hi], q, 10, "cosine", flags_of("scan_fp32"))


# --------------------------------------------------------------------------- bf16 database
@pytest.mark.parametrize("metric", METRICS)
def test_bf16_database_recall_with_fp32_rescoring(make_store, metric):
    """Config C style: scan the bf16 shadow, rescore candidates in fp32 (K5).  Reported as
    recall@k against the oracle ids; rescored scores are exact fp32."""
    n, d, b, k = 60000, 256, 8, 100
    db = datasets.make_db(n, d)
    q = datasets.make_queries(b, d)
    st = make_store(d, metric)
    st.add_vectors(db, [{} for _ in range(n)])
    ref_ids, ref_scores, S = vs_oracle.search(q, db, k, metric)
    ids, scores = st.search_arrays(q, k, flags=flags_of("scan_bf16"))
    rec = compare.recall_at_k(ref_ids, ids)
    assert rec >= 0.99, rec
    got = np.take_along_axis(S, ids.astype(np.int64), axis=1)
    np.testing.assert_allclose(scores, got, atol=1e-5, rtol=1e-5)


def test_rescore_is_bit_identical_to_scan(make_store):
    """K5 uses the scan's accumulation order: same ids in -> bit-identical scores out."""
    n, d = 30000, 384
    db = datasets.make_db(n, d)
    q = datasets.make_queries(4, d)
    from b200vs import _cabi
    import ctypes as C
    for metric in METRICS:
        st = make_store(d, metric)
        st.add_vectors(db, [{} for _ in range(n)])
        ids, scores = st.search_arrays(q, 16, flags=flags_of("scan_fp32"))
        dq = torch.from_numpy(q).cuda()
        dc = torch.from_numpy(ids[:, ::-1].copy()).cuda()     # shuffled candidate order
        os_ = torch.empty((4, 16), dtype=torch.float32, device="cuda")
        oi = torch.empty((4, 16), dtype=torch.int32, device="cuda")
        _cabi.check(_cabi.lib().vs_rescore(st._handle, C.c_void_p(dq.data_ptr()), 4,
                                           C.c_void_p(dc.data_ptr()), 16, 16,
                                           C.c_void_p(os_.data_ptr()), C.c_void_p(oi.data_ptr()),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(oi.cpu().numpy(), ids)
        np.testing.assert_array_equal(os_.cpu().numpy(), scores)


# --------------------------------------------------------------------------- merge (K4)
def test_merge_of_shard_results_equals_global(make_store):
    """Row-shard a database over G stores, merge local top-k with vs_merge: identical to the
    unsharded search (what the multi-GPU path does after the all-gather)."""
    from b200vs import _cabi
    import ctypes as C
    n, d, b, k, G = 40000, 128, 6, 10, 4
    db = datasets.make_db(n, d)
    q = datasets.make_queries(b, d)
    for metric in METRICS:
        whole = make_store(d, metric)
        whole.add_vectors(db, [{} for _ in range(n)])
        wi, ws = whole.search_arrays(q, k)
        cs, ci = [], []
        bounds = np.linspace(0, n, G + 1).astype(int)
        for g in range(G):
            sh = make_store(d, metric)
            lo, hi = bounds[g], bounds[g + 1]
            sh.add_vectors(db[lo:hi], [{} for _ in range(hi - lo)])
            li, ls = sh.search_arrays(q, k)
            ci.append(li + lo)
            cs.append(ls)
        dcs = torch.from_numpy(np.stack(cs)).cuda()
        dci = torch.from_numpy(np.stack(ci).astype(np.int32)).cuda()
        os_ = torch.empty((b, k), dtype=torch.float32, device="cuda")
        oi = torch.empty((b, k), dtype=torch.int32, device="cuda")
        _cabi.check(_cabi.lib().vs_merge(0, _cabi.METRICS[metric], C.c_void_p(dcs.data_ptr()),
                                         C.c_void_p(dci.data_ptr()), G, b, k, 0,
                                         C.c_void_p(os_.data_ptr()), C.c_void_p(oi.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(oi.cpu().numpy(), wi)
        np.testing.assert_array_equal(os_.cpu().numpy(), ws)


# --------------------------------------------------------------------------- larger property checks
def test_self_match_and_sortedness_at_scale(make_store):
    """Size-independent properties at a size the oracle cannot sort quickly: every stored
    row queried against the store returns itself first with similarity ~1, results are
    sorted, ids unique."""
    n, d = 400000, 256
    g = torch.Generator(device="cuda").manual_seed(7)
    db = torch.randn((n, d), generator=g, device="cuda")
    st = make_store(d)
    st.add_vectors(db, [])
    pick = np.array([0, 1, 12345, 199999, n - 1])
    q = db[torch.from_numpy(pick).cuda()].cpu().numpy()
    for variant in ("ldg", "tma"):
        ids, scores = st.search_arrays(q, 10, flags=flags_of("scan_fp32", variant))
        assert ids[:, 0].tolist() == pick.tolist()
        assert (scores[:, 0] > 0.9999).all()
        assert (np.diff(scores, axis=1) <= 0).all()
        assert all(len(set(r.tolist())) == 10 for r in ids)
    # sub-sampled oracle check: rows of the top-10 rescored on the CPU
    rows = st._read_rows(0, n)
    S = vs_oracle.cosine_similarity_batch(q, rows)
    ref = np.argsort(-S, axis=1, kind="stable")[:, :10]
    rep = compare.compare_topk(ref, np.take_along_axis(S, ref, 1), ids, scores, S)
    assert rep.ok, f"{rep}"
