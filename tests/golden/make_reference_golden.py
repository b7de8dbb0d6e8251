"""Generate tests/golden/ref_*.npz / ref_*.json by EXECUTING THE REFERENCE'S OWN MODULES.

    python tests/golden/make_reference_golden.py          # needs /root/reference (this container only)

`service/optimized_vector_store.py` and `performance/mlx_optimized.py` are imported unmodified
from /root/reference; their only missing dependency, `mlx.core`, is satisfied by the NumPy
stand-in under tests/golden/mlx_standin/ (read its header for exactly what that does and does
not pin), and `hnswlib` by an import stub (the exact path never calls it).  Nothing of the
reference is copied: this script only calls its public functions and records inputs + outputs.

The fixtures therefore pin, against the reference's own code: op order and clamps of the
cosine / euclidean / dot scoring functions, `argsort(-s)[:k]` slicing, k > N and k <= 0
behaviour, the (1, D) / (D,) query handling, add_vectors id assignment over several batches,
the metadata filter's AND semantics and local->global id mapping, empty results, and the
on-disk layout (`vectors.npz` key `vectors` + `metadata.jsonl`).  tests/test_reference_golden.py
checks the oracle against them on CPU; tests/test_reference_golden_gpu.py checks the CUDA
engine against them on the B200.
"""
import json
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REFERENCE = Path("/root/reference")
if not REFERENCE.exists():
    sys.exit("the reference tree is not mounted; fixtures can only be regenerated in the build container")
# the stand-in and the reference come FIRST so `service` / `performance` resolve to the reference,
# not to this repo's same-named import shims
sys.path.insert(0, str(HERE / "mlx_standin"))
sys.path.insert(0, str(REFERENCE))

import mlx.core as mx  # noqa: E402  (the stand-in)
import performance.mlx_optimized as ref_ops  # noqa: E402
import service.optimized_vector_store as ref_store  # noqa: E402

assert Path(ref_ops.__file__).is_relative_to(REFERENCE) and Path(ref_store.__file__).is_relative_to(REFERENCE)


def rng_db(seed, n, d, dist):
    r = np.random.default_rng(seed)
    x = r.standard_normal((n, d)) if dist == "normal" else r.random((n, d))
    return x.astype(np.float32)


def adversarial(seed, n, d):
    """duplicated rows (exact ties), a zero row (1e-8 clamp), scaled copies (equal cosine),
    queries equal to stored rows (self-match pin)."""
    db = rng_db(seed, n, d, "normal")
    db[7] = 0.0
    db[20] = db[3]
    db[21] = db[3]
    db[40] = 2.5 * db[11]
    db[n - 1] = db[0]
    q = np.stack([db[3], db[11], db[n - 2], rng_db(seed + 1, 1, d, "normal")[0], np.zeros(d, np.float32)])
    return db, q.astype(np.float32)


def ops_case(name, db, q, k):
    """The functions of performance/mlx_optimized.py on one (db, q, k)."""
    blob = {"db": db, "q": q, "k": np.int32(k)}
    dbm = mx.array(db)
    # single-query functions, one query at a time (that is all they accept)
    cs, eu, dp, si, ss = [], [], [], [], []
    for row in q:
        cs.append(np.asarray(ref_ops.compute_cosine_similarity_single(mx.array(row), dbm)))
        eu.append(np.asarray(ref_ops.compute_euclidean_distance(mx.array(row), dbm)))
        dp.append(np.asarray(ref_ops.compute_dot_product(mx.array(row), dbm)))
        i_, s_ = ref_ops.optimized_similarity_search(mx.array(row), dbm, k)
        si.append(np.asarray(i_, dtype=np.int64))
        ss.append(np.asarray(s_, dtype=np.float32))
    blob["cosine_single"] = np.stack(cs)
    blob["euclidean"] = np.stack(eu)
    blob["dot"] = np.stack(dp)
    blob["search_ids"] = np.stack(si)
    blob["search_scores"] = np.stack(ss)
    blob["cosine_batch"] = np.asarray(ref_ops.compute_cosine_similarity_batch(mx.array(q), dbm))
    bi, bs = ref_ops.optimized_batch_similarity_search(mx.array(q), dbm, k)
    blob["batch_ids"] = np.asarray(bi, dtype=np.int64)
    blob["batch_scores"] = np.asarray(bs, dtype=np.float32)
    blob["normalized"] = np.asarray(ref_ops.normalize_vectors(dbm))
    blob["topk_of_first"] = np.asarray(ref_ops.fast_top_k_indices(mx.array(cs[0]), k), dtype=np.int64)
    half = db.shape[0] // 2
    blob["added_normalized"] = np.asarray(
        ref_ops.optimized_vector_addition(mx.array(db[:half]), mx.array(db[half:]), normalize=True))
    np.savez_compressed(HERE / f"ref_ops_{name}.npz", **blob)
    print("ops  ", name, db.shape, q.shape, k)


def store_case(name, db, q, k, metric, batches):
    """MLXVectorStore end to end: several add_vectors calls, plain and filtered queries, reload."""
    tmp = Path(tempfile.mkdtemp(prefix="refgold_"))
    try:
        cfg = ref_store.MLXVectorStoreConfig(dimension=db.shape[1], metric=metric)
        st = ref_store.MLXVectorStore(str(tmp), cfg)
        meta = [{"id": f"doc_{i}", "group": int(i % 5), "parity": "even" if i % 2 == 0 else "odd"}
                for i in range(db.shape[0])]
        record = {"metric": metric, "k": k, "batches": batches, "adds": [], "queries": [], "empty_before_add":
                  [list(x) for x in st.query(q[0], k=k)]}
        lo = 0
        for m in batches:
            record["adds"].append(st.add_vectors(db[lo:lo + m], meta[lo:lo + m]))
            lo += m
        assert lo == db.shape[0]
        filters = [None, {"group": 2}, {"group": 3, "parity": "odd"}, {"group": 99}]
        for qi, row in enumerate(q):
            for f in filters:
                # k > N, k = 0 and a negative k (the reference slices `argsort(...)[:k]`, so -3 drops the last 3)
                for kk in (k, db.shape[0] + 5, 0, -3) if (qi == 0 and f in (None, {"group": 2})) else (k,):
                    ids, scores, metas = st.query(row, k=kk, filter_metadata=f)
                    record["queries"].append({"q": qi, "k": kk, "filter": f, "ids": [int(i) for i in ids],
                                              "scores": [float(np.float32(s)) for s in scores],
                                              "meta_ids": [m_["id"] for m_ in metas]})
        record["stats"] = st.get_stats()
        # a (1, D) query is accepted like a (D,) one (optimized_vector_store.py:32,44)
        ids2, scores2, _ = st.query(q[0][None, :], k=k)
        record["query_2d_first"] = {"ids": [int(i) for i in ids2], "scores": [float(np.float32(s)) for s in scores2]}
        # what the reference left on disk, and that a fresh instance reloads it
        with np.load(tmp / "vectors.npz") as z:
            record["disk_keys"] = list(z.files)
            assert np.array_equal(z["vectors"], db)
        record["disk_metadata_lines"] = sum(1 for _ in open(tmp / "metadata.jsonl"))
        st2 = ref_store.MLXVectorStore(str(tmp), cfg)
        ids3, scores3, _ = st2.query(q[0], k=k)
        record["reloaded_first"] = {"ids": [int(i) for i in ids3], "scores": [float(np.float32(s)) for s in scores3]}
        np.savez_compressed(HERE / f"ref_store_{name}.npz", db=db, q=q)
        (HERE / f"ref_store_{name}.json").write_text(json.dumps(record, indent=1))
        print("store", name, db.shape, q.shape, k, metric, len(record["queries"]), "queries")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def error_cases():
    """Exception types / degenerate returns of the ops module (mlx_optimized.py:38,66-72,97-105,117,135,209,224-228)."""
    db = mx.array(rng_db(5, 6, 4, "normal"))
    out = {}

    def kind(fn):
        try:
            r = fn()
            return {"ok": [list(np.asarray(x).shape) for x in (r if isinstance(r, tuple) else (r,))]}
        except Exception as e:  # noqa: BLE001
            return {"raises": type(e).__name__}

    out["cosine_single_2row_query"] = kind(lambda: ref_ops.compute_cosine_similarity_single(mx.array(np.zeros((2, 4), np.float32)), db))
    out["cosine_batch_1d_query"] = kind(lambda: ref_ops.compute_cosine_similarity_batch(mx.array(np.zeros(4, np.float32)), db))
    out["cosine_batch_dim_mismatch"] = kind(lambda: ref_ops.compute_cosine_similarity_batch(mx.array(np.zeros((2, 5), np.float32)), db))
    out["topk_2d_scores"] = kind(lambda: ref_ops.fast_top_k_indices(mx.array(np.zeros((2, 3), np.float32)), 2))
    out["topk_k0"] = kind(lambda: ref_ops.fast_top_k_indices(mx.array(np.arange(5, dtype=np.float32)), 0))
    out["topk_k_gt_n"] = kind(lambda: ref_ops.fast_top_k_indices(mx.array(np.arange(5, dtype=np.float32)), 9))
    out["normalize_1d"] = kind(lambda: ref_ops.normalize_vectors(mx.array(np.zeros(4, np.float32))))
    out["normalize_empty"] = kind(lambda: ref_ops.normalize_vectors(mx.array(np.zeros((0, 4), np.float32))))
    out["concat_dim_mismatch"] = kind(lambda: ref_ops.fast_vector_concatenation(db, mx.array(np.zeros((2, 5), np.float32))))
    out["concat_empty_left"] = kind(lambda: ref_ops.fast_vector_concatenation(mx.array(np.zeros((0, 4), np.float32)), db))
    out["search_2row_query"] = kind(lambda: ref_ops.optimized_similarity_search(mx.array(np.zeros((2, 4), np.float32)), db, 3))
    out["batch_search_empty_db"] = kind(lambda: ref_ops.optimized_batch_similarity_search(
        mx.array(np.zeros((3, 4), np.float32)), mx.array(np.zeros((0, 4), np.float32)), 3))
    out["batch_search_k_gt_n"] = kind(lambda: ref_ops.optimized_batch_similarity_search(
        mx.array(np.ones((3, 4), np.float32)), db, 50))
    # the store: unsupported metric / jit_compile=False leave no scoring function -> RuntimeError
    # on query (optimized_vector_store.py:153-154, 211-213)
    tmp = Path(tempfile.mkdtemp(prefix="refgold_"))
    try:
        for label, cfg in (("store_dot_product", ref_store.MLXVectorStoreConfig(dimension=4, metric="dot_product")),
                           ("store_no_jit", ref_store.MLXVectorStoreConfig(dimension=4, jit_compile=False))):
            st = ref_store.MLXVectorStore(str(tmp / label), cfg)
            st.add_vectors(np.asarray(db), [{} for _ in range(6)])
            out[label + "_query"] = kind(lambda: st.query(np.zeros(4, np.float32), k=2))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    (HERE / "ref_errors.json").write_text(json.dumps(out, indent=1))
    print("errors", len(out))


if __name__ == "__main__":
    ops_case("normal_400x40", rng_db(11, 400, 40, "normal"), rng_db(12, 5, 40, "normal"), 10)
    ops_case("uniform_250x96", rng_db(13, 250, 96, "uniform"), rng_db(14, 4, 96, "uniform"), 7)
    db, q = adversarial(15, 130, 32)
    ops_case("adversarial_130x32", db, q, 12)
    ops_case("k_gt_n_9x16", rng_db(16, 9, 16, "normal"), rng_db(17, 2, 16, "normal"), 20)

    store_case("cosine_300x48", rng_db(21, 300, 48, "uniform"), rng_db(22, 4, 48, "uniform"), 5, "cosine", [100, 1, 150, 49])
    store_case("euclidean_200x24", rng_db(23, 200, 24, "normal"), rng_db(24, 3, 24, "normal"), 8, "euclidean", [120, 80])
    db, q = adversarial(25, 130, 32)
    store_case("cosine_adversarial_130x32", db, q, 6, "cosine", [65, 65])
    store_case("euclidean_adversarial_130x32", db, q, 6, "euclidean", [130])
    error_cases()
