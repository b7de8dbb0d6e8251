"""CPU: a NumPy model of K3's certification (csrc/gemm_topk.cu `finish_kernel`, the bounds written by
K1 in csrc/store.cu, the query preparation in `prep_queries_bf16_kernel`).

K3 proposes candidates from 16-bit (fp16 for cosine) tensor-core scores and claims the exact fp32 top-k is
inside the candidate set whenever

    exact k-th  >  beta + E,     E = |u|*max|v - v^| + |u - u^|*max|v^| + slack*(1 + |u|*max|v^|)

(u, v = prepared query / row, ^ = rounded to 16 bits, beta = 16-bit score no row outside the candidate
set exceeds).  These tests check the MATHS on the CPU, independent of any kernel: (1) E really bounds
|exact - 16-bit| for every (query, row) pair, including fp32 accumulation in any order; (2) a certificate
is never issued when the candidate set misses a true top-k row -- on i.i.d. data, on the reference tests'
U[0,1) data and on adversarial near-duplicate clusters; (3) the bound is not vacuous (i.i.d. data
certifies).  The GPU suite (tests/test_gemm_gpu.py) checks that the kernels implement this and that
uncertified queries fall back to the exact scan.
"""
import numpy as np
import pytest

SAFETY = np.float32(1.00001)          # the kernels inflate every norm they store by this factor


def prepare_rows(db):
    """K1 (store.cu): norms clamped at 1e-8, fp16 shadow of x/|x|, bounds = max|v - v^|, max|v^|."""
    db = db.astype(np.float32)
    nrm = np.maximum(np.sqrt((db * db).sum(1, dtype=np.float32)), np.float32(1e-8))
    v = (db / nrm[:, None]).astype(np.float32)
    vh = v.astype(np.float16).astype(np.float32)
    e_max = np.sqrt(((v - vh) ** 2).sum(1, dtype=np.float32)).max() * SAFETY
    s_max = np.sqrt((vh * vh).sum(1, dtype=np.float32)).max() * SAFETY
    return nrm, v, vh, np.float32(e_max), np.float32(s_max)


def prepare_queries(q):
    """prep_queries_bf16_kernel (cosine): u = q/|q|, u^ = fp16(u), qerr = |u - u^|, qlen = |u|."""
    q = q.astype(np.float32)
    nrm = np.maximum(np.sqrt((q * q).sum(1, dtype=np.float32)), np.float32(1e-8))
    u = (q / nrm[:, None]).astype(np.float32)
    uh = u.astype(np.float16).astype(np.float32)
    qerr = np.sqrt(((u - uh) ** 2).sum(1, dtype=np.float32)) * SAFETY
    qlen = np.sqrt((u * u).sum(1, dtype=np.float32)) * SAFETY
    return u, uh, qerr.astype(np.float32), qlen.astype(np.float32)


def slack_of(dim):
    """gemm_block: 4 * D * 2^-24 + 1e-6 (the two fp32 accumulations)."""
    return np.float32(4.0 * dim * 5.9604645e-8 + 1e-6)


def model_search(db, q, k, kc, tau_sample=0.1, seed=0):
    """Candidates = rows whose 16-bit score reaches tau (a lower bound from a sample), best kc of them by
    16-bit score; exact fp32 rescoring; certificate.  Returns (exact top-k ids per query from the candidate
    set, certified flags, true exact top-k ids from the full exact scores)."""
    nrm, v, vh, e_max, s_max = prepare_rows(db)
    u, uh, qerr, qlen = prepare_queries(q)
    n, dim = db.shape
    S16 = (uh @ vh.T).astype(np.float32)                       # fp32 accumulation of 16-bit operands
    exact = ((u @ db.T.astype(np.float32)) / nrm[None, :]).astype(np.float32)   # K2: dot(u, x) / |x|
    rng = np.random.default_rng(seed)
    sample = rng.choice(n, size=min(n, max(kc, int(tau_sample * n))), replace=False)
    out_ids, certified, truth = [], [], []
    slack = slack_of(dim)
    for b in range(q.shape[0]):
        # kc sampled rows reach tau; a store with fewer rows has no threshold (every row is a candidate)
        tau = np.sort(S16[b, sample])[-kc] if len(sample) >= kc else np.float32(-np.inf)
        passed = np.flatnonzero(S16[b] >= tau)
        order = passed[np.argsort(-S16[b, passed], kind="stable")][:kc]
        full = len(order) == kc
        beta = S16[b, order[-1]] if full else tau
        ex = exact[b, order]
        rank = order[np.lexsort((order, -ex))]                 # exact score desc, id asc
        kk = min(k, n)
        top = rank[:kk]
        E = qlen[b] * e_max + qerr[b] * s_max + slack * (np.float32(1) + qlen[b] * s_max)
        ok = len(top) == kk and exact[b, top[-1]] > beta + E
        if not full and n <= kc:
            ok = True
        out_ids.append(top)
        certified.append(bool(ok))
        truth.append(np.lexsort((np.arange(n), -exact[b]))[:kk])
    return out_ids, certified, truth


@pytest.mark.parametrize("dim", [32, 128, 384, 1536])
@pytest.mark.parametrize("dist", ["normal", "uniform"])
def test_error_bound_holds_for_every_pair(dim, dist):
    rng = np.random.default_rng(dim)
    gen = rng.standard_normal if dist == "normal" else rng.random
    db = gen((3000, dim)).astype(np.float32)
    db[5] = 0.0                                                 # zero row: 1e-8 clamp
    db[6] *= 1e-3
    db[7] *= 1e3
    q = gen((16, dim)).astype(np.float32)
    nrm, v, vh, e_max, s_max = prepare_rows(db)
    u, uh, qerr, qlen = prepare_queries(q)
    true16 = uh.astype(np.float64) @ vh.astype(np.float64).T    # what the tensor core approximates
    true_exact = u.astype(np.float64) @ v.astype(np.float64).T
    E_math = (qlen[:, None] * e_max + qerr[:, None] * s_max).astype(np.float64)
    assert (np.abs(true_exact - true16) <= E_math).all()        # Cauchy-Schwarz part of the bound
    # fp32 accumulation, two different orders (BLAS, and a strictly sequential one), stays within slack
    slack = float(slack_of(dim)) * (1.0 + qlen[:, None].astype(np.float64) * float(s_max))
    blas = (uh @ vh.T).astype(np.float64)
    seq = np.zeros((4, 64), np.float32)
    for j in range(dim):
        seq += np.float32(1) * uh[:4, j:j + 1] * vh[None, :64, j]
    assert (np.abs(blas - true16) <= slack).all()
    assert (np.abs(seq.astype(np.float64) - true16[:4, :64]) <= slack[:4]).all()
    # and the exact path (K2: fp32 dot of u with the raw row, divided by its norm) vs the real value
    k2 = ((u @ db.T) / nrm[None, :]).astype(np.float64)
    assert (np.abs(k2 - true_exact) <= slack + 1e-6).all()


def assert_no_false_certificate(db, q, k, kc):
    ids, cert, truth = model_search(db, q, k, kc)
    for b, ok in enumerate(cert):
        if ok:
            assert list(ids[b]) == list(truth[b]), f"query {b}: certified but wrong"
    return sum(cert)


@pytest.mark.parametrize("shape", [(20000, 128, 10, 32), (8000, 384, 10, 32), (5000, 1536, 100, 200), (3000, 64, 1, 23)])
def test_iid_data_certifies_and_is_right(shape):
    n, dim, k, kc = shape
    rng = np.random.default_rng(n + dim)
    db = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((24, dim)).astype(np.float32)
    n_ok = assert_no_false_certificate(db, q, k, kc)
    assert n_ok >= 20, f"the bound should not be vacuous on i.i.d. data: {n_ok}/24 certified"


def test_reference_test_distribution():
    """U[0,1) rows (the reference tests' np.random.rand): all-positive, scores cluster near 0.75."""
    rng = np.random.default_rng(7)
    db = rng.random((20000, 128)).astype(np.float32)
    q = rng.random((24, 128)).astype(np.float32)
    assert_no_false_certificate(db, q, 10, 32)


def test_near_duplicate_clusters_never_certify_wrongly():
    """Rows that differ by less than the 16-bit rounding: the 16-bit order inside a cluster is arbitrary, so
    with more than kc cluster members above everything else the candidate set can miss true top-k rows --
    the certificate must then be refused."""
    rng = np.random.default_rng(11)
    dim, k, kc = 128, 10, 32
    base = rng.standard_normal((40, dim)).astype(np.float32)
    rows = [rng.standard_normal((4000, dim)).astype(np.float32)]
    for c in base:                                             # 40 clusters of 60 near-copies, spread 2e-5
        rows.append(c[None, :] + np.float32(2e-5) * rng.standard_normal((60, dim)).astype(np.float32))
    db = np.concatenate(rows)
    q = base[:24] + np.float32(1e-3) * rng.standard_normal((24, dim)).astype(np.float32)
    ids, cert, truth = model_search(db, q, k, kc)
    wrong_sets = 0
    for b in range(24):
        if set(ids[b]) != set(truth[b]):
            wrong_sets += 1
            assert not cert[b], f"query {b}: candidate set misses a true top-k row but was certified"
        if cert[b]:
            assert list(ids[b]) == list(truth[b])
    assert wrong_sets > 0, "the adversarial set should defeat the 16-bit prefilter at least once"


def test_exact_duplicates_and_small_stores():
    rng = np.random.default_rng(13)
    db = rng.standard_normal((50, 96)).astype(np.float32)
    db[10:20] = db[3]                                          # exact ties
    q = np.concatenate([db[3:4], rng.standard_normal((3, 96)).astype(np.float32)])
    ids, cert, truth = model_search(db, q, 5, 64)              # n <= kc: every row is a candidate
    for b in range(4):
        assert cert[b] and list(ids[b]) == list(truth[b])
