"""GPU: BASELINE.json's configurations at their FULL sizes, pinned to the oracle.

The oracle cannot fully sort (B, N) scores for B = 1024 at N = 10^7 in test time, so:
  * config A (100K x 384, batch 1024) is checked in full: every query, `vs_oracle.search`;
  * 1M x 768, 1M x 1536, 5M x 384 and 10M x 128 run the batch of 1024 through the engine and check
    the first 8 queries against a chunk-free exact oracle: the oracle's full score row
    (`cosine_similarity_batch`, reference op order) -> argpartition to the best 4k -> stable
    (score desc, id asc) order of those.  That equals `argsort(-s, stable)[:k]`
    (service/optimized_vector_store.py:176-184) because the k-th score is far inside the 4k best.
    The same 8 queries are also run one at a time (batch 1: AUTO and the fp32 scan).
Contract (BASELINE.json): ids exact outside 1e-6 ties, scores within 1e-5 -- oracle/compare.py."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import compare, vs_oracle

pytestmark = pytest.mark.gpu


def _host_db(n, d, seed=1234, blocks=8):
    return [np.random.default_rng(seed + b).standard_normal((n // blocks, d), dtype=np.float32) for b in range(blocks)]


def _oracle_topk(q, db, k):
    """Exact `argsort(-s, stable)[:k]` per query without sorting N values."""
    S = vs_oracle.cosine_similarity_batch(q, db)
    ids = np.empty((q.shape[0], k), np.int32)
    for b in range(q.shape[0]):
        part = np.argpartition(-S[b], 4 * k)[:4 * k]
        ids[b] = part[np.lexsort((part, -S[b][part]))][:k]
    return ids, np.take_along_axis(S, ids.astype(np.int64), axis=1), S


def test_config_a_100k_x_384_batch_1024_full_oracle(make_store):
    from b200vs import _cabi
    n, d, B, k = 100_000, 384, 1024, 10
    db = np.concatenate(_host_db(n, d))
    q = np.random.default_rng(4321).standard_normal((B, d), dtype=np.float32)
    st = make_store(d, "cosine")
    st.add_vectors(db, [])
    ref_ids, ref_scores, S = vs_oracle.search(q, db, k, "cosine")
    ids, scores = st.search_arrays(q, k)                       # AUTO -> K3
    rep = compare.compare_topk(ref_ids, ref_scores, ids, scores, S)
    assert rep.ok, f"{rep}"
    assert int(_cabi.lib().vs_fallback_count(st._handle)) <= B // 16


@pytest.mark.parametrize("n,d", [(1_000_000, 768), (1_000_000, 1536), (5_000_000, 384), (10_000_000, 128)],
                         ids=["1Mx768", "1Mx1536", "5Mx384", "10Mx128"])
def test_full_size_against_oracle(native_lib, n, d):
    from b200vs import _cabi
    from b200vs.sharded import ShardedVectorStore
    dev = torch.device("cuda", 0)
    free, _ = torch.cuda.mem_get_info()
    if free < (n * d * 6 + (4 << 30)):
        pytest.skip("not enough free device memory for the full-size store")
    B, k, nq = 1024, 10, 8
    blocks = _host_db(n, d)
    st = ShardedVectorStore(d, "cosine", device=dev, max_vectors_per_shard=n + 16)
    try:
        for blk in blocks:
            st.add_vectors(torch.from_numpy(blk).to(dev))
        db = np.concatenate(blocks)
        del blocks
        q = np.random.default_rng(4321).standard_normal((B, d), dtype=np.float32)
        ref_ids, ref_scores, S = _oracle_topk(q[:nq], db, k)
        del db
        qd = torch.from_numpy(q).to(dev)
        ids, scores = st.search(qd, k)                         # AUTO: K3, certified
        rep = compare.compare_topk(ref_ids, ref_scores, ids[:nq].cpu().numpy(), scores[:nq].cpu().numpy(), S)
        assert rep.ok, f"batch {B}: {rep}"
        assert int(_cabi.lib().vs_fallback_count(st.shard.handle)) <= B // 16
        for mode in ("auto", "scan_fp32"):                     # single queries
            st.shard.flags = _cabi.SEARCH_MODES[mode]
            got_i = np.empty((nq, k), np.int32)
            got_s = np.empty((nq, k), np.float32)
            for b in range(nq):
                i1, s1 = st.search(qd[b:b + 1].contiguous(), k)
                got_i[b], got_s[b] = i1[0].cpu().numpy(), s1[0].cpu().numpy()
            rep = compare.compare_topk(ref_ids, ref_scores, got_i, got_s, S)
            assert rep.ok, f"batch 1 ({mode}): {rep}"
    finally:
        st.close()


@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_two_shards_of_70k_rows_take_the_gemm_path(native_lib, metric):
    """Sharded K3: each emulated rank holds >= 65 536 rows, so its local search is the tensor-core
    path with local -> global id mapping, then the gathered packs are merged (K4)."""
    from b200vs import _cabi
    from b200vs.sharded import NativeShard, split_batch
    from oracle import datasets
    G, n, d, B, k = 2, 150_000, 128, 96, 10
    db = datasets.make_db(n, d)
    q = datasets.make_queries(B, d)
    dev = torch.device("cuda", 0)
    shards = [NativeShard(d, metric, dev, True, n, "gemm") for _ in range(G)]
    try:
        total = 0
        for lo, hi in ((0, 50_000), (50_000, 50_001), (50_001, n)):
            m = hi - lo
            for r, sh in enumerate(shards):
                a, b = split_batch(m, G, r)
                if b > a:
                    sh.append(db[lo + a:lo + b], total + a)
            total += m
        assert min(sh.count() for sh in shards) >= 70_000
        qd = shards[0].prepare_queries(q)
        packs = []
        for sh in shards:
            p = sh.new_pack(B, k)
            t = sh.submit_into(qd, k, p)
            sh.complete(t)
            packs.append(p)
        gathered = torch.stack(packs).contiguous()
        ids, scores = shards[0].merge(gathered, G, B, k)
        torch.cuda.synchronize()
        ref_ids, ref_scores, S = vs_oracle.search(q, db, k, metric)
        rep = compare.compare_topk(ref_ids, ref_scores, ids.cpu().numpy(), scores.cpu().numpy(), S)
        assert rep.ok, f"{rep}"
        assert sum(int(_cabi.lib().vs_fallback_count(sh.handle)) for sh in shards) <= B // 4
    finally:
        for sh in shards:
            sh.close()


@pytest.mark.parametrize("d", [1536, 4096])
def test_tensor_core_accumulation_error_is_inside_the_certification_slack(make_store, d):
    """The certification bound charges the tensor core's fp32 accumulation `slack * (1 + ||u|| ||v||)`
    with slack = 4 D 2^-24 + 1e-6 (gemm_topk.cu).  Measure it: K3's raw scores against float64 dot
    products of the SAME bf16 operands (dot_product stores raw rows: the operands are exactly
    torch's bf16 roundings, products of two bf16 are exact in fp32, so the only error is the
    accumulation)."""
    from b200vs import _cabi
    n, B = 4096, 256
    rng = np.random.default_rng(3)
    db = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((B, d), dtype=np.float32)
    st = make_store(d, "dot_product")
    st.add_vectors(db, [])
    qd = torch.from_numpy(q).cuda()
    out = torch.empty((B, n), dtype=torch.float32, device="cuda")
    _cabi.check(_cabi.lib().vs_debug_gemm_scores(st._handle, C.c_void_p(qd.data_ptr()), B, C.c_void_p(out.data_ptr()),
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    qb = qd.to(torch.bfloat16).double()
    xb = torch.from_numpy(db).cuda().to(torch.bfloat16).double()
    exact = qb @ xb.T
    err = (out.double() - exact).abs()
    slack = 4.0 * d * 2.0 ** -24 + 1e-6
    bound = slack * (1.0 + qb.norm(dim=1, keepdim=True) * xb.norm(dim=1, keepdim=True).T)
    worst = float((err / bound).max())
    assert worst <= 1.0, f"tensor-core accumulation error reaches {worst:.3f} of the certified slack at D={d}"
