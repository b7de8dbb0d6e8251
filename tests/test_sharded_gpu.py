"""GPU: the row-sharded path with the real kernels.  One GPU emulates G ranks as G native
shards in one process: each gets its `split_batch` slice of every append with global ids
(vs_append_ids), searches locally, the packs are concatenated exactly as the NCCL
all-gather lays them out, and K4 (vs_merge with a group stride) produces the global top-k.
Result must equal the unsharded oracle (ids exact outside 1e-6 ties, scores within 1e-5)."""
import numpy as np
import pytest
import torch

from oracle import compare, datasets, vs_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric", ["cosine", "euclidean", "dot_product"])
@pytest.mark.parametrize("G", [2, 8])
def test_emulated_ranks_match_oracle(native_lib, metric, G):
    from b200vs.sharded import NativeShard, split_batch
    n, d, B, k = 30011, 128, 6, 10
    db = datasets.make_db(n, d)
    db[20000] = db[5]                     # tie across shards
    q = datasets.make_queries(B, d)
    q[0] = db[5]
    dev = torch.device("cuda", 0)
    shards = [NativeShard(d, metric, dev, True, n, "auto") for _ in range(G)]
    try:
        total = 0
        cuts = [0, 3, 1000, 17777, n]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            m = hi - lo
            for r, sh in enumerate(shards):
                a, b = split_batch(m, G, r)
                if b > a:
                    sh.append(db[lo + a:lo + b], total + a)
            total += m
        assert sum(sh.count() for sh in shards) == n
        qd = shards[0].prepare_queries(q)
        packs = []
        for sh in shards:
            p = sh.new_pack(B, k)
            sh.search_into(qd, k, p)
            packs.append(p)
        gathered = torch.stack(packs).contiguous()           # (G, 2, B, k) like all_gather
        ids, scores = shards[0].merge(gathered, G, B, k)
        torch.cuda.synchronize()
        ref_ids, ref_scores, S = vs_oracle.search(q, db, k, metric)
        rep = compare.compare_topk(ref_ids, ref_scores, ids.cpu().numpy(), scores.cpu().numpy(), S)
        assert rep.ok, f"{rep}"
        if metric == "cosine":
            assert ids[0, :2].tolist() == [5, 20000]
    finally:
        for sh in shards:
            sh.close()


def test_sharded_store_world1_equals_plain_store(make_store):
    from b200vs.sharded import ShardedVectorStore
    n, d = 20000, 96
    db = datasets.make_db(n, d)
    q = datasets.make_queries(9, d)
    st = ShardedVectorStore(d, "cosine", device=torch.device("cuda", 0), max_vectors_per_shard=n)
    st.add_vectors(db[:7000])
    st.add_vectors(torch.from_numpy(db[7000:]).cuda())
    ids, scores = st.search(q, 10)
    plain = make_store(d)
    plain.add_vectors(db, [{} for _ in range(n)])
    pi, ps = plain.search_arrays(q, 10)
    np.testing.assert_array_equal(ids.cpu().numpy(), pi)
    np.testing.assert_array_equal(scores.cpu().numpy(), ps)
    st.close()
