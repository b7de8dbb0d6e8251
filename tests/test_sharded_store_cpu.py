"""CPU, world_size 2 over gloo: `ShardedMLXVectorStore` -- the reference's store surface
(service/optimized_vector_store.py:96-145) over a row-sharded database -- returns on every rank
what the unsharded oracle store returns: global ids, scores, metadata, filters, k > N, negative
k, persistence (per-rank logs replayed at the same world size; `optimize()` -> one vectors.npz in
the reference's format that the ORACLE store loads).  The CUDA shard is replaced by
tests/_oracle_shard.py; tests/test_sharded_store_gpu.py runs the same surface on the kernels."""
import json
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _data():
    from oracle import datasets
    d = 16
    db = datasets.make_db(403, d)
    db[300] = db[7]
    meta = [{"id": i, "category": "A" if i % 3 == 0 else "B", "lang": "en" if i % 2 else "de", "tags": [i % 5]}
            for i in range(403)]
    q = datasets.make_queries(4, d)
    q[0] = db[7]
    return d, db, meta, q


def _worker(rank, world, port, out_dir, store_dir):
    for p in (str(ROOT), str(ROOT / "mlx-vector-db_b200"), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200vs.sharded_store import ShardedMLXVectorStore
    from b200vs.store import MLXVectorStoreConfig
    from _oracle_shard import OracleShard
    d, db, meta, q = _data()
    cfg = MLXVectorStoreConfig(dimension=d, metric="cosine", persist=True)

    def make():
        return ShardedMLXVectorStore(store_dir, cfg, device=torch.device("cpu"), shard_factory=OracleShard)

    st = make()
    out = {}
    cuts = [0, 1, 50, 51, 403]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        r = st.add_vectors(db[lo:hi], meta[lo:hi])
        assert r == {"vectors_added": hi - lo, "total_vectors": hi}
    out["q"] = st.query(q[0], k=5)
    out["batch"] = st.batch_query(q, k=6)
    out["flt"] = st.query(q[1], k=10, filter_metadata={"category": "A", "lang": "en"})
    out["flt_generic"] = st.query(q[1], k=4, filter_metadata={"tags": [2]})
    out["flt_none"] = st.query(q[1], k=4, filter_metadata={"category": "Z"})
    out["k_big"] = st.query(q[2], k=1000)
    out["k_neg"] = st.query(q[2], k=-400)
    out["stats"] = {k_: v for k_, v in st.get_stats().items() if k_ not in ("memory_usage_mb", "local_vectors")}
    assert st.health_check() == {"healthy": True, "issues": []}
    st.close()
    dist.barrier()
    st = make()                                    # replay the per-rank logs
    out["reloaded"] = st.query(q[3], k=5, filter_metadata={"lang": "de"})
    st.optimize()                                  # -> reference format
    extra = np.random.default_rng(11).standard_normal((3, d)).astype(np.float32)
    st.add_vectors(extra, [{"id": 1000 + i} for i in range(3)])     # appends continue after the snapshot
    st.close()
    dist.barrier()
    st = make()                                    # snapshot + one logged batch
    out["after_optimize"] = st.query(db[1], k=3)
    out["count"] = st.get_stats()["vector_count"]
    st.close()
    with open(os.path.join(out_dir, f"out_{rank}.json"), "w") as f:
        json.dump(out, f)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_store_surface_equals_oracle_store(tmp_path):
    from oracle import vs_oracle
    world = 2
    store_dir = tmp_path / "store"
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), str(store_dir)), nprocs=world, join=True)
    d, db, meta, q = _data()
    ora = vs_oracle.OracleVectorStore(None, dimension=d, metric="cosine")
    ora.add_vectors(db, meta)

    def norm(t):
        return [list(t[0]), [round(float(x), 5) for x in t[1]], list(t[2])]

    outs = [json.load(open(tmp_path / f"out_{r}.json")) for r in range(world)]
    assert outs[0] == outs[1]                       # SPMD: identical on every rank
    o = outs[0]
    assert norm(o["q"]) == norm(ora.query(q[0], k=5))
    assert o["q"][0][:2] == [7, 300]                # tie across shards -> lower global id first
    for b in range(4):
        assert norm(o["batch"][b]) == norm(ora.query(q[b], k=6))
    assert norm(o["flt"]) == norm(ora.query(q[1], k=10, filter_metadata={"category": "A", "lang": "en"}))
    assert norm(o["flt_generic"]) == norm(ora.query(q[1], k=4, filter_metadata={"tags": [2]}))
    assert o["flt_none"] == [[], [], []]
    assert norm(o["k_big"]) == norm(ora.query(q[2], k=1000)) and len(o["k_big"][0]) == 403
    assert norm(o["k_neg"]) == norm(ora.query(q[2], k=-400)) and len(o["k_neg"][0]) == 3
    assert o["stats"] == {"vector_count": 403, "dimension": d, "metric": "cosine", "index_type": "flat", "shards": 2}
    assert norm(o["reloaded"]) == norm(ora.query(q[3], k=5, filter_metadata={"lang": "de"}))
    # the snapshot is the reference's format: the oracle store (a restatement of the reference's
    # _load_store, :225-239) loads it
    snap = np.load(store_dir / "vectors.npz")
    assert list(snap.keys()) == ["vectors"]
    np.testing.assert_array_equal(snap["vectors"], db)
    ora.add_vectors(np.random.default_rng(11).standard_normal((3, d)).astype(np.float32),
                    [{"id": 1000 + i} for i in range(3)])
    assert o["count"] == 406
    assert norm(o["after_optimize"]) == norm(ora.query(db[1], k=3))
