"""CPU, world_size 2 and 3 over gloo: the row-sharded store's host logic (batch splitting,
global ids, all-gather layout, merge order) gives exactly the unsharded oracle result.
The CUDA shard is replaced by tests/_oracle_shard.py here; the same logic runs with the real
kernels in tests/test_sharded_gpu.py."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
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


def _worker(rank, world, port, metric, out_dir):
    for p in (str(ROOT), str(ROOT / "mlx-vector-db_b200"), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200vs.sharded import ShardedVectorStore
    from _oracle_shard import OracleShard
    from oracle import datasets
    d, k = 24, 7
    db = datasets.make_db(1003, d)
    db[500] = db[3]          # exact tie across shards -> lower global id first
    q = datasets.make_queries(5, d)
    q[0] = db[3]
    st = ShardedVectorStore(d, metric, device=torch.device("cpu"), shard_factory=OracleShard)
    # ragged streaming appends, each split over the ranks
    cuts = [0, 1, 10, 333, 334, 1003]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        r = st.add_vectors(db[lo:hi])
        assert r == {"vectors_added": hi - lo, "total_vectors": hi}
    ids, scores = st.search(q, k)
    ids2, _ = st.search(q, 2000)          # k > N
    assert ids2.shape == (5, 1003)
    np.save(os.path.join(out_dir, f"ids_{rank}.npy"), ids.numpy())
    np.save(os.path.join(out_dir, f"scores_{rank}.npy"), scores.numpy())
    np.save(os.path.join(out_dir, f"count_{rank}.npy"), np.array([st.shard.count()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_sharded_equals_unsharded_oracle(tmp_path, world, metric):
    from oracle import datasets, vs_oracle
    port = _free_port()
    mp.spawn(_worker, args=(world, port, metric, str(tmp_path)), nprocs=world, join=True)
    d, k = 24, 7
    db = datasets.make_db(1003, d)
    db[500] = db[3]
    q = datasets.make_queries(5, d)
    q[0] = db[3]
    ref_ids, ref_scores, _ = vs_oracle.search(q, db, k, metric)
    counts = 0
    for r in range(world):
        ids = np.load(tmp_path / f"ids_{r}.npy")
        scores = np.load(tmp_path / f"scores_{r}.npy")
        np.testing.assert_array_equal(ids, ref_ids)              # every rank, global ids
        np.testing.assert_allclose(scores, ref_scores, atol=2e-6)
        counts += int(np.load(tmp_path / f"count_{r}.npy")[0])
    assert counts == 1003
    if metric == "cosine":
        assert ref_ids[0, :2].tolist() == [3, 500]


def test_split_batch_is_balanced_and_contiguous():
    sys.path.insert(0, str(ROOT / "mlx-vector-db_b200"))
    from b200vs.sharded import split_batch
    for m in (0, 1, 7, 8, 1001):
        for world in (1, 2, 3, 8):
            edges = [split_batch(m, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == m
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
