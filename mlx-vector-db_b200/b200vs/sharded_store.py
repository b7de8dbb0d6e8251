"""The reference's store surface over a row-sharded database: one process per GPU, every rank
calls every method with the same arguments (SPMD) and gets the same return value.

`ShardedMLXVectorStore` has the methods of `MLXVectorStore` (service/optimized_vector_store.py:
59-246; `b200vs.store` is the single-GPU drop-in): `add_vectors(vectors, metadata)`,
`query` / `batch_query` returning `(indices, scores, metadata)` lists with GLOBAL insertion-order
row ids, `filter_metadata`, `clear`, `get_stats`, `optimize`, `health_check`, and the reference's
on-disk format.  Rows are split over the ranks by `split_batch` (rank r keeps a contiguous slice
of every appended batch), searches are the local K2/K3 kernels + one all-gather of the (B, k)
candidates + K4 (`b200vs.sharded.ShardedVectorStore`).  Metadata is replicated on every rank (it
is host-side Python data, like in the reference); the vectors are not.

Persistence (one box, shared filesystem): rank r logs ITS rows under
`<store_path>/shard_<r>_of_<world>/segments/`, rank 0 appends `metadata.jsonl`; `optimize()`
assembles the reference's single `vectors.npz` (key `vectors`, global row order) + `metadata.jsonl`
and drops the logs.  A store written by the reference, by `MLXVectorStore` or by a sharded store of
any world size loads at any world size from that snapshot; the per-rank logs only replay at the
world size that wrote them.
"""
from __future__ import annotations

import json
import logging
import shutil
import threading
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .sharded import ShardedVectorStore, split_batch
from .store import MLXVectorStoreConfig, _MetadataIndex, _to_host_f32

logger = logging.getLogger("b200vs.sharded_store")


class ShardedMLXVectorStore:
    def __init__(self, store_path: str, config: Optional[MLXVectorStoreConfig] = None, group=None,
                 device: Optional[torch.device] = None, shard_factory=None):
        self.store_path = Path(store_path).expanduser()
        self.config = config or MLXVectorStoreConfig()
        self.group = group
        self._lock = threading.RLock()
        if self.config.metric not in ("cosine", "euclidean", "dot_product") or not self.config.jit_compile:
            self._metric_ok = False          # reference :153-154: query raises RuntimeError
            metric = "cosine"
        else:
            self._metric_ok = True
            metric = self.config.metric
        if device is None:
            device = torch.device("cuda", int(self.config.device))
        self._sv = ShardedVectorStore(self.config.dimension, metric, device=device, group=group,
                                      shadow_bf16=self.config.shadow_bf16,
                                      max_vectors_per_shard=self.config.max_vectors,
                                      search_mode=self.config.search_mode, shard_factory=shard_factory)
        self.rank, self.world = self._sv.rank, self._sv.world
        self._metadata: List[Dict] = []
        self._index = _MetadataIndex()
        self._runs: List[Tuple[int, int]] = []     # (first global id, rows) of this rank's local rows, in local order
        self._version = 0
        self._mask_cache: Dict[Any, Tuple[int, Any, int]] = {}
        self._segments = 0
        if self.rank == 0:
            self.store_path.mkdir(parents=True, exist_ok=True)
        self._barrier()
        self._load_store()

    # ------------------------------------------------------------------ helpers
    def _barrier(self):
        if self._sv.distributed:
            dist.barrier(group=self.group)

    @property
    def _vector_count(self) -> int:
        return self._sv.total

    def _shard_dir(self) -> Path:
        return self.store_path / f"shard_{self.rank}_of_{self.world}"

    def _check_dim(self, shape):
        if len(shape) != 2 or shape[1] != self.config.dimension:
            raise ValueError(f"vectors must have shape (m, {self.config.dimension}), got {tuple(shape)}")

    # ------------------------------------------------------------------ add
    def add_vectors(self, vectors, metadata: List[Dict]):
        """service/optimized_vector_store.py:96-114: every rank is handed the same batch and keeps
        its `split_batch` slice; global ids continue the insertion order."""
        with self._lock:
            v = vectors if isinstance(vectors, torch.Tensor) else _to_host_f32(vectors)
            if v.ndim == 1:
                v = v.reshape(1, -1)
            self._check_dim(tuple(v.shape))
            m = int(v.shape[0])
            metadata = list(metadata)
            if len(metadata) != m:
                logger.warning("%d vectors but %d metadata entries", m, len(metadata))
                metadata = metadata[:m] + [{} for _ in range(m - len(metadata))]
            first = self._sv.total
            lo, hi = split_batch(m, self.world, self.rank)
            self._metadata.extend(metadata)
            try:
                self._sv.add_vectors(v)
            except Exception:
                del self._metadata[first:]
                raise
            self._index.extend(first, metadata)
            if hi > lo:
                self._runs.append((first + lo, hi - lo))
            self._version += 1
            self._mask_cache.clear()
            if self.config.persist:
                self._persist_append(first, lo, hi, v, metadata)
            return {"vectors_added": m, "total_vectors": self._sv.total}

    # ------------------------------------------------------------------ query
    def query(self, query_vector, k: int = 10, filter_metadata: Optional[Dict] = None,
              use_hnsw: bool = True) -> Tuple:
        """service/optimized_vector_store.py:116-192 -> (indices, scores, metadata), best first."""
        if self._sv.total == 0:
            return [], [], []
        if not self._metric_ok:
            raise RuntimeError("no compiled similarity function available")
        q = _to_host_f32(query_vector).reshape(-1)
        if q.shape[0] != self.config.dimension:
            raise ValueError(f"query must have {self.config.dimension} components, got {q.shape[0]}")
        return self._search(q.reshape(1, -1), k, filter_metadata)[0]

    def batch_query(self, queries, k: int = 10, filter_metadata: Optional[Dict] = None):
        """List of per-query `(indices, scores, metadata)` tuples (api/routes/vectors.py:291)."""
        q = _to_host_f32(queries)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.ndim != 2 or q.shape[1] != self.config.dimension:
            raise ValueError(f"queries must have shape (B, {self.config.dimension}), got {q.shape}")
        if self._sv.total == 0:
            return [([], [], []) for _ in range(q.shape[0])]
        if not self._metric_ok:
            raise RuntimeError("no compiled similarity function available")
        return self._search(q, k, filter_metadata)

    def _local_mask(self, filter_metadata: Dict):
        """(this rank's row bitmap as the shard wants it, global number of hits); cached."""
        try:
            ckey = tuple(sorted(filter_metadata.items(), key=lambda kv: repr(kv[0])))
            hash(ckey)
        except TypeError:
            ckey = None
        if ckey is not None:
            ent = self._mask_cache.get(ckey)
            if ent is not None and ent[0] == self._version:
                return ent[1], ent[2], ent[3]
        n = self._sv.total
        hit = None
        generic = {}
        for key, val in filter_metadata.items():
            h = self._index.lookup(n, key, val)
            if h is None:
                generic[key] = val
            else:
                hit = h if hit is None else (hit & h)
        if generic:
            g = np.fromiter((all(m.get(key) == val for key, val in generic.items()) for m in self._metadata[:n]),
                            dtype=np.bool_, count=n)
            hit = g if hit is None else (hit & g)
        local = np.concatenate([hit[g0:g0 + c] for g0, c in self._runs]) if self._runs else np.zeros(0, np.bool_)
        mask = self._sv.shard.make_row_mask(local)
        n_live, local_live = int(hit.sum()), int(local.sum())
        if ckey is not None:
            if len(self._mask_cache) >= 16:
                self._mask_cache.clear()
            self._mask_cache[ckey] = (self._version, mask, n_live, local_live)
        return mask, n_live, local_live

    def _search(self, q: np.ndarray, k: int, filter_metadata: Optional[Dict]):
        B = q.shape[0]
        k = int(k)
        empty = [([], [], []) for _ in range(B)]
        if k == 0:
            return empty
        with self._lock:
            mask, local_live = None, -1
            n_live = self._sv.total
            if filter_metadata:
                mask, n_live, local_live = self._local_mask(filter_metadata)
                if n_live == 0:
                    return empty
            kk = min(k, n_live) if k > 0 else max(0, n_live + k)     # reference slices argsort(...)[:k]
            if kk == 0:
                return empty
            ids, scores = self._sv.result(self._sv.submit(q, kk, row_mask=mask, mask_live=local_live))
        ids_h = ids.cpu().numpy()
        sc_h = scores.cpu().numpy()
        meta = self._metadata
        out = []
        for b in range(B):
            idx = [i for i in ids_h[b].tolist() if i >= 0]
            out.append((idx, sc_h[b, :len(idx)].tolist(), [meta[i] for i in idx]))
        return out

    # ------------------------------------------------------------------ misc surface
    def clear(self):
        """service/optimized_vector_store.py:198-209."""
        with self._lock:
            self._barrier()
            if self.rank == 0 and self.store_path.exists():
                shutil.rmtree(self.store_path)
                self.store_path.mkdir(parents=True, exist_ok=True)
            self._barrier()
            self._sv.reset()
            self._metadata, self._runs, self._segments = [], [], 0
            self._index = _MetadataIndex()
            self._version += 1
            self._mask_cache.clear()

    def get_stats(self) -> Dict[str, Any]:
        """service/optimized_vector_store.py:241-242 (+ memory_usage_mb of THIS rank's shard)."""
        return {"vector_count": self._sv.total, "dimension": self.config.dimension, "metric": self.config.metric,
                "index_type": "flat", "memory_usage_mb": self._sv.shard.memory_bytes() / 2**20,
                "shards": self.world, "local_vectors": self._sv.shard.count()}

    def health_check(self) -> Dict[str, Any]:
        issues = []
        if len(self._metadata) != self._sv.total:
            issues.append(f"{self._sv.total} vectors but {len(self._metadata)} metadata entries")
        if sum(c for _, c in self._runs) != self._sv.shard.count():
            issues.append("local row runs do not add up to the shard's row count")
        if not self._metric_ok:
            issues.append("no similarity function (jit_compile=False or unknown metric)")
        return {"healthy": not issues, "issues": issues}

    def close(self):
        self._sv.close()

    # ------------------------------------------------------------------ persistence
    def _persist_append(self, first: int, lo: int, hi: int, v, metadata: List[Dict]):
        # one segment per rank and batch (also when this rank's slice is empty), named after the
        # BATCH: seg_<first global id of the batch>_<rows in the batch>.npy holds this rank's slice
        rows = v[lo:hi]
        rows = rows.detach().cpu().numpy() if isinstance(rows, torch.Tensor) else rows
        seg_dir = self._shard_dir() / "segments"
        seg_dir.mkdir(parents=True, exist_ok=True)
        tmp = seg_dir / f"tmp_{first:012d}.npy"
        np.save(tmp, np.ascontiguousarray(rows, dtype=np.float32).reshape(hi - lo, self.config.dimension))
        tmp.replace(seg_dir / f"seg_{first:012d}_{v.shape[0]}.npy")
        if self.rank == 0:
            with open(self.store_path / "metadata.jsonl", "a") as f:
                f.write("".join(json.dumps(m) + "\n" for m in metadata))

    def optimize(self):
        """Fold the per-rank logs into the reference's single `vectors.npz` + `metadata.jsonl`
        (api/routes/vectors.py:425).  Every rank writes its rows into one memory-mapped (N, D) file at
        their global positions; rank 0 packs it."""
        with self._lock:
            n, d = self._sv.total, self.config.dimension
            if n == 0:
                return
            raw = self.store_path / "vectors.tmp.npy"
            if self.rank == 0:
                np.lib.format.open_memmap(raw, mode="w+", dtype=np.float32, shape=(n, d)).flush()
            self._barrier()
            mm = np.lib.format.open_memmap(raw, mode="r+")
            local = 0
            for g0, c in self._runs:
                mm[g0:g0 + c] = self._sv.shard.read_rows(local, c)
                local += c
            mm.flush()
            del mm
            self._barrier()
            if self.rank == 0:
                mtmp = self.store_path / "metadata.tmp.jsonl"
                with open(mtmp, "w") as f:
                    for m in self._metadata[:n]:
                        f.write(json.dumps(m) + "\n")
                tmp = self.store_path / "vectors.tmp.npz"
                np.savez(str(tmp), vectors=np.load(raw, mmap_mode="r"))
                tmp.replace(self.store_path / "vectors.npz")
                mtmp.replace(self.store_path / "metadata.jsonl")
                raw.unlink()
                for p in self.store_path.glob("shard_*_of_*"):
                    shutil.rmtree(p)
            self._barrier()

    def _load_store(self):
        try:
            covered = 0
            vp = self.store_path / "vectors.npz"
            if vp.exists():
                snap = np.load(str(vp))["vectors"]
                self._check_dim(snap.shape)
                lo, hi = split_batch(snap.shape[0], self.world, self.rank)
                self._sv.add_vectors(np.ascontiguousarray(snap, dtype=np.float32))
                if hi > lo:
                    self._runs.append((lo, hi - lo))
                covered = snap.shape[0]
                del snap
            # per-rank logs (written at this world size): every rank holds one segment per batch
            seg_dir = self._shard_dir() / "segments"
            batches = []
            for seg in (sorted(seg_dir.glob("seg_*.npy")) if seg_dir.exists() else []):
                first, m = (int(x) for x in seg.stem.split("_")[1:3])
                if first + m <= covered:
                    continue                     # inside the snapshot (interrupted optimize())
                batches.append((first, m, seg))
            if self._sv.distributed:
                seen = [None] * self.world
                dist.all_gather_object(seen, [(f, m) for f, m, _ in batches], group=self.group)
                if any(x != seen[0] for x in seen):
                    raise ValueError("per-rank segment logs disagree (written at a different world size?)")
            for first, m, seg in batches:
                if first != self._sv.total:
                    raise ValueError(f"segment {seg.name} does not continue the store at row {self._sv.total}")
                lo, hi = split_batch(m, self.world, self.rank)
                rows = np.load(str(seg))
                if rows.shape[0] != hi - lo:
                    raise ValueError(f"segment {seg.name} holds {rows.shape[0]} rows, this rank's slice has {hi - lo}")
                self._sv.add_local(rows, first + lo, m)
                if hi > lo:
                    self._runs.append((first + lo, hi - lo))
            meta = []
            mp = self.store_path / "metadata.jsonl"
            if mp.exists():
                with open(mp) as f:
                    meta = [json.loads(line) for line in f if line.strip()]
            n = self._sv.total
            if len(meta) != n:
                logger.error("%s: %d vectors but %d metadata entries", self.store_path, n, len(meta))
                meta = meta[:n] + [{} for _ in range(n - len(meta))]
            self._metadata = meta
            self._index = _MetadataIndex()
            self._index.extend(0, meta)
            self._version += 1
        except Exception as e:   # reference :237-239: log and start empty
            logger.error("loading %s failed, starting empty: %s", self.store_path, e)
            self._sv.reset()
            self._metadata, self._runs = [], []
            self._index = _MetadataIndex()


def create_sharded_vector_store(store_path: str, dimension: int = 384, jit_compile: bool = True,
                                enable_hnsw: bool = False, **kwargs) -> ShardedMLXVectorStore:
    """`create_optimized_vector_store` (service/optimized_vector_store.py:244-246) for the sharded store."""
    group = kwargs.pop("group", None)
    config = MLXVectorStoreConfig(dimension=dimension, jit_compile=jit_compile, enable_hnsw=enable_hnsw, **kwargs)
    return ShardedMLXVectorStore(store_path, config, group=group)
