"""Row-sharded exact search over the GPUs of one box: one process per GPU
(`torch.distributed`, NCCL over NVLink), each rank owns a contiguous slice of every appended
batch, searches it with the local kernels (K2/K3 -> local top-k with GLOBAL ids), then one
all-gather of the (B, k) candidates per rank and a K4 merge give every rank the global top-k.

The reference is single-device (no collective anywhere); this layer is new.
Global ids are insertion-order row numbers exactly as in the single-store case
(service/optimized_vector_store.py:96-114), so results are identical to an unsharded store.

There is no collective on the data path other than that one all-gather: the database rows
never move between GPUs, queries are replicated (every rank is handed the same batch).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi


class NativeShard:
    """This rank's shard: a libb200vs store on one CUDA device (no CPU fallback)."""

    def __init__(self, dimension: int, metric: str, device: torch.device, shadow_bf16: bool,
                 max_vectors: int, search_mode: str):
        if device.type != "cuda":
            raise RuntimeError("b200vs shards live on CUDA devices only (no CPU fallback)")
        self.device = device
        self.metric = _cabi.METRICS[metric]
        self.flags = _cabi.SEARCH_MODES[search_mode]
        self.handle = C.c_void_p()
        _cabi.check(_cabi.lib().vs_create(device.index or 0, dimension, self.metric,
                                          int(shadow_bf16),      # bool, or a VS_SHADOW_* mask
                                          int(max_vectors), C.byref(self.handle)))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def append(self, rows, first_global_id: int) -> None:
        if isinstance(rows, torch.Tensor) and rows.is_cuda:
            r = rows.detach().to(torch.float32).contiguous()
            _cabi.check(_cabi.lib().vs_append_ids(self.handle, C.c_void_p(r.data_ptr()), r.shape[0], 1,
                                                  int(first_global_id), self._stream()))
        else:
            if isinstance(rows, torch.Tensor):
                rows = rows.detach().cpu().numpy()
            r = np.ascontiguousarray(rows, dtype=np.float32)
            _cabi.check(_cabi.lib().vs_append_ids(self.handle, r.ctypes.data_as(C.c_void_p), r.shape[0],
                                                  0, int(first_global_id), None))

    def count(self) -> int:
        return int(_cabi.lib().vs_count(self.handle))

    def new_pack(self, B: int, k: int) -> torch.Tensor:
        return torch.empty((2, B, k), dtype=torch.int32, device=self.device)

    def search_into(self, q: torch.Tensor, k: int, pack: torch.Tensor) -> None:
        """pack[0] <- fp32 scores (bit pattern), pack[1] <- int32 global ids; (B, k) each."""
        B = q.shape[0]
        _cabi.check(_cabi.lib().vs_search(self.handle, C.c_void_p(q.data_ptr()), B, k, self.flags, None,
                                          C.c_void_p(pack[0].data_ptr()), C.c_void_p(pack[1].data_ptr()),
                                          self._stream()))

    def merge(self, gathered: torch.Tensor, G: int, B: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """gathered: (G, 2, B, k) int32 as produced by the all-gather of `pack`."""
        out_s = torch.empty((B, k), dtype=torch.float32, device=self.device)
        out_i = torch.empty((B, k), dtype=torch.int32, device=self.device)
        base = gathered.data_ptr()
        _cabi.check(_cabi.lib().vs_merge(self.device.index or 0, self.metric, C.c_void_p(base),
                                         C.c_void_p(base + 4 * B * k), G, B, k, 2 * B * k,
                                         C.c_void_p(out_s.data_ptr()), C.c_void_p(out_i.data_ptr()),
                                         self._stream()))
        return out_i, out_s

    def prepare_queries(self, queries) -> torch.Tensor:
        if isinstance(queries, torch.Tensor):
            return queries.detach().to(self.device, torch.float32, non_blocking=True).contiguous()
        return torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32)).to(self.device)

    def close(self) -> None:
        if self.handle.value:
            _cabi.lib().vs_destroy(self.handle)
            self.handle = C.c_void_p()


def split_batch(m: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) of an m-row batch that `rank` of `world` owns: contiguous, sizes differ
    by at most one, so scan time stays balanced as the store grows."""
    base, rem = divmod(m, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedVectorStore:
    """Collective store: every rank calls every method with the same arguments."""

    def __init__(self, dimension: int, metric: str = "cosine", device: Optional[torch.device] = None,
                 group=None, shadow_bf16: bool = True, max_vectors_per_shard: int = 0,
                 search_mode: str = "auto", shard_factory=None):
        self.group = group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.dimension = dimension
        self.metric = metric
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        # `shard_factory` exists so the host logic (splitting, id assignment, gather, merge
        # order) can be exercised by multi-process CPU tests with a stand-in shard; the
        # default -- the only one the package ships -- is the CUDA engine.
        factory = shard_factory or NativeShard
        self.shard = factory(dimension, metric, device, shadow_bf16, max_vectors_per_shard, search_mode)
        self.total = 0
        self._bufs = {}
        # B200VS_SHARD_SYNC=1: synchronise the stream before returning (measured: no effect on throughput)
        import os
        self.sync_each_search = (self.world > 1 and device.type == "cuda" and
                                 os.environ.get("B200VS_SHARD_SYNC", "0") == "1")

    # ------------------------------------------------------------------ add
    def add_vectors(self, vectors) -> dict:
        """Every rank is handed the same (m, D) batch and keeps its slice; global ids continue
        the insertion order: total_before + row number in the batch."""
        m = int(vectors.shape[0])
        lo, hi = split_batch(m, self.world, self.rank)
        if hi > lo:
            self.shard.append(vectors[lo:hi], self.total + lo)
        self.total += m
        return {"vectors_added": m, "total_vectors": self.total}

    def add_local(self, rows, first_global_id: int, batch_rows: int) -> None:
        """This rank's slice of a batch of `batch_rows` rows that was produced shard by shard
        (e.g. generated on the device); slices must follow `split_batch`."""
        lo, hi = split_batch(batch_rows, self.world, self.rank)
        assert rows.shape[0] == hi - lo and first_global_id == self.total + lo
        if hi > lo:
            self.shard.append(rows, first_global_id)
        self.total += batch_rows

    # ------------------------------------------------------------------ search
    def search(self, queries, k: int = 10) -> Tuple[torch.Tensor, torch.Tensor]:
        """(ids (B, kk) int32, scores (B, kk) fp32) on this rank's device, kk = min(k, total);
        identical on every rank and to an unsharded store."""
        q = self.shard.prepare_queries(queries)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.ndim != 2 or q.shape[1] != self.dimension:
            raise ValueError(f"queries must have shape (B, {self.dimension}), got {tuple(q.shape)}")
        B = q.shape[0]
        kk = max(0, min(int(k), self.total))
        if B == 0 or kk == 0:
            return (torch.zeros((B, 0), dtype=torch.int32, device=q.device),
                    torch.zeros((B, 0), dtype=torch.float32, device=q.device))
        # The buffers the collective touches are kept per (B, k) instead of being re-allocated:
        # tensors used on NCCL's stream go back to torch's caching allocator only after that
        # stream has passed them, so a host that runs ahead of the GPU would otherwise fall
        # through to cudaMalloc (a device-wide synchronisation) every few steps.
        key = (B, kk)
        bufs = self._bufs.get(key)
        if bufs is None:
            pack = self.shard.new_pack(B, kk)
            flat = torch.empty((self.world * pack.numel(),), dtype=pack.dtype, device=pack.device)
            if len(self._bufs) > 8:
                self._bufs.clear()
            bufs = self._bufs[key] = (pack, flat)
        pack, flat = bufs
        self.shard.search_into(q, kk, pack)
        if self.world == 1:
            return pack[1].clone(), pack[0].view(torch.float32).clone()
        dist.all_gather_into_tensor(flat, pack.view(-1), group=self.group)
        out = self.shard.merge(flat.view((self.world,) + tuple(pack.shape)), self.world, B, kk)
        if self.sync_each_search:
            torch.cuda.current_stream(self.device).synchronize()
        return out

    def close(self) -> None:
        self.shard.close()
