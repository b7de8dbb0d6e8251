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
import os
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
        self.dimension = dimension
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
        """(2, B, k) int32 view of a buffer padded to whole 16-byte words (vs_exchange_push)."""
        words = 2 * B * k
        raw = torch.zeros(((words + 3) // 4 * 4,), dtype=torch.int32, device=self.device)
        return raw[:words].view(2, B, k)

    def search_into(self, q: torch.Tensor, k: int, pack: torch.Tensor) -> None:
        """pack[0] <- fp32 scores (bit pattern), pack[1] <- int32 global ids; (B, k) each."""
        B = q.shape[0]
        _cabi.check(_cabi.lib().vs_search(self.handle, C.c_void_p(q.data_ptr()), B, k, self.flags, None, -1,
                                          C.c_void_p(pack[0].data_ptr()), C.c_void_p(pack[1].data_ptr()),
                                          self._stream()))

    def submit_into(self, q: torch.Tensor, k: int, pack: torch.Tensor, row_mask=None,
                    mask_live: int = -1) -> C.c_void_p:
        """`search_into` without the host wait: everything is enqueued, the certification check
        stays pending in the returned ticket until `complete` (vs_search_submit).
        row_mask: int32 device bitmap over this shard's LOCAL rows (make_row_mask)."""
        ticket = C.c_void_p()
        _cabi.check(_cabi.lib().vs_search_submit(
            self.handle, C.c_void_p(q.data_ptr()), q.shape[0], k, self.flags,
            None if row_mask is None else C.c_void_p(row_mask.data_ptr()), int(mask_live),
            C.c_void_p(pack[0].data_ptr()), C.c_void_p(pack[1].data_ptr()), self._stream(), C.byref(ticket)))
        return ticket

    def make_row_mask(self, hit: np.ndarray) -> torch.Tensor:
        """bool (local rows,) -> device bitmap (whole 32-bit words + one spare word)."""
        bits = np.packbits(hit, bitorder="little")
        buf = np.zeros(((hit.shape[0] + 31) // 32 + 1) * 4, np.uint8)
        buf[:bits.size] = bits
        mask = torch.from_numpy(buf.view(np.int32)).to(self.device)
        torch.cuda.current_stream(self.device).synchronize()
        return mask

    def reset(self) -> None:
        _cabi.check(_cabi.lib().vs_reset(self.handle))

    def memory_bytes(self) -> int:
        return int(_cabi.lib().vs_memory_bytes(self.handle))

    def read_rows(self, first: int, m: int) -> np.ndarray:
        out = np.empty((m, self.dimension), np.float32)
        if m:
            _cabi.check(_cabi.lib().vs_read_rows(self.handle, first, m, out.ctypes.data_as(C.c_void_p), 0, None))
        return out

    def complete(self, ticket) -> None:
        lib = _cabi.lib()
        before = lib.vs_retry_count(self.handle) + lib.vs_fallback_count(self.handle)
        _cabi.check(lib.vs_search_complete(self.handle, ticket))
        self._completed_launches = lib.vs_retry_count(self.handle) + lib.vs_fallback_count(self.handle) - before

    def last_complete_enqueued_work(self) -> bool:
        """True when the last `complete` had to re-run uncertified queries (kernels enqueued)."""
        return bool(getattr(self, "_completed_launches", 0))

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


class PeerExchange:
    """Exchange buffers of one (B, k) shape in symmetric memory (every rank of the group has every
    rank's buffer mapped): `depth` slots of `world` candidate blocks + `world` flag words each, and
    the block-completion counter of the push kernel.  Creation is collective.  csrc/exchange.cu."""
    DEPTH = 4

    def __init__(self, B: int, k: int, device: torch.device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.B, self.k = B, k
        self.block_bytes = (2 * B * k * 4 + 15) // 16 * 16
        self.slot_bytes = self.world * self.block_bytes
        self.flags_off = self.DEPTH * self.slot_bytes
        self.counter_off = self.flags_off + self.DEPTH * self.world * 4
        total = (self.counter_off + 16 + 255) // 256 * 256
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        torch.cuda.current_stream(device).synchronize()
        self.handle = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        ptrs = [int(x) for x in self.handle.buffer_ptrs]
        assert len(ptrs) == self.world and ptrs[self.rank] == self.buf.data_ptr()
        self.base = ptrs[self.rank]
        P = C.c_void_p * self.world
        # per slot: where this rank's block / flag word lives in every rank's buffer
        self.dst = [P(*[x + s * self.slot_bytes + self.rank * self.block_bytes for x in ptrs])
                    for s in range(self.DEPTH)]
        self.flag = [P(*[x + self.flags_off + (s * self.world + self.rank) * 4 for x in ptrs])
                     for s in range(self.DEPTH)]
        self.step = 0
        dist.barrier(group)                      # every buffer is zeroed and mapped before the first push

    def result(self, shard, ticket, pack: torch.Tensor, xs, cur) -> Tuple[torch.Tensor, torch.Tensor]:
        """vs_exchange_result: finish the search behind `ticket` (host wait for its certification count),
        then push / wait / merge on stream `xs` and make `cur` wait for the merge: (ids, scores)."""
        self.step += 1
        slot = self.step % self.DEPTH
        B, k = self.B, self.k
        out_s = torch.empty((B, k), dtype=torch.float32, device=shard.device)
        out_i = torch.empty((B, k), dtype=torch.int32, device=shard.device)
        _cabi.check(_cabi.lib().vs_exchange_result(
            shard.handle, ticket, C.c_void_p(pack.data_ptr()), self.block_bytes, self.dst[slot], self.flag[slot],
            self.world, self.step, C.c_void_p(self.base + self.counter_off),
            C.c_void_p(self.base + self.flags_off + slot * self.world * 4), C.c_void_p(self.base + slot * self.slot_bytes),
            B, k, C.c_void_p(out_s.data_ptr()), C.c_void_p(out_i.data_ptr()),
            C.c_void_p(xs.cuda_stream), C.c_void_p(cur.cuda_stream)))
        return out_i, out_s

    def exchange(self, shard, pack: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Push this rank's (2, B, k) block to every rank, wait for everybody's, merge: (ids, scores).
        Everything is enqueued on the current stream."""
        lib = _cabi.lib()
        self.step += 1
        slot = self.step % self.DEPTH
        dev = shard.device.index or 0
        stream = shard._stream()
        # `pack` is a view of a buffer padded to whole 16-byte words (NativeShard.new_pack)
        _cabi.check(lib.vs_exchange_push(dev, C.c_void_p(pack.data_ptr()), self.block_bytes, self.dst[slot],
                                         self.flag[slot], self.world, self.step,
                                         C.c_void_p(self.base + self.counter_off), stream))
        B, k = self.B, self.k
        out_s = torch.empty((B, k), dtype=torch.float32, device=shard.device)
        out_i = torch.empty((B, k), dtype=torch.int32, device=shard.device)
        flags = C.c_void_p(self.base + self.flags_off + slot * self.world * 4)
        base = self.base + slot * self.slot_bytes
        if self.world * k <= 256:
            # wait + merge in one kernel that needs no shared memory: runs next to the next search's GEMM
            _cabi.check(lib.vs_exchange_wait_merge(dev, shard.metric, flags, self.world, self.step, C.c_void_p(base),
                                                   self.block_bytes // 4, B, k, C.c_void_p(out_s.data_ptr()),
                                                   C.c_void_p(out_i.data_ptr()), stream))
        else:
            _cabi.check(lib.vs_exchange_wait(dev, flags, self.world, self.step, stream))
            _cabi.check(lib.vs_merge(dev, shard.metric, C.c_void_p(base), C.c_void_p(base + 4 * B * k), self.world, B, k,
                                     self.block_bytes // 4, C.c_void_p(out_s.data_ptr()), C.c_void_p(out_i.data_ptr()),
                                     stream))
        return out_i, out_s


class PendingSearch:
    """Handle returned by ShardedVectorStore.submit()."""
    __slots__ = ("q", "B", "kk", "bufs", "ticket", "done", "stream", "ex")

    def __init__(self, q, B, kk, bufs, ticket, done=None, stream=None, ex=None):
        self.q, self.B, self.kk, self.bufs, self.ticket, self.done = q, B, kk, bufs, ticket, done
        self.stream = stream                     # the side stream the search was enqueued on
        self.ex = ex                             # PeerExchange: the search went through vs_search_submit_on


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
                 search_mode: str = "auto", shard_factory=None, overlap_streams: Optional[bool] = None):
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
        self._xstream = None
        # submit() alternates between two side streams: the small kernels around a search's GEMM
        # (query preparation, threshold selection, rescoring) then overlap the neighbouring
        # search's GEMM instead of leaving the tensor pipe idle between two searches
        self._sstreams = None
        self._nsubmit = 0
        # Default: on for world > 1, where a shard's GEMM is short and those kernels are a quarter of
        # the step; off for a single GPU, where they are ~5 % and a single stream keeps every kernel's
        # CUDA-event bracket free of queueing time (bench.py's roofline).  B200VS_OVERLAP=0|1 overrides.
        if overlap_streams is None:
            env = os.environ.get("B200VS_OVERLAP", "")
            overlap_streams = (env == "1") if env in ("0", "1") else self.world > 1
        self.overlap_streams = bool(overlap_streams)
        # candidate exchange at world > 1: "p2p" = direct stores into every rank's symmetric-memory
        # buffer (PeerExchange, csrc/exchange.cu), "nccl" = all_gather_into_tensor.  p2p falls back to
        # nccl (on every rank together) when the symmetric-memory rendezvous is not possible.
        self.exchange_mode = os.environ.get("B200VS_EXCHANGE", "p2p")
        self._exchanges = {}
        # searches that may be in flight between submit() and result(); their buffers are recycled
        # in a ring of this depth
        self.pipeline_depth = 4
        # B200VS_SHARD_SYNC=1: synchronise the stream before returning (measured: no effect on throughput)
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
        identical on every rank and to an unsharded store.  = result(submit(...))."""
        return self.result(self.submit(queries, k))

    def reset(self) -> None:
        """Forget all rows on every rank (MLXVectorStore.clear's in-memory part)."""
        self.shard.reset()
        self.total = 0

    def submit(self, queries, k: int = 10, row_mask=None, mask_live: int = -1) -> "PendingSearch":
        """Enqueue this rank's local search (K2/K3 -> local top-k with global ids) on the current
        stream without waiting for the GPU; `result` finishes it.  Several searches may be in
        flight (a server keeps the GPU busy with batch i+1 while batch i's certification count
        travels to the host and its candidates cross NVLink); results come back in submit order."""
        q = self.shard.prepare_queries(queries)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.ndim != 2 or q.shape[1] != self.dimension:
            raise ValueError(f"queries must have shape (B, {self.dimension}), got {tuple(q.shape)}")
        B = q.shape[0]
        kk = max(0, min(int(k), self.total))
        if B == 0 or kk == 0:
            return PendingSearch(q, B, kk, None, None)
        # The buffers the collective touches are kept per (B, k) in a small ring instead of being
        # re-allocated: tensors used on NCCL's stream go back to torch's caching allocator only
        # after that stream has passed them, so a host that runs ahead of the GPU would otherwise
        # fall through to cudaMalloc (a device-wide synchronisation) every few steps.
        key = (B, kk)
        ring = self._bufs.get(key)
        if ring is None:
            if len(self._bufs) > 8:
                self._bufs.clear()
            ring = self._bufs[key] = {"slots": [], "next": 0}
        if len(ring["slots"]) < self.pipeline_depth:
            pack = self.shard.new_pack(B, kk)
            flat = torch.empty((self.world * pack.numel(),), dtype=pack.dtype, device=pack.device) \
                if self.world > 1 else None
            ring["slots"].append((pack, flat))
        pack, flat = ring["slots"][ring["next"] % len(ring["slots"])]
        ring["next"] += 1
        if self.device.type != "cuda":           # CPU stand-in shards (gloo tests)
            if row_mask is not None and mask_live == 0:
                pack[0].zero_()
                pack[1].fill_(-1)
                ticket = None
            elif row_mask is not None:
                ticket = self.shard.submit_into(q, kk, pack, row_mask, mask_live)
            else:
                ticket = self.shard.submit_into(q, kk, pack)
            return PendingSearch(q, B, kk, (pack, flat), ticket, None)
        cur = torch.cuda.current_stream(self.device)
        if self.overlap_streams:
            if self._sstreams is None:
                self._sstreams = [torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)]
            ss = self._sstreams[self._nsubmit % 2]
            self._nsubmit += 1
        else:
            ss = cur
        # world > 1 with the peer-memory exchange: the whole stream plumbing of a step lives in two C calls
        # (vs_search_submit_on here, vs_exchange_result in result()); a step's host cost otherwise bounds
        # the throughput at small batches
        ex = self._peer_exchange(B, kk) if self.world > 1 else None
        if ex is not None and not (row_mask is not None and mask_live == 0):
            ticket = C.c_void_p()
            _cabi.check(_cabi.lib().vs_search_submit_on(
                self.shard.handle, C.c_void_p(q.data_ptr()), B, kk, self.shard.flags,
                None if row_mask is None else C.c_void_p(row_mask.data_ptr()), int(mask_live),
                C.c_void_p(pack.data_ptr()), C.c_void_p(pack.data_ptr() + 4 * B * kk),
                C.c_void_p(cur.cuda_stream), C.c_void_p(ss.cuda_stream), C.byref(ticket)))
            if ss is not cur:
                q.record_stream(ss)
            return PendingSearch(q, B, kk, (pack, flat), ticket, None, ss, ex)
        if ss is not cur:
            ss.wait_stream(cur)                  # the queries (and an earlier reader of `pack`) are on `cur`
        with torch.cuda.stream(ss):
            if row_mask is not None and mask_live == 0:
                # no local row takes part: this rank contributes an empty candidate block
                pack[0].zero_()
                pack[1].fill_(-1)
                ticket = None
            elif row_mask is not None:
                ticket = self.shard.submit_into(q, kk, pack, row_mask, mask_live)
            else:
                ticket = self.shard.submit_into(q, kk, pack)
            done = torch.cuda.Event()
            done.record(ss)
        if ss is not cur:
            q.record_stream(ss)
        return PendingSearch(q, B, kk, (pack, flat), ticket, done, ss)

    def result(self, pending: "PendingSearch") -> Tuple[torch.Tensor, torch.Tensor]:
        """Finish a submitted search: wait for ITS certification count (re-running what could not
        be certified), then one all-gather of the packed (2, B, k) block per rank and the K4 merge,
        both on a side stream so the searches enqueued behind it are not held up."""
        q, B, kk = pending.q, pending.B, pending.kk
        if pending.bufs is None:
            return (torch.zeros((B, 0), dtype=torch.int32, device=q.device),
                    torch.zeros((B, 0), dtype=torch.float32, device=q.device))
        pack, flat = pending.bufs
        if pending.ex is not None:
            cur = torch.cuda.current_stream(self.device)
            if self._xstream is None:
                self._xstream = torch.cuda.Stream(self.device)
            return pending.ex.result(self.shard, pending.ticket, pack, self._xstream, cur)
        if pending.ticket is not None:
            self.shard.complete(pending.ticket)
        if self.device.type != "cuda":          # CPU stand-in shards (gloo tests)
            if self.world == 1:
                return pack[1].clone(), pack[0].view(torch.float32).clone()
            dist.all_gather_into_tensor(flat, pack.view(-1), group=self.group)
            return self.shard.merge(flat.view((self.world,) + tuple(pack.shape)), self.world, B, kk)
        cur = torch.cuda.current_stream(self.device)
        # a query that was re-run wrote its rows (on the search's stream) after `done` was recorded
        redone = pending.ticket is not None and self.shard.last_complete_enqueued_work()
        if self.world == 1:
            if pending.stream is not cur:
                if redone:
                    cur.wait_stream(pending.stream)
                else:
                    cur.wait_event(pending.done)
            return pack[1].clone(), pack[0].view(torch.float32).clone()
        if self._xstream is None:
            self._xstream = torch.cuda.Stream(self.device)
        xs = self._xstream
        if redone:
            xs.wait_stream(pending.stream)
        else:
            xs.wait_event(pending.done)
        with torch.cuda.stream(xs):
            ex = self._peer_exchange(B, kk)
            if ex is not None:
                out = ex.exchange(self.shard, pack)
            else:
                dist.all_gather_into_tensor(flat, pack.reshape(-1), group=self.group)
                out = self.shard.merge(flat.view((self.world,) + tuple(pack.shape)), self.world, B, kk)
        cur.wait_stream(xs)      # the caller consumes the results on its own stream
        for t in out:
            t.record_stream(cur)
        if self.sync_each_search:
            cur.synchronize()
        return out

    def _peer_exchange(self, B: int, kk: int):
        """The PeerExchange of this result shape, created (collectively) on first use; None when
        the exchange runs over NCCL."""
        if self.exchange_mode != "p2p":
            return None
        key = (B, kk)
        if key not in self._exchanges:
            if len(self._exchanges) > 8:
                self._exchanges.clear()
            ex, ok = None, 1
            try:
                ex = PeerExchange(B, kk, self.device, self.group)
            except Exception as e:       # noqa: BLE001 -- no symmetric memory on this box / build
                ok = 0
                import logging
                logging.getLogger(__name__).warning("peer-memory exchange unavailable (%s): using NCCL", e)
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) == 0:    # all ranks take the same path
                ex = None
                self.exchange_mode = "nccl"
            self._exchanges[key] = ex
        return self._exchanges[key]

    def close(self) -> None:
        self._exchanges.clear()
        self.shard.close()
