"""Stand-in for b200vs.sharded.NativeShard used ONLY by the multi-process CPU tests: the
oracle scores this rank's rows, so that the host logic of ShardedVectorStore (batch
splitting, global-id assignment, all-gather layout, merge order) runs under gloo without a
GPU.  Test infrastructure; the product path never imports this."""
import numpy as np
import torch

from oracle import vs_oracle


class OracleShard:
    def __init__(self, dimension, metric, device, shadow_bf16, max_vectors, search_mode):
        self.dimension, self.metric = dimension, metric
        self.rows = np.zeros((0, dimension), np.float32)
        self.gids = np.zeros((0,), np.int32)

    def append(self, rows, first_global_id):
        r = vs_oracle.to_f32(rows)
        self.rows = np.concatenate([self.rows, r], axis=0)
        self.gids = np.concatenate([self.gids, np.arange(first_global_id, first_global_id + r.shape[0],
                                                         dtype=np.int32)])

    def count(self):
        return self.rows.shape[0]

    def prepare_queries(self, queries):
        return torch.from_numpy(vs_oracle.to_f32(queries))

    def new_pack(self, B, k):
        return torch.empty((2, B, k), dtype=torch.int32)

    def search_into(self, q, k, pack, row_mask=None):
        B = q.shape[0]
        ids = np.full((B, k), -1, np.int32)
        sc = np.zeros((B, k), np.float32)
        rows, gids = self.rows, self.gids
        if row_mask is not None:          # the reference gathers the filtered rows (:167)
            rows, gids = rows[row_mask], gids[row_mask]
        if rows.shape[0]:
            li, ls, _ = vs_oracle.search(q.numpy(), rows, k, self.metric)
            ids[:, :li.shape[1]] = gids[li]
            sc[:, :li.shape[1]] = ls
        pack[0].copy_(torch.from_numpy(sc.view(np.int32)))
        pack[1].copy_(torch.from_numpy(ids))

    def submit_into(self, q, k, pack, row_mask=None, mask_live=-1):
        self.search_into(q, k, pack, row_mask)
        return None

    def make_row_mask(self, hit):
        return np.asarray(hit, dtype=np.bool_).copy()

    def reset(self):
        self.rows = np.zeros((0, self.dimension), np.float32)
        self.gids = np.zeros((0,), np.int32)

    def memory_bytes(self):
        return int(self.rows.nbytes)

    def read_rows(self, first, m):
        return self.rows[first:first + m].copy()

    def complete(self, ticket):
        pass

    def last_complete_enqueued_work(self):
        return False

    def merge(self, gathered, G, B, k):
        g = gathered.numpy()
        sc = g[:, 0].view(np.float32).transpose(1, 0, 2).reshape(B, G * k)
        ids = g[:, 1].transpose(1, 0, 2).reshape(B, G * k)
        key = np.where(ids >= 0, sc if self.metric == "euclidean" else -sc, np.inf)
        order = np.lexsort((ids, key), axis=1)[:, :k]       # key asc, then id asc
        return (torch.from_numpy(np.take_along_axis(ids, order, 1)),
                torch.from_numpy(np.take_along_axis(sc, order, 1)))

    def close(self):
        pass
