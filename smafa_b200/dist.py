"""Multi-GPU query with one process per GPU (torchrun): the db is row-sharded across ranks, queries are
replicated, every rank scans its shard, and the per-shard answers are exchanged and merged INSIDE the library
(csrc/sharded.cu: one ncclAllGather of fixed-capacity blocks on the call's stream + the sort-free merge of
csrc/merge.cu) -- SURVEY.md 8e.

Local answers are a superset of the global one: Mode A emits the local minimum and its ties, Mode B everything
<= min(local k-th distance, --max-divergence); the local cutoff is never below the global one.  Shards are
contiguous row ranges and report global subject indices, so the merged (distance, subject) order equals the
reference's print order (src/lib.rs:250,307).

torch / torch.distributed are plumbing only: device buffers, the stream, and the one-time broadcast of the
communicator id.  No tensor operation sits on the per-step data path.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import api


def shard_bounds(D, world_size, rank):
    """Contiguous row range [lo, hi) of `rank`; the first D % world_size shards get one extra row
    (the same split smafa_ctx_create_multi uses inside one process)."""
    base, extra = divmod(D, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_comm(ctx, rank, world_size, group=None):
    """Rank 0 draws the communicator id, torch.distributed (any backend) carries it to the others, every rank joins."""
    box = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.comm_init(box[0], rank, world_size)


class ShardedSearcher:
    """Holds this rank's db shard on its GPU and answers replicated query batches (collective calls)."""

    def __init__(self, ctx, db_words, L, world_size=1, rank=0, group=None, hits_capacity=1 << 22, presharded=False,
                 shard_offset=None, total_rows=None, group_rows=True):
        self.ctx, self.L, self.group = ctx, L, group
        self.world_size, self.rank = world_size, rank
        self.W = (L + 11) // 12
        D = db_words.shape[0]
        if presharded:
            # shards built rank-locally: equal ones unless the caller gives this shard's first global row and the db's total
            self.lo = rank * D if shard_offset is None else int(shard_offset)
            shard = db_words
            self.D_total = D * world_size if total_rows is None else int(total_rows)
        else:
            self.lo, hi = shard_bounds(D, world_size, rank)
            shard = db_words[self.lo:hi]
            self.D_total = D
        self.device = torch.device("cuda", ctx.device)
        self.clusters = 0
        if world_size > 1:
            init_comm(ctx, rank, world_size, group)
            if not presharded and group_rows:
                # The WHOLE db is grouped (every rank computes the same order on its own GPU: the greedy is exact and
                # deterministic) and the grouped order is cut into the shards, so every rank holds whole families of similar
                # windows -- what the wide union rows of the tcgen05 scan need.  Rows keep their subject numbers.
                perm, self.clusters = ctx.group_order(db_words, L)
            if self.clusters:
                self.db = ctx.upload_mapped(np.ascontiguousarray(db_words[perm[self.lo:hi]]), L, perm[self.lo:hi], self.D_total)
            else:
                self.db = ctx.upload_shard(np.ascontiguousarray(shard), L, self.lo, self.D_total)
        else:
            self.db = ctx.upload(np.ascontiguousarray(shard), L, subject_offset=self.lo)
        self.hits = torch.empty((hits_capacity, 3), dtype=torch.int32, device=self.device)
        self.last_stats = None
        self.last_launches = 0

    def query_dev(self, q_dev, max_divergence=None, max_num_hits=None):
        """q_dev: int64 [Q, W] tensor on this rank's GPU (bit pattern of the u64 words).
        Returns an int32 [n, 3] device tensor of (query, subject, distance) rows in print order, identical on
        every rank -- a view of this searcher's result buffer, valid until its next call."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        call = self.ctx.query_sharded_dev if self.world_size > 1 else self.ctx.query_dev
        while True:
            try:
                n, st = call(self.db, q_dev.data_ptr(), q_dev.shape[0], self.L, self.hits.data_ptr(), self.hits.shape[0],
                             max_divergence, max_num_hits, stream=stream)
                break
            except api.SmafaCapacityError as e:
                # the answer has more rows than the buffer: every rank sees the same count, so all of them grow and repeat
                self.hits = torch.empty((int(e.needed * 1.25) + 1024, 3), dtype=torch.int32, device=self.device)
        self.last_stats = st
        self.last_launches = st["kernel_launches"]
        return self.hits[:n]

    def query_host(self, q_pinned, max_divergence=None, max_num_hits=None):
        """End to end through the C ABI: host words in (pinned or not), host rows out; H2D, scan, exchange, merge and
        D2H all happen inside smafa_query_sharded / smafa_query."""
        ptr, Q = q_pinned.data_ptr(), q_pinned.shape[0]
        if self.world_size > 1:
            rows, st = self.ctx.query_sharded_ptr(self.db, ptr, Q, self.L, max_divergence, max_num_hits)
        else:
            rows, st = self.ctx.query_ptr(self.db, ptr, Q, self.L, max_divergence, max_num_hits)
        self.last_stats = st
        self.last_launches = st["kernel_launches"]
        return rows

    def close(self):
        self.db.close()
