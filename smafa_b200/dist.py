"""Multi-GPU query: the db is row-sharded across ranks (one process per GPU), queries are
replicated, every rank scans its shard and the per-query local candidates are merged with one
all-gather of fixed-capacity candidate blocks (the row count rides in a header row) -- SURVEY.md 8e.

Local candidates are a superset of the global answer: Mode A emits the local minimum and its
ties, Mode B everything <= min(local k-th distance, --max-divergence); the local cutoff is never
below the global one.  Shards are contiguous row ranges and report global subject indices, so the
merged (distance, subject) order equals the reference's print order (src/lib.rs:250,307).

torch / torch.distributed are plumbing only (device buffers, streams, NCCL); the scan and the
merge are the library's own kernels (smafa_query_dev / smafa_merge_dev).
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

MAX_SLAB = 1 << 20  # queries per exchange (candidate keys carry 20 query bits)


def shard_bounds(D, world_size, rank):
    """Contiguous row range [lo, hi) of `rank`; the first D % world_size shards get one extra row."""
    base, extra = divmod(D, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _pow2_at_least(n):
    return 1 << max(0, int(n) - 1).bit_length()


class _ExchangeBuffers:
    """Send block and receive area of the candidate all-gather, kept between steps (allocating and zero-filling
    them every step showed up as launches in front of the collective)."""

    def __init__(self):
        self.cap, self.block, self.gathered = 0, None, None

    def get(self, cap, ws, dev):
        if self.cap != cap or self.block is None or self.block.device != dev:
            self.block = torch.zeros((cap + 1, 3), dtype=torch.int32, device=dev)
            self.gathered = torch.empty((ws * (cap + 1), 3), dtype=torch.int32, device=dev)
            self.cap = cap
        return self.block, self.gathered


def exchange_candidates(local_rows, group=None, capacity=None, buffers=None):
    """All-gathers ragged [n_r, 3] int32 candidate blocks; returns (concatenation in rank order,
    largest per-rank count).  ONE collective on the common path: every rank contributes a block of
    `capacity` rows behind a header row holding its true count, so no separate count exchange is
    needed; only if some rank had more rows than `capacity` (seen by every rank in the gathered
    headers) is the gather repeated with a capacity that fits.  Rows of a block beyond its count are
    stale and never read.  Works on any backend (NCCL on GPUs, gloo in the CPU tests)."""
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    n_local = int(local_rows.shape[0])
    if ws == 1:
        return local_rows, n_local
    dev = local_rows.device
    cap = max(int(capacity or 0), 16)
    buffers = buffers or _ExchangeBuffers()
    while True:
        block, gathered = buffers.get(cap, ws, dev)
        block[0, 0] = n_local
        keep = min(n_local, cap)
        block[1:1 + keep] = local_rows[:keep]
        dist.all_gather_into_tensor(gathered, block, group=group)
        view = gathered.view(ws, cap + 1, 3)
        counts = view[:, 0, 0].tolist()  # the one host read-back of the exchange
        if max(counts) <= cap:
            break
        cap = _pow2_at_least(max(counts))  # identical on every rank: all of them saw the same headers
    return torch.cat([view[r, 1:1 + c] for r, c in enumerate(counts)], dim=0), max(counts)


class ShardedSearcher:
    """Holds this rank's db shard on its GPU and answers replicated query batches."""

    def __init__(self, ctx, db_words, L, world_size=1, rank=0, group=None, hits_capacity=1 << 24, presharded=False,
                 shard_offset=None, total_rows=None):
        self.ctx, self.L, self.group = ctx, L, group
        self.world_size, self.rank = world_size, rank
        self.W = (L + 11) // 12
        D = db_words.shape[0]
        if presharded:
            # shards built rank-locally: equal ones (bench weak scaling) unless the caller gives this shard's first
            # global row and the db's total
            self.lo = rank * D if shard_offset is None else int(shard_offset)
            shard = db_words
            self.D_total = D * world_size if total_rows is None else int(total_rows)
        else:
            self.lo, hi = shard_bounds(D, world_size, rank)
            shard = db_words[self.lo:hi]
            self.D_total = D
        self.db = ctx.upload(np.ascontiguousarray(shard), L, subject_offset=self.lo)
        self.device = torch.device("cuda", ctx.device)
        self.hits = torch.empty((hits_capacity, 3), dtype=torch.int32, device=self.device)
        self.last_stats = None
        self._exchange_cap = 0  # rows per rank in the candidate all-gather (adapts to the workload)
        self._exchange_buffers = _ExchangeBuffers()
        self._timing = bool(os.environ.get("SMAFA_TIMING"))
        self.phase_ms = [0.0, 0.0, 0.0]  # local scan + selection, candidate exchange, merge (SMAFA_TIMING=1)

    def _local(self, q_dev, m, k):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        n, st = self.ctx.query_dev(self.db, q_dev.data_ptr(), q_dev.shape[0], self.L, self.hits.data_ptr(),
                                   self.hits.shape[0], m, k, stream=stream)
        self.last_stats = st
        return self.hits[:n]

    def query_dev(self, q_dev, max_divergence=None, max_num_hits=None):
        """q_dev: int64 [Q, W] tensor on this rank's GPU (bit pattern of the u64 words).
        Returns an int32 [n, 3] device tensor of (query, subject, distance) rows in print order,
        identical on every rank."""
        out = []
        launches = 0
        timing = self._timing
        for s0 in range(0, q_dev.shape[0], MAX_SLAB):
            slab = q_dev[s0:s0 + MAX_SLAB]
            t0 = time.perf_counter()
            rows = self._local(slab, max_divergence, max_num_hits)
            launches += self.last_stats["kernel_launches"]
            if self.world_size > 1:
                if timing:
                    torch.cuda.synchronize()
                    t1 = time.perf_counter()
                cap = self._exchange_cap or _pow2_at_least(max(4096, 2 * slab.shape[0]))
                union, biggest = exchange_candidates(rows, self.group, cap, self._exchange_buffers)
                self._exchange_cap = _pow2_at_least(max(4096, 2 * biggest))
                if timing:
                    torch.cuda.synchronize()
                    t2 = time.perf_counter()
                stream = torch.cuda.current_stream(self.device).cuda_stream
                n = self.ctx.merge_dev(union.data_ptr(), union.shape[0], max_divergence, max_num_hits, stream=stream)
                launches += 8
                rows = union[:n]  # a fresh tensor (the concatenation), merged in place
                if timing:  # SMAFA_TIMING=1: host clock per phase (with a device sync after each: measurement aid only)
                    torch.cuda.synchronize()
                    t3 = time.perf_counter()
                    self.phase_ms = [x + 1e3 * y for x, y in zip(self.phase_ms, (t1 - t0, t2 - t1, t3 - t2))]
            else:
                rows = rows.clone()  # self.hits is overwritten by the next call
            if s0:
                rows[:, 0] += s0
            out.append(rows)
        self.last_launches = launches
        return out[0] if len(out) == 1 else torch.cat(out, dim=0)

    def query_host(self, q_pinned, max_divergence=None, max_num_hits=None):
        """End-to-end: pinned host words in, host rows out (H2D and D2H inside)."""
        q_dev = q_pinned.to(self.device, non_blocking=True)
        rows = self.query_dev(q_dev, max_divergence, max_num_hits)
        host = rows.cpu()
        return host.numpy().view(np.uint32)

    def close(self):
        self.db.close()
