"""Multi-GPU parity check (run under torchrun, one rank per GPU): row-sharded db + the library's NCCL block exchange +
sort-free device merge (csrc/sharded.cu, merge.cu) must equal the oracle on the whole db -- through the host entry
point (smafa_query_sharded) and the device one (smafa_query_sharded_dev), in every selection mode, with both kernels,
with the overflow re-send forced, with shards that are empty, and across a slab boundary (> 2^20 queries)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smafa_b200
from oracle import c_oracle
from smafa_b200 import synth
from smafa_b200.dist import ShardedSearcher

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local_rank = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
dist.init_process_group("nccl", device_id=dev)
L = 60
D, Q = int(os.environ.get("CHECK_D", "200001")), int(os.environ.get("CHECK_Q", "3000"))
db_sym = synth.make_db(D, L=L, seed=41)
db = synth.pack_symbols(db_sym)
q = synth.pack_symbols(synth.make_queries(db_sym, Q, seed=42))
threads = os.cpu_count() or 1
ok = True
report = {}


def check(name, s, dbw, qw, m, k):
    """host and device entry points against the oracle (rank 0 compares; every rank must hold the same rows)"""
    global ok
    qp = torch.from_numpy(qw.view(np.int64)).pin_memory()
    got = s.query_host(qp, m, k)
    retries = s.last_stats["retries"]
    got_dev = s.query_dev(qp.to(dev), m, k).cpu().numpy().view(np.uint32)
    retries += s.last_stats["retries"]
    same = got.shape == got_dev.shape and bool((got == got_dev).all())
    digest = torch.tensor([int(got.astype(np.uint64).sum() % (1 << 62)), got.shape[0]], dtype=torch.int64, device=dev)
    lo_d, hi_d = digest.clone(), digest.clone()
    dist.all_reduce(lo_d, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_d, op=dist.ReduceOp.MAX)
    same = same and bool((lo_d == hi_d).all().item())
    if rank == 0:
        want = c_oracle.query(dbw, L, qw, L, m, k, None, threads=threads)
        same = same and got.shape == want.shape and bool((got == want).all())
        report[name] = [same, int(got.shape[0]), retries]
        if not same:
            print("MISMATCH:", name, got.shape, got_dev.shape, want.shape, file=sys.stderr, flush=True)
    ok = ok and same


for kernel in ("mma", "popc"):
    ctx = smafa_b200.Context(local_rank, kernel)
    s = ShardedSearcher(ctx, db, L, world_size=world, rank=rank)
    for m, k in [(5, None), (None, None), (5, 10), (None, 10), (3, 1), (60, 3), (8, 200001)]:
        check(f"{kernel} m={m} k={k}", s, db, q, m, k)
    s.close()
    ctx.close()

# the overflow re-send: first block capacity 64 rows
os.environ["SMAFA_XCHG_CAP"] = "64"
ctx = smafa_b200.Context(local_rank, "auto")
s = ShardedSearcher(ctx, db, L, world_size=world, rank=rank)
check("overflow m=5", s, db, q, 5, None)
ok = ok and (rank != 0 or report["overflow m=5"][2] >= 1)
check("overflow m=None k=50", s, db, q, None, 50)
s.close()
del os.environ["SMAFA_XCHG_CAP"]

# fewer windows than ranks: some shards are empty, the db is not
tiny = db[: max(1, world - 1)]
s = ShardedSearcher(ctx, tiny, L, world_size=world, rank=rank)
for m, k in [(None, None), (20, 5), (None, 1)]:
    check(f"tiny D={tiny.shape[0]} m={m} k={k}", s, tiny, q[:300], m, k)
s.close()

# more than 2^20 queries: two slabs, query numbers continue across the boundary
small = db[:2001]
big_q = np.ascontiguousarray(np.tile(q, ((1 << 20) // Q + 2, 1))[: (1 << 20) + 777])
s = ShardedSearcher(ctx, small, L, world_size=world, rank=rank)
check("two slabs m=6", s, small, big_q, 6, None)
s.close()
ctx.close()

t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"ok": bool(t.item()), "world": world, "cases": report}))
    bad = [k for k, v in report.items() if not v[0]]
    print("failed cases:", bad, "overflow retries:", report.get("overflow m=5"), file=sys.stderr, flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
