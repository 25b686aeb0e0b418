"""Multi-GPU parity check (run under torchrun, one rank per GPU): row-sharded db + NCCL candidate
all-gather + device merge must equal the oracle on the whole db."""
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
ok = True
report = {}
for kernel in ("mma", "popc"):
    ctx = smafa_b200.Context(local_rank, kernel)
    s = ShardedSearcher(ctx, db, L, world_size=world, rank=rank)
    qp = torch.from_numpy(q.view(np.int64)).pin_memory()
    for m, k in [(5, None), (None, None), (5, 10), (None, 10), (3, 1)]:
        got = s.query_host(qp, m, k)
        if rank == 0:
            want = c_oracle.query(db, L, q, L, m, k, None, threads=os.cpu_count() or 1)
            same = got.shape == want.shape and bool((got == want).all())
            report[f"{kernel} m={m} k={k}"] = [same, int(got.shape[0])]
            ok = ok and same
    s.close()
    ctx.close()
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"ok": bool(t.item()), "world": world, "cases": report}))
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
