"""Per-phase host timings of one multi-GPU step (run under torchrun with SMAFA_TIMING=1): local scan + selection,
candidate exchange (NCCL all-gather), merge.  Measurement aid: a device sync follows every phase."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smafa_b200
from smafa_b200 import synth
from smafa_b200.dist import ShardedSearcher

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
L = 60
db0 = synth.make_db(1_000_000, L=L, seed=synth.SEED_DB)
q = synth.pack_symbols(synth.make_queries(db0, 100_000, seed=synth.SEED_QUERY))
shard = db0 if rank == 0 else synth.make_db(1_000_000, L=L, seed=synth.SEED_DB + 7919 * rank)
ctx = smafa_b200.Context(lr, "auto")
s = ShardedSearcher(ctx, synth.pack_symbols(shard), L, world_size=world, rank=rank, presharded=True)
qd = torch.from_numpy(q.view(np.int64)).to(dev)
for _ in range(3):
    s.query_dev(qd, 5, None)
s.phase_ms = [0.0, 0.0, 0.0]
N = 10
for _ in range(N):
    s.query_dev(qd, 5, None)
print(rank, "per step ms: local %.3f exchange %.3f merge %.3f; scan_ms %.3f total_ms (local call, events) %.3f" % (
    s.phase_ms[0] / N, s.phase_ms[1] / N, s.phase_ms[2] / N, s.last_stats["scan_ms"], s.last_stats["total_ms"]))
dist.destroy_process_group()
