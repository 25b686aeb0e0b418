"""Where a multi-GPU step spends its time (run under torchrun): per rank, the local scan (smafa_stats.scan_ms), the
whole call (total_ms, CUDA events) and the exchange + merge (exchange_ms: from the end of the rank's local part to the end
of the merge -- includes waiting for the slowest rank, so the minimum over ranks is the pure cost)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smafa_b200
from smafa_b200 import synth
from smafa_b200.dist import ShardedSearcher, shard_bounds

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
L = 60
D = 1_000_000 * world
db_sym = synth.make_db(D, L=L, seed=synth.SEED_DB)
q = synth.pack_symbols(synth.make_queries(db_sym, 100_000, seed=synth.SEED_QUERY))
lo, hi = shard_bounds(D, world, rank)
ctx = smafa_b200.Context(lr, "auto")
s = ShardedSearcher(ctx, synth.pack_symbols(db_sym[lo:hi]), L, world_size=world, rank=rank, presharded=True, shard_offset=lo, total_rows=D)
del db_sym
qd = torch.from_numpy(q.view(np.int64)).to(dev)
for _ in range(5):
    s.query_dev(qd, 5, None)
N = 20
acc = np.zeros(3)
for _ in range(N):
    rows = s.query_dev(qd, 5, None)
    st = s.last_stats
    acc += [st["scan_ms"], st["total_ms"], st["exchange_ms"]]
acc /= N
t = torch.tensor(acc, dtype=torch.float64, device=dev)
lo_t, hi_t = t.clone(), t.clone()
dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
print(f"rank {rank}: per step ms: scan {acc[0]:.3f} call {acc[1]:.3f} exchange+merge {acc[2]:.3f}; rows {rows.shape[0]}", flush=True)
dist.barrier()
if rank == 0:
    print(f"world {world}: scan min/max {lo_t[0]:.3f}/{hi_t[0]:.3f} ms, call min/max {lo_t[1]:.3f}/{hi_t[1]:.3f} ms, "
          f"exchange+merge min/max {lo_t[2]:.3f}/{hi_t[2]:.3f} ms (min = pure exchange + merge)")
dist.destroy_process_group()
