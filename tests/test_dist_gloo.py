"""World-size-2 CPU (gloo) test of the multi-GPU plumbing in smafa_b200/dist.py: row-shard bounds,
ragged candidate all-gather and the superset-merge argument.  No GPU: per-shard
candidates come from the oracle, the merge is the oracle's finalize -- what is under test is the
sharding / exchange / global-index logic the NCCL path shares."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %(root)r)
    from oracle import c_oracle, np_oracle
    from smafa_b200 import synth
    from smafa_b200.dist import shard_bounds, exchange_candidates

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = 60
    db = synth.pack_symbols(synth.make_db(2001, L=L, seed=5))       # odd size: ragged shards
    q = synth.pack_symbols(synth.make_queries(synth.make_db(2001, L=L, seed=5), 150, seed=6))
    lo, hi = shard_bounds(db.shape[0], world, rank)
    ok = True
    for m, k in [(None, None), (5, None), (None, 10), (7, 10), (None, 3000), (0, None)]:
        local = c_oracle.query(db[lo:hi], L, q, L, m, k, None).astype(np.int64)
        local[:, 1] += lo                                            # global subject indices
        if rank == 1 and m == 0:
            local = local[:0]                                        # an empty block must survive the exchange
            want_local_dropped = True
        # capacity 64 forces the overflow retry for the big cases, the default path for the small ones
        union, biggest = exchange_candidates(torch.from_numpy(local.astype(np.int32)), capacity=64 if k != 10 else 1 << 16)
        ok = ok and biggest >= local.shape[0]
        merged = np_oracle.finalize_candidates([tuple(int(x) for x in r) for r in union.numpy()], m, k)
        if m == 0:
            continue                                                 # rank 1 withheld rows on purpose
        want = c_oracle.query(db, L, q, L, m, k, None)
        same = len(merged) == want.shape[0] and all(tuple(int(x) for x in w) == g for w, g in zip(want, merged))
        ok = ok and same
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"ok": bool(t.item()), "shards": [shard_bounds(db.shape[0], world, r) for r in range(world)]}))
    dist.destroy_process_group()
""")


def test_shard_bounds_cover_and_are_contiguous():
    sys.path.insert(0, ROOT)
    from smafa_b200.dist import shard_bounds
    for D in (0, 1, 7, 8, 1000001):
        for ws in (1, 2, 3, 8):
            spans = [shard_bounds(D, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == D
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_candidate_merge_matches_single_db(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    import json
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["ok"], res
    assert res["shards"] == [[0, 1001], [1001, 2001]]
