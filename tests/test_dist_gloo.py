"""World-size-2 CPU (gloo) test of the multi-GPU protocol: row-shard bounds, the fixed-capacity block
all-gather with its overflow re-send and status propagation, and the sort-free merge -- as restated in
tests/merge_model.py from smafa_b200/csrc/sharded.cu and merge.cu.  No GPU: per-shard answers come from the oracle, the
blocks travel over gloo, the merged rows must equal the oracle's on the whole db.  (The CUDA merge itself is compared
with the oracle by the -m gpu tests and by bench.py at every N.)"""
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
    sys.path.insert(0, os.path.join(%(root)r, "tests"))
    from oracle import c_oracle
    from smafa_b200 import synth
    from smafa_b200.dist import shard_bounds
    import merge_model as mm

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = 60
    db_sym = synth.make_db(2001, L=L, seed=5)                        # odd size: ragged shards
    db = synth.pack_symbols(db_sym)
    q = synth.pack_symbols(synth.make_queries(db_sym, 150, seed=6))
    lo, hi = shard_bounds(db.shape[0], world, rank)
    ok, regrown, failed_together = True, 0, False
    for m, k in [(None, None), (5, None), (None, 10), (7, 10), (None, 3000), (0, None), (3, 1), ("fail", None)]:
        status = 0
        if m == "fail":                                              # rank 1's local part fails: everybody must learn it
            m, status = 5, 11 if rank == 1 else 0
        local = c_oracle.query(db[lo:hi], L, q, L, m, k, None).astype(np.int64)
        local[:, 1] += lo                                            # global subject indices
        if status:
            local = local[:0]
        cap = 64                                                     # forces the overflow path for the big cases
        while True:
            block = torch.from_numpy(mm.make_block(local, cap, status).view(np.int64))
            gathered = torch.empty(world * (2 + cap), dtype=torch.int64)
            dist.all_gather_into_tensor(gathered, block)
            rows, need, st, who = mm.merge_blocks(gathered.numpy().view(np.uint64).reshape(world, 2 + cap), cap, 1 if k in (None, 1) else k)
            if need > cap and not st:
                cap = mm.next_cap(need)                              # identical on every rank: all saw the same headers
                regrown += 1
                continue
            break
        if status or st:
            failed_together = (st == 11 and who == 1)
            continue
        want = c_oracle.query(db, L, q, L, m, k, None)
        ok = ok and rows.shape == want.shape and bool((rows == want).all())
    t = torch.tensor([1 if ok and failed_together else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"ok": bool(t.item()), "regrown": regrown,
                          "shards": [shard_bounds(db.shape[0], world, r) for r in range(world)]}))
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


def test_merge_model_single_process():
    """The merge of 1..5 shards (some empty, D < shards included) equals the oracle on the whole db in every mode."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import merge_model as mm
    from oracle import c_oracle
    from smafa_b200 import synth
    from smafa_b200.dist import shard_bounds
    c_oracle.build()
    L = 33
    for D, R in [(700, 1), (701, 3), (3, 5), (2000, 4)]:
        db_sym = synth.make_db(D, L=L, seed=50 + D, family=8, max_subs=4)
        db = synth.pack_symbols(db_sym)
        q = synth.pack_symbols(synth.make_queries(db_sym, 90, seed=60 + D, max_subs=5))
        for m, k in [(None, None), (4, None), (None, 7), (6, 7), (None, 5000), (2, 1)]:
            blocks = []
            for r in range(R):
                lo, hi = shard_bounds(D, R, r)
                if hi == lo:                                         # an empty shard of a non-empty db sends no rows
                    blocks.append(mm.make_block(np.zeros((0, 3)), 1 << 16))
                    continue
                local = c_oracle.query(db[lo:hi], L, q, L, m, k, None).astype(np.int64)
                local[:, 1] += lo
                blocks.append(mm.make_block(local, 1 << 16))
            rows, need, st, _ = mm.merge_blocks(np.stack(blocks), 1 << 16, 1 if k in (None, 1) else k)
            want = c_oracle.query(db, L, q, L, m, k, None)
            assert st == 0 and need <= 1 << 16
            assert rows.shape == want.shape and (rows == want).all(), (D, R, m, k)


def test_two_rank_block_exchange_matches_single_db(tmp_path):
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
    assert res["regrown"] >= 2                                       # the overflow path really ran
    assert res["shards"] == [[0, 1001], [1001, 2001]]
