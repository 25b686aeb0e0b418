"""Round-2 validation of the similarity-grouped db order (SMAFA_DB_GROUP=1, api.cu group_order).
NOT part of tests/: this path has not run on a GPU yet.  Run under gpurun on one B200:

    python scripts/grouped_rows_check.py            # parity vs the oracle, then timings next to the default order

Checks: rows of every selection mode equal the oracle's (whole output on a 200 k x 3 k case, query subsample at
1 M x 100 k), smafa_distances is un-permuted correctly, and prints which union degree each scan picked."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smafa_b200
from oracle import c_oracle
from smafa_b200 import synth


def context(group):
    old = os.environ.get("SMAFA_DB_GROUP")
    os.environ["SMAFA_DB_GROUP"] = "1" if group else "0"
    try:
        return smafa_b200.Context(0, "mma")
    finally:
        if old is None:
            os.environ.pop("SMAFA_DB_GROUP", None)
        else:
            os.environ["SMAFA_DB_GROUP"] = old


def main():
    c_oracle.build()
    L = 60
    ok = True
    g = context(True)
    # 1. whole-output parity
    db_sym = synth.make_db(200_001, L=L, seed=41)
    db = synth.pack_symbols(db_sym)
    q = synth.pack_symbols(synth.make_queries(db_sym, 3000, seed=42))
    t0 = time.perf_counter()
    d = g.upload(db, L)
    print("grouped upload of 200 k windows: %.3f s" % (time.perf_counter() - t0))
    for m, k, r in [(5, None, None), (None, None, None), (5, 10, None), (None, 10, None), (3, 1, None), (8, 25, 2), (0, None, None)]:
        got = g.query(d, q, L, max_divergence=m, max_num_hits=k, limit_per_sequence=r)
        want = c_oracle.query(db, L, q, L, m, k, r, threads=os.cpu_count() or 1)
        same = got.shape == want.shape and bool((got == want).all())
        print(f"m={m} k={k} r={r}: {'ok' if same else 'MISMATCH'} rows={got.shape[0]} K/window of the last scan={g.last_mma_k}")
        ok &= same
    dist = g.distances(d, q[:4], L)
    for i in range(4):
        ok &= bool((dist[i].astype(np.int64) == c_oracle.distances(db, q[i])).all())
    print("distances un-permuted:", ok)
    d.close()
    # 2. configs[1] shape: subsample parity + timing next to the plain order
    db_sym = synth.make_db(1_000_000, L=L, seed=synth.SEED_DB)
    db = synth.pack_symbols(db_sym)
    q = synth.pack_symbols(synth.make_queries(db_sym, 100_000, seed=synth.SEED_QUERY))
    sub = np.arange(0, len(q), 1009)
    want = c_oracle.query(db, L, q[sub], L, 5, None, None, threads=os.cpu_count() or 1)
    for name, ctx in (("grouped", g), ("plain", context(False))):
        t0 = time.perf_counter()
        d = ctx.upload(db, L)
        t_up = time.perf_counter() - t0
        for _ in range(3):
            got, st = ctx.query(d, q, L, max_divergence=5, return_stats=True)
        rows = got[np.isin(got[:, 0], sub)].copy()
        rows[:, 0] = np.searchsorted(sub, rows[:, 0])
        same = rows.shape == want.shape and bool((rows == want).all())
        ok &= same
        print(f"{name}: upload {t_up:.3f} s, scan {st['scan_ms']:.3f} ms, candidates {st['candidates']}, K/window {ctx.last_mma_k}, "
              f"subsample {'ok' if same else 'MISMATCH'}")
        d.close()
    print("ALL OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
