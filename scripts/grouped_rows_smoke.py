"""30-second GPU smoke of the grouped db order: the same queries against the same db in plain and in grouped order
(SMAFA_DB_GROUP=1) must give identical rows.  See scripts/grouped_rows_check.py for the full check against the oracle."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smafa_b200
from smafa_b200 import synth

L = 60
db_sym = synth.make_db(200_001, L=L, seed=41)
db = synth.pack_symbols(db_sym)
q = synth.pack_symbols(synth.make_queries(db_sym, 20_000, seed=42))
res = {}
for group in ("0", "1"):
    os.environ["SMAFA_DB_GROUP"] = group
    c = smafa_b200.Context(0, "mma")
    t0 = time.perf_counter()
    d = c.upload(db, L)
    t_up = time.perf_counter() - t0
    for m, k in [(5, None), (5, 10), (None, None)]:
        got, st = c.query(d, q, L, max_divergence=m, max_num_hits=k, return_stats=True)
        res[(group, m, k)] = got
        print(f"group={group} m={m} k={k}: rows={got.shape[0]} scan_ms={st['scan_ms']:.3f} cands={st['candidates']} K/window={c.last_mma_k} upload={t_up:.3f}s",
              flush=True)
    d.close()
    c.close()
ok = all(res[("0", m, k)].shape == res[("1", m, k)].shape and bool((res[("0", m, k)] == res[("1", m, k)]).all())
         for m, k in [(5, None), (5, 10), (None, None)])
print("grouped == plain:", ok)
