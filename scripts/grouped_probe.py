"""Scan time of the similarity-grouped db order (SMAFA_DB_GROUP=1) at forced union degrees next to the plain order, with
the sampled passing fractions (SMAFA_UNION_DEBUG): bench db (1 M windows in families of 16), 100 k queries."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smafa_b200
from smafa_b200 import synth

L = 60
db_sym = synth.make_db(1_000_000, L=L, seed=synth.SEED_DB)
db = synth.pack_symbols(db_sym)
q = synth.pack_symbols(synth.make_queries(db_sym, 100_000, seed=synth.SEED_QUERY))
for m in [int(x) for x in os.environ.get("PROBE_M", "5,10").split(",")]:
    for group, force in [("0", "0"), ("1", "0"), ("1", "3"), ("1", "4"), ("1", "8"), ("1", "16")]:
        os.environ["SMAFA_DB_GROUP"] = group
        os.environ["SMAFA_MMA_UNION_FORCE"] = force
        os.environ["SMAFA_UNION_DEBUG"] = "1"
        c = smafa_b200.Context(0, "mma")
        t0 = time.perf_counter()
        d = c.upload(db, L)
        t_up = time.perf_counter() - t0
        t = []
        for i in range(4):
            got, st = c.query(d, q, L, max_divergence=m, return_stats=True)
            t.append(round(st["scan_ms"], 3))
            os.environ.pop("SMAFA_UNION_DEBUG", None)
        print(f"m={m} group={group} force={force}: scan_ms {t} degree={st['union_degree']} cands={st['candidates']} rows={got.shape[0]} upload={t_up:.2f}s",
              flush=True)
        d.close()
        c.close()
