"""Grouped db order at several clustering thresholds (SMAFA_DB_GROUP_T) x forced union degrees: bench db, m = 5."""
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
m = 5
for tg in os.environ.get("PROBE_T", "15,20,24").split(","):
    os.environ["SMAFA_DB_GROUP"] = "1"
    os.environ["SMAFA_DB_GROUP_T"] = tg
    for force in ("0", "4", "8", "16"):
        os.environ["SMAFA_MMA_UNION_FORCE"] = force
        c = smafa_b200.Context(0, "mma")
        t0 = time.perf_counter()
        if force == "0":
            os.environ["SMAFA_UNION_DEBUG"] = "1"
        d = c.upload(db, L)
        os.environ.pop("SMAFA_UNION_DEBUG", None)
        t_up = time.perf_counter() - t0
        t = []
        for i in range(4):
            if force == "0" and i == 0:
                os.environ["SMAFA_UNION_DEBUG"] = "1"
            got, st = c.query(d, q, L, max_divergence=m, return_stats=True)
            os.environ.pop("SMAFA_UNION_DEBUG", None)
            t.append(round(st["scan_ms"], 3))
        print(f"t_group={tg} force={force}: scan_ms {t} degree={st['union_degree']} cands={st['candidates']} rows={got.shape[0]} upload={t_up:.2f}s",
              flush=True)
        d.close()
        c.close()
