"""Steady-state scan time of one batch at forced union degrees (calibration of pick_union_degree's cost per verified
window) on a family-dense db: 200 k windows in 12.5 k families, 20 k queries, --max-divergence 5."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smafa_b200
from smafa_b200 import synth

L = 60
db_sym = synth.make_db(200_001, L=L, seed=41)
db = synth.pack_symbols(db_sym)
q = synth.pack_symbols(synth.make_queries(db_sym, 20_000, seed=42))
for force in ("0", "1", "2", "3"):
    os.environ["SMAFA_MMA_UNION_FORCE"] = force
    c = smafa_b200.Context(0, "mma")
    d = c.upload(db, L)
    ms = []
    for _ in range(4):
        got, st = c.query(d, q, L, max_divergence=5, return_stats=True)
        ms.append(round(st["scan_ms"], 3))
    print(f"force={force}: scan_ms {ms} K/window={c.last_mma_k} cands={st['candidates']} rows={got.shape[0]}", flush=True)
    d.close()
    c.close()
