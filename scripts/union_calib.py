"""Calibration data for pick_union_degree (api.cu): steady-state scan time at forced union degrees 1, 2, 3 and at the
library's own choice (force=0), with the sampled passing fractions (SMAFA_UNION_DEBUG), on db shapes that differ in how
often a union row passes: the bench generator (families of 16 spread over the db), a family-dense small db, unrelated
uniform windows, and skewed base compositions."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smafa_b200
from smafa_b200 import synth

L = 60
rng = np.random.default_rng(7)


def skewed(n, pA):
    r = (1 - pA) / 3
    return rng.choice(4, size=(n, L), p=[pA, r, r, r]).astype(np.uint8)


def shapes():
    db = synth.make_db(1_000_000, L=L, seed=synth.SEED_DB)
    yield "bench 1M/100k", db, synth.make_queries(db, 100_000, seed=synth.SEED_QUERY), [5, 10, 15, 20]
    db = synth.make_db(200_001, L=L, seed=41)
    yield "family-dense 200k/20k", db, synth.make_queries(db, 20_000, seed=42), [5, 10]
    db = rng.integers(0, 4, size=(500_000, L), dtype=np.uint8)
    yield "uniform 500k/50k", db, synth._mutate(rng, db[rng.integers(0, len(db), size=50_000)], 4, 0.01), [5, 10, 15]
    for pA in (0.55, 0.7, 0.85):
        db = skewed(500_000, pA)
        yield f"{int(pA * 100)}% A 500k/50k", db, synth._mutate(rng, db[rng.integers(0, len(db), size=50_000)], 4, 0.01), [5, 10]


only = os.environ.get("CALIB_ONLY")
for name, db_sym, q_sym, ms in shapes():
    if only and only not in name:
        continue
    db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
    for m in ms:
        for force in ("1", "2", "3", "0"):
            os.environ["SMAFA_MMA_UNION_FORCE"] = force
            if force == "0":
                os.environ["SMAFA_UNION_DEBUG"] = "1"
            else:
                os.environ.pop("SMAFA_UNION_DEBUG", None)
            c = smafa_b200.Context(0, "mma")
            d = c.upload(db, L)
            t = []
            for i in range(4):
                if i == 1:
                    os.environ.pop("SMAFA_UNION_DEBUG", None)
                got, st = c.query(d, q, L, max_divergence=m, return_stats=True)
                t.append(round(st["scan_ms"], 3))
            print(f"{name} m={m} force={force}: scan_ms {t} K/window={c.last_mma_k} cands={st['candidates']} rows={got.shape[0]}", flush=True)
            d.close()
            c.close()
