"""CPU model (numpy, no GPU) of how often a union row passes the filter at --max-divergence 5 when its windows are
(a) neighbours in db order (unrelated, what v9 does) or (b) members of one family (what a similarity-ordered db would
give).  Synthetic db of SURVEY 8d at 1/16 of configs[1] (64 k windows, families of 16), 3000 queries.  Printed: the
fraction of (query, row) pairs whose accumulator would be >= 0, i.e. verified windows per comparison.
Used for DESIGN.md section 11, item 1."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smafa_b200 import synth

L, D, need = 60, 64000, 55
db = synth.make_db(D, L=L, seed=synth.SEED_DB)          # window i descends from root i mod R
q = synth.make_queries(db, 3000, seed=synth.SEED_QUERY)
R = D // 16


def onehot(s):
    return (s[..., None] == np.arange(4)).astype(np.uint8)  # N -> all zero


Q1 = onehot(q).reshape(len(q), -1).astype(np.int32)
nNq = (q == 4).sum(1)


def passing(rows_idx):
    U = np.zeros((rows_idx.shape[0], L * 4), dtype=np.int32)
    for i in range(rows_idx.shape[1]):
        U |= onehot(db[rows_idx[:, i]]).reshape(-1, L * 4)
    return float(((Q1 @ U.T) + nNq[:, None] >= need).mean())


for u in (1, 2, 3, 4, 8, 16):
    n_rows = min(D // u, 4000)
    adj = np.arange(n_rows * u).reshape(n_rows, u)
    fam = np.arange(n_rows)[:, None] % R + (np.arange(u)[None, :] + (np.arange(n_rows)[:, None] // R) * u) % 16 * R
    print("u=%2d  neighbours in db order: %.2e   one family per row: %.2e" % (u, passing(adj), passing(fam)))
