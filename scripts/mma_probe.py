"""GPU probe for the tcgen05 kernel: raw accumulators vs numpy, then a small query vs the POPC kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SMAFA_MMA_UNION", "1")  # this probe models the single-window operands
import smafa_b200
from smafa_b200 import synth

L = int(os.environ.get("PROBE_L", "60"))
ctx = smafa_b200.Context(0, "mma")
db_sym = synth.make_db(1000, L=L, seed=1)
q_sym = synth.make_queries(db_sym, 256, seed=2)
db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
d = ctx.upload(db, L)
bound = 7
acc = ctx.debug_mma_dump(d, q, bound)
nsym = int(os.environ.get("SMAFA_MMA_NSYM", "4"))
eq = db_sym[:128, None, :] == q_sym[None, :, :]
if nsym == 5:
    matches = eq.sum(axis=2).astype(np.int32)
    want = matches - (L - bound)
else:  # base-base matches only; bias = -(need - nN_q), clamped at 0
    matches = (eq & (q_sym[None, :, :] < 4)).sum(axis=2).astype(np.int32)
    want = matches - np.maximum(0, (L - bound) - (q_sym == 4).sum(axis=1))[None, :].astype(np.int32)
ok = (acc == want)
print("accumulator tile exact:", bool(ok.all()), "mismatching cells:", int((~ok).sum()))
if not ok.all():
    print("got[0,:8]", acc[0, :8], "want[0,:8]", want[0, :8])
    print("got[:8,0]", acc[:8, 0], "want[:8,0]", want[:8, 0])
    print("got==want.T?", bool((acc[:128, :128] == want[:128, :128].T).all()))
    print("distribution of got-want:", np.unique(acc - want, return_counts=True))
for m, k in [(5, None), (None, None), (5, 10), (None, 10)]:
    ctx.set_kernel("mma")
    a, st = ctx.query(d, q, L, m, k, return_stats=True)
    ctx.set_kernel("popc")
    b = ctx.query(d, q, L, m, k)
    print("query m=%s k=%s: mma==popc %s rows %d/%d kernel_used=%d cands=%d" % (
        m, k, a.shape == b.shape and bool((a == b).all()), a.shape[0], b.shape[0], st["kernel_used"], st["candidates"]))
sys.exit(0 if ok.all() else 3)
