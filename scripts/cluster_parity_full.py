"""`smafa cluster` against the oracle on the WHOLE list (BASELINE configs[4] shape at CLUSTER_N sequences, default 500 k):
membership of every sequence, centroid count and the reference's comparison count must be identical (src/cluster.rs:45-74)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smafa_b200
from oracle import c_oracle
from smafa_b200 import synth

n = int(os.environ.get("CLUSTER_N", "500000"))
L, t = 60, 3
c_oracle.build()
sym = synth.make_cluster_input(n, L=L, seed=synth.SEED_CLUSTER)
enc_all = synth.pack_symbols(sym)
t0 = time.perf_counter()
want_cof, want_nc, want_cmp = c_oracle.cluster(enc_all, L, t)
t_cpu = time.perf_counter() - t0
keep = want_cof >= 0                      # the oracle marks duplicates (src/cluster.rs:46-48) with -1
remap = np.cumsum(keep) - 1
ctx = smafa_b200.Context(0)
for rep in range(2):
    t0 = time.perf_counter()
    cof, nc, ncmp, st = ctx.cluster(enc_all[keep], L, t, return_stats=True)
    t_gpu = time.perf_counter() - t0
same = nc == want_nc and ncmp == want_cmp and bool((cof.astype(np.int64) == remap[want_cof[keep]]).all())
print(f"cluster {n} sequences ({int(keep.sum())} unique), t = {t}: {nc} centroids, {ncmp} reference comparisons; "
      f"oracle (1 thread) {t_cpu:.1f} s, GPU {t_gpu:.3f} s ({ncmp / t_gpu:.3e} cmp/s); membership of ALL sequences identical: {same}")
ctx.close()
sys.exit(0 if same else 1)
