"""GPU cluster benchmark (config 5 shape): smafa_cluster on synthetic 60-nt windows, t=3, with an oracle
parity check on a prefix and the reference's own work count (sum_i |centroids before i|)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smafa_b200
from oracle import c_oracle
from smafa_b200 import synth

n = int(os.environ.get("CLUSTER_N", "500000"))
n_check = int(os.environ.get("CLUSTER_CHECK", "60000"))
L, t = 60, 3
sym = synth.make_cluster_input(n, L=L)
enc_all = synth.pack_symbols(sym)
_, first = np.unique(enc_all, axis=0, return_index=True)   # host-side dedup (src/cluster.rs:46-48)
enc = enc_all[np.sort(first)]
print(f"n={n} unique={enc.shape[0]}")
for kernel in ("mma", "popc"):
    ctx = smafa_b200.Context(0, kernel)
    ctx.cluster(enc[:5000], L, t)  # warm-up (allocations)
    t0 = time.perf_counter()
    cof, nc, ncmp, st = ctx.cluster(enc, L, t, return_stats=True)
    dt = time.perf_counter() - t0
    print(f"{kernel}: {dt:.3f} s wall, {nc} centroids, reference work {ncmp:.3e} comparisons -> {ncmp/dt:.3e} cmp/s; "
          f"GPU evaluated {st['pairs']:.3e} pairs, scan {st['scan_ms']:.1f} ms, launches {st['kernel_launches']}")
    if n_check:
        sub = enc[:n_check]
        t0 = time.perf_counter()
        want_cof, want_nc, want_cmp = c_oracle.cluster(sub, L, t)
        dto = time.perf_counter() - t0
        got_cof, got_nc, got_cmp = ctx.cluster(sub, L, t)
        ok = got_nc == want_nc and got_cmp == want_cmp and (got_cof.astype(np.int64) == want_cof).all()
        print(f"  parity on first {n_check}: {ok}; oracle 1 thread {dto:.2f} s = {want_cmp/dto:.3e} cmp/s")
    ctx.close()
