"""CPU check of the inequality the union-row filter of the tcgen05 scan rests on (scan_mma.cu UPR, guess.cu
union_sample_kernel), stated on the reference's own word layout (src/lib.rs:29-52, 167-184) and checked against the
oracle's distances:

    for every window w_i of a union row:   L - distance(q, w_i)  <=  popcount(q & (w_1 | ... | w_u) & BASE) + nN_q

so "sum >= need" (the sign test of the accumulator, D >= 0) can never reject a window with distance <= bound.  No GPU."""
import numpy as np
import pytest

from oracle import np_oracle
from smafa_b200 import synth

NBITS = np.uint64(0x0084210842108421)   # bit 0 of every 5-bit group: code 1 = N / gap / IUPAC


def popcount(x):
    return np.unpackbits(np.ascontiguousarray(x).view(np.uint8), axis=-1).reshape(*x.shape, 64).sum(axis=-1)


@pytest.mark.parametrize("L,u,noise", [(60, 2, 0.02), (60, 3, 0.02), (20, 3, 0.1), (63, 2, 0.3), (9, 3, 0.0), (31, 2, 0.05)])
def test_union_row_sum_bounds_every_window(L, u, noise):
    db_sym = synth.make_db(60 * u, L=L, seed=5 + L, family=6, max_subs=min(4, L), noise=noise)
    q_sym = synth.make_queries(db_sym, 40, seed=6 + L, max_subs=min(6, L), noise=noise)
    db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
    rows = db.reshape(60, u, -1)
    union = np.bitwise_or.reduce(rows, axis=1)                                   # [rows][W]
    base = ~NBITS
    nN_q = popcount(q & NBITS).sum(axis=1)                                       # [Q]
    s = popcount(q[:, None, :] & union[None, :, :] & base).sum(axis=2) + nN_q[:, None]   # [Q][rows]
    for qi in range(len(q)):
        dist = np_oracle.distances(db, q[qi]).reshape(60, u)                     # [rows][u]
        matches = L - dist
        assert (matches <= s[qi][:, None]).all()
        # and for u = 1 the sum is exact up to the N-N bound: matches - nNN <= base matches
    single = popcount(q[:, None, :] & db[None, :, :] & base).sum(axis=2)         # base matches, one window
    for qi in range(len(q)):
        matches = L - np_oracle.distances(db, q[qi])
        nNN = popcount(q[qi][None, :] & db & NBITS).sum(axis=1)
        assert (matches == single[qi] + nNN).all()                               # N-vs-N counts as a match


def test_union_pass_probability_of_unrelated_uniform_windows():
    """The thresholds small batches use (need >= 3L/4 for two windows, 7L/8 for three) come from the per-position pass
    probability 1 - (3/4)^u of unrelated uniform windows: check the model against a sample."""
    rng = np.random.default_rng(3)
    L, n = 60, 20000
    q = rng.integers(0, 4, size=(n, L))
    for u, p in [(1, 0.25), (2, 7 / 16), (3, 37 / 64)]:
        w = rng.integers(0, 4, size=(u, n, L))
        hit = (w == q[None]).any(axis=0).sum(axis=1)
        assert abs(hit.mean() - p * L) < 0.15
        sigma = (L * p * (1 - p)) ** 0.5
        assert abs(hit.std() - sigma) < 0.1
