"""GPU parity of the union-row operands of the tcgen05 scan (scan_mma.cu, UPR = 2, 3): several db windows share one
accumulator, a surviving row sends all of them to the exact re-check.  Everything is compared with the CPU oracle;
`ctx.last_mma_k` tells which operands the last scan used (union rows: K / degree per window).  The degree is forced
(SMAFA_MMA_UNION_FORCE) where a test wants a particular one, and picked by the library elsewhere."""
import os

import numpy as np
import pytest

import smafa_b200
from oracle import c_oracle
from smafa_b200 import synth

pytestmark = pytest.mark.gpu


def _context(**env):
    """A context created under SMAFA_* settings (they are read by smafa_ctx_create)."""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return smafa_b200.Context(0, "mma")
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.fixture(scope="module", params=[2, 3])
def uctx(request):
    from smafa_b200 import build
    build.build()
    c_oracle.build()
    c = _context(SMAFA_MMA_UNION=3, SMAFA_MMA_UNION_FORCE=request.param)
    c.degree = request.param
    yield c
    c.close()


def union_k(L, degree):
    return 4 * (32 if L <= 31 else 64) // degree


def check(c, db, q, L, m, k, r=None):
    d = c.upload(db, L)
    try:
        got, st = c.query(d, q, L, max_divergence=m, max_num_hits=k, limit_per_sequence=r, return_stats=True)
    finally:
        d.close()
    want = c_oracle.query(db, L, q, L, m, k, r)
    assert got.shape == want.shape, (L, m, k, r, got.shape, want.shape)
    assert (got == want).all(), (L, m, k, r)
    return st


@pytest.mark.parametrize("L,noise", [(20, 0.05), (31, 0.1), (60, 0.05), (63, 0.3)])
def test_union_accumulators_exact(uctx, L, noise):
    """Raw accumulators of the first tile: row r = windows u*r .. u*r+u-1; D = #positions where the query base equals
    any of their bases, minus max(0, need - nN_q).  The db ends inside the tile with a partly filled row."""
    u = uctx.degree
    n_db = 128 * u - u - 1                           # the last row holds u - 1 windows, one row is padding
    db_sym = synth.make_db(n_db, L=L, seed=11, noise=noise)
    q_sym = synth.make_queries(db_sym, 256, seed=12, noise=noise)
    d = uctx.upload(synth.pack_symbols(db_sym), L)
    bound = L // 5
    acc = uctx.debug_mma_dump(d, synth.pack_symbols(q_sym), bound)
    assert uctx.last_mma_k == union_k(L, u)
    padded = np.full((128 * u, L), 255, dtype=db_sym.dtype)   # 255: matches no query symbol
    padded[:n_db] = db_sym
    rows = padded.reshape(128, u, L)
    qb = q_sym[None, :, :]
    hit = np.zeros((128, 256, L), dtype=bool)
    for i in range(u):
        hit |= qb == rows[:, i][:, None, :]
    hit &= qb < 4
    nq = (q_sym == 4).sum(axis=1)
    want = hit.sum(axis=2).astype(np.int32) - np.maximum(0, np.minimum((L - bound) - nq, 127))[None, :].astype(np.int32)
    assert (acc == want).all()
    for i in range(u):                               # conservative for every window of a row
        dist = (rows[:, i][:, None, :] != q_sym[None, :, :]).sum(axis=2)
        assert (acc[dist <= bound] >= 0).all()
    d.close()


MODES = [(3, None, None), (0, None, None), (5, 10, None), (2, 50, None), (None, None, None), (None, 10, None),
         (8, 25, 2), (99, 99, None)]


@pytest.mark.parametrize("L", [9, 20, 31, 32, 33, 60, 63])
def test_union_query_matches_oracle(uctx, L):
    db_sym = synth.make_db(3001, L=L, seed=300 + L, family=8, max_subs=min(4, L), noise=0.03)   # not a multiple of 2 or 3
    q_sym = synth.make_queries(db_sym, 300, seed=400 + L, max_subs=min(6, L), noise=0.03)
    db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
    used = set()
    for m, k, r in MODES + [(L // 4, 5000, None), (L // 2, 40, None)]:
        st = check(uctx, db, q, L, m, k, r)
        assert st["kernel_used"] == 2
        used.add(uctx.last_mma_k)
    assert union_k(L, uctx.degree) in used           # every mode with a bound below L ran on the union rows (forced)
    assert len(used) == 2                            # the others (need < L/2) on the single-window operands


def test_union_ties_floods_and_tiny_shapes(uctx):
    L = 60
    one = synth.random_symbols(1, L, seed=1)
    same = synth.pack_symbols(np.repeat(one, 701, axis=0))   # every pair survives: flood path, both windows of every row
    q_sym = np.repeat(one, 320, axis=0)
    q_sym[::4] = synth.random_symbols(len(q_sym[::4]), L, seed=22)
    q = synth.pack_symbols(q_sym)
    for m, k, r in [(0, None, None), (4, 2000, None), (3, 3, 2), (5, 700, None)]:
        check(uctx, same, q, L, m, k, r)
        assert uctx.last_mma_k == union_k(L, uctx.degree)
    check(uctx, synth.pack_symbols(one), q, L, 2, None)        # D = 1: one row, one window
    check(uctx, same[:2], q, L, 2, 5)                          # D = 2
    check(uctx, same[:3], q, L, 2, 5)                          # D = 3


def test_union_survivor_rings_under_pressure(uctx):
    L = 60
    fam = synth.make_db(40_001, L=L, seed=23, family=400, max_subs=6, noise=0.0)
    qf = synth.make_queries(fam, 512, seed=24, max_subs=4, noise=0.0)
    dbw, qw = synth.pack_symbols(fam), synth.pack_symbols(qf)
    for m, k in [(12, 40_000), (14, None), (15, 300)]:
        st = check(uctx, dbw, qw, L, m, k)
        assert uctx.last_mma_k == union_k(L, uctx.degree)
    assert st["candidates"] >= 300 * 512
    uctx.set_candidate_capacity(4096)
    try:
        st = check(uctx, dbw, qw, L, 10, 6000)                 # overflow -> the batch is split and scanned again
        assert st["retries"] > 0
    finally:
        uctx.set_candidate_capacity(0)


@pytest.mark.parametrize("L,t,n", [(60, 3, 20000), (20, 1, 5000), (33, 6, 4000)])
def test_union_cluster_matches_oracle(uctx, L, t, n):
    """The centroid db grows by appends of any size: a union row is re-packed when its next window arrives."""
    sym = synth.make_cluster_input(n, L=L, seed=700 + L, family=10, max_subs=min(3, L))
    enc_all = synth.pack_symbols(sym)
    want_cof, want_nc, want_cmp = c_oracle.cluster(enc_all, L, t)
    keep = want_cof >= 0
    enc = enc_all[keep]
    remap = np.cumsum(keep) - 1
    cof, nc, ncmp = uctx.cluster(enc, L, t)
    assert nc == want_nc and ncmp == want_cmp
    assert (cof.astype(np.int64) == remap[want_cof[keep]]).all()


def test_union_append_parity(uctx):
    """smafa_db_append at row boundaries and inside rows equals one upload of the whole db."""
    L = 60
    db_sym = synth.make_db(1500, L=L, seed=41)
    q = synth.pack_symbols(synth.make_queries(db_sym, 128, seed=42))
    db = synth.pack_symbols(db_sym)
    d = uctx.upload(db[:301], L)
    for a, b in [(301, 302), (302, 304), (304, 555), (555, 1000), (1000, 1500)]:
        d.append(db[a:b])
    got = uctx.query(d, q, L, max_divergence=6, max_num_hits=7)
    assert uctx.last_mma_k == union_k(L, uctx.degree)
    d.close()
    want = c_oracle.query(db, L, q, L, 6, 7, None)
    assert got.shape == want.shape and (got == want).all()


def test_degree_is_picked_per_scan():
    """Library defaults.  Small batches follow the unrelated-window model (a row of u windows passes when
    Binomial(L, 1 - (3/4)^u) >= need: degree 3 while need >= 47 of 60, degree 2 while need >= 39); a batch large enough
    to be sampled gets what the sample says (cost model calibrated in profiles/r02_union_calib.log): windows with a
    skewed base composition match a union row far more often -- at --max-divergence 5 a db with 85 % A stays on
    single-window operands, one with 70 % A gets degree 2, uniform windows get degree 3."""
    L = 60
    c = _context(SMAFA_DB_GROUP=0)   # plain db order: degrees 1..3 (the grouped order has its own tests, test_gpu_grouped.py)
    try:
        db_sym = synth.make_db(2000, L=L, seed=51)
        db = synth.pack_symbols(db_sym)
        q = synth.pack_symbols(synth.make_queries(db_sym, 200, seed=52))
        for m, k_want in [(5, 256 // 3), (13, 256 // 3), (14, 128), (21, 128), (22, 192), (40, 192)]:
            st = check(c, db, q, L, m, 10)
            assert c.last_mma_k == k_want, (m, c.last_mma_k)
            assert st["union_degree"] == {256 // 3: 3, 128: 2, 192: 1}[k_want]
        rng = np.random.default_rng(61)

        def skewed(pA):
            r = (1 - pA) / 3
            return rng.choice(4, size=(70_000, L), p=[pA, r, r, r]).astype(np.uint8)

        got = dbc = qc = None
        for pA, k_want in [(0.85, 192), (0.70, 128), (0.25, 256 // 3)]:
            sym = skewed(pA)
            qs = synth._mutate(rng, sym[rng.integers(0, len(sym), size=32_000)], 4, 0.01)        # 2.2e9 pairs: sampled
            dbw, qw = synth.pack_symbols(sym), synth.pack_symbols(qs)
            d = c.upload(dbw, L)
            rows = c.query(d, qw, L, max_divergence=5, max_num_hits=3)
            assert c.last_mma_k == k_want, (pA, c.last_mma_k)
            rows2 = c.query(d, qw, L, max_divergence=5, max_num_hits=3)                            # the kept verdict: same degree, same rows
            assert c.last_mma_k == k_want and rows.shape == rows2.shape and (rows == rows2).all()
            d.close()
            if pA == 0.85:
                got, dbc, qc = rows, dbw, qw
        sub = np.arange(0, len(qc), 997)
        want = c_oracle.query(dbc, L, qc[sub], L, 5, 3, None)
        rows = got[np.isin(got[:, 0], sub)].copy()
        rows[:, 0] = np.searchsorted(sub, rows[:, 0])
        assert rows.shape == want.shape and (rows == want).all()
    finally:
        c.close()


def test_single_rows_still_selectable():
    """SMAFA_MMA_UNION=1 keeps the +-1 feature operands for every bound."""
    L = 60
    c = _context(SMAFA_MMA_UNION=1)
    try:
        db_sym = synth.make_db(2000, L=L, seed=51)
        q = synth.pack_symbols(synth.make_queries(db_sym, 200, seed=52))
        check(c, synth.pack_symbols(db_sym), q, L, 5, 10)
        assert c.last_mma_k == 192
    finally:
        c.close()
