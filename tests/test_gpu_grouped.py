"""GPU parity of the similarity-grouped db order (api.cu group_order) and of the wide union rows it enables (scan_mma.cu,
UPR = 4, 8, 16: up to sixteen db windows behind one accumulator).  A db of >= 65536 nucleotide windows is stored on the
device with similar windows next to each other; every row keeps its subject number, so results must not change: everything
here is compared with the CPU oracle on the db in its ORIGINAL order."""
import os

import numpy as np
import pytest

import smafa_b200
from oracle import c_oracle
from smafa_b200 import synth

pytestmark = pytest.mark.gpu
L = 60


def _context(**env):
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


@pytest.fixture(scope="module")
def data():
    from smafa_b200 import build
    build.build()
    c_oracle.build()
    db_sym = synth.make_db(100_003, L=L, seed=141)                 # families of 16 spread over the db, not a multiple of 16
    q_sym = synth.make_queries(db_sym, 1000, seed=142)
    return synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)


MODES = [(5, None, None), (None, None, None), (5, 10, None), (None, 10, None), (3, 1, None), (8, 25, 2), (0, None, None),
         (12, 7, None), (60, 3, None)]
_WANT = {}


def oracle(db, q, m, k, r=None):
    """the oracle's rows on the db in its original order, computed once per mode (all host threads)"""
    key = (db.shape[0], q.shape[0], m, k, r)
    if key not in _WANT:
        _WANT[key] = c_oracle.query(db, L, q, L, m, k, r, threads=os.cpu_count() or 1)
    return _WANT[key]


@pytest.mark.parametrize("force", [0, 4, 8, 16])
def test_grouped_query_matches_oracle(data, force):
    """Library's own degree choice (0) and every wide degree forced: rows equal the oracle's in every selection mode."""
    db, q = data
    c = _context(SMAFA_MMA_UNION_FORCE=force) if force else smafa_b200.Context(0, "mma")
    try:
        d = c.upload(db, L)
        degrees = set()
        for m, k, r in MODES:
            got, st = c.query(d, q, L, max_divergence=m, max_num_hits=k, limit_per_sequence=r, return_stats=True)
            want = oracle(db, q, m, k, r)
            assert got.shape == want.shape and (got == want).all(), (force, m, k, r)
            degrees.add(st["union_degree"])
        if force:
            assert force in degrees                                  # bounds below L/2 ran on the forced degree
        dist = c.distances(d, q[:5], L)                              # get_distances comes back in subject order
        for i in range(5):
            assert (dist[i].astype(np.int64) == c_oracle.distances(db, q[i])).all()
        d.close()
    finally:
        c.close()


def test_group_order_is_a_row_aligned_permutation(data):
    db, _ = data
    c = smafa_b200.Context(0, "mma")
    try:
        perm, clusters = c.group_order(db, L)
        assert clusters > 0 and sorted(perm.tolist()) == list(range(db.shape[0]))
        # the generator's families (window i belongs to root i mod R) come out as runs of consecutive rows ...
        R = db.shape[0] // 16
        fam = (perm.astype(np.int64) % R)
        runs = 1 + int((np.diff(fam) != 0).sum())
        assert runs <= 1.2 * R
        # ... and 16-row operand rows rarely straddle two of them
        rows = fam[: len(fam) // 16 * 16].reshape(-1, 16)
        mixed = (rows != rows[:, :1]).any(axis=1).mean()
        assert mixed < 0.25, mixed
        small = synth.pack_symbols(synth.random_symbols(70_000, L, seed=5))      # unrelated windows: nothing to group
        perm2, clusters2 = c.group_order(small, L)
        assert clusters2 == 0 and (perm2 == np.arange(70_000)).all()
    finally:
        c.close()


def test_grouped_db_ties_floods_overflow_and_append(data):
    db, q = data
    c = smafa_b200.Context(0, "mma")
    try:
        # many exact duplicates of a few windows: ties at every distance, whole families pass together
        rng = np.random.default_rng(7)
        dup_sym = synth.make_db(500, L=L, seed=143)[rng.integers(0, 500, size=80_000)]
        dup = synth.pack_symbols(dup_sym)
        d = c.upload(dup, L)
        qs = synth.pack_symbols(synth.make_queries(dup_sym, 300, seed=144, max_subs=4))
        for m, k in [(0, None), (3, None), (6, 40), (None, 5)]:
            got, st = c.query(d, qs, L, max_divergence=m, max_num_hits=k, return_stats=True)
            want = c_oracle.query(dup, L, qs, L, m, k, None, threads=os.cpu_count() or 1)
            assert got.shape == want.shape and (got == want).all(), (m, k)
        c.set_candidate_capacity(4096)                               # candidate overflow: the batch is split and scanned again
        try:
            got, st = c.query(d, qs, L, max_divergence=4, max_num_hits=2000, return_stats=True)
            assert st["retries"] > 0
            want = c_oracle.query(dup, L, qs, L, 4, 2000, None, threads=os.cpu_count() or 1)
            assert got.shape == want.shape and (got == want).all()
        finally:
            c.set_candidate_capacity(0)
        d.close()
        # windows appended to a grouped db go behind it under the next subject numbers
        d = c.upload(db[:90_000], L)
        d.append(db[90_000:90_007])
        d.append(db[90_007:])
        got = c.query(d, q, L, max_divergence=6, max_num_hits=4)
        want = oracle(db, q, 6, 4)
        assert got.shape == want.shape and (got == want).all()
        d.close()
    finally:
        c.close()


def test_mapped_shards_cover_the_db(data):
    """What a one-process-per-GPU run does (smafa_b200/dist.py): the whole db is grouped once, the grouped order is cut
    into shards, every shard is uploaded under its rows' subject numbers.  The union of the shards' answers, cut at the
    global k-th distance, is the oracle's answer on the whole db."""
    db, q = data
    c = smafa_b200.Context(0, "mma")
    try:
        perm, clusters = c.group_order(db, L)
        assert clusters > 0
        D = db.shape[0]
        cuts = [0, D // 3, D // 3 + 70_001 if D // 3 + 70_001 < D else D - 1, D]
        shards = [c.upload_mapped(np.ascontiguousarray(db[perm[a:b]]), L, perm[a:b], D) for a, b in zip(cuts, cuts[1:])]
        for m, k in [(5, None), (None, 10), (12, 7)]:
            parts = [c.query(s, q, L, max_divergence=m, max_num_hits=k) for s in shards]
            rows = np.concatenate(parts).astype(np.int64)
            order = np.lexsort((rows[:, 1], rows[:, 2], rows[:, 0]))
            rows = rows[order]
            keep = np.zeros(len(rows), dtype=bool)
            kk = 1 if k in (None, 1) else k
            start = 0
            for qn in np.unique(rows[:, 0]):
                seg = rows[rows[:, 0] == qn]
                cutoff = seg[min(kk, len(seg)) - 1, 2]
                keep[start:start + len(seg)] = seg[:, 2] <= cutoff
                start += len(seg)
            got = rows[keep].astype(np.uint32)
            want = oracle(db, q, m, k)
            assert got.shape == want.shape and (got == want).all(), (m, k)
        for s in shards:
            s.close()
    finally:
        c.close()


def test_plain_order_still_selectable(data):
    """SMAFA_DB_GROUP=0: no grouping, union rows of at most three windows, same rows."""
    db, q = data
    c = _context(SMAFA_DB_GROUP=0)
    try:
        d = c.upload(db, L)
        got, st = c.query(d, q, L, max_divergence=5, return_stats=True)
        assert st["union_degree"] <= 3
        want = oracle(db, q, 5, None)
        assert got.shape == want.shape and (got == want).all()
        d.close()
    finally:
        c.close()
