"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the C ABI
(libsmafa_b200.so via ctypes) or the `smafa` CLI and is compared bit-exactly with the CPU oracle
or with the reference's golden vectors."""
import os
import subprocess

import numpy as np
import pytest

import smafa_b200
from oracle import c_oracle
from smafa_b200 import api, synth

pytestmark = pytest.mark.gpu

KERNELS = ["popc", "mma"]


@pytest.fixture(scope="module")
def ctx():
    from smafa_b200 import build
    build.build()
    c_oracle.build()
    c = smafa_b200.Context(0)
    yield c
    c.close()


def cli(*args):
    return subprocess.run([api.CLI_PATH, *map(str, args)], capture_output=True, text=True)


def check_query(ctx, db, q, L, m, k, r=None, kernel="popc"):
    ctx.set_kernel(kernel)
    d = ctx.upload(db, L)
    try:
        got, st = ctx.query(d, q, L, max_divergence=m, max_num_hits=k, limit_per_sequence=r, return_stats=True)
    finally:
        d.close()
    want = c_oracle.query(db, L, q, L, m, k, r, threads=os.cpu_count() or 1)
    assert got.shape == want.shape, (L, m, k, r, kernel, got.shape, want.shape)
    assert (got == want).all(), (L, m, k, r, kernel)
    return st


# ---- the reference's own golden vectors, through the CLI --------------------------------------

def test_reference_query_kats(ctx, kats, kat_dir, tmp_path):
    for kernel in KERNELS:
        for case in kats["query"]:
            if "makedb_from" in case:
                db = tmp_path / (case["name"] + ".db")
                assert cli("makedb", "-i", kat_dir / case["makedb_from"], "-d", db).returncode == 0
            else:
                db = kat_dir / case["db"]
            r = cli("query", "-d", db, "-q", kat_dir / case["query"], "--kernel", kernel, *case["args"])
            assert r.returncode == 0, (case["name"], r.stderr)
            assert r.stdout == case["stdout"], (case["name"], kernel)


def test_reference_cluster_kats(ctx, kats, kat_dir):
    for case in kats["cluster"]:
        r = cli("cluster", "-i", kat_dir / case["input"], "-d", case["t"])
        assert r.returncode == 0, r.stderr
        assert r.stdout == case["stdout"], case["name"]


def test_cli_panics_match_reference(ctx, kat_dir, tmp_path):
    db = kat_dir / "random_3_2.fna.smafadb"
    q = kat_dir / "random_3_2.fna"
    ragged = tmp_path / "ragged.fna"
    ragged.write_text(">a\nCTT\n>b\nACGT\n")
    r = cli("query", "-d", db, "-q", ragged)
    assert r.returncode == 101
    assert "Cannot compute distances between seq of length 4 and windows of lengths 3" in r.stderr
    assert r.stdout == "0\t0\t0\tCTT\n"  # the first record was answered before the panic
    r = cli("query", "-d", db, "-q", q, "--limit-per-sequence", "1")
    assert r.returncode == 101 and "limit_per_sequence" in r.stderr
    r = cli("query", "-d", db, "-q", q, "--max-num-hits", "0")
    assert r.returncode == 101
    bad = tmp_path / "bad.fna"
    bad.write_text(">ok\nCTT\n>x\nCEG\n")
    r = cli("query", "-d", db, "-q", bad)
    assert r.returncode == 101 and "Byte 69 cannot be interpreted as nucleotide" in r.stderr
    assert r.stdout == "0\t0\t0\tCTT\n"
    r = cli("cluster", "-i", ragged, "-d", "1")
    assert r.returncode == 101 and r.stdout == "CTT\tCTT\n"


def test_cli_full_output_diff_60nt(ctx, tmp_path):
    # config 1: 1k x 10k 60-nt file, complete stdout diff against the oracle CLI
    db_sym = synth.make_db(10000, L=60, seed=21)
    q_sym = synth.make_queries(db_sym, 1000, seed=22)
    synth.write_fasta(tmp_path / "db.fna", synth.to_ascii(db_sym))
    synth.write_fasta(tmp_path / "q.fna", synth.to_ascii(q_sym))
    assert cli("makedb", "-i", tmp_path / "db.fna", "-d", tmp_path / "db").returncode == 0
    for args in [[], ["--max-divergence", "5"], ["--max-num-hits", "10"],
                 ["--max-num-hits", "10", "--max-divergence", "7", "--limit-per-sequence", "1"]]:
        want = subprocess.run([c_oracle.CLI, "query", "-d", tmp_path / "db", "-q", tmp_path / "q.fna", *args],
                              capture_output=True, text=True)
        for kernel in KERNELS:
            got = cli("query", "-d", tmp_path / "db", "-q", tmp_path / "q.fna", "--kernel", kernel, *args)
            assert got.returncode == 0, got.stderr
            assert got.stdout == want.stdout, (args, kernel)
    want = subprocess.run([c_oracle.CLI, "cluster", "-i", tmp_path / "q.fna", "-d", "3"], capture_output=True, text=True)
    got = cli("cluster", "-i", tmp_path / "q.fna", "-d", "3")
    assert got.returncode == 0 and got.stdout == want.stdout


# ---- get_distances -------------------------------------------------------------------------------

@pytest.mark.parametrize("L", [1, 3, 4, 6, 9, 12, 13, 20, 32, 33, 60, 61, 64, 65, 100])
def test_distances_bit_exact(ctx, L):
    db_sym = synth.make_db(777, L=L, seed=100 + L, family=8, max_subs=min(4, L), noise=0.05)
    q_sym = synth.make_queries(db_sym, 33, seed=200 + L, max_subs=min(6, L), noise=0.05)
    db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
    d = ctx.upload(db, L)
    got = ctx.distances(d, q, L)
    d.close()
    for i in range(q.shape[0]):
        assert (got[i].astype(np.int64) == c_oracle.distances(db, q[i])).all()


# ---- raw tcgen05 accumulators (pins operand layout + descriptors independently of the epilogue) ----

def _mma_accumulator_model(db_sym, q_sym, L, bound, enc):
    """What the tcgen05 accumulators must hold for operand encoding `enc` (scan_mma.cu, "operand encodings")."""
    need = L - bound
    eq = db_sym[:, None, :] == q_sym[None, :, :]
    q_base = (q_sym < 4)[None, :, :]
    nq, nd = (q_sym == 4).sum(axis=1), (db_sym == 4).sum(axis=1)
    if enc == 5:    # D = matches - need
        return eq.sum(axis=2) - need
    if enc == 4:    # D = base-base matches - max(0, need - nN_q)
        return (eq & q_base).sum(axis=2) - np.maximum(0, need - nq)[None, :]
    PB = 32 if L <= 30 else 64
    alpha = 2 if enc == 2 else 1
    w5, T = alpha + 4, enc * (PB - L) - 3
    h = np.array([1, 1, -1, -1, 0])
    l = np.array([1, -1, 1, -1, 0])
    feats = [h, l] if enc == 2 else [h, l, h * l]
    S = sum((f[db_sym][:, None, :] * f[q_sym][None, :, :]).sum(axis=2) for f in feats)
    over = np.maximum(0, nq - (T - 1))
    c = np.minimum(254, alpha * (L - nq) + np.maximum(0, w5 * over - 127) - 4 * need)
    thermo = w5 * np.minimum(np.minimum(nq[None, :], nd[:, None]), T - 1) + \
        np.minimum(127, w5 * over)[None, :] * (nd[:, None] >= T)
    return S + c[None, :] - alpha * nd[:, None] + thermo


@pytest.mark.parametrize("nsym", [2, 3, 4, 5])
@pytest.mark.parametrize("L,noise", [(20, 0.05), (60, 0.05), (62, 0.6), (31, 0.1)])
def test_mma_accumulators_exact(L, noise, nsym, monkeypatch):
    monkeypatch.setenv("SMAFA_MMA_NSYM", str(nsym))
    c = smafa_b200.Context(0, "mma")
    try:
        db_sym = synth.make_db(1000, L=L, seed=1, noise=noise)
        q_sym = synth.make_queries(db_sym, 256, seed=2, noise=noise)
        d = c.upload(synth.pack_symbols(db_sym), L)
        bound = 7
        acc = c.debug_mma_dump(d, synth.pack_symbols(q_sym), bound)
        want = _mma_accumulator_model(db_sym[:128], q_sym, L, bound, nsym)
        assert (acc == want.astype(np.int32)).all()
        # the filter is conservative: every pair within the bound has a non-negative accumulator
        dist = (db_sym[:128, None, :] != q_sym[None, :, :]).sum(axis=2)
        assert (acc[dist <= bound] >= 0).all()
        # and every variant still answers queries exactly
        got = c.query(d, synth.pack_symbols(q_sym), L, max_divergence=9, max_num_hits=5)
        want_rows = c_oracle.query(synth.pack_symbols(db_sym), L, synth.pack_symbols(q_sym), L, 9, 5, None)
        assert got.shape == want_rows.shape and (got == want_rows).all()
        d.close()
    finally:
        c.close()


# ---- query selection -----------------------------------------------------------------------------

MODES = [(None, None, None), (3, None, None), (0, None, None), (None, 1, None), (None, 2, None), (None, 10, None),
         (5, 10, None), (2, 50, None), (None, 5000, None), (6, 5000, None), (None, 10, 1), (8, 25, 2), (99, 99, None)]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("L", [1, 3, 9, 12, 20, 32, 33, 60, 63, 64])
def test_query_matches_oracle(ctx, L, kernel):
    db_sym = synth.make_db(3000, L=L, seed=300 + L, family=8, max_subs=min(4, L), noise=0.03)
    q_sym = synth.make_queries(db_sym, 300, seed=400 + L, max_subs=min(6, L), noise=0.03)
    db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
    for m, k, r in MODES:
        check_query(ctx, db, q, L, m, k, r, kernel)


@pytest.mark.parametrize("kernel", KERNELS)
def test_query_long_windows_generic_path(ctx, kernel):
    for L in (65, 100, 130):
        db_sym = synth.make_db(1500, L=L, seed=500 + L)
        q_sym = synth.make_queries(db_sym, 100, seed=600 + L)
        db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
        for m, k, r in [(None, None, None), (5, None, None), (None, 7, None), (9, 7, 1)]:
            check_query(ctx, db, q, L, m, k, r, kernel)


def test_query_invalid_codes_fall_back_to_reference_layout(ctx):
    # a hand-made db may hold words that are not one-hot; the reference just XORs and popcounts
    rng = np.random.default_rng(5)
    L = 24
    db = rng.integers(0, 1 << 60, size=(500, 2), dtype=np.uint64)
    q = rng.integers(0, 1 << 60, size=(40, 2), dtype=np.uint64)
    ctx.set_kernel("auto")
    d = ctx.upload(db, L)
    got = ctx.query(d, q, L, max_num_hits=5)
    dist = ctx.distances(d, q, L)
    d.close()
    want = c_oracle.query(db, L, q, L, None, 5, None)
    assert got.shape == want.shape and (got == want).all()
    assert (dist[0].astype(np.int64) == c_oracle.distances(db, q[0])).all()


@pytest.mark.parametrize("kernel", KERNELS)
def test_ties_duplicates_and_tiny_shapes(ctx, kernel):
    L = 60
    one = synth.pack_symbols(synth.random_symbols(1, L, seed=1))
    same = np.repeat(one, 700, axis=0)  # every window ties
    q = synth.pack_symbols(synth.random_symbols(5, L, seed=2))
    q[0] = one[0]
    for m, k, r in [(None, None, None), (None, 3, None), (None, 3, 2), (None, 800, None), (0, None, None)]:
        check_query(ctx, same, q, L, m, k, r, kernel)
    check_query(ctx, one, q, L, None, None, None, kernel)    # D = 1
    check_query(ctx, one, q, L, None, 2, None, kernel)       # k > D
    check_query(ctx, same, q[:1], L, None, 10, None, kernel)  # Q = 1
    ctx.set_kernel(kernel)
    d = ctx.upload(same, L)
    assert ctx.query(d, q[:0], L).shape == (0, 3)            # no queries: no output, no panic
    d.close()


def test_reference_panics_through_the_abi(ctx):
    L = 12
    db = synth.pack_symbols(synth.random_symbols(10, L, seed=3))
    q = synth.pack_symbols(synth.random_symbols(2, L, seed=4))
    d = ctx.upload(db, L)
    with pytest.raises(smafa_b200.SmafaPanic, match="Cannot compute distances between seq of length 13 and windows of lengths 12"):
        ctx.query(d, np.zeros((1, 2), dtype=np.uint64), 13)
    with pytest.raises(smafa_b200.SmafaPanic, match="SMAFA_E_BAD_K"):
        ctx.query(d, q, L, max_num_hits=0)
    with pytest.raises(smafa_b200.SmafaPanic, match="limit_per_sequence"):
        ctx.query(d, q, L, limit_per_sequence=1)
    d.close()
    empty = ctx.upload(np.zeros((0, 1), dtype=np.uint64), L)
    with pytest.raises(smafa_b200.SmafaPanic, match="SMAFA_E_EMPTY_DB"):
        ctx.query(empty, q, L)
    with pytest.raises(smafa_b200.SmafaPanic, match="SMAFA_E_EMPTY_DB"):
        ctx.query(empty, q, L, max_num_hits=3)
    empty.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_candidate_overflow_retries_are_exact(ctx, kernel):
    L = 60
    db_sym = synth.make_db(5000, L=L, seed=31)
    q_sym = synth.make_queries(db_sym, 700, seed=32)
    db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
    ctx.set_candidate_capacity(4096)
    try:
        st = check_query(ctx, db, q, L, None, 20, None, kernel)   # floods the buffer -> split batches
        assert st["retries"] > 0
        check_query(ctx, db, q, L, None, 6000, None, kernel)      # one query emits D rows > capacity
    finally:
        ctx.set_candidate_capacity(0)


def test_sort_free_selection_and_its_fallback(ctx):
    """The common batch runs speculatively (api.cu run_batch_fast: tcgen05 scan + the sort-free "bucket" selection of
    finalize.cu, one host read-back).  Small buckets: no retry.  A query with more than 256 candidates, a candidate
    overflow or invalid query codes break the speculation: the batch runs again through the sort -- same rows.  With
    SMAFA_NO_FAST_FINALIZE=1 everything takes the sort."""
    L = 60
    db_sym = synth.make_db(30_000, L=L, seed=91, family=40, max_subs=5)
    q_sym = synth.make_queries(db_sym, 900, seed=92, max_subs=4)
    db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
    for m, k, r in [(5, None, None), (4, 10, None), (9, 30, 2), (0, None, None)]:
        st = check_query(ctx, db, q, L, m, k, r, "mma")
        assert st["retries"] == 0, (m, k, st["retries"])
    big_sym = synth.make_db(30_000, L=L, seed=93, family=600, max_subs=4)     # 50 families of 600 near-copies
    big, qb = synth.pack_symbols(big_sym), synth.pack_symbols(synth.make_queries(big_sym, 900, seed=94, max_subs=3))
    st = check_query(ctx, big, qb, L, 14, 25_000, None, "mma")               # every query keeps its whole family: buckets of 600
    assert st["retries"] > 0
    st = check_query(ctx, db, q, L, 6, 3, None, "mma")                       # ... after which the next batches skip the speculation
    assert st["retries"] == 0
    one = synth.random_symbols(1, L, seed=1)
    same = synth.pack_symbols(np.repeat(one, 5000, axis=0))        # 5000 ties per query
    st = check_query(ctx, same, synth.pack_symbols(np.repeat(one, 100, axis=0)), L, 0, None, None, "mma")
    old = os.environ.get("SMAFA_NO_FAST_FINALIZE")
    os.environ["SMAFA_NO_FAST_FINALIZE"] = "1"
    try:
        c2 = smafa_b200.Context(0, "mma")
    finally:
        if old is None:
            os.environ.pop("SMAFA_NO_FAST_FINALIZE")
        else:
            os.environ["SMAFA_NO_FAST_FINALIZE"] = old
    try:
        for m, k in [(5, None), (4, 10), (None, 10)]:
            st = check_query(c2, db, q, L, m, k, None, "mma")
            assert st["retries"] == 0
    finally:
        c2.close()


# ---- cluster -------------------------------------------------------------------------------------

@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("L,t,n", [(60, 3, 20000), (9, 2, 3000), (20, 1, 5000), (60, 0, 2000), (33, 6, 4000)])
def test_cluster_matches_oracle(ctx, L, t, n, kernel):
    ctx.set_kernel(kernel)
    sym = synth.make_cluster_input(n, L=L, seed=700 + L, family=10, max_subs=min(3, L))
    enc_all = synth.pack_symbols(sym)
    want_cof, want_nc, want_cmp = c_oracle.cluster(enc_all, L, t)
    keep = want_cof >= 0  # the host removes duplicate encodings (src/cluster.rs:46-48)
    enc = enc_all[keep]
    remap = np.cumsum(keep) - 1
    cof, nc, ncmp = ctx.cluster(enc, L, t)
    assert nc == want_nc and ncmp == want_cmp
    assert (cof.astype(np.int64) == remap[want_cof[keep]]).all()


def test_cluster_cli_parallel_host_stages(ctx, tmp_path):
    """`smafa cluster` end to end on 120 k sequences (10 % duplicate encodings, some in lower case / IUPAC): enough
    records for the multi-threaded de-duplication and output formatting; stdout must equal the oracle CLI's and must
    not depend on the number of host threads."""
    L, n = 60, 120_000
    sym = synth.make_cluster_input(n, L=L, seed=77, dup_fraction=0.10)
    seqs = synth.to_ascii(sym)
    seqs[5] = seqs[5].lower()                      # same encoding as seqs[5] upper-cased
    seqs[11] = seqs[11].replace(b"N", b"R")          # IUPAC ambiguity encodes like N
    fa = tmp_path / "in.fna"
    synth.write_fasta(fa, seqs)
    want = subprocess.run([c_oracle.CLI, "cluster", "-i", str(fa), "-d", "3"], capture_output=True, text=True)
    assert want.returncode == 0, want.stderr
    outs = []
    for threads in ("1", "3", ""):
        env = dict(os.environ)
        if threads:
            env["SMAFA_HOST_THREADS"] = threads
        else:
            env.pop("SMAFA_HOST_THREADS", None)
        r = subprocess.run([api.CLI_PATH, "cluster", "-i", str(fa), "-d", "3"], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout)
    assert outs[0] == want.stdout
    assert outs[1] == outs[0] and outs[2] == outs[0]


# ---- tcgen05 kernel under candidate pressure: survivor rings, verifier warps, flood path ----------------

def test_mma_survivor_rings_under_pressure(ctx):
    """Dense survivor patterns for the tcgen05 kernel (>= 64 queries, so AUTO/MMA really runs it): (a) every pair
    survives -- the epilogue's flood path verifies straight from the masks; (b) big families under a loose bound --
    tens of survivors per 64-column chunk, so the per-warp rings fill faster than the verifier warps drain them and
    the epilogue has to wait for room; (c) the same with the k-th tightening active."""
    L = 60
    ctx.set_kernel("mma")
    one = synth.random_symbols(1, L, seed=21)
    same = synth.pack_symbols(np.repeat(one, 1500, axis=0))
    q_sym = np.repeat(one, 320, axis=0)
    q_sym[::4] = synth.random_symbols(len(q_sym[::4]), L, seed=22)
    q = synth.pack_symbols(q_sym)
    for m, k in [(None, None), (0, None), (None, 3), (4, 2000), (None, 1499)]:
        st = check_query(ctx, same, q, L, m, k, None, "mma")            # (a)
        assert st["kernel_used"] == 2
    fam = synth.make_db(40_000, L=L, seed=23, family=400, max_subs=6, noise=0.0)
    qf = synth.make_queries(fam, 512, seed=24, max_subs=4, noise=0.0)
    dbw, qw = synth.pack_symbols(fam), synth.pack_symbols(qf)
    for m, k in [(12, 40_000), (14, None), (16, 300), (None, 250)]:
        st = check_query(ctx, dbw, qw, L, m, k, None, "mma")           # (b), (c)
        assert st["kernel_used"] == 2
    assert st["candidates"] >= 250 * 512


# ---- optimistic first pass under a guessed bound (csrc/guess.cu) -----------------------------------

@pytest.mark.parametrize("kernel", KERNELS + ["mma-union2", "mma-union3"])
@pytest.mark.parametrize("guess", [0, 2, 7, 25])
def test_guessed_bound_pass_is_exact(guess, kernel, monkeypatch):
    """Unbounded selection scans the batch under a guessed bound first and re-scans only the queries that did not find
    their k windows within it.  SMAFA_FORCE_GUESS pins the guess (the sampled estimate only switches on for >= 2e9
    comparisons), so every split between the two passes is exercised: nothing finished (0), some, all (25).  The
    mma-union variants force the first pass (and every scan that starts at need >= L/2) onto union rows of that degree."""
    monkeypatch.setenv("SMAFA_FORCE_GUESS", str(guess))
    if kernel.startswith("mma-union"):
        monkeypatch.setenv("SMAFA_MMA_UNION_FORCE", kernel[-1])
        kernel = "mma"
    c = smafa_b200.Context(0, kernel)
    try:
        for L in (20, 60):
            db_sym = synth.make_db(6000, L=L, seed=700 + L, family=8, max_subs=min(5, L), noise=0.03)
            q_sym = synth.make_queries(db_sym, 700, seed=800 + L, max_subs=min(8, L), noise=0.03)
            # a third of the queries are unrelated to the db: they never finish under a small guess
            q_sym[::3] = synth.random_symbols(len(q_sym[::3]), L, seed=900 + L)
            db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
            d = c.upload(db, L)
            rescanned = []
            for m, k, r in [(None, None, None), (None, 1, None), (None, 2, None), (None, 10, None), (None, 10, 1),
                            (L - 1, 25, None), (None, 5999, None), (None, 7000, None), (3, 10, None)]:
                got, st = c.query(d, q, L, max_divergence=m, max_num_hits=k, limit_per_sequence=r, return_stats=True)
                want = c_oracle.query(db, L, q, L, m, k, r, threads=os.cpu_count() or 1)
                assert got.shape == want.shape and (got == want).all(), (L, m, k, r, guess)
                rescanned.append((st["guess_bound"], st["rescanned"]))
            d.close()
            # the pass ran for the loose modes (not for k > D, which keeps everything, nor for --max-divergence 3)
            want_g = guess if guess < L else -1  # a guess that is not below the caller's bound is pointless
            assert rescanned[0][0] == want_g and rescanned[3][0] == want_g
            assert rescanned[7][0] == -1 and rescanned[8][0] == -1
            if guess == 0:
                assert rescanned[3][1] > 0
    finally:
        c.close()


def test_guessed_bound_from_the_sample(ctx):
    """Large enough for the sampled estimate (no forcing): unbounded top-10 and best-hit on 20 k x 200 k, oracle on a
    query subsample, and the same rows with the pass switched off."""
    L, D, Q = 60, 200_000, 20_000
    db_sym = synth.make_db(D, L=L, seed=11)
    q_sym = synth.make_queries(db_sym, Q, seed=12)
    q_sym[::7] = synth.random_symbols(len(q_sym[::7]), L, seed=13)  # queries without relatives in the db
    db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
    ctx.set_kernel("auto")
    d = ctx.upload(db, L)
    for k in (None, 10):
        got, st = ctx.query(d, q, L, max_num_hits=k, return_stats=True)
        assert 0 <= st["guess_bound"] < L and st["rescanned"] >= Q // 10   # a seventh of the queries has no relatives
        want = c_oracle.query(db, L, q[:140], L, None, k, None, threads=os.cpu_count() or 1)
        sub = got[got[:, 0] < 140]
        assert sub.shape == want.shape and (sub == want).all()
        x = np.bitwise_count(db[got[:, 1]] ^ q[got[:, 0]]).sum(axis=1) // 2
        assert (x == got[:, 2]).all()
        os.environ["SMAFA_NO_GUESS"] = "1"
        try:
            c2 = smafa_b200.Context(0, "auto")
            d2 = c2.upload(db, L)
            plain, st2 = c2.query(d2, q, L, max_num_hits=k, return_stats=True)
            d2.close()
            c2.close()
        finally:
            del os.environ["SMAFA_NO_GUESS"]
        assert st2["guess_bound"] == -1
        assert plain.shape == got.shape and (plain == got).all()
    d.close()


# ---- full-size checks (BASELINE config 2 shape): oracle on a query subsample + properties -------

@pytest.mark.parametrize("kernel", KERNELS)
def test_config2_scale_subsample_and_properties(ctx, kernel):
    L, D, Q = 60, 1_000_000, 100_000
    db_sym = synth.make_db(D, L=L)
    q_sym = synth.make_queries(db_sym, Q)
    db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
    ctx.set_kernel(kernel)
    d = ctx.upload(db, L)
    got, st = ctx.query(d, q, L, max_divergence=5, return_stats=True)
    assert st["pairs"] == Q * D
    # (1) bit-exact against the oracle on the first 300 queries
    want = c_oracle.query(db, L, q[:300], L, 5, None, None, threads=os.cpu_count() or 1)
    sub = got[got[:, 0] < 300]
    assert sub.shape == want.shape and (sub == want).all()
    # (2) size-independent properties: sorted print order, distances within the bound, one distance
    # per query, and every reported distance re-derived from the encodings
    assert (np.diff(got[:, 0].astype(np.int64)) >= 0).all()
    assert (got[:, 2] <= 5).all()
    same_q = got[1:, 0] == got[:-1, 0]
    assert (got[1:, 2][same_q] == got[:-1, 2][same_q]).all()
    assert (got[1:, 1][same_q] > got[:-1, 1][same_q]).all()
    x = np.bitwise_count(db[got[:, 1]] ^ q[got[:, 0]]).sum(axis=1) // 2
    assert (x == got[:, 2]).all()
    # (3) self-hit: querying db windows finds them at distance 0 (Mode B, k=3 exercises the k-th path)
    probe = db[::5000]
    self_hits = ctx.query(d, probe, L, max_num_hits=3)
    first = self_hits[np.unique(self_hits[:, 0], return_index=True)[1]]
    assert (first[:, 2] == 0).all()
    want = c_oracle.query(db, L, probe[:40], L, None, 3, None, threads=os.cpu_count() or 1)
    sub = self_hits[self_hits[:, 0] < 40]
    assert sub.shape == want.shape and (sub == want).all()
    d.close()


# ---- BASELINE config 3 at full size on ONE GPU: 1M queries x 10M windows (the north-star target shape) ----

def test_config3_full_size_single_gpu(ctx):
    """1 M x 10 M = 10^13 comparisons through smafa_query (default kernel), --max-divergence 5.  The oracle needs
    ~10^5 core-seconds for the whole job, so parity is checked bit-exactly on a query subsample spread over the
    whole batch and, for every one of the rows, through properties that do not depend on the size."""
    L, D, Q = 60, 10_000_000, 1_000_000
    db_sym = synth.make_db(D, L=L)
    q_sym = synth.make_queries(db_sym, Q)
    db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
    del db_sym, q_sym
    ctx.set_kernel("auto")
    d = ctx.upload(db, L)
    got, st = ctx.query(d, q, L, max_divergence=5, return_stats=True)
    assert st["pairs"] == Q * D and st["kernel_used"] == 2
    # (1) bit-exact on 96 queries taken from the start, the middle and the end of the batch
    pick = np.concatenate([np.arange(32), Q // 2 + np.arange(32), Q - 32 + np.arange(32)])
    want = c_oracle.query(db, L, q[pick], L, 5, None, None, threads=os.cpu_count() or 1)
    want[:, 0] = pick[want[:, 0]]
    sub = got[np.isin(got[:, 0], pick)]
    assert sub.shape == want.shape and (sub == want).all()
    # (2) properties over all rows: print order, bound, one distance per query, ascending subjects within a query,
    # and every reported distance re-derived from the encodings
    assert (np.diff(got[:, 0].astype(np.int64)) >= 0).all()
    assert (got[:, 2] <= 5).all()
    same_q = got[1:, 0] == got[:-1, 0]
    assert (got[1:, 2][same_q] == got[:-1, 2][same_q]).all()
    assert (got[1:, 1][same_q] > got[:-1, 1][same_q]).all()
    x = np.bitwise_count(db[got[:, 1]] ^ q[got[:, 0]]).sum(axis=1) // 2
    assert (x == got[:, 2]).all()
    assert got.shape[0] > Q // 4  # about half of the queries sit within 5 of their source window
    d.close()
