"""Protein windows (BASELINE.json configs[3]): an EXTENSION of this build -- the reference panics on amino-acid
bytes (src/lib.rs:35-42), so there is no reference parity to pin.  These tests pin the GPU path to
the build's own CPU restatement of the extension (oracle, alphabet = 1): distance = positions whose symbols
differ, selection rules unchanged (src/lib.rs:242-314)."""
import os
import subprocess

import numpy as np
import pytest

import smafa_b200
from smafa_b200 import api
from oracle import c_oracle, np_oracle
from smafa_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def aa_oracle():
    c_oracle.set_alphabet(1)
    yield
    c_oracle.set_alphabet(0)


@pytest.fixture()
def ctx():
    c = smafa_b200.Context(0)
    c.set_alphabet("protein")
    yield c
    c.close()


@pytest.mark.parametrize("L", [1, 12, 20, 30, 31, 45, 62, 63, 64, 70])
def test_protein_distances_bit_exact(ctx, aa_oracle, L):
    db_sym = synth.make_db_aa(700, L=L, seed=11, noise=0.05)
    q_sym = synth.make_queries_aa(db_sym, 9, seed=12, noise=0.05)
    db, q = synth.pack_symbols_aa(db_sym), synth.pack_symbols_aa(q_sym)
    d = ctx.upload(db, L)
    got = ctx.distances(d, q, L)
    want = (db_sym[None, :, :] != q_sym[:, None, :]).sum(axis=2)
    assert (got.astype(np.int64) == want).all()
    for i in range(q.shape[0]):
        assert (c_oracle.distances(db, q[i]) == want[i]).all()          # the oracle agrees with the definition
        assert (np_oracle.distances(db, q[i], alphabet=1) == want[i]).all()
    d.close()


MODES = [(None, None), (3, None), (0, None), (None, 1), (None, 10), (4, 10), (None, 3), (2, 50), (None, 100000)]


@pytest.mark.parametrize("kernel", ["popc", "mma"])
@pytest.mark.parametrize("L", [20, 33, 60])
def test_protein_query_matches_oracle(ctx, aa_oracle, kernel, L):
    ctx.set_kernel(kernel)
    db_sym = synth.make_db_aa(6000, L=L, seed=21, noise=0.03)
    q_sym = synth.make_queries_aa(db_sym, 300, seed=22, noise=0.03)
    db, q = synth.pack_symbols_aa(db_sym), synth.pack_symbols_aa(q_sym)
    d = ctx.upload(db, L)
    for m, k in MODES:
        got, st = ctx.query(d, q, L, max_divergence=m, max_num_hits=k, return_stats=True)
        want = c_oracle.query(db, L, q, L, m, k, None)
        assert got.shape == want.shape and (got == want).all(), (kernel, L, m, k)
        if not (k is not None and k >= db.shape[0] and m is None):  # an all-admitting fixed bound always runs on POPC
            assert st["kernel_used"] == {"popc": 1, "mma": 2}[kernel]
    d.close()


def test_protein_mma_accumulators_exact(ctx):
    """ENC 20 (scan_mma.cu): D = amino-acid matches - max(0, need - nX_q) for the first db tile."""
    L, bound = 20, 6
    db_sym = synth.make_db_aa(1000, L=L, seed=1, noise=0.05)
    q_sym = synth.make_queries_aa(db_sym, 256, seed=2, noise=0.05)
    ctx.set_kernel("mma")
    d = ctx.upload(synth.pack_symbols_aa(db_sym), L)
    assert d.mma_k == 416
    acc = ctx.debug_mma_dump(d, synth.pack_symbols_aa(q_sym), bound)
    eq = (db_sym[:128, None, :] == q_sym[None, :, :]) & (q_sym[None, :, :] < 20)
    want = eq.sum(axis=2) - np.maximum(0, (L - bound) - (q_sym >= 20).sum(axis=1))[None, :]
    assert (acc == want.astype(np.int32)).all()
    dist = (db_sym[:128, None, :] != q_sym[None, :, :]).sum(axis=2)
    assert (acc[dist <= bound] >= 0).all()
    d.close()


def test_protein_config4_shape_top10(ctx, aa_oracle):
    """configs[3] scaled to one test: 20-aa windows, --max-num-hits 10, no --max-divergence; a query subsample is
    compared with the oracle and the whole answer is checked through size-independent properties."""
    L, D, Q = 20, 400_000, 20_000
    db_sym = synth.make_db_aa(D, L=L)
    q_sym = synth.make_queries_aa(db_sym, Q)
    db, q = synth.pack_symbols_aa(db_sym), synth.pack_symbols_aa(q_sym)
    d = ctx.upload(db, L)
    got = ctx.query(d, q, L, max_divergence=None, max_num_hits=10)
    sub = 64
    want = c_oracle.query(db, L, q[:sub], L, None, 10, None, threads=os.cpu_count() or 1)
    head = got[got[:, 0] < sub]
    assert head.shape == want.shape and (head == want).all()
    # properties: sorted by (query, distance, subject); >= 10 rows per query; distances re-derived from the symbols
    key = got[:, 0].astype(np.int64) << 40 | got[:, 2].astype(np.int64) << 32 | got[:, 1].astype(np.int64)
    assert (np.diff(key) > 0).all()
    assert (np.bincount(got[:, 0], minlength=Q) >= 10).all()
    pick = np.random.default_rng(5).integers(0, got.shape[0], size=5000)
    rows = got[pick]
    assert ((db_sym[rows[:, 1]] != q_sym[rows[:, 0]]).sum(axis=1) == rows[:, 2]).all()
    d.close()


def test_protein_cluster_matches_oracle(ctx, aa_oracle):
    L = 20
    rng = np.random.default_rng(31)
    roots = rng.integers(0, 20, size=(150, L), dtype=np.uint8)
    sym = synth._mutate_aa(rng, roots[np.arange(3000) % 150], 2, 0.0)[rng.permutation(3000)]
    enc = synth.pack_symbols_aa(sym)
    _, first = np.unique(enc, axis=0, return_index=True)
    enc = enc[np.sort(first)]
    for kernel in ("popc", "mma"):
        ctx.set_kernel(kernel)
        cof, nc, _ = ctx.cluster(enc, L, 2)
        want_cof, want_nc, _ = c_oracle.cluster(enc, L, 2)
        assert nc == want_nc and (cof.astype(np.int64) == want_cof).all()


def test_protein_cli_matches_oracle_cli(tmp_path, aa_oracle):
    L = 20
    db_sym = synth.make_db_aa(3000, L=L, seed=41, noise=0.02)
    q_sym = synth.make_queries_aa(db_sym, 200, seed=42, noise=0.02)
    synth.write_fasta(tmp_path / "db.faa", synth.to_ascii_aa(db_sym))
    synth.write_fasta(tmp_path / "q.faa", synth.to_ascii_aa(q_sym))
    c_oracle.build()
    outs = []
    for exe, db in ((api.CLI_PATH, "b200.db"), (c_oracle.CLI, "orc.db")):
        subprocess.run([exe, "makedb", "--protein", "-i", str(tmp_path / "db.faa"), "-d", str(tmp_path / db)], check=True)
        r = subprocess.run([exe, "query", "--protein", "-d", str(tmp_path / db), "-q", str(tmp_path / "q.faa"),
                            "--max-num-hits", "10"], capture_output=True, check=True)
        outs.append(r.stdout)
    assert (tmp_path / "b200.db").read_bytes() == (tmp_path / "orc.db").read_bytes()
    assert outs[0] == outs[1] and outs[0].count(b"\n") >= 2000
    # an amino-acid file without --protein panics like the reference does on non-nucleotide bytes
    r = subprocess.run([api.CLI_PATH, "makedb", "-i", str(tmp_path / "db.faa"), "-d", str(tmp_path / "x.db")],
                       capture_output=True)
    assert r.returncode == 101 and b"cannot be interpreted as nucleotide" in r.stderr
