"""Pins the CPU oracle (oracle/) against every golden vector the reference's own tests hold
for the query/cluster path, and the C and numpy restatements against each
other.  CPU only."""
import base64
import subprocess

import numpy as np
import pytest

from oracle import c_oracle, np_oracle
from smafa_b200 import synth


@pytest.fixture(scope="module", autouse=True)
def _built():
    c_oracle.build()


def run_cli(*args):
    return subprocess.run([c_oracle.CLI, *map(str, args)], capture_output=True, text=True)


def test_encoding_kat(kats, kat_dir):
    # src/lib.rs:357-366
    fx = (kat_dir / kats["encoding"]["input"]).read_text()
    seqs = [l.encode() for l in fx.splitlines() if not l.startswith(">")]
    enc = np_oracle.encode(seqs)
    assert enc.tolist() == kats["encoding"]["words"]


@pytest.mark.parametrize("i", [0, 1])
def test_makedb_bytes(kats, kat_dir, tmp_path, i):
    case = kats["makedb"][i]
    out = tmp_path / "db"
    r = run_cli("makedb", "-i", kat_dir / case["input"], "-d", out)
    assert r.returncode == 0, r.stderr
    assert out.read_bytes() == base64.b64decode(kats["binary_files_b64"][case["db"]])


def test_query_kats(kats, kat_dir, tmp_path):
    for case in kats["query"]:
        if "makedb_from" in case:
            db = tmp_path / (case["name"] + ".db")
            assert run_cli("makedb", "-i", kat_dir / case["makedb_from"], "-d", db).returncode == 0
        else:
            db = kat_dir / case["db"]
        r = run_cli("query", "-d", db, "-q", kat_dir / case["query"], *case["args"])
        assert r.returncode == 0, (case["name"], r.stderr)
        assert r.stdout == case["stdout"], case["name"]


def test_old_db_rejected(kats, kat_dir):
    c = kats["old_db"]
    r = run_cli("query", "-d", kat_dir / c["db"], "-q", kat_dir / c["query"])
    assert r.returncode == 101
    assert c["stderr_contains"] in r.stderr


def test_cluster_kats(kats, kat_dir):
    for case in kats["cluster"]:
        r = run_cli("cluster", "-i", kat_dir / case["input"], "-d", case["t"])
        assert r.returncode == 0, r.stderr
        assert r.stdout == case["stdout"], case["name"]


def test_count_kats(kats, kat_dir):
    for case in kats["count"]:
        p = str(kat_dir / case["input"])
        r = run_cli("count", "-i", p)
        assert r.stdout == '[{"path":"%s","num_reads":%d,"num_bases":%d}]\n' % (
            p, case["num_reads"], case["num_bases"])


def test_panics(kat_dir, tmp_path):
    bad = tmp_path / "bad.fna"
    bad.write_text(">x desc\nACGE\n")
    r = run_cli("makedb", "-i", bad, "-d", tmp_path / "o")
    assert r.returncode == 101
    assert 'Byte 69 cannot be interpreted as nucleotide, in sequence "x desc" at position 3' in r.stderr
    ragged = tmp_path / "ragged.fna"
    ragged.write_text(">a\nACG\n>b\nACGT\n")
    r = run_cli("makedb", "-i", ragged, "-d", tmp_path / "o")
    assert r.returncode == 101 and "WindowSet seq length is 3, got a new sequence of length 4" in r.stderr
    r = run_cli("query", "-d", kat_dir / "random_3_2.fna.smafadb", "-q", ragged)
    assert r.returncode == 101
    assert "Cannot compute distances between seq of length 4 and windows of lengths 3" in r.stderr
    assert r.stdout.startswith("0\t")  # the first (valid) query was answered before the panic
    r = run_cli("query", "-d", kat_dir / "random_3_2.fna.smafadb", "-q", kat_dir / "random_3_2.fna",
                "--limit-per-sequence", "1")
    assert r.returncode == 101 and "limit_per_sequence" in r.stderr
    r = run_cli("query", "-d", kat_dir / "random_3_2.fna.smafadb", "-q", kat_dir / "random_3_2.fna",
                "--max-num-hits", "0")
    assert r.returncode == 101
    r = run_cli("query", "-d", tmp_path / "missing.db", "-q", kat_dir / "random_3_2.fna")
    assert r.returncode == 1 and r.stderr.startswith("Error:")


@pytest.mark.parametrize("L", [1, 3, 12, 13, 20, 60, 61])
def test_c_vs_numpy_query(L):
    db_sym = synth.make_db(300, L=L, seed=11 + L, family=8, max_subs=min(4, L), noise=0.05)
    q_sym = synth.make_queries(db_sym, 40, seed=5 + L, max_subs=min(5, L), noise=0.05)
    db, q = synth.pack_symbols(db_sym), synth.pack_symbols(q_sym)
    assert (db == np_oracle.encode_codes(np.array([16, 8, 4, 2, 1], dtype=np.uint8)[db_sym])).all()
    for i in range(3):
        assert (c_oracle.distances(db, q[i]) == np_oracle.distances(db, q[i])).all()
    for m, k, r in [(None, None, None), (2, None, None), (None, 1, None), (None, 5, None),
                    (3, 5, None), (None, 1000, None), (None, 7, 1), (4, 20, 2)]:
        got = c_oracle.query(db, L, q, L, m, k, r)
        want = np_oracle.query(db, L, q, L, m, k, r)
        assert got.tolist() == [list(h) for h in want], (L, m, k, r)
    # threads>1 must not change the answer
    assert (c_oracle.query(db, L, q, L, 3, 5, None, threads=4) == c_oracle.query(db, L, q, L, 3, 5, None)).all()


@pytest.mark.parametrize("L,t", [(9, 2), (60, 3), (20, 1)])
def test_c_vs_numpy_cluster(L, t):
    sym = synth.make_cluster_input(400, L=L, seed=3 + L, family=10, max_subs=min(3, L))
    enc = synth.pack_symbols(sym)
    cof_c, nc_c, cmp_c = c_oracle.cluster(enc, L, t)
    cof_n, nc_n = np_oracle.cluster(enc, t)
    assert (cof_c == cof_n).all() and nc_c == nc_n
    assert (cof_c == -1).sum() > 0  # duplicates are present and suppressed
    assert cmp_c > 0


def test_decode_roundtrip():
    sym = synth.random_symbols(5, 61, seed=1)
    enc = synth.pack_symbols(sym)
    for i in range(5):
        assert np_oracle.decode(enc[i], 61) == synth.to_ascii(sym[i:i + 1], gap_fraction=0)[0].decode()
