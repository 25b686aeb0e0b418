"""CPU-only checks of the protein extension (no reference parity exists for it: the reference panics on
amino-acid bytes, src/lib.rs:35-42): the host encoder / decoder / makedb of the B200 build against the
oracle's restatement of the same definition, and the two oracles against the definition itself."""
import ctypes as C
import subprocess

import numpy as np

from oracle import c_oracle, np_oracle
from smafa_b200 import api, synth


def test_aa_symbol_table_and_roundtrip():
    l = api.load_library()
    letters = b"ACDEFGHIKLMNPQRSTVWY"
    for i, b in enumerate(letters):
        assert l.smafa_encode_symbol_alphabet(b, 1) == i + 1
        assert l.smafa_encode_symbol_alphabet(b + 32, 1) == i + 1          # lower case
    for b in b"XBZJUOxbzjuo":
        assert l.smafa_encode_symbol_alphabet(b, 1) == 21
    assert l.smafa_encode_symbol_alphabet(ord("-"), 1) == 22 and l.smafa_encode_symbol_alphabet(ord("*"), 1) == 23
    assert l.smafa_encode_symbol_alphabet(ord("1"), 1) == 0 and l.smafa_encode_symbol_alphabet(ord("A"), 0) == 16
    sym = synth.make_db_aa(50, L=37, seed=3, noise=0.1)
    words = synth.pack_symbols_aa(sym)
    for row, s in zip(words, synth.to_ascii_aa(sym)):
        out = np.zeros(4, dtype=np.uint64)
        bad = C.c_size_t(0)
        assert l.smafa_encode_window_alphabet(s, len(s), out.ctypes.data, C.byref(bad), 1) == 0
        assert (out == row).all()
        buf = C.create_string_buffer(len(s))
        assert l.smafa_decode_window_alphabet(out.ctypes.data, len(s), buf, 1) == 0
        assert buf.raw == s


def test_aa_oracles_agree_with_definition():
    c_oracle.set_alphabet(1)
    try:
        db_sym = synth.make_db_aa(400, L=20, seed=5, noise=0.05)
        q_sym = synth.make_queries_aa(db_sym, 6, seed=6, noise=0.05)
        db, q = synth.pack_symbols_aa(db_sym), synth.pack_symbols_aa(q_sym)
        for i in range(q.shape[0]):
            want = (db_sym != q_sym[i][None, :]).sum(axis=1)
            assert (c_oracle.distances(db, q[i]) == want).all()
            assert (np_oracle.distances(db, q[i], alphabet=1) == want).all()
        # Mode B on the definition: everything <= the 10th smallest distance, in (distance, subject) order
        rows = c_oracle.query(db, 20, q, 20, None, 10, None)
        for i in range(q.shape[0]):
            d = (db_sym != q_sym[i][None, :]).sum(axis=1)
            cut = np.sort(d)[9]
            idx = np.nonzero(d <= cut)[0]
            idx = idx[np.lexsort((idx, d[idx]))]
            mine = rows[rows[:, 0] == i]
            assert (mine[:, 1] == idx).all() and (mine[:, 2] == d[idx]).all()
    finally:
        c_oracle.set_alphabet(0)


def test_aa_makedb_bytes_equal_oracle_cli(tmp_path):
    c_oracle.build()
    sym = synth.make_db_aa(300, L=20, seed=8, noise=0.05)
    synth.write_fasta(tmp_path / "p.faa", synth.to_ascii_aa(sym))
    assert subprocess.run([api.CLI_PATH, "makedb", "--protein", "-i", tmp_path / "p.faa", "-d", tmp_path / "a.db"]).returncode == 0
    assert subprocess.run([c_oracle.CLI, "makedb", "--protein", "-i", tmp_path / "p.faa", "-d", tmp_path / "b.db"]).returncode == 0
    assert (tmp_path / "a.db").read_bytes() == (tmp_path / "b.db").read_bytes()
    r = subprocess.run([api.CLI_PATH, "makedb", "-i", tmp_path / "p.faa", "-d", tmp_path / "c.db"], capture_output=True)
    assert r.returncode == 101 and b"cannot be interpreted as nucleotide" in r.stderr
