"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol
include/smafa_b200.h declares, makedb/count are byte-exact against the reference's fixtures, the
CLI mirrors the reference's exit codes, and compute entry points fail loudly without a GPU."""
import base64
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import smafa_b200
from smafa_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    from smafa_b200 import build
    build.build()


def cli(*args):
    return subprocess.run([api.CLI_PATH, *map(str, args)], capture_output=True, text=True)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "smafa_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(smafa_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    lib = ctypes.CDLL(api.lib_path())
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert smafa_b200.load_library().smafa_abi_version() == 4


def test_makedb_bytes_match_reference_fixtures(kats, kat_dir, tmp_path):
    for case in kats["makedb"]:
        out = tmp_path / "db"
        smafa_b200.makedb(kat_dir / case["input"], out)
        assert out.read_bytes() == base64.b64decode(kats["binary_files_b64"][case["db"]])
        r = cli("makedb", "-i", kat_dir / case["input"], "-d", out)
        assert r.returncode == 0 and out.read_bytes() == base64.b64decode(kats["binary_files_b64"][case["db"]])


def test_makedb_large_roundtrip_matches_oracle(tmp_path):
    from oracle import c_oracle
    c_oracle.build()
    sym = synth.make_db(3000, L=60, seed=9)
    fa = tmp_path / "db.fna"
    synth.write_fasta(fa, synth.to_ascii(sym))
    smafa_b200.makedb(fa, tmp_path / "a.db")
    r = subprocess.run([c_oracle.CLI, "makedb", "-i", fa, "-d", tmp_path / "b.db"])
    assert r.returncode == 0
    assert (tmp_path / "a.db").read_bytes() == (tmp_path / "b.db").read_bytes()


def test_encode_decode_helpers():
    l = smafa_b200.load_library()
    # reference src/lib.rs:167-184
    for ch, code in [("A", 16), ("c", 8), ("G", 4), ("u", 2), ("T", 2), ("N", 1), ("-", 1), ("y", 1), ("E", 0), ("*", 0)]:
        assert l.smafa_encode_symbol(ord(ch)) == code
    seq = b"ACGTNacgtn-RYUu"
    words = np.zeros(2, dtype=np.uint64)
    bad = ctypes.c_size_t(0)
    assert l.smafa_encode_window(seq, len(seq), words.ctypes.data, ctypes.byref(bad)) == 0
    out = ctypes.create_string_buffer(len(seq))
    assert l.smafa_decode_window(words.ctypes.data, len(seq), out) == 0
    assert out.raw == b"ACGTNACGTNNNNTT"
    assert l.smafa_encode_window(b"ACXG", 4, words.ctypes.data, ctypes.byref(bad)) != 0 and bad.value == 2
    sym = synth.random_symbols(4, 61, seed=2)
    enc = synth.pack_symbols(sym)
    for i, s in enumerate(synth.to_ascii(sym)):
        w = np.zeros(6, dtype=np.uint64)
        assert l.smafa_encode_window(s, len(s), w.ctypes.data, ctypes.byref(bad)) == 0
        assert (w == enc[i]).all()


def test_count_json(kats, kat_dir):
    for case in kats["count"]:
        p = str(kat_dir / case["input"])
        r = cli("count", "-i", p)
        assert r.returncode == 0
        assert r.stdout == '[{"path":"%s","num_reads":%d,"num_bases":%d}]\n' % (p, case["num_reads"], case["num_bases"])


def test_cli_error_paths_without_device_work(kats, kat_dir, tmp_path):
    c = kats["old_db"]  # tests/test_cmdline.rs:27-41: version gate fires before anything else
    r = cli("query", "-d", kat_dir / c["db"], "-q", kat_dir / c["query"])
    assert r.returncode == 101 and c["stderr_contains"] in r.stderr
    r = cli("query", "-d", tmp_path / "nope.db", "-q", kat_dir / c["query"])
    assert r.returncode == 1 and r.stderr.startswith("Error:")
    bad = tmp_path / "bad.fna"
    bad.write_text(">x desc\nACGE\n")
    r = cli("makedb", "-i", bad, "-d", tmp_path / "o")
    assert r.returncode == 101
    assert 'Byte 69 cannot be interpreted as nucleotide, in sequence "x desc" at position 3' in r.stderr
    ragged = tmp_path / "ragged.fna"
    ragged.write_text(">a\nACG\n>b\nACGT\n")
    r = cli("makedb", "-i", ragged, "-d", tmp_path / "o")
    assert r.returncode == 101 and "WindowSet seq length is 3, got a new sequence of length 4" in r.stderr
    r = cli("cluster", "-i", ragged)  # src/main.rs:43 unwrap on missing -d
    assert r.returncode == 101
    assert cli("frobnicate").returncode == 2


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_compute_fails_loudly_without_gpu(kat_dir):
    with pytest.raises(smafa_b200.SmafaError) as e:
        smafa_b200.Context(0)
    assert "no CPU fallback" in str(e.value)
    r = cli("query", "-d", kat_dir / "random_3_2.fna.smafadb", "-q", kat_dir / "random_3_2.fna")
    assert r.returncode == 1 and "SMAFA_E_CUDA" in r.stderr and r.stdout == ""


def test_limit_per_sequence_host_filter():
    # run-length semantics of src/lib.rs:269-289 on a hand-made hit list
    l = smafa_b200.load_library()
    db = np.array([[5], [7], [7], [5], [7]], dtype=np.uint64)
    rows = [(0, 1, 0), (0, 2, 0), (0, 4, 0), (0, 0, 1), (0, 3, 1), (1, 1, 2), (1, 2, 2)]
    hits = (api.Hit * len(rows))(*[api.Hit(*r) for r in rows])
    n = l.smafa_apply_limit_per_sequence(hits, len(rows), db.ctypes.data, 1, 0, 2)
    got = [(hits[i].query, hits[i].subject, hits[i].distance) for i in range(n)]
    assert got == [(0, 1, 0), (0, 2, 0), (0, 0, 1), (0, 3, 1), (1, 1, 2), (1, 2, 2)]
    hits = (api.Hit * len(rows))(*[api.Hit(*r) for r in rows])
    n = l.smafa_apply_limit_per_sequence(hits, len(rows), db.ctypes.data, 1, 0, 1)
    got = [(hits[i].query, hits[i].subject, hits[i].distance) for i in range(n)]
    assert got == [(0, 1, 0), (0, 0, 1), (1, 1, 2)]


def test_cli_accepts_clap_value_syntax(kat_dir, tmp_path):
    """clap (src/main.rs:64-116) takes `--name=value`, `-n=value` and `-nvalue` as well as `--name value`."""
    fa = kat_dir / "random_3_2.fna"
    outs = []
    for i, args in enumerate([["-i", fa, "-d", None], [f"--input={fa}", "--database", None], [f"-i{fa}", "-d=", None]]):
        out = tmp_path / f"o{i}.db"
        args = [a if a is not None else out for a in args]
        if str(args[-2]) == "-d=":
            args = args[:-2] + [f"-d={out}"]
        r = cli("makedb", *args)
        assert r.returncode == 0, r.stderr
        outs.append(out.read_bytes())
    assert outs[0] == outs[1] == outs[2] == (kat_dir / "random_3_2.fna.smafadb").read_bytes()
    r = cli("count", f"-i{fa}")
    assert r.returncode == 0 and '"num_reads":2' in r.stdout
    assert cli("makedb", "--input").returncode == 2


def test_group_layout_keeps_clusters_together_and_row_aligned():
    """smafa_group_layout (host half of smafa_group_order, csrc/api.cu): clusters stay contiguous, members keep their
    input order, and clusters are arranged so that they straddle as few 16-wide operand rows as their sizes allow."""
    import ctypes as C
    lib = smafa_b200.load_library()
    lib.smafa_group_layout.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    rng = np.random.default_rng(3)

    def layout(sizes):
        # windows of the clusters interleaved at random; the first window of a cluster is its founder
        labels = rng.permutation(np.repeat(np.arange(len(sizes)), sizes))
        first = {}
        cof = np.zeros(len(labels), dtype=np.uint32)
        for i, c in enumerate(labels):
            first.setdefault(c, i)
            cof[i] = first[c]
        perm = np.zeros(len(labels), dtype=np.uint32)
        assert lib.smafa_group_layout(cof.ctypes.data, len(labels), perm.ctypes.data) == 0
        assert sorted(perm.tolist()) == list(range(len(labels)))
        lab = labels[perm]
        runs = 1 + int((np.diff(lab) != 0).sum())
        assert runs == len(sizes)                                   # every cluster is one run of rows
        for c in range(len(sizes)):                                 # members in input order
            rows = perm[lab == c]
            assert (np.diff(rows.astype(np.int64)) > 0).all()
        # rows touched by each cluster vs the fewest its size allows
        start = {}
        for r, c in enumerate(lab):
            start.setdefault(c, r)
        extra = sum(((start[c] + s - 1) // 16 - start[c] // 16 + 1) - (s + 15) // 16 for c, s in enumerate(sizes))
        return extra

    assert layout([16] * 40 + [32] * 5 + [48]) == 0                 # multiples of 16: perfectly aligned
    assert layout([16] * 100 + [17] * 3) <= 3                       # three odd clusters must not shift the hundred others
    assert layout([15, 1, 14, 2, 13, 3, 8, 8, 16, 16, 31, 1]) == 0  # remainders that pair up to full rows
    sizes = rng.integers(1, 40, size=400).tolist()
    assert layout(sizes) <= 0.35 * len(sizes)                       # random sizes: far fewer straddles than clusters
    bad = np.array([1, 1, 2], dtype=np.uint32)                      # window 0 claims a founder that comes after it
    out = np.zeros(3, dtype=np.uint32)
    assert lib.smafa_group_layout(bad.ctypes.data, 3, out.ctypes.data) != 0
