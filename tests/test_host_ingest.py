"""Host ingest paths of the B200 build: the multi-threaded LEB128 db decode, FASTA parse and
window encoding must give byte-identical results to the single-threaded pass for every thread count, and the same
errors on malformed input.  CPU only."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import smafa_b200
from smafa_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LOADER = """
import sys, hashlib, json
sys.path.insert(0, %r)
from smafa_b200 import api
try:
    w, L = api.load_db_file(sys.argv[1])
    print(json.dumps({"n": int(w.shape[0]), "W": int(w.shape[1]) if w.ndim == 2 else 0, "L": L,
                      "sha": hashlib.sha256(w.tobytes()).hexdigest()}))
except Exception as e:
    print(json.dumps({"error": type(e).__name__, "text": str(e)}))
""" % ROOT


def _load(path, threads):
    env = dict(os.environ, SMAFA_HOST_THREADS=str(threads))
    r = subprocess.run([sys.executable, "-c", LOADER, str(path)], env=env, capture_output=True, text=True, check=True)
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.fixture(scope="module")
def big_db(tmp_path_factory):
    d = tmp_path_factory.mktemp("ingest")
    sym = synth.make_db(200_000, L=60, seed=77, noise=0.02)
    synth.write_fasta(d / "db.fna", synth.to_ascii(sym))
    for threads, name in ((1, "db1"), (7, "db7")):   # makedb itself: parallel parse + encode vs one thread
        env = dict(os.environ, SMAFA_HOST_THREADS=str(threads))
        subprocess.run([api.CLI_PATH, "makedb", "-i", d / "db.fna", "-d", d / name], env=env, check=True)
    return d, synth.pack_symbols(sym)


def test_makedb_bytes_do_not_depend_on_thread_count(big_db):
    d, _ = big_db
    assert (d / "db1").read_bytes() == (d / "db7").read_bytes()


@pytest.mark.parametrize("threads", [1, 2, 5, 16])
def test_parallel_db_decode_matches_input(big_db, threads):
    import hashlib
    d, words = big_db
    got = _load(d / "db1", threads)
    assert got == {"n": words.shape[0], "W": 5, "L": 60, "sha": hashlib.sha256(words.tobytes()).hexdigest()}


def test_malformed_db_errors_do_not_depend_on_thread_count(big_db):
    d, _ = big_db
    raw = (d / "db1").read_bytes()
    cases = {"truncated": raw[: len(raw) // 2], "no_tail": raw[:-2],
             "bad_inner_len": raw[:1000] + bytes([raw[1000] ^ 0x01]) + raw[1001:]}
    for name, blob in cases.items():
        (d / name).write_bytes(blob)
        one, many = _load(d / name, 1), _load(d / name, 8)
        assert one == many, name
    assert "error" in _load(d / "truncated", 8)


def _varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def test_inconsistent_hand_made_db_is_refused(tmp_path):
    """`len` and the windows' word counts are separate fields of the file (src/lib.rs:54-60).  A hand-made db that
    states a length its windows cannot hold (or none at all) must be refused at load -- the device upload, the TSV decode
    and --limit-per-sequence all derive the row stride from `len` -- and a window count the file cannot hold is a
    truncated file, not an allocation."""
    def db_bytes(n_words, stated_len, n_windows=3):
        body = b"".join(_varint(n_words) + b"".join(_varint(16) for _ in range(n_words)) for _ in range(n_windows))
        tail = b"\x00" if stated_len is None else b"\x01" + _varint(stated_len)
        return _varint(2) + _varint(n_windows) + body + tail
    good = tmp_path / "good.db"
    good.write_bytes(db_bytes(1, 1))
    words, L = smafa_b200.load_db_file(good)
    assert L == 1 and words.shape == (3, 1) and (words == 16).all()
    for name, blob in {"len_too_long": db_bytes(1, 60), "len_too_short": db_bytes(5, 12), "len_none": db_bytes(1, None),
                       "count_beyond_file": _varint(2) + _varint(1 << 40) + _varint(1) + _varint(16) + b"\x01\x01"}.items():
        p = tmp_path / name
        p.write_bytes(blob)
        with pytest.raises(smafa_b200.SmafaError) as e:
            smafa_b200.load_db_file(p)
        assert e.value.status == "SMAFA_E_IO", name
        q = tmp_path / "q.fna"
        q.write_text(">a\nA\n")
        r = subprocess.run([api.CLI_PATH, "query", "-d", p, "-q", q], capture_output=True, text=True)
        assert r.returncode == 1 and r.stdout == "", (name, r.stderr)       # refused before any device work


def test_malformed_fastq_is_rejected_like_needletail(tmp_path):
    """needletail rejects a FASTQ record without a '+' line, with quality and sequence of different lengths, or cut off
    before its fourth line; the reference then panics on that record (src/lib.rs:149,234; src/cluster.rs:39) -- after the
    records before it -- and `count` returns the error (src/lib.rs:385, exit code 1).  Unpinned by reference fixtures:
    the message text is ours, the exit codes and the position of the failure are the reference's."""
    ok = "@r0\nACGT\n+\nIIII\n"
    cases = {"no_plus": ok + "@r1\nACGT\nIIII\nIIII\n", "qual_len": ok + "@r1\nACGT\n+\nIII\n",
             "truncated": ok + "@r1\nACGT\n+\n", "bad_start": ok + "r1\nACGT\n+\nIIII\n"}
    for name, text in cases.items():
        p = tmp_path / (name + ".fq")
        p.write_text(text)
        r = subprocess.run([api.CLI_PATH, "makedb", "-i", p, "-d", tmp_path / "x.db"], capture_output=True, text=True)
        assert r.returncode == 101 and "valid record" in r.stderr, (name, r.stderr)
        r = subprocess.run([api.CLI_PATH, "count", "-i", p], capture_output=True, text=True)
        assert r.returncode == 1 and r.stdout == "", (name, r.stderr)
    good = tmp_path / "good.fq"
    good.write_text(ok + "@r1\nTTGA\n+r1\nIIII")                        # '+' may repeat the id; no final newline
    r = subprocess.run([api.CLI_PATH, "count", "-i", good], capture_output=True, text=True)
    assert r.returncode == 0 and '"num_reads":2,"num_bases":8' in r.stdout


def test_first_bad_record_wins_for_every_thread_count(tmp_path):
    sym = synth.make_db(70_000, L=60, seed=5)
    seqs = synth.to_ascii(sym)
    seqs[40_000] = seqs[40_000][:10] + b"!" + seqs[40_000][11:]
    seqs[65_000] = seqs[65_000][:-1]                       # a later length mismatch must not be reported
    synth.write_fasta(tmp_path / "bad.fna", seqs)
    outs = []
    for threads in (1, 8):
        env = dict(os.environ, SMAFA_HOST_THREADS=str(threads))
        r = subprocess.run([api.CLI_PATH, "makedb", "-i", tmp_path / "bad.fna", "-d", tmp_path / "x"], env=env,
                           capture_output=True, text=True)
        assert r.returncode == 101
        outs.append(r.stderr)
    assert outs[0] == outs[1] and 'Byte 33 cannot be interpreted as nucleotide, in sequence "seq40000" at position 10' in outs[0]


DEDUP = """
import sys, ctypes, hashlib, numpy as np
sys.path.insert(0, %r)
import smafa_b200
from smafa_b200 import synth
lib = smafa_b200.load_library()
sym = synth.make_cluster_input(150_000, L=60, seed=9, dup_fraction=0.2)
w = synth.pack_symbols(sym)
first = np.zeros(w.shape[0], dtype=np.uint8)
lib.smafa_mark_first_occurrences.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p]
assert lib.smafa_mark_first_occurrences(w.ctypes.data, w.shape[0], w.shape[1], first.ctypes.data) == 0
_, idx = np.unique(w, axis=0, return_index=True)
want = np.zeros_like(first); want[idx] = 1
print("OK" if (first == want).all() else "MISMATCH", hashlib.sha256(first.tobytes()).hexdigest(), int(first.sum()))
""" % ROOT


def test_cluster_deduplication_matches_numpy_for_every_thread_count():
    """src/cluster.rs:46-48 (skip an encoding seen before): the hash-partitioned host pass must mark exactly the first
    occurrences, whatever the number of host threads."""
    outs = []
    for threads in (1, 3, 8, 32):
        env = dict(os.environ, SMAFA_HOST_THREADS=str(threads))
        r = subprocess.run([sys.executable, "-c", DEDUP], env=env, capture_output=True, text=True, check=True)
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0].startswith("OK") and len(set(outs)) == 1, outs


def _py_encode(seqs, L):
    """Reference word layout (src/lib.rs:29-52) of already-clean sequences, for the format tests below."""
    code = {**{c: 16 for c in "Aa"}, **{c: 8 for c in "Cc"}, **{c: 4 for c in "Gg"}, **{c: 2 for c in "TtUu"}}
    W = (L + 11) // 12
    out = np.zeros((len(seqs), W), dtype=np.uint64)
    for i, s in enumerate(seqs):
        assert len(s) == L
        for p, ch in enumerate(s):
            out[i, p // 12] |= np.uint64(code.get(ch, 1)) << np.uint64(5 * (p % 12))
    return out


@pytest.mark.parametrize("threads", [1, 4])
def test_fasta_layout_variants_parse_alike(tmp_path, threads):
    """Line structure must not change what is encoded: multi-line sequences, CRLF line ends, blank lines, '>' inside a
    header, lower case, no newline at the end of the file.  Only the single-line / no-trailing-newline shapes are pinned
    by reference fixtures (tests/data/subjects.fa); the others follow needletail's documented behaviour
    and are checked against the oracle CLI and a Python model of the encoding."""
    rng = np.random.default_rng(3)
    L, n = 60, 9000
    seqs = ["".join(rng.choice(list("ACGTNacgt-RY"), size=L)) for _ in range(n)]
    want = _py_encode(seqs, L)
    variants = {
        "plain": "".join(f">s{i} d>x\n{s}\n" for i, s in enumerate(seqs)),
        "no_final_newline": "".join(f">s{i}\n{s}\n" for i, s in enumerate(seqs))[:-1],
        "multiline": "".join(f">s{i}\n{s[:17]}\n{s[17:40]}\n{s[40:]}\n" for i, s in enumerate(seqs)),
        "crlf": "".join(f">s{i}\r\n{s}\r\n" for i, s in enumerate(seqs)),
        "crlf_multiline": "".join(f">s{i}\r\n{s[:30]}\r\n{s[30:]}\r\n" for i, s in enumerate(seqs)),
        "blank_lines": "".join(f">s{i}\n{s}\n\n" for i, s in enumerate(seqs)),
    }
    env = dict(os.environ, SMAFA_HOST_THREADS=str(threads))
    from oracle import c_oracle
    c_oracle.build()
    for name, text in variants.items():
        fa = tmp_path / f"{name}.fna"
        reps = 1 if threads == 1 else 20  # > 8 MB: the multi-threaded parse
        sep = "\n" if name == "no_final_newline" else ""
        fa.write_bytes(sep.join([text] * reps).encode())
        r = subprocess.run([api.CLI_PATH, "makedb", "-i", fa, "-d", tmp_path / f"{name}.db"], env=env, capture_output=True, text=True)
        assert r.returncode == 0, (name, r.stderr)
        w, got_L = api.load_db_file(str(tmp_path / f"{name}.db"))
        assert got_L == L and w.shape == (n * reps, 5), name
        assert (w.reshape(reps, n, 5) == want[None]).all(), name
        if threads == 1 and name in ("plain", "no_final_newline", "multiline", "crlf"):
            o = subprocess.run([c_oracle.CLI, "makedb", "-i", fa, "-d", tmp_path / f"{name}.odb"], capture_output=True, text=True)
            assert o.returncode == 0, (name, o.stderr)
            assert (tmp_path / f"{name}.odb").read_bytes() == (tmp_path / f"{name}.db").read_bytes(), name
