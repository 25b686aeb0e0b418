"""Host ingest paths of the B200 build: the multi-threaded LEB128 db decode, FASTA parse and
window encoding must give byte-identical results to the single-threaded pass for every thread count, and the same
errors on malformed input.  CPU only."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

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
