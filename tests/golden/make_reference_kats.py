#!/usr/bin/env python3
"""Regenerates tests/golden/reference_kats.json from the reference checkout.

Run in the authoring container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/make_reference_kats.py

The *inputs* are read from the reference's fixture files (tests/data/); the *expected
outputs* are the literal strings the reference's own tests assert, transcribed here with
their file:line so a reviewer can check them.  Nothing is computed by our own code, so the
JSON is an independent pin for both the oracle and the CUDA path.
"""
import base64
import gzip
import json
import os

REF = os.environ.get("SMAFA_REFERENCE", "/root/reference")
DATA = os.path.join(REF, "tests", "data")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")


def read(name, binary=False):
    with open(os.path.join(DATA, name), "rb") as f:
        b = f.read()
    return b if binary else b.decode()


def b64(name):
    return base64.b64encode(read(name, binary=True)).decode()


files = {n: read(n) for n in [
    "random_3_2.fna", "random_3_2_one_repeated.fna", "degenerate.fna", "subjects.fa",
    "cluster_dummy1.fna", "cluster_bug1.fna", "cluster_best_hit_changes.fna"]}
binary_files = {n: b64(n) for n in [
    "random_3_2.fna.smafadb", "random_3_2_one_repeated.fna.smafadb",
    "random_3_2.fna.v1.smafadb", "random_30_4.fq.gz"]}

Q = "random_3_2.fna"
kats = {
    "files": files,
    "binary_files_b64": binary_files,
    # src/lib.rs:357-366  test_makedb: tests/data/subjects.fa -> one word per window
    "encoding": {"input": "subjects.fa", "words": [[16], [8], [4], [2], [1]], "cite": "src/lib.rs:357-366"},
    # db fixtures consumed directly by tests/test_cmdline.rs:77-181,204-247; makedb of the
    # matching .fna must reproduce them byte for byte
    "makedb": [
        {"input": "random_3_2.fna", "db": "random_3_2.fna.smafadb"},
        {"input": "random_3_2_one_repeated.fna", "db": "random_3_2_one_repeated.fna.smafadb"},
    ],
    "query": [
        {"name": "dna_makedb_and_query", "cite": "tests/test_cmdline.rs:9-25",
         "makedb_from": "random_3_2.fna", "query": Q, "args": [],
         "stdout": "0\t0\t0\tCTT\n1\t1\t0\tAGG\n"},
        {"name": "degenerate_makedb_and_query", "cite": "tests/test_cmdline.rs:43-74",
         "makedb_from": "degenerate.fna", "query": "degenerate.fna", "args": ["--max-num-hits", "99"],
         "stdout": "0\t0\t0\tCTTNGG\n0\t1\t5\tAGGTGA\n0\t2\t6\tNACTTT\n"
                   "1\t1\t0\tAGGTGA\n1\t0\t5\tCTTNGG\n1\t2\t5\tNACTTT\n"
                   "2\t2\t0\tNACTTT\n2\t1\t5\tAGGTGA\n2\t0\t6\tCTTNGG\n"},
        {"name": "max_divergence_unlimited", "cite": "tests/test_cmdline.rs:76-97",
         "db": "random_3_2.fna.smafadb", "query": Q,
         "args": ["--max-divergence", "99", "--max-num-hits", "99"],
         "stdout": "0\t0\t0\tCTT\n0\t1\t3\tAGG\n1\t1\t0\tAGG\n1\t0\t3\tCTT\n"},
        {"name": "max_divergence_limited", "cite": "tests/test_cmdline.rs:99-118",
         "db": "random_3_2.fna.smafadb", "query": Q,
         "args": ["--max-divergence", "2", "--max-num-hits", "99"],
         "stdout": "0\t0\t0\tCTT\n1\t1\t0\tAGG\n"},
        {"name": "max_divergence_equal", "cite": "tests/test_cmdline.rs:120-141",
         "db": "random_3_2.fna.smafadb", "query": Q,
         "args": ["--max-divergence", "3", "--max-num-hits", "99"],
         "stdout": "0\t0\t0\tCTT\n0\t1\t3\tAGG\n1\t1\t0\tAGG\n1\t0\t3\tCTT\n"},
        {"name": "max_num_hits1", "cite": "tests/test_cmdline.rs:143-160",
         "db": "random_3_2.fna.smafadb", "query": Q, "args": ["--max-num-hits", "1"],
         "stdout": "0\t0\t0\tCTT\n1\t1\t0\tAGG\n"},
        {"name": "max_num_hits_more", "cite": "tests/test_cmdline.rs:162-181",
         "db": "random_3_2.fna.smafadb", "query": Q, "args": ["--max-num-hits", "99"],
         "stdout": "0\t0\t0\tCTT\n0\t1\t3\tAGG\n1\t1\t0\tAGG\n1\t0\t3\tCTT\n"},
        {"name": "limit_per_sequence_no_limit", "cite": "tests/test_cmdline.rs:203-224",
         "db": "random_3_2_one_repeated.fna.smafadb", "query": Q, "args": ["--max-num-hits", "99"],
         "stdout": "0\t0\t0\tCTT\n0\t1\t3\tAGG\n0\t2\t3\tAGG\n1\t1\t0\tAGG\n1\t2\t0\tAGG\n1\t0\t3\tCTT\n"},
        {"name": "limit_per_sequence_limit1", "cite": "tests/test_cmdline.rs:226-247",
         "db": "random_3_2_one_repeated.fna.smafadb", "query": Q,
         "args": ["--max-num-hits", "99", "--limit-per-sequence", "1"],
         "stdout": "0\t0\t0\tCTT\n0\t1\t3\tAGG\n1\t1\t0\tAGG\n1\t0\t3\tCTT\n"},
    ],
    # tests/test_cmdline.rs:27-41: must fail, stderr contains this substring
    "old_db": {"db": "random_3_2.fna.v1.smafadb", "query": Q,
               "stderr_contains": "Unsupported db file version: 1.", "cite": "tests/test_cmdline.rs:27-41"},
    "cluster": [
        {"name": "simple", "cite": "src/cluster.rs:101-112", "input": "cluster_dummy1.fna", "t": 1,
         "stdout": "ATGC\tATGC\nATGG\tATGC\nAAAA\tAAAA\n"},
        {"name": "bug1", "cite": "src/cluster.rs:114-124", "input": "cluster_bug1.fna", "t": 2,
         "stdout": "ATGCAAAAA\tATGCAAAAA\nATAAAAAAA\tATGCAAAAA\nTTAAAAAAA\tTTAAAAAAA\n"},
        {"name": "best_hit_changes", "cite": "src/cluster.rs:126-143",
         "input": "cluster_best_hit_changes.fna", "t": 2,
         "stdout": "ATGCAAAAA\tATGCAAAAA\nATAAAAAAA\tATGCAAAAA\nTTAAAAAAA\tTTAAAAAAA\n"},
    ],
    # tests/test_cmdline.rs:183-201; the reference prints the path exactly as given
    "count": [
        {"input": "random_3_2.fna", "num_reads": 2, "num_bases": 6, "cite": "tests/test_cmdline.rs:183-191"},
        {"input": "random_30_4.fq.gz", "num_reads": 4, "num_bases": 120, "cite": "tests/test_cmdline.rs:193-201"},
    ],
}

# sanity: the gz fixture really holds 4 x 30 nt
assert sum(1 for _ in gzip.decompress(read("random_30_4.fq.gz", True)).splitlines()) == 16

with open(OUT, "w") as f:
    json.dump(kats, f, indent=1, sort_keys=True)
print("wrote", OUT)
