"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise).  Both ways of using several GPUs must reproduce the oracle's
output on the whole db:
  * one process per GPU (torchrun): smafa_ctx_comm_init + smafa_db_upload_shard + smafa_query_sharded[_dev]
    (scripts/dist_check.py: every mode, both kernels, overflow re-send, empty shards, two slabs);
  * one process, several GPUs: smafa_ctx_create_multi behind the ordinary calls and behind the CLI's --devices
    (the reference's golden vectors and the 1 k x 10 k full-stdout diff)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")


def _devices():
    return list(range(min(_n_gpus(), 4)))


@needs2
def test_sharded_query_matches_oracle():
    n = min(_n_gpus(), 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "scripts", "dist_check.py")], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ok"], res


@needs2
def test_multi_device_context_matches_oracle():
    import smafa_b200
    from oracle import c_oracle
    from smafa_b200 import build, synth
    build.build()
    c_oracle.build()
    L = 60
    db_sym = synth.make_db(100_003, L=L, seed=71)
    db = synth.pack_symbols(db_sym)
    q = synth.pack_symbols(synth.make_queries(db_sym, 1000, seed=72))
    threads = os.cpu_count() or 1
    wants = {}
    c = smafa_b200.Context(_devices(), "auto")
    try:
        d = c.upload(db, L)
        assert d.size == db.shape[0]
        for kernel in ("mma", "popc"):
            c.set_kernel(kernel)
            for m, k, r in [(5, None, None), (None, None, None), (5, 10, None), (None, 10, None), (8, 25, 2), (3, 1, None),
                            (60, 100_003, None)]:
                got, st = c.query(d, q, L, max_divergence=m, max_num_hits=k, limit_per_sequence=r, return_stats=True)
                if (m, k, r) not in wants:
                    wants[(m, k, r)] = c_oracle.query(db, L, q, L, m, k, r, threads=threads)
                want = wants[(m, k, r)]
                assert got.shape == want.shape and (got == want).all(), (kernel, m, k, r)
                assert st["pairs"] == q.shape[0] * db.shape[0]
        # get_distances gathers the shards' columns
        got = c.distances(d, q[:7], L)
        for i in range(7):
            assert (got[i] == c_oracle.distances(db, q[i])).all()
        # appended rows join the last shard and keep their global numbers
        extra = synth.pack_symbols(synth.make_db(501, L=L, seed=73))
        d.append(extra)
        both = np.concatenate([db, extra])
        got = c.query(d, q, L, max_divergence=6, max_num_hits=4)
        want = c_oracle.query(both, L, q, L, 6, 4, None, threads=threads)
        assert got.shape == want.shape and (got == want).all()
        d.close()
        # fewer windows than devices; the reference's panics keep their order
        tiny = c.upload(db[:1], L)
        got = c.query(tiny, q[:50], L)
        want = c_oracle.query(db[:1], L, q[:50], L, None, None, None)
        assert (got == want).all()
        tiny.close()
        empty = c.upload(db[:0], L)
        with pytest.raises(smafa_b200.SmafaPanic):
            c.query(empty, q[:5], L)
        empty.close()
        # cluster: the greedy runs on the first device
        sym = synth.make_cluster_input(6000, L=L, seed=74)
        enc_all = synth.pack_symbols(sym)
        want_cof, want_nc, want_cmp = c_oracle.cluster(enc_all, L, 3)
        keep = want_cof >= 0
        remap = np.cumsum(keep) - 1
        cof, nc, ncmp = c.cluster(enc_all[keep], L, 3)
        assert nc == want_nc and ncmp == want_cmp and (cof.astype(np.int64) == remap[want_cof[keep]]).all()
    finally:
        c.close()


@needs2
def test_cli_devices_flag(kats, kat_dir, tmp_path):
    from oracle import c_oracle
    from smafa_b200 import api, build, synth
    build.build()
    c_oracle.build()
    devs = ",".join(map(str, _devices()))

    def cli(*args):
        return subprocess.run([api.CLI_PATH, *map(str, args)], capture_output=True, text=True)

    # the reference's golden vectors (dbs of 2-5 windows: most shards are empty)
    for case in kats["query"]:
        if "makedb_from" in case:
            db = tmp_path / (case["name"] + ".db")
            assert cli("makedb", "-i", kat_dir / case["makedb_from"], "-d", db).returncode == 0
        else:
            db = kat_dir / case["db"]
        r = cli("query", "-d", db, "-q", kat_dir / case["query"], "--devices", devs, *case["args"])
        assert r.returncode == 0, (case["name"], r.stderr)
        assert r.stdout == case["stdout"], case["name"]
    for case in kats["cluster"]:
        r = cli("cluster", "-i", kat_dir / case["input"], "-d", case["t"], "--devices", devs)
        assert r.returncode == 0 and r.stdout == case["stdout"], case["name"]
    r = cli("query", "-d", kat_dir / "random_3_2.fna.v1.smafadb", "-q", kat_dir / "random_3_2.fna", "--devices", devs)
    assert r.returncode != 0 and "Unsupported db file version: 1." in r.stderr
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
        got = cli("query", "-d", tmp_path / "db", "-q", tmp_path / "q.fna", "--devices", devs, *args)
        assert got.returncode == 0, got.stderr
        assert got.stdout == want.stdout, args
