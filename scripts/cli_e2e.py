"""End-to-end `smafa` CLI run at BASELINE config 2 scale (files in, TSV out) with per-stage host timings
(SMAFA_TIMING=1), next to the oracle CLI on a query subsample (full-output equality on that subsample)."""
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle
from smafa_b200 import api, synth

D = int(os.environ.get("E2E_D", "1000000"))
Q = int(os.environ.get("E2E_Q", "100000"))
SUB = int(os.environ.get("E2E_SUB", "2000"))
SUB_B = int(os.environ.get("E2E_SUB_B", "50"))   # Mode B: the reference sorts all D distances per query (~1 s each at 10 M)
c_oracle.build()
tmp = tempfile.mkdtemp(prefix="smafa_e2e_")
db_sym = synth.make_db(D, L=60)
q_sym = synth.make_queries(db_sym, Q)
synth.write_fasta(f"{tmp}/db.fna", synth.to_ascii(db_sym))
synth.write_fasta(f"{tmp}/q.fna", synth.to_ascii(q_sym))
synth.write_fasta(f"{tmp}/qsub.fna", synth.to_ascii(q_sym[:SUB]))
synth.write_fasta(f"{tmp}/qsubb.fna", synth.to_ascii(q_sym[:SUB_B]))
env = dict(os.environ, SMAFA_TIMING="1")
DEV = ["--devices", os.environ["E2E_DEVICES"]] if os.environ.get("E2E_DEVICES") else []   # row-shard the db over several GPUs


def run(cmd, **kw):
    t = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, env=env, **kw)
    dt = time.perf_counter() - t
    if r.returncode != 0:
        print("FAILED", cmd, r.stderr.decode()[-2000:])
        sys.exit(1)
    return r, dt


print(f"host threads available: {os.cpu_count()}; D={D} Q={Q}")
r, dt = run([api.CLI_PATH, "makedb", "-i", f"{tmp}/db.fna", "-d", f"{tmp}/db.smafadb"])
print(f"smafa makedb: {dt:.3f} s wall\n{r.stderr.decode()}")
for args in (["--max-divergence", "5"], ["--max-divergence", "5", "--max-num-hits", "10"]):
    r, dt = run([api.CLI_PATH, "query", "-d", f"{tmp}/db.smafadb", "-q", f"{tmp}/q.fna", *DEV, *args])
    full = r.stdout
    print(f"smafa query {' '.join(DEV + args)}: {dt:.3f} s wall, {full.count(10)} lines, {Q * D / dt:.3e} comparisons/s end to end "
          f"(process start, CUDA init, file I/O included)\n{r.stderr.decode()}")
    sub, subf = (SUB_B, "qsubb.fna") if "--max-num-hits" in args else (SUB, "qsub.fna")
    rs, dts = run([api.CLI_PATH, "query", "-d", f"{tmp}/db.smafadb", "-q", f"{tmp}/{subf}", *DEV, *args])
    ro, dto = run([c_oracle.CLI, "query", "-d", f"{tmp}/db.smafadb", "-q", f"{tmp}/{subf}", *args])
    same = rs.stdout == ro.stdout
    head = b"".join(l + b"\n" for l in full.split(b"\n") if l and int(l.split(b"\t", 1)[0]) < sub)
    print(f"  oracle CLI (1 thread) on the first {sub} queries: {dto:.3f} s wall = {sub * D / dto:.3e} comparisons/s; "
          f"stdout identical to the GPU CLI: {same}; identical to the head of the full run: {head == ro.stdout}")
    if not same:
        sys.exit(2)
