"""End-to-end `smafa cluster` CLI run at BASELINE config 5 scale (FASTA in, TSV out) with per-stage host timings
(SMAFA_TIMING=1), next to the oracle CLI on a prefix of the input (full-output equality on that prefix: the greedy is
order dependent, so a prefix of the input gives a prefix of the output)."""
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle
from smafa_b200 import api, synth

N = int(os.environ.get("CLUSTER_N", "5000000"))
SUB = int(os.environ.get("CLUSTER_SUB", "60000"))
T = "3"
c_oracle.build()
tmp = tempfile.mkdtemp(prefix="smafa_cluster_e2e_")
t0 = time.perf_counter()
sym = synth.make_cluster_input(N, L=60)
synth.write_fasta(f"{tmp}/in.fna", synth.to_ascii(sym))
synth.write_fasta(f"{tmp}/sub.fna", synth.to_ascii(sym[:SUB]))
print(f"host threads available: {os.cpu_count()}; n={N}; input generated in {time.perf_counter() - t0:.1f} s")
env = dict(os.environ, SMAFA_TIMING="1")


def run(cmd):
    t = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, env=env)
    dt = time.perf_counter() - t
    if r.returncode != 0:
        print("FAILED", cmd, r.stderr.decode()[-2000:])
        sys.exit(1)
    return r, dt


r, dt = run([api.CLI_PATH, "cluster", "-i", f"{tmp}/in.fna", "-d", T])
full = r.stdout
print(f"smafa cluster -d {T}: {dt:.3f} s wall, {full.count(10)} lines (process start, CUDA init, file I/O included)\n{r.stderr.decode()}")
rs, dts = run([api.CLI_PATH, "cluster", "-i", f"{tmp}/sub.fna", "-d", T])
ro, dto = run([c_oracle.CLI, "cluster", "-i", f"{tmp}/sub.fna", "-d", T])
same = rs.stdout == ro.stdout
head = full[:len(ro.stdout)] == ro.stdout
print(f"  oracle CLI (1 thread) on the first {SUB} sequences: {dto:.3f} s wall; stdout identical to the GPU CLI: {same}; "
      f"identical to the head of the full run: {head}")
sys.exit(0 if same and head else 2)
