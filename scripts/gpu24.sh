set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CLUSTER_N=5000000 CLUSTER_CHECK=0 SMAFA_TIMING=1 timeout 600 python scripts/cluster_bench.py > gpurun_out/cluster_5m.log 2>&1; cat gpurun_out/cluster_5m.log
python - <<'PY'
import sys, subprocess, os, tempfile
sys.path.insert(0, ".")
from smafa_b200 import api, synth
tmp = tempfile.mkdtemp()
db = synth.make_db(100000, L=60); q = synth.make_queries(db, 1000)
synth.write_fasta(f"{tmp}/db.fna", synth.to_ascii(db)); synth.write_fasta(f"{tmp}/q.fna", synth.to_ascii(q))
subprocess.run([api.CLI_PATH, "makedb", "-i", f"{tmp}/db.fna", "-d", f"{tmp}/db"], check=True)
env = dict(os.environ, SMAFA_TIMING="1")
for i in range(2):
    r = subprocess.run(["bash", "-c", f"time {api.CLI_PATH} query -d {tmp}/db -q {tmp}/q.fna --max-divergence 5 > /dev/null"], env=env, capture_output=True, text=True)
    print(r.stderr)
PY
echo "CUDA_VISIBLE_DEVICES=$CUDA_VISIBLE_DEVICES"; nvidia-smi -L | head -3; nvidia-smi --query-gpu=persistence_mode --format=csv
