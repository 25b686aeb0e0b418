set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=2
for MODE in a b; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 --mode $MODE > gpurun_out/bench_${N}gpu_$MODE.json 2> gpurun_out/bench_${N}gpu_$MODE.err; echo "bench exit=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${N}gpu_$MODE.json").read().strip().splitlines()[-1])
print("RESULT N=$N mode=$MODE value=%.3e e2e=%.3e ms=%.2f scan_ms=%.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["scan_ms_per_step"]))
PY
done
CLUSTER_N=5000000 CLUSTER_CHECK=0 SMAFA_TIMING=1 timeout 600 python scripts/cluster_bench.py > gpurun_out/cluster_5m.log 2>&1; cat gpurun_out/cluster_5m.log
CLUSTER_N=500000 CLUSTER_CHECK=40000 timeout 600 python scripts/cluster_bench.py > gpurun_out/cluster_500k.log 2>&1; cat gpurun_out/cluster_500k.log
