cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CLUSTER_N=500000 timeout 900 python scripts/cluster_bench.py 2>&1 | tee gpurun_out/cluster_500k.log
CLUSTER_N=5000000 CLUSTER_CHECK=0 timeout 1200 python scripts/cluster_bench.py 2>&1 | tee gpurun_out/cluster_5m.log
