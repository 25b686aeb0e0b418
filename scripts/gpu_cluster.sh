# Cluster at BASELINE config 5 scale (run under gpurun): API-level bench of both kernels with stage timers, then the CLI
# end to end (FASTA in, TSV out) next to the oracle CLI on a prefix.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
CLUSTER_N=${CLUSTER_N:-5000000} CLUSTER_CHECK=60000 SMAFA_TIMING=1 timeout 800 python scripts/cluster_bench.py > gpurun_out/cluster_bench.log 2>&1; echo "cluster_bench exit=$?"; grep -v "cluster: [28] batches" gpurun_out/cluster_bench.log
CLUSTER_N=${CLUSTER_N:-5000000} timeout 800 python scripts/cli_cluster_e2e.py > gpurun_out/cli_cluster_e2e.log 2>&1; echo "cli exit=$?"; cat gpurun_out/cli_cluster_e2e.log
