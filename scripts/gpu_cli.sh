# CLI end to end (run under gpurun): config 2 and config 3 query runs and the config 5 cluster run, stage timers on,
# each next to the oracle CLI on a subsample.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python scripts/cli_e2e.py > gpurun_out/cli_e2e_config2.log 2>&1; echo "config2 exit=$?"; cat gpurun_out/cli_e2e_config2.log
if [ -n "$WITH_CONFIG3" ]; then
E2E_D=10000000 E2E_Q=1000000 E2E_SUB=300 E2E_SUB_B=12 timeout 1500 python scripts/cli_e2e.py > gpurun_out/cli_e2e_config3.log 2>&1; echo "config3 exit=$?"; cat gpurun_out/cli_e2e_config3.log
fi
timeout 800 python scripts/cli_cluster_e2e.py > gpurun_out/cli_cluster_e2e.log 2>&1; echo "cluster exit=$?"; cat gpurun_out/cli_cluster_e2e.log
timeout 600 python -m pytest tests/test_gpu_query.py -m gpu -x -q -k "kats or panics or cli" > gpurun_out/pytest_cli.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest_cli.log
