set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
E2E_D=10000000 E2E_Q=1000000 E2E_SUB=300 E2E_SUB_B=12 timeout 1500 python scripts/cli_e2e.py > gpurun_out/cli_e2e_c3.log 2>&1; echo "e2e exit=$?"; cat gpurun_out/cli_e2e_c3.log
