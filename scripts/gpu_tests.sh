# GPU parity suite (run under gpurun): every -m gpu test, log under gpurun_out/.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/pytest_gpu_all.log
