set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_query.py -m gpu -x -q -k config3 ) > gpurun_out/pytest_config3.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_config3.log; tail -12 gpurun_out/pytest_config3.log
