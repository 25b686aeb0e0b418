# Round 2, GPU call 1: picker calibration data, 4-stage union kernel under the whole parity suite.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python scripts/union_calib.py > gpurun_out/r02_union_calib.log 2>&1; echo "calib exit=$?"
SMAFA_MMA_UNION_STAGES4=1 CALIB_ONLY=bench timeout 600 python scripts/union_calib.py > gpurun_out/r02_union_calib_stages4.log 2>&1; echo "calib4 exit=$?"
SMAFA_MMA_UNION_STAGES4=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_stages4.log 2>&1; echo "pytest(stages4) exit=$?"; tail -3 gpurun_out/r02_pytest_gpu_stages4.log
tail -40 gpurun_out/r02_union_calib.log
