cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
C3_D=2000000 C3_Q=200000 timeout 600 python scripts/config3_multi.py > gpurun_out/c3_small_1gpu.json 2> gpurun_out/c3_small_1gpu.err; echo "exit=$?"; cat gpurun_out/c3_small_1gpu.json; tail -5 gpurun_out/c3_small_1gpu.err
