set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --no-cpu-baseline --queries 25600 --db-per-gpu 262144"
timeout 300 python bench.py --kernel mma $ARGS > gpurun_out/plain_mma.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan_mma -s 1 -c 1 -o gpurun_out/prof_mma python bench.py --kernel mma $ARGS > gpurun_out/ncu_mma.log 2>&1
timeout 300 python bench.py --kernel popc $ARGS > gpurun_out/plain_popc.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan_popc -s 1 -c 1 -o gpurun_out/prof_popc2 python bench.py --kernel popc $ARGS > gpurun_out/ncu_popc.log 2>&1
cat gpurun_out/plain_mma.log gpurun_out/plain_popc.log | cut -c1-400
tail -3 gpurun_out/ncu_mma.log
