set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ARGS="--kernel mma --mode b --max-divergence none --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 python bench.py $ARGS > gpurun_out/plain_b.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_b_none.csv python bench.py $ARGS > gpurun_out/ncu_b1.log 2>&1
timeout 300 python bench.py $ARGS > gpurun_out/plain_b2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan_mma -s 1 -c 1 -o gpurun_out/prof_mma_b_none python bench.py $ARGS > gpurun_out/ncu_b2.log 2>&1
tail -2 gpurun_out/ncu_b2.log
