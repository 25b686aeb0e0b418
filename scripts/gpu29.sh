set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --mode b --max-divergence none --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_ub.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_mma -s 1 -c 1 -f -o gpurun_out/prof_mma_unbounded_b_v8 python bench.py --mode b --max-divergence none --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ub.log 2>&1
echo "ncu exit=$?"; tail -5 gpurun_out/ncu_ub.log
