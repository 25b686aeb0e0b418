# Default bench, then its ncu launch list and one full capture of the scan kernel (run under gpurun, one GPU).
# ncu runs only after the same command has exited 0 without it; numbers printed under ncu are never bench values.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 1 --no-cpu-baseline ${BENCH_ARGS:-}"
python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit=$?"; tail -c 3000 gpurun_out/bench_default.json
python bench.py $ARGS > gpurun_out/bench_plain.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_mma -s 1 -c 1 -f -o gpurun_out/prof_scan_mma python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
echo "ncu exit=$?"
