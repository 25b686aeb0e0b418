set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 180 python scripts/mma_probe.py > gpurun_out/probe.log 2>&1; P=$?; echo "probe exit=$P" >> gpurun_out/probe.log; tail -6 gpurun_out/probe.log
timeout 120 python -c "
import smafa_b200
c=smafa_b200.Context(0)
for n in (2000,20000,100000): print('int8 peak TOP/s', n, c.mma_peak_tops(n))
" > gpurun_out/peak.log 2>&1; cat gpurun_out/peak.log
timeout 900 python -m pytest tests -m gpu -x -q -k "mma or kats or cli" > gpurun_out/pytest_mma.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_mma.log; tail -5 gpurun_out/pytest_mma.log
timeout 600 python bench.py --kernel mma --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mma_a.json 2> gpurun_out/bench_mma_a.err; cat gpurun_out/bench_mma_a.json; tail -3 gpurun_out/bench_mma_a.err
timeout 600 python bench.py --kernel mma --mode b --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mma_b.json 2> gpurun_out/bench_mma_b.err; cat gpurun_out/bench_mma_b.json
ARGS="--steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 python bench.py --kernel mma $ARGS > gpurun_out/plain_mma.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan_mma -s 1 -c 1 -o gpurun_out/prof_mma2 python bench.py --kernel mma $ARGS > gpurun_out/ncu_mma.log 2>&1
tail -2 gpurun_out/ncu_mma.log
