set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 180 python scripts/mma_probe.py > gpurun_out/probe.log 2>&1; P=$?; echo "probe exit=$P" >> gpurun_out/probe.log; tail -6 gpurun_out/probe.log
timeout 120 python -c "
import smafa_b200
c=smafa_b200.Context(0)
for n in (20000,100000,100000): print('int8 peak TOP/s', n, c.mma_peak_tops(n))
" > gpurun_out/peak.log 2>&1; cat gpurun_out/peak.log
timeout 600 python bench.py --kernel mma --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mma_a.json 2> gpurun_out/bench_mma_a.err; cat gpurun_out/bench_mma_a.json; tail -3 gpurun_out/bench_mma_a.err
timeout 600 python bench.py --kernel mma --mode b --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mma_b.json 2> gpurun_out/bench_mma_b.err; cat gpurun_out/bench_mma_b.json
