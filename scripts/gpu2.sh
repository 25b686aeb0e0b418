set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 180 python scripts/mma_probe.py > gpurun_out/probe.log 2>&1; P=$?
echo "probe exit=$P" >> gpurun_out/probe.log
if [ $P -ne 0 ]; then
  SMAFA_MMA_SWAP_LBO_SBO=1 timeout 180 python scripts/mma_probe.py > gpurun_out/probe_swap.log 2>&1; echo "probe_swap exit=$?" >> gpurun_out/probe_swap.log
  cat gpurun_out/probe_swap.log | tail -20
fi
cat gpurun_out/probe.log | tail -20
PROBE_L=20 timeout 180 python scripts/mma_probe.py > gpurun_out/probe20.log 2>&1; echo "probe20 exit=$?" >> gpurun_out/probe20.log; tail -8 gpurun_out/probe20.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
timeout 600 python bench.py --kernel popc --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_popc_a.json 2> gpurun_out/bench_popc_a.err; cat gpurun_out/bench_popc_a.json
if [ $P -eq 0 ]; then
timeout 600 python bench.py --kernel mma --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mma_a.json 2> gpurun_out/bench_mma_a.err; cat gpurun_out/bench_mma_a.json; tail -3 gpurun_out/bench_mma_a.err
timeout 600 python bench.py --kernel mma --mode b --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mma_b.json 2> gpurun_out/bench_mma_b.err; cat gpurun_out/bench_mma_b.json
fi
