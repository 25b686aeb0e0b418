set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
timeout 600 python bench.py --kernel popc --steps 3 --warmup 3 > gpurun_out/bench_popc_a.json 2> gpurun_out/bench_popc_a.err; echo "bench exit=$?"
cat gpurun_out/bench_popc_a.json
timeout 600 python bench.py --kernel popc --mode b --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_popc_b.json 2> gpurun_out/bench_popc_b.err
cat gpurun_out/bench_popc_b.json
SMALL="--kernel popc --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 python bench.py $SMALL > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_popc.csv python bench.py $SMALL > gpurun_out/ncu1.log 2>&1
timeout 300 python bench.py $SMALL > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_popc -s 1 -c 1 -o gpurun_out/prof_popc python bench.py $SMALL > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
