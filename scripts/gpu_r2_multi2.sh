# Round 2, 2-GPU call: whole parity suite (multi-GPU tests included), bench at N = 1 and N = 2, phase timings.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_${N}gpu.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/r02_pytest_gpu_${N}gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench1 exit=$?"; tail -c 1500 gpurun_out/r02_bench_1gpu.json; tail -3 gpurun_out/r02_bench_1gpu.err
timeout 600 $TR --master-port 29503 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench$N exit=$?"; tail -c 1500 gpurun_out/r02_bench_${N}gpu.json; tail -5 gpurun_out/r02_bench_${N}gpu.err
timeout 300 $TR --master-port 29504 scripts/dist_phases.py > gpurun_out/r02_dist_phases_${N}gpu.log 2>&1; echo "phases exit=$?"; tail -4 gpurun_out/r02_dist_phases_${N}gpu.log
