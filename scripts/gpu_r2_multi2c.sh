# Round 2, 2-GPU call c (grouped db order as the default): dist_check, whole parity suite, bench N = 1, 2, launch list.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29501 scripts/dist_check.py > gpurun_out/r02_dist_check_${N}gpu_grouped.log 2> gpurun_out/r02_dist_check_${N}gpu_grouped.err; echo "dist_check exit=$?"; tail -c 2500 gpurun_out/r02_dist_check_${N}gpu_grouped.log; grep "failed cases\|MISMATCH\|Error" gpurun_out/r02_dist_check_${N}gpu_grouped.err | tail -5
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_${N}gpu_grouped.log 2>&1; echo "pytest exit=$?"; tail -12 gpurun_out/r02_pytest_gpu_${N}gpu_grouped.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_v11.json 2> gpurun_out/r02_bench_1gpu_v11.err; echo "bench1 exit=$?"; tail -c 700 gpurun_out/r02_bench_1gpu_v11.json; tail -3 gpurun_out/r02_bench_1gpu_v11.err
timeout 600 $TR --master-port 29503 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu_v11.json 2> gpurun_out/r02_bench_${N}gpu_v11.err; echo "bench$N exit=$?"; tail -c 700 gpurun_out/r02_bench_${N}gpu_v11.json; grep -v "^\[W\|^W\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/r02_bench_${N}gpu_v11.err | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_v11.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1; echo "ncu exit=$?"
