# Round 2, N-GPU call (gpurun --gpus N with NGPU=N): parity of the sharded path at N ranks, the bench line at N, BASELINE
# configs[2] (1 M x 10 M) through bench.py at N = 8, phase timings.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29501 scripts/dist_check.py > gpurun_out/r02_dist_check_${N}gpu_v14.log 2> gpurun_out/r02_dist_check_${N}gpu_v14.err; echo "dist_check exit=$?"; tail -c 2500 gpurun_out/r02_dist_check_${N}gpu_v14.log; grep "failed cases" gpurun_out/r02_dist_check_${N}gpu_v14.err
timeout 600 $TR --master-port 29503 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu_v14.json 2> gpurun_out/r02_bench_${N}gpu_v14.err; echo "bench$N exit=$?"; tail -c 900 gpurun_out/r02_bench_${N}gpu_v14.json; grep -v "^\[W\|^W\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/r02_bench_${N}gpu_v14.err | tail -5
if [ "$N" = "8" ]; then
timeout 900 $TR --master-port 29505 bench.py --gpus $N --steps 5 --warmup 3 --db-per-gpu 1250000 --queries 1000000 > gpurun_out/r02_config3_1Mx10M_${N}gpu_v14.json 2> gpurun_out/r02_config3_${N}gpu_v14.err; echo "config3 exit=$?"; tail -c 900 gpurun_out/r02_config3_1Mx10M_${N}gpu_v14.json; grep -v "^\[W\|^W\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/r02_config3_${N}gpu_v14.err | tail -5
true
fi
true
