# Multi-GPU parity + bench (run under gpurun --gpus N with NGPU=N): small-db parity of every mode and both kernels,
# configs[2] (1 M x 10 M unless C3_D / C3_Q say otherwise) with the oracle check on rank 0, then bench.py.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -z "$SKIP_DIST_CHECK" ]; then
timeout 600 $TR --master-port 29501 scripts/dist_check.py > gpurun_out/dist_check_${N}gpu.log 2>&1; echo "dist_check exit=$?"; tail -1 gpurun_out/dist_check_${N}gpu.log
fi
timeout 900 $TR --master-port 29502 scripts/config3_multi.py > gpurun_out/c3_${N}gpu.json 2> gpurun_out/c3_${N}gpu.err; echo "c3 exit=$?"; tail -1 gpurun_out/c3_${N}gpu.json; tail -3 gpurun_out/c3_${N}gpu.err
for mode in ${MODES:-a}; do
timeout 600 $TR --master-port 29503 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --mode $mode > gpurun_out/bench_${N}gpu_mode$mode.json 2> gpurun_out/bench_${N}gpu_mode$mode.err; echo "bench exit=$?"; tail -1 gpurun_out/bench_${N}gpu_mode$mode.json
done
