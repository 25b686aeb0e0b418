# 2-GPU check of the round's final tree (run under gpurun --gpus 2): NCCL parity on a small db, then the weak-scaling bench.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( CHECK_Q=1500 timeout 60 $TR --master-port 29501 scripts/dist_check.py > gpurun_out/dist_check_2gpu_v9.log 2>&1; echo "dist_check exit=$?"; tail -1 gpurun_out/dist_check_2gpu_v9.log | cut -c1-400 ) &
wait
timeout 60 $TR --master-port 29503 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_v9.json 2> gpurun_out/bench_2gpu_v9.err; echo "bench exit=$?"; tail -1 gpurun_out/bench_2gpu_v9.json | cut -c1-700
