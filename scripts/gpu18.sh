set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${NG:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check_${N}gpu.log 2>&1; echo "dist_check exit=$?"; tail -3 gpurun_out/dist_check_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench exit=$?"; tail -1 gpurun_out/bench_${N}gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --mode b > gpurun_out/bench_${N}gpu_modeb.json 2> gpurun_out/bench_${N}gpu_modeb.err; echo "bench exit=$?"; tail -1 gpurun_out/bench_${N}gpu_modeb.json
