cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29502 scripts/config3_multi.py > gpurun_out/c3_${N}gpu.json 2> gpurun_out/c3_${N}gpu.err; echo "c3 exit=$?"; tail -1 gpurun_out/c3_${N}gpu.json; tail -3 gpurun_out/c3_${N}gpu.err
timeout 300 $TR --master-port 29503 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_modea.json 2> gpurun_out/bench_${N}gpu_modea.err; echo "bench exit=$?"; tail -1 gpurun_out/bench_${N}gpu_modea.json
nvidia-smi --query-gpu=index,clocks.sm,power.draw,power.limit,temperature.gpu --format=csv > gpurun_out/smi_8gpu.csv
