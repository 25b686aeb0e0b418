# Round 2, 2-GPU call b: dist_check with its whole report, the rest of the parity suite, grouped-rows validation on GPU 1 meanwhile.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29501 scripts/dist_check.py > gpurun_out/r02_dist_check_${N}gpu.log 2> gpurun_out/r02_dist_check_${N}gpu.err; echo "dist_check exit=$?"; tail -c 3000 gpurun_out/r02_dist_check_${N}gpu.log; grep -v "^\[W\|^W\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/r02_dist_check_${N}gpu.err | tail -8
(CUDA_VISIBLE_DEVICES=1 timeout 900 python scripts/grouped_rows_check.py > gpurun_out/r02_grouped_rows_check.log 2>&1; echo "grouped exit=$?" >> gpurun_out/r02_grouped_rows_check.log) &
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_${N}gpu.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/r02_pytest_gpu_${N}gpu.log
wait
tail -30 gpurun_out/r02_grouped_rows_check.log
