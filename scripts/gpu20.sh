set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_protein.py -m gpu -x -q > gpurun_out/pytest_protein.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_protein.log; tail -25 gpurun_out/pytest_protein.log
for K in mma; do timeout 600 python bench.py --alphabet protein --kernel $K --steps 3 --warmup 2 --cpu-seconds 5 > gpurun_out/bench_protein_$K.json 2> gpurun_out/bench_protein_$K.err; tail -c 2500 gpurun_out/bench_protein_$K.json; tail -3 gpurun_out/bench_protein_$K.err; done
timeout 600 python bench.py --alphabet protein --kernel mma --max-divergence 4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_protein_mma_m4.json 2> gpurun_out/bench_protein_mma_m4.err; tail -c 1500 gpurun_out/bench_protein_mma_m4.json
SMAFA_MMA_NSYM=3 timeout 600 python bench.py --alphabet protein --kernel mma --max-divergence 4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_protein_mma_m4_cls.json 2> gpurun_out/bench_protein_mma_m4_cls.err; tail -c 1500 gpurun_out/bench_protein_mma_m4_cls.json
