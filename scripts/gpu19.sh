set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_protein.py -m gpu -x -q > gpurun_out/pytest_protein.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_protein.log; tail -25 gpurun_out/pytest_protein.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_protein.py > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
for K in mma popc; do timeout 600 python bench.py --alphabet protein --kernel $K --steps 3 --warmup 2 --cpu-seconds 5 > gpurun_out/bench_protein_$K.json 2> gpurun_out/bench_protein_$K.err; tail -c 2500 gpurun_out/bench_protein_$K.json; tail -3 gpurun_out/bench_protein_$K.err; done
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json
