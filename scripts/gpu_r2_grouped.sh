# Round 2: grouped db order as the default -- forced-degree probe, then the whole parity suite, then the bench line.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
PROBE_T=24 timeout 600 python scripts/grouped_probe2.py 2>&1 | grep -v "need 4[0-9]\|need 3[0-9]" > gpurun_out/r02_grouped_probe3.log; tail -8 gpurun_out/r02_grouped_probe3.log
PROBE_M=5,10,15 timeout 600 python scripts/grouped_probe.py 2>&1 | grep -v "need 4[0-9]\|need 3[0-9]" > gpurun_out/r02_grouped_probe4.log; tail -30 gpurun_out/r02_grouped_probe4.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_grouped.log 2>&1; echo "pytest exit=$?"; tail -12 gpurun_out/r02_pytest_gpu_grouped.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_grouped.json 2> gpurun_out/r02_bench_1gpu_grouped.err; echo "bench1 exit=$?"; tail -c 1200 gpurun_out/r02_bench_1gpu_grouped.json; tail -3 gpurun_out/r02_bench_1gpu_grouped.err
