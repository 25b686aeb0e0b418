# Round 2, 1-GPU call: whole parity suite on the speculative fast path + grouped order, bench, launch list, int8 peak cross-check.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_fast.log 2>&1; echo "pytest exit=$?"; grep -v "^  File\|^$" gpurun_out/r02_pytest_gpu_fast.log | tail -40 | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_v12.json 2> gpurun_out/r02_bench_1gpu_v12.err; echo "bench1 exit=$?"; tail -c 600 gpurun_out/r02_bench_1gpu_v12.json; tail -3 gpurun_out/r02_bench_1gpu_v12.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_v12.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1; echo "ncu exit=$?"
timeout 300 python scripts/int8_peak_crosscheck.py > gpurun_out/r02_int8_peak_crosscheck.txt 2>&1; echo "int8 exit=$?"; cat gpurun_out/r02_int8_peak_crosscheck.txt
