# Round 2, 1-GPU call: parity suite with durations, bench default + the CLI-default (unbounded) modes, launch list of a step, cluster parity.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r02_pytest_gpu_v13.log 2>&1; echo "pytest exit=$?"; grep -v "^  File\|^$" gpurun_out/r02_pytest_gpu_v13.log | tail -32 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_v13.json 2> gpurun_out/r02_bench_1gpu_v13.err; echo "bench1 exit=$?"; tail -c 500 gpurun_out/r02_bench_1gpu_v13.json; tail -3 gpurun_out/r02_bench_1gpu_v13.err
for mode in a b; do
timeout 600 python bench.py --steps 10 --warmup 3 --mode $mode --max-divergence none --cpu-seconds 4 > gpurun_out/r02_bench_v13_mode_unbounded_$mode.json 2> gpurun_out/r02_bench_v13_mode_unbounded_$mode.err; echo "unbounded $mode exit=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_v13_mode_unbounded_$mode.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.3f scan %.3f rows %d cands %d guess %s rescanned %s degree %s match %s" % (d["value"], d["ms_per_step"], d["scan_ms_per_step"], d["config"]["hit_rows"], d["config"]["candidates_per_step"], d["config"]["guess_bound"], d["config"]["rescanned_queries"], d["config"].get("union_degree"), d["cpu_baseline"]["matches_gpu_rows"]))
PY
done
timeout 600 python bench.py --steps 10 --warmup 3 --mode b --cpu-seconds 4 > gpurun_out/r02_bench_v13_mode_top10_m5.json 2>/dev/null; echo "top10 m5 exit=$?"; tail -c 300 gpurun_out/r02_bench_v13_mode_top10_m5.json | head -c 300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_v13.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1; echo "ncu exit=$?"
CLUSTER_N=500000 timeout 900 python scripts/cluster_parity_full.py > gpurun_out/r02_cluster_500k_parity.log 2>&1; echo "cluster exit=$?"; tail -2 gpurun_out/r02_cluster_500k_parity.log
