# Round 2, very last 1-GPU call: whole parity suite with the capped guess, the two CLI-default modes.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_v17.log 2>&1; echo "pytest exit=$?"; grep -v "^  File\|^$" gpurun_out/r02_pytest_gpu_v17.log | tail -12 | cut -c1-300
for mode in b a; do
timeout 120 python bench.py --steps 10 --warmup 3 --mode $mode --max-divergence none --cpu-seconds 3 > gpurun_out/r02_bench_v17_mode_unbounded_$mode.json 2>/dev/null; echo "unbounded $mode exit=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_v17_mode_unbounded_$mode.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.3f scan %.3f rows %d cands %d guess %s rescanned %s degree %s match %s" % (d["value"], d["ms_per_step"], d["scan_ms_per_step"], d["config"]["hit_rows"], d["config"]["candidates_per_step"], d["config"]["guess_bound"], d["config"]["rescanned_queries"], d["config"].get("union_degree"), d["cpu_baseline"]["matches_gpu_rows"]))
PY
done
