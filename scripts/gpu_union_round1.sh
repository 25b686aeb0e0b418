# One GPU call: parity of the union-row operands (tests/test_gpu_union.py), then benches with the degree picked per scan.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi -L
timeout 500 python -m pytest tests/test_gpu_union.py -m gpu -x -q > gpurun_out/pytest_union.log 2>&1; echo "pytest union exit=$?"; tail -15 gpurun_out/pytest_union.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { name=$1; shift; env "$@" > gpurun_out/union_bench_$name.json 2> gpurun_out/union_bench_$name.err; echo "$name exit=$?"; python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/union_bench_$name.json").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("  $name value=%.3e e2e=%.3e scan_ms=%.3f ms_per_step=%.3f cands=%s rows=%s k=%s frac_exec=%.3f" % (j["value"], j["e2e"]["value"], j.get("scan_ms_per_step", -1), j["ms_per_step"], j["config"].get("candidates_per_step"), j["config"].get("hit_rows"), r.get("executed_ops_per_comparison"), r.get("frac_executed", -1)))
except Exception as e:
    print("  $name unreadable:", e)
PY
}
run m5_besthit_auto        SMAFA_MMA_UNION=3 $B
run m5_besthit_force2      SMAFA_MMA_UNION_FORCE=2 $B
run m5_top10_auto          SMAFA_MMA_UNION=3 $B --mode b
run m10_top10_auto         SMAFA_MMA_UNION=3 $B --mode b --max-divergence 10
run m15_top10_auto         SMAFA_MMA_UNION=3 $B --mode b --max-divergence 15
run unbounded_besthit_auto SMAFA_MMA_UNION=3 $B --max-divergence none
run unbounded_top10_auto   SMAFA_MMA_UNION=3 $B --mode b --max-divergence none
run l30_m5_auto            SMAFA_MMA_UNION=3 $B --window-length 30
