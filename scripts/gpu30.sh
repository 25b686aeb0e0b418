set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_query.py -m gpu -x -q > gpurun_out/pytest_query.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/pytest_query.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nt_m5_a.json 2>gpurun_out/err.log; 
for mode in a b; do
  python bench.py --mode $mode --max-divergence none --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nt_unbounded_$mode.json 2> gpurun_out/bench_nt_unbounded_$mode.err
done
python bench.py --alphabet protein --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_aa_unbounded_b.json 2>&1
python bench.py --mode b --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nt_m5_b.json 2>&1
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    for line in open(f):
        if line.startswith("{"):
            j=json.loads(line)
            print(f, "%.3g"%j["value"], "ms", round(j["ms_per_step"],2), "scan", round(j["scan_ms_per_step"],2), "rows",j["config"]["hit_rows"], "cand", j["config"]["candidates_per_step"], j["gpu_launches"])
P
