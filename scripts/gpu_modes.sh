# One-GPU bench of every selection mode (run under gpurun): default (--max-divergence 5, best hit), top-10 with the bound,
# both without --max-divergence, and the protein top-10 shape.  One JSON line per mode under gpurun_out/.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
rm -f gpurun_out/bench_mode_*.json
python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline > gpurun_out/bench_mode_m5_besthit.json 2> gpurun_out/bench_modes.err
python bench.py --mode b --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mode_m5_top10.json 2>> gpurun_out/bench_modes.err
python bench.py --mode a --max-divergence none --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mode_unbounded_besthit.json 2>> gpurun_out/bench_modes.err
python bench.py --mode b --max-divergence none --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mode_unbounded_top10.json 2>> gpurun_out/bench_modes.err
python bench.py --alphabet protein --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mode_protein_top10.json 2>> gpurun_out/bench_modes.err
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_mode_*.json")):
    for line in open(f):
        if line.startswith("{"):
            j=json.loads(line)
            print(f, "%.3g"%j["value"], "ms", round(j["ms_per_step"],2), "scan", round(j["scan_ms_per_step"],2), "rows",j["config"]["hit_rows"], "cand", j["config"]["candidates_per_step"], j["gpu_launches"])
P
