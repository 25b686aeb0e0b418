set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_protein.py -m gpu -x -q > gpurun_out/pytest_protein.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_protein.log; tail -5 gpurun_out/pytest_protein.log
for K in mma popc; do timeout 600 python bench.py --alphabet protein --kernel $K --steps 3 --warmup 2 --cpu-seconds 5 > gpurun_out/bench_protein_$K.json 2> gpurun_out/bench_protein_$K.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_protein_$K.json").read().strip().splitlines()[-1])
print("RESULT protein $K value=%.3e e2e=%.3e ms=%.2f scan_ms=%.2f cands=%d rows=%d cpu=%.3e match=%s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["scan_ms_per_step"],d["config"]["candidates_per_step"],d["config"]["hit_rows"],d["cpu_baseline"]["value"],d["cpu_baseline"]["matches_gpu_rows"]))
PY
done
