set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_query.py -m gpu -x -q -k "not config2" > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
for K in mma popc; do for M in "a none" "b none" "a 5"; do set -- $M; timeout 600 python bench.py --kernel $K --mode $1 --max-divergence $2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_${K}_$1_$2.json 2> gpurun_out/bench_${K}_$1_$2.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${K}_$1_$2.json").read().strip().splitlines()[-1])
    print("RESULT ${K} mode=$1 m=$2 value=%.3e e2e=%.3e ms=%.2f scan_ms=%.2f cands=%d rows=%d"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["scan_ms_per_step"],d["config"]["candidates_per_step"],d["config"]["hit_rows"]))
except Exception as e:
    print("RESULT ${K} mode=$1 m=$2 FAILED", e); print(open("gpurun_out/bench_${K}_$1_$2.err").read()[-800:])
PY
done; done
