set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "accumulators" > gpurun_out/pytest_acc.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_acc.log; tail -15 gpurun_out/pytest_acc.log
for NS in 3 2 4; do for M in "a 5" "a none" "b none" "b 5"; do set -- $M; SMAFA_MMA_NSYM=$NS timeout 600 python bench.py --kernel mma --mode $1 --max-divergence $2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ns${NS}_$1_$2.json 2> gpurun_out/bench_ns${NS}_$1_$2.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_ns${NS}_$1_$2.json").read().strip().splitlines()[-1])
    print("RESULT nsym=${NS} mode=$1 m=$2 value=%.3e e2e=%.3e ms=%.2f scan_ms=%.2f cands=%d rows=%d"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["scan_ms_per_step"],d["config"]["candidates_per_step"],d["config"]["hit_rows"]))
except Exception as e:
    print("RESULT nsym=${NS} mode=$1 m=$2 FAILED", e); print(open("gpurun_out/bench_ns${NS}_$1_$2.err").read()[-800:])
PY
done; done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest.log; tail -6 gpurun_out/pytest.log
