set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for V in "8 0" "8 1" "16 0" "16 1"; do set -- $V; for M in "a 5" "b none" "a none"; do set -- $V $M; SMAFA_MMA_EPI=$1 SMAFA_MMA_PACK16=$2 timeout 600 python bench.py --kernel mma --mode $3 --max-divergence $4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e$1_p$2_$3_$4.json 2> gpurun_out/bench_e$1_p$2_$3_$4.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_e$1_p$2_$3_$4.json").read().strip().splitlines()[-1])
    print("RESULT epi=$1 pack16=$2 mode=$3 m=$4 value=%.3e e2e=%.3e ms=%.2f scan_ms=%.2f cands=%d rows=%d"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["scan_ms_per_step"],d["config"]["candidates_per_step"],d["config"]["hit_rows"]))
except Exception as e:
    print("RESULT epi=$1 pack16=$2 FAILED", e); print(open("gpurun_out/bench_e$1_p$2_$3_$4.err").read()[-800:])
PY
done; done
SMAFA_MMA_EPI=16 SMAFA_MMA_PACK16=1 timeout 900 python -m pytest tests -m gpu -x -q -k "mma or kats or config2" > gpurun_out/pytest_e16p1.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_e16p1.log; tail -4 gpurun_out/pytest_e16p1.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
