set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for K in mma popc; do for M in "a 5" "a none" "b none" "a 15"; do set -- $M; timeout 600 python bench.py --kernel $K --mode $1 --max-divergence $2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_${K}_$1_$2.json 2> gpurun_out/bench_${K}_$1_$2.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${K}_$1_$2.json").read().strip().splitlines()[-1])
    print("RESULT ${K} mode=$1 m=$2 value=%.3e e2e=%.3e ms=%.2f cands=%d rows=%d"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["config"]["candidates_per_step"],d["config"]["hit_rows"]))
except Exception as e:
    print("RESULT ${K} mode=$1 m=$2 FAILED", e); print(open("gpurun_out/bench_${K}_$1_$2.err").read()[-800:])
PY
done; done
ARGS="--steps 2 --warmup 1 --no-cpu-baseline"
timeout 300 python bench.py $ARGS > gpurun_out/plain_auto.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_auto.csv python bench.py $ARGS > gpurun_out/ncu_auto.log 2>&1
