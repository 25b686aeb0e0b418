set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, env..., then bench args
  name=$1; shift
  env "$@" timeout 600 python bench.py --kernel mma --steps 5 --warmup 3 --no-cpu-baseline $BARGS > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("RESULT $name value=%.3e e2e=%.3e ms=%.2f scan_ms=%.2f cands=%d rows=%d frac=%.3f frac_exec=%.3f peak=%.0f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["scan_ms_per_step"],d["config"]["candidates_per_step"],d["config"]["hit_rows"],r["frac"],r.get("frac_executed",0),r["peak"]))
except Exception as e:
    print("RESULT $name FAILED", e); print(open("gpurun_out/bench_$name.err").read()[-800:])
PY
}
for M in "a 5" "b none" "a none" "b 5"; do set -- $M; BARGS="--mode $1 --max-divergence $2"; run n3_$1_$2 SMAFA_MMA_NSYM=3; done
for M in "a 5" "a none"; do set -- $M; BARGS="--mode $1 --max-divergence $2"; run n2_$1_$2 SMAFA_MMA_NSYM=2; run n3e16_$1_$2 SMAFA_MMA_EPI=16;  done
BARGS="--mode a --max-divergence 5"; run n4_a_5 SMAFA_MMA_NSYM=4; run n5_a_5 SMAFA_MMA_NSYM=5
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
ARGS="--kernel mma --mode a --max-divergence 5 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan_mma -s 1 -c 1 -o gpurun_out/prof_mma_v7 python bench.py $ARGS > gpurun_out/ncu_v7.log 2>&1
tail -2 gpurun_out/ncu_v7.log
