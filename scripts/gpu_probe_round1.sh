# One GPU call: sparse / split-N probes for the next kernel iteration (results under gpurun_out/).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi -L
# 1. instruction-shape issue rates and the sparse metadata decode, one process per stage
for sh in 0 1 2; do timeout 90 python scripts/sparse_probe.py rate $sh > gpurun_out/sparse_rate_$sh.log 2>&1; echo "rate $sh exit=$?"; tail -2 gpurun_out/sparse_rate_$sh.log; done
timeout 60 python scripts/sparse_probe.py decode 0 > gpurun_out/sparse_decode_st.log 2>&1; echo "decode st exit=$?"; head -3 gpurun_out/sparse_decode_st.log
timeout 60 python scripts/sparse_probe.py decode 1 > gpurun_out/sparse_decode_cp.log 2>&1; echo "decode cp exit=$?"; head -3 gpurun_out/sparse_decode_cp.log
# 2. split-N hand-over: parity subset, then the benches next to the default kernel on the same box
SMAFA_MMA_SPLIT_N=1 timeout 400 python -m pytest tests/test_gpu_query.py -m gpu -x -q \
  -k "test_query_matches_oracle or test_mma_survivor_rings or test_ties or test_candidate_overflow or test_guessed_bound_pass" \
  > gpurun_out/pytest_split_n.log 2>&1; echo "pytest split exit=$?"; tail -2 gpurun_out/pytest_split_n.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { name=$1; shift; env "$@" > gpurun_out/probe_bench_$name.json 2> gpurun_out/probe_bench_$name.err; echo "$name exit=$?"; python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/probe_bench_$name.json").read().strip().splitlines()[-1])
    print("  $name value=%.3e scan_ms=%.3f ms_per_step=%.3f cands=%s" % (j["value"], j.get("scan_ms_per_step", -1), j["ms_per_step"], j["config"].get("candidates_per_step")))
except Exception as e:
    print("  $name unreadable:", e)
PY
}
run l60_default          SMAFA_MMA_SPLIT_N=0 $B
run l60_split            SMAFA_MMA_SPLIT_N=1 $B
run l60_unbounded_top10_default SMAFA_MMA_SPLIT_N=0 $B --mode b --max-divergence none
run l60_unbounded_top10_split   SMAFA_MMA_SPLIT_N=1 $B --mode b --max-divergence none
run l30_default          SMAFA_MMA_SPLIT_N=0 $B --window-length 30
run l30_split            SMAFA_MMA_SPLIT_N=1 $B --window-length 30
run l60_enc2_default     SMAFA_MMA_SPLIT_N=0 SMAFA_MMA_NSYM=2 $B
run l60_enc2_split       SMAFA_MMA_SPLIT_N=1 SMAFA_MMA_NSYM=2 $B
