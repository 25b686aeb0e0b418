# Round-end validation on one B200 (run under gpurun): the whole -m gpu suite, smoke(), the default bench, the ncu
# launch list and one full capture of the scan kernel, the other selection modes, one ablation, the cluster bench.
# Stops after the test suite if it fails (the rest would measure a broken build).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_all.log 2>&1; rc=$?; echo "pytest exit=$rc"; tail -6 gpurun_out/pytest_gpu_all.log
if [ $rc -ne 0 ]; then exit $rc; fi
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit=$?"; tail -c 2500 gpurun_out/bench_default.json
ARGS="--steps 2 --warmup 1 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/bench_plain.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_mma -s 1 -c 1 -f -o gpurun_out/prof_scan_mma python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
echo "ncu exit=$?"
STEPS=5 bash scripts/gpu_modes.sh
SMAFA_MMA_UNION_STAGES4=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_union_stages4.json 2> gpurun_out/bench_union_stages4.err; echo "stages4 exit=$?"
python - <<'P'
import json
for f in ("gpurun_out/bench_default.json", "gpurun_out/bench_union_stages4.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.4g" % j["value"], "e2e %.4g" % j["e2e"]["value"], "scan_ms", round(j["scan_ms_per_step"], 3), j["roofline"]["operands"])
    except Exception as e:
        print(f, "unreadable", e)
P
SMAFA_MMA_UNION_STAGES4=1 timeout 200 python -m pytest tests/test_gpu_union.py -m gpu -x -q -k "query_matches or pressure or ties" > gpurun_out/pytest_union_stages4.log 2>&1; echo "pytest stages4 exit=$?"; tail -2 gpurun_out/pytest_union_stages4.log
CLUSTER_N=5000000 CLUSTER_CHECK=60000 SMAFA_TIMING=1 timeout 240 python scripts/cluster_bench.py > gpurun_out/cluster_bench.log 2>&1; echo "cluster_bench exit=$?"; grep -v "cluster: [28] batches" gpurun_out/cluster_bench.log | tail -8
