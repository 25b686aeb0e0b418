# Last GPU call of the round: the union-row tests and the default bench on the final tree (sample size / loose-bound rule changed
# after the full-suite run of scripts/gpu_final_round1.sh).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_union.py -m gpu -x -q > gpurun_out/pytest_union_final.log 2>&1; echo "pytest union exit=$?"; tail -3 gpurun_out/pytest_union_final.log
python bench.py > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; echo "bench exit=$?"; tail -c 2600 gpurun_out/bench_default_final.json
python bench.py --mode b --max-divergence none --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_unbounded_top10_final.json 2>/dev/null; python -c "
import json; j=json.loads(open('gpurun_out/bench_unbounded_top10_final.json').read().strip().splitlines()[-1]); print('unbounded top10 %.3e scan %.3f' % (j['value'], j['scan_ms_per_step']))"
