# Round 2, 2-GPU call d: whole parity suite (multi-GPU tests included) on the final code, bench N = 1 and 2, phase timings.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_${N}gpu_v14.log 2>&1; echo "pytest exit=$?"; grep -v "^  File\|^$" gpurun_out/r02_pytest_gpu_${N}gpu_v14.log | tail -25 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_v14.json 2> gpurun_out/r02_bench_1gpu_v14.err; echo "bench1 exit=$?"; tail -3 gpurun_out/r02_bench_1gpu_v14.err
timeout 600 $TR --master-port 29503 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu_v14.json 2> gpurun_out/r02_bench_${N}gpu_v14.err; echo "bench$N exit=$?"; grep -v "^\[W\|^W\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/r02_bench_${N}gpu_v14.err | tail -5
timeout 300 $TR --master-port 29504 scripts/dist_phases.py > gpurun_out/r02_dist_phases_${N}gpu_v14.log 2>&1; echo "phases exit=$?"; tail -1 gpurun_out/r02_dist_phases_${N}gpu_v14.log
python - <<PY
import json
for f in ["gpurun_out/r02_bench_1gpu_v14.json","gpurun_out/r02_bench_${N}gpu_v14.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms/step %.3f scan %.3f e2e %.4g rows %d launches %d degree %s upload %s match %s same %s frac %.3f" % (d["value"], d["ms_per_step"], d["scan_ms_per_step"], d["e2e"]["value"], d["config"]["hit_rows"], d["gpu_launches"], d["config"].get("union_degree"), d["config"].get("db_upload_s"), d["cpu_baseline"]["matches_gpu_rows"], d["config"]["rows_identical_on_all_ranks"], d["roofline"]["frac"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
