# Round 2, last 2-GPU call: smoke(), the CLI end to end at configs[2] scale (1 M x 10 M) with --devices 0,1, protein 1 M x 10 M top-10.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/r02_smoke.log
E2E_DEVICES=0,1 E2E_D=10000000 E2E_Q=1000000 E2E_SUB=200 E2E_SUB_B=8 timeout 420 python scripts/cli_e2e.py > gpurun_out/r02_cli_e2e_config3_2gpu.log 2>&1; echo "cli e2e exit=$?"; grep -v "^\[smafa timing\] ctx" gpurun_out/r02_cli_e2e_config3_2gpu.log | tail -40 | cut -c1-260
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29507 bench.py --gpus 2 --steps 2 --warmup 3 --alphabet protein --db-per-gpu 5000000 --queries 1000000 --cpu-seconds 6 > gpurun_out/r02_protein_1Mx10M_2gpu.json 2> gpurun_out/r02_protein_2gpu.err; echo "protein exit=$?"; grep -v "^\[W\|^W\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/r02_protein_2gpu.err | tail -4
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_protein_1Mx10M_2gpu.json").read().strip().splitlines()[-1])
    print("protein: value %.4g ms/step %.3f scan %.3f e2e %.4g rows %d cands %d guess %s rescanned %s match %s (%d queries checked)" % (d["value"], d["ms_per_step"], d["scan_ms_per_step"], d["e2e"]["value"], d["config"]["hit_rows"], d["config"]["candidates_per_step"], d["config"]["guess_bound"], d["config"]["rescanned_queries"], d["cpu_baseline"]["matches_gpu_rows"], d["cpu_baseline"]["checked_queries"]))
except Exception as e:
    print("protein unreadable", e)
PY
