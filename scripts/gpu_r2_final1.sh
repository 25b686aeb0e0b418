# Round 2, last 1-GPU call: whole parity suite on the final code, bench line, launch list of a step.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_v15.log 2>&1; echo "pytest exit=$?"; grep -v "^  File\|^$" gpurun_out/r02_pytest_gpu_v15.log | tail -25 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_v15.json 2> gpurun_out/r02_bench_1gpu_v15.err; echo "bench1 exit=$?"; tail -3 gpurun_out/r02_bench_1gpu_v15.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_1gpu_v15.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.3f scan %.3f e2e %.4g rows %d launches %d degree %s match %s frac %.3f" % (d["value"], d["ms_per_step"], d["scan_ms_per_step"], d["e2e"]["value"], d["config"]["hit_rows"], d["gpu_launches"], d["config"].get("union_degree"), d["cpu_baseline"]["matches_gpu_rows"], d["roofline"]["frac"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_v15.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1; echo "ncu exit=$?"
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r02_launches_v15.csv')))
hdr=None; seq=[]
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        dd=dict(zip(hdr,r))
        try: t=float(dd['Metric Value'].replace(',',''))
        except: continue
        u=dd['Metric Unit']
        if u=='ns': t/=1e3
        elif u=='ms': t*=1e3
        seq.append((dd['Kernel Name'][:70],t))
idx=[i for i,(n,t) in enumerate(seq) if 'scan_mma_kernel<8, 4, 4, 8, 1, 1, 0, 16>' in n]
if idx:
    i=idx[-2]
    for n,t in seq[i-4:i+12]: print("%9.2f us  %s"%(t,n))
PY
