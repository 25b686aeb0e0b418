set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CLUSTER_N=500000 CLUSTER_CHECK=0 SMAFA_TIMING=1 timeout 600 python scripts/cluster_bench.py > gpurun_out/cluster_500k.log 2>&1; cat gpurun_out/cluster_500k.log
CLUSTER_N=500000 CLUSTER_CHECK=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_cluster500k.csv python scripts/cluster_bench.py > gpurun_out/ncu_cluster.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_cluster500k.csv")) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
agg=collections.defaultdict(lambda:[0,0.0,0.0])
for r in rows[1:]:
    try: v=float(r[ix["Metric Value"]])
    except: continue
    u=r[ix["Metric Unit"]]
    ms = v/1e6 if u in ("ns","nsecond") else (v/1e3 if u in ("us","usecond") else (v if u in ("ms","msecond") else v*1e3))
    k=r[ix["Kernel Name"]][:70]
    a=agg[k]; a[0]+=1; a[1]+=ms; a[2]=max(a[2],ms)
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:12]:
    print("%-72s n=%4d total=%9.3f ms max=%8.3f ms"%(k,a[0],a[1],a[2]))
PY
