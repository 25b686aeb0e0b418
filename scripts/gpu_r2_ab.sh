# A/B on one box: item order of the wide-row scan (SMAFA_MMA_QT_MAJOR) and the speculative batch (SMAFA_NO_FAST_FINALIZE).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1: value %.4g ms/step %.3f scan %.3f e2e %.4g launches %d degree %s' % (d['value'], d['ms_per_step'], d['scan_ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['config'].get('union_degree')))"; }
( SMAFA_MMA_QT_MAJOR=0 run "chunk-major fast"
  SMAFA_MMA_QT_MAJOR=1 run "qt-major    fast"
  SMAFA_MMA_QT_MAJOR=0 SMAFA_NO_FAST_FINALIZE=1 run "chunk-major sort"
  SMAFA_MMA_QT_MAJOR=0 run "chunk-major fast (again)"
  SMAFA_MMA_QT_MAJOR=0 SMAFA_MMA_UNION_STAGES4=0 run "chunk-major fast, 2+2 stages" ) > gpurun_out/r02_ab_item_order.log 2>&1
cat gpurun_out/r02_ab_item_order.log
