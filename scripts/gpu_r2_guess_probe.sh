# How the bound of the optimistic first pass moves unbounded top-10 (100 k x 1 M): forced guesses next to the sampled one.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
run() { python bench.py --steps 6 --warmup 3 --mode b --max-divergence none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1: value %.4g ms/step %.3f scan %.3f cands %d guess %s rescanned %s degree(last scan) %s rows %d' % (d['value'], d['ms_per_step'], d['scan_ms_per_step'], d['config']['candidates_per_step'], d['config']['guess_bound'], d['config']['rescanned_queries'], d['config'].get('union_degree'), d['config']['hit_rows']))"; }
( run "sampled guess"
  for g in 16 19 22 25; do SMAFA_FORCE_GUESS=$g run "forced guess $g"; done ) > gpurun_out/r02_guess_probe.log 2>&1
cat gpurun_out/r02_guess_probe.log
