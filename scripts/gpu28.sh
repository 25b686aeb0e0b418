set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for mode in a b; do
  python bench.py --mode $mode --max-divergence none --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nt_unbounded_$mode.json 2> gpurun_out/bench_nt_unbounded_$mode.err; echo "exit=$?"; cat gpurun_out/bench_nt_unbounded_$mode.json
done
python bench.py --alphabet protein --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_aa_unbounded_b.json 2>&1; cat gpurun_out/bench_aa_unbounded_b.json
python bench.py --mode b --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nt_m5_b.json 2>&1; cat gpurun_out/bench_nt_m5_b.json
