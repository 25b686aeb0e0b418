# One ncu --set full capture of the dominant kernel (wide-row tcgen05 scan, degree 16) in a bench step.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:scan_mma_kernelILi8ELi4ELi4ELi8ELb1ELi1ELb0ELi16 --launch-skip 4 -c 1 -f -o gpurun_out/r02_ncu_mma_u16 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu exit=$?"; tail -3 gpurun_out/r02_ncu_full.log; ls -la gpurun_out/r02_ncu_mma_u16.ncu-rep
