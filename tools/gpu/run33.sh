set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 120 python tools/ch_solve_once.py 1966 1916 20 3 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_solve_kernel -s 1 -c 1 -o gpurun_out/r2_solve_strips python tools/ch_solve_once.py 1966 1916 20 2 > gpurun_out/ncu_solve.log 2>&1; tail -3 gpurun_out/ncu_solve.log
ls -la gpurun_out/*.ncu-rep
