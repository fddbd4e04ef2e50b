set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_gpu_ot.py tests/test_gpu_tc_sweep.py tests/test_gpu_tc.py tests/test_gpu_libot.py tests/test_clustering_rule.py -x -q -m gpu 2>&1 | tail -8
python tools/ch_time.py > gpurun_out/r2_ch_time.txt 2>&1; cat gpurun_out/r2_ch_time.txt
python tools/ch_solve_once.py 1966 1916 20 3 > gpurun_out/ch_once.log 2>&1 && cat gpurun_out/ch_once.log && \
ncu --set full --clock-control none --import-source on -k regex:sinkhorn_solve_kernel -s 1 -c 1 -o gpurun_out/r2_solve_kernel python tools/ch_solve_once.py 1966 1916 20 2 > gpurun_out/ncu_solve.log 2>&1; tail -3 gpurun_out/ncu_solve.log
