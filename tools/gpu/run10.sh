set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_gpu_ot.py tests/test_gpu_tc_sweep.py tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -8
python tools/ch_time.py > gpurun_out/r2_ch_time_b.txt 2>&1; cat gpurun_out/r2_ch_time_b.txt
SDB_RESIDENT_TILES=0 python tools/ch_time.py 2>&1 | grep wall
python tools/ch_solve_once.py 1966 1916 20 3; python tools/ch_solve_once.py 747 1966 20 3; python tools/ch_solve_once.py 2500 3000 32 3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
