set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ot.py -q -m gpu -x -k "whole_solve" 2>&1 | tail -6
timeout 120 python tools/ch_time.py 2>&1 | tail -6 | tee gpurun_out/r2_ch_time.txt
timeout 200 python tools/ch_time.py 2>&1 | tail -6
