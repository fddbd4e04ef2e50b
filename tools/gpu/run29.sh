set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_gpu_ot.py -q -m gpu -x 2>&1 | tail -5
python tools/ch_time.py 2>&1 | tail -6 | tee gpurun_out/r2_ch_time.txt
SDB_STRIP_FORM=0 python tools/ch_time.py 2>&1 | tail -6
