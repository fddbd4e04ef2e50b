set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_gpu_ot.py -x -q -m gpu 2>&1 | tail -4
for r in 1 0; do echo "RESIDENT=$r"; SDB_RESIDENT_TILES=$r python tools/ch_time.py 2>&1 | grep -E "wall|solve"; done
python tools/size_scan.py --sizes 747x1966,1966x1916,2500x3000 --d 20
