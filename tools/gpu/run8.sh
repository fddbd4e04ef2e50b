set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for env in "A=1" "SDB_TC_FUSED_UPDATE=0" "SDB_PREDICTED_MAX=0" "SDB_TC_FUSED_UPDATE=0 SDB_PREDICTED_MAX=0"; do
  echo "== $env"; env $env python tools/pair_time.py 77369 102519 20 2>&1 | tail -3
done
echo "== d=32"; python tools/pair_time.py 77369 102519 32 2>&1 | tail -2
python tools/libot_bench.py 2>&1 | tail -6
