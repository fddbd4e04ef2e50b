set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for st in 0 600 1000 1400; do echo "STAGGER=$st"; SDB_TC_STAGGER_CLK=$st python tools/size_scan.py --sizes 131072x131072,262144x262144 --sweeps 10 2>&1 | tail -2; done
