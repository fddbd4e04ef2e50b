set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for p in 0 16 8 4; do echo "POLY=$p"; SDB_TC_POLY=$p python tools/size_scan.py --sizes 131072x131072,262144x262144 --sweeps 10 2>&1 | tail -2 | cut -c1-400; done
python tools/syn_t_bench.py --host-profile > gpurun_out/r2_syn_t_breakdown.json 2> gpurun_out/r2_syn_t_hostprofile.txt; tail -1 gpurun_out/r2_syn_t_breakdown.json | cut -c1-700; head -70 gpurun_out/r2_syn_t_hostprofile.txt | cut -c1-200
python -m pytest tests/test_gpu_ot.py tests/test_gat.py -q -m gpu -x 2>&1 | tail -3
