set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_gpu_ot.py tests/test_gpu_tc_sweep.py tests/test_gpu_libot.py tests/test_graph.py tests/test_model.py tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -15
python tools/ch_time.py > gpurun_out/r2_ch_time.txt 2>&1; cat gpurun_out/r2_ch_time.txt
python tools/size_scan.py --sizes 747x1966,1966x1916,4096x4096,8192x8192 > gpurun_out/r2_size_scan_small.jsonl 2>&1; cat gpurun_out/r2_size_scan_small.jsonl
for v in 0 1; do SDB_TC_RUNSUM=$v python tools/size_scan.py --sizes 131072x131072,262144x262144 --sweeps 10 > gpurun_out/r2_runsum_$v.jsonl 2>&1; cat gpurun_out/r2_runsum_$v.jsonl; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
