set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gat.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r36_test_gat.txt; tail -5 gpurun_out/r36_test_gat.txt
timeout 200 python tools/gat_kernel_bench.py 100000 > gpurun_out/r36_gat_bench.json 2> gpurun_out/r36_gat_bench.err; cat gpurun_out/r36_gat_bench.json; tail -3 gpurun_out/r36_gat_bench.err
timeout 300 python tools/syn_t_bench.py > gpurun_out/r36_syn_t.json 2> gpurun_out/r36_syn_t.err; cat gpurun_out/r36_syn_t.json; tail -3 gpurun_out/r36_syn_t.err
SDB_GAT_TILES=0 timeout 300 python tools/syn_t_bench.py > gpurun_out/r36_syn_t_pernode.json 2>> gpurun_out/r36_syn_t.err; cat gpurun_out/r36_syn_t_pernode.json
timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r36_pytest_gpu.txt; tail -4 gpurun_out/r36_pytest_gpu.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gat_tile -s 18 -c 3 -o gpurun_out/r36_gat_tiles python tools/gat_kernel_bench.py 100000 > gpurun_out/r36_ncu.log 2>&1; tail -3 gpurun_out/r36_ncu.log
ls -la gpurun_out/
