set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gat.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r39_test_gat.txt; tail -4 gpurun_out/r39_test_gat.txt
timeout 200 python tools/gat_kernel_bench.py 100000 > gpurun_out/r39_gat_bench.json 2> gpurun_out/r39_gat_bench.err; cat gpurun_out/r39_gat_bench.json; tail -3 gpurun_out/r39_gat_bench.err
timeout 300 python tools/syn_t_bench.py --batches 10 > gpurun_out/r39_syn_t.json 2> gpurun_out/r39_syn_t.err; cat gpurun_out/r39_syn_t.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gat_tile -s 18 -c 3 -o gpurun_out/r39_gat_tiles python tools/gat_kernel_bench.py 100000 > gpurun_out/r39_ncu.log 2>&1; tail -3 gpurun_out/r39_ncu.log
