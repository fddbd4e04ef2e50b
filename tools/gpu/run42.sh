set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gat.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r42_test_gat.txt; tail -12 gpurun_out/r42_test_gat.txt
timeout 300 python tools/syn_t_bench.py --batches 10 > gpurun_out/r42_syn_t.json 2> gpurun_out/r42_syn_t.err; cat gpurun_out/r42_syn_t.json; tail -3 gpurun_out/r42_syn_t.err
