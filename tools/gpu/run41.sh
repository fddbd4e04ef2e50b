set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gat.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r41_test_gat.txt; tail -5 gpurun_out/r41_test_gat.txt
timeout 300 python tools/syn_t_bench.py --batches 10 > gpurun_out/r41_syn_t.json 2> gpurun_out/r41_syn_t.err; cat gpurun_out/r41_syn_t.json; tail -3 gpurun_out/r41_syn_t.err
