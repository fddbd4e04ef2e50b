set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gat.py tests/test_model.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r47_test_gat.txt; tail -6 gpurun_out/r47_test_gat.txt
timeout 200 python tools/syn_t_bench.py --batches 8 --top-kernels 12 > gpurun_out/r47_syn_t.json 2> gpurun_out/r47_syn_t.err; cut -c1-900 gpurun_out/r47_syn_t.json; grep " ms  x" gpurun_out/r47_syn_t.err | cut -c1-160
