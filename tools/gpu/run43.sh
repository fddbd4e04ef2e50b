set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python tools/syn_t_bench.py --batches 4 --top-kernels 28 > gpurun_out/r43_syn_t.json 2> gpurun_out/r43_syn_t.err; grep " ms  x" gpurun_out/r43_syn_t.err
