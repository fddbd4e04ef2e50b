set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python tools/syn_t_bench.py --batches 10 --host-profile > gpurun_out/r38_syn_t.json 2> gpurun_out/r38_syn_t.err; cat gpurun_out/r38_syn_t.json; grep -v Warning gpurun_out/r38_syn_t.err | head -70
