set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_gat.py tests/test_model.py tests/test_svgp.py tests/test_graph.py -q -m gpu -x 2>&1 | tail -15
python tools/syn_t_bench.py > gpurun_out/r2_syn_t_breakdown.json 2> gpurun_out/r2_syn_t.err; tail -1 gpurun_out/r2_syn_t_breakdown.json | cut -c1-2200; tail -3 gpurun_out/r2_syn_t.err
