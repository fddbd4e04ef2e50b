cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gat.py tests/test_graph.py -m gpu -q -k "prebuilt or two_hop or gat_on_batch" 2>&1 | tail -6
