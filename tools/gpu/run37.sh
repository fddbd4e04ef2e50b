set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gat.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r37_test_gat.txt; tail -4 gpurun_out/r37_test_gat.txt
timeout 300 python tools/gat_variant_sweep.py 100000 > gpurun_out/r37_gat_variants.jsonl 2> gpurun_out/r37_gat_variants.err; cat gpurun_out/r37_gat_variants.jsonl; tail -3 gpurun_out/r37_gat_variants.err
timeout 300 python tools/syn_t_bench.py > gpurun_out/r37_syn_t.json 2> gpurun_out/r37_syn_t.err; cat gpurun_out/r37_syn_t.json; tail -3 gpurun_out/r37_syn_t.err
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r37_pytest_gpu.txt; tail -4 gpurun_out/r37_pytest_gpu.txt
