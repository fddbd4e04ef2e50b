set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu_final.txt; cat gpurun_out/r2_pytest_gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -2 gpurun_out/r2_bench_final.err; cut -c1-600 gpurun_out/r2_bench_final.json
timeout 300 python tools/libot_bench.py 2>&1 | tail -6
