set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 400 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r44_smoke.txt 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r44_smoke.txt
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r44_pytest_gpu.txt; tail -4 gpurun_out/r44_pytest_gpu.txt
timeout 200 python tools/gat_kernel_bench.py 100000 > gpurun_out/r44_gat_bench.json 2> gpurun_out/r44_gat_bench.err; cut -c1-300 gpurun_out/r44_gat_bench.json
