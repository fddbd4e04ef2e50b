set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu_final.txt; cat gpurun_out/r2_pytest_gpu_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -2 gpurun_out/r2_bench_final.err; cut -c1-1200 gpurun_out/r2_bench_final.json
python tools/syn_t_bench.py > gpurun_out/r2_syn_t_breakdown.json 2>&1; tail -3 gpurun_out/r2_syn_t_breakdown.json | cut -c1-2500
# the per-rank shape of 250k x 250k on 8 GPUs, on one GPU
python tools/sweep_bench.py --rows 31250 --cols 250000 --stages > gpurun_out/r2_sweep_31250x250k_1gpu.txt 2>&1; tail -16 gpurun_out/r2_sweep_31250x250k_1gpu.txt | cut -c1-500
