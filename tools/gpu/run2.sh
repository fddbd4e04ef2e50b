set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -5 gpurun_out/r2_bench_a.err; cat gpurun_out/r2_bench_a.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref_a.json 2> gpurun_out/r2_bench_ref_a.err; tail -5 gpurun_out/r2_bench_ref_a.err; cat gpurun_out/r2_bench_ref_a.json
