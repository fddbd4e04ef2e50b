set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR tests/dist_gpu_check.py > gpurun_out/r2_dist_check_${N}gpu.txt 2>&1; tail -3 gpurun_out/r2_dist_check_${N}gpu.txt
$TR tools/sweep_bench.py --rows 250000 --cols 250000 > gpurun_out/r2_sweep250k_${N}gpu.jsonl 2>&1; tail -2 gpurun_out/r2_sweep250k_${N}gpu.jsonl | cut -c1-600
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; tail -2 gpurun_out/r2_bench_${N}gpu.err; cut -c1-1500 gpurun_out/r2_bench_${N}gpu.json
$TR tools/sweep_bench.py --rows 1000000 --cols 1000000 > gpurun_out/r2_full_solve_1Mx1M_${N}gpu.jsonl 2>&1; tail -2 gpurun_out/r2_full_solve_1Mx1M_${N}gpu.jsonl | cut -c1-600
