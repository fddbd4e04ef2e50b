set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR tools/mosta_bench.py > gpurun_out/r2_mosta_${N}gpu.jsonl 2> gpurun_out/r2_mosta_${N}gpu.err; tail -2 gpurun_out/r2_mosta_${N}gpu.err; cut -c1-130 gpurun_out/r2_mosta_${N}gpu.jsonl
$TR tools/sweep_bench.py --rows 250000 --cols 250000 2>&1 | grep '"stage"' > gpurun_out/r2_sweep250k_${N}gpu.jsonl; cut -c1-330 gpurun_out/r2_sweep250k_${N}gpu.jsonl
$TR tests/dist_gpu_check.py 2>&1 | grep "dist check" > gpurun_out/r2_dist_check_${N}gpu.txt; cat gpurun_out/r2_dist_check_${N}gpu.txt
