set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR tests/dist_gpu_check.py > gpurun_out/r2b_dist_check_${N}gpu.txt 2>&1; tail -4 gpurun_out/r2b_dist_check_${N}gpu.txt
$TR tools/sweep_bench.py --rows 250000 --cols 250000 > gpurun_out/r2b_sweep250k_${N}gpu.jsonl 2>&1; tail -2 gpurun_out/r2b_sweep250k_${N}gpu.jsonl | cut -c1-600
SDB_NATIVE_DIST=0 $TR tools/sweep_bench.py --rows 250000 --cols 250000 2>&1 | tail -1 | cut -c1-400
