set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR tools/allreduce_time.py 2>&1 | grep doubles > gpurun_out/r2_allreduce_${N}gpu.jsonl; cat gpurun_out/r2_allreduce_${N}gpu.jsonl
$TR tools/sweep_bench.py --rows 250000 --cols 250000 --repeat 3 2>&1 | grep stage > gpurun_out/r2_sweep250k_x3_${N}gpu.jsonl; cut -c1-330 gpurun_out/r2_sweep250k_x3_${N}gpu.jsonl
$TR tools/sweep_bench.py --rows 250000 --cols 250000 --stages 2>&1 | grep -i "stage" > gpurun_out/r2_sweep250k_stages_${N}gpu.txt; cut -c1-330 gpurun_out/r2_sweep250k_stages_${N}gpu.txt
