set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${NGPU:-4}
python -m pytest tests/test_gpu_ot.py -q -m gpu -k "whole_solve_in_one_launch_matches and 1-1-3" 2>&1 | grep -E "^E|assert|passed|failed" | head -30
python -m pytest tests/test_gat.py -q -m gpu 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR tools/allreduce_time.py 2>&1 | grep doubles > gpurun_out/r2_allreduce_${N}gpu.jsonl; cat gpurun_out/r2_allreduce_${N}gpu.jsonl
$TR tools/sweep_bench.py --rows 250000 --cols 250000 --stages > gpurun_out/r2_sweep250k_stages_${N}gpu.txt 2>&1; tail -16 gpurun_out/r2_sweep250k_stages_${N}gpu.txt | cut -c1-500
SDB_NATIVE_DIST=0 $TR tools/sweep_bench.py --rows 250000 --cols 250000 --stages 2>&1 | tail -14 | cut -c1-400
