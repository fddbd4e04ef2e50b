set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${NGPU:-2}
python -m pytest tests/test_gpu_ot.py -q -m gpu -x -k "radix or median" 2>&1 | tail -3
python tools/ch_time.py 2>&1 | tail -6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR tools/allreduce_time.py 2>&1 | grep doubles > gpurun_out/r2_allreduce_${N}gpu.jsonl; cat gpurun_out/r2_allreduce_${N}gpu.jsonl
