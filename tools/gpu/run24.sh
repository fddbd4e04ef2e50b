set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for N in ${NLIST:-1 2 4}; do
  if [ "$N" = 1 ]; then python tools/mosta_bench.py > gpurun_out/r2_mosta_${N}gpu.jsonl 2> gpurun_out/r2_mosta_${N}gpu.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/mosta_bench.py > gpurun_out/r2_mosta_${N}gpu.jsonl 2> gpurun_out/r2_mosta_${N}gpu.err; fi
  tail -3 gpurun_out/r2_mosta_${N}gpu.err; tail -1 gpurun_out/r2_mosta_${N}gpu.jsonl | cut -c1-600
done
