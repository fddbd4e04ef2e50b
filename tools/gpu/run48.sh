set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 200 python tools/syn_t_bench.py --batches 6 --epoch2 > gpurun_out/r48_syn_t.json 2> gpurun_out/r48_syn_t.err; python -c "
import json
d=json.load(open('gpurun_out/r48_syn_t.json'))
print(d['train_step_s'], d['step_times_s'], d['sample_batch_s'])
print(d['second_epoch'])
"; tail -2 gpurun_out/r48_syn_t.err
