set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r45_bench.json 2> gpurun_out/r45_bench.err; echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r45_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
print(json.dumps(d.get('aux_syn_t_step'))[:1500])
print(json.dumps(d.get('aux_train_epoch'))[:600])
"; tail -3 gpurun_out/r45_bench.err
