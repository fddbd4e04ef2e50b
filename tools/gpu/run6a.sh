set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_gpu_ot.py tests/test_gpu_tc.py tests/test_graph.py tests/test_clustering_rule.py tests/test_gat.py tests/test_kmeans.py -x -q -m gpu 2>&1 | tail -8
python tools/ch_time.py > gpurun_out/r2_ch_time.txt 2>&1; cat gpurun_out/r2_ch_time.txt
python tools/sweep_bench.py --rows 250000 --cols 250000 > gpurun_out/r2_sweep250k_1gpu.jsonl 2>&1; cat gpurun_out/r2_sweep250k_1gpu.jsonl
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0,'.')
from spadot_b200 import graph
pts=np.random.default_rng(0).uniform(0,30000,size=(100000,2))
for method in ("grid","brute"):
    graph.knn(pts,30,method=method); torch.cuda.synchronize(); t0=time.perf_counter(); graph.knn(pts,30,method=method); torch.cuda.synchronize()
    print("knn 100k k=30", method, time.perf_counter()-t0)
PY
