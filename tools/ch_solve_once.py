"""One ChickenHeart-sized duality-gap solve (the one-launch kernel), for ncu captures and timing."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ot_dense
from spadot_b200 import sinkhorn
from spadot_b200.cuda_ops import CudaOps
n, m, d = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (1966, 1916, 20)))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
cfg = dict(ot_dense.DEFAULT_OT_CONFIG)
a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=n)
ops = CudaOps(a, b)
med = sinkhorn.median_cost(ops)
ops.set_median(med)
for r in range(reps):
    info = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st, eps = sinkhorn.solve_duality_gap(ops, np.ones(n), info=info, **cfg)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{n}x{m} d={d}: solve {1e3 * (t1 - t0):.3f} ms, iters {info['iters_per_stage']}, direct={ops._direct}", flush=True)
