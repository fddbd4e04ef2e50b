"""Time of repeated full solves of one pair (growth-loop style) with the knobs of the tensor-core loop, to attribute changes."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from spadot_b200 import ot_solvers, sinkhorn
from spadot_b200.cuda_ops import CudaOps
n, m, d = (int(v) for v in sys.argv[1:4])
x, y = bench.synth(n, m, d)
cfg = dict(ot_solvers.default_config, epsilon=0.05, lambda1=0.1, lambda2=5.0, tau=10000.0)
ops = CudaOps(x, y)
med = sinkhorn.median_cost(ops)
for r in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); l0 = ops.launches
    cp = ot_solvers.solve_coupling(x, y, cfg, median=med, ops=ops, dist=sinkhorn.Dist(enabled=False))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps(dict(n=n, m=m, d=d, run=r, seconds=dt, iters=cp.info["iters_per_stage"], launches=ops.launches - l0,
                          pred_ok=bool(ops._pred is not None and ops._pred.ok), fused=ops.FUSED_UPDATE, predict=ops.PREDICT,
                          ideal_s=2.0 * n * m * cp.info["total_iters"] / (148 * 16 * 1965e6))), flush=True)
