"""Wall time of one duality-gap solve (95 iterations) across sizes: where host overhead stops mattering."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from spadot_b200 import ot_solvers, sinkhorn
from spadot_b200.cuda_ops import CudaOps
cfg = dict(ot_solvers.default_config, lambda1=0.1, lambda2=5.0)
for n in (747, 1966, 4096, 8192, 16384, 32768):
    x, y = bench.synth(n, n, 20 if n < 4000 else 32)
    ops = CudaOps(x, y); ops.set_median(100.0 if n < 4000 else 160.0)
    single = sinkhorn.Dist(enabled=False)
    sinkhorn.solve_duality_gap(ops, np.ones(n), dist=single, **cfg)
    torch.cuda.synchronize(); t0 = time.perf_counter(); info = {}
    sinkhorn.solve_duality_gap(ops, np.ones(n), dist=single, info=info, **cfg)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"n={n} tc={ops.use_tc} iters={info['total_iters']} solve={dt*1e3:.2f} ms  per-iter={dt/info['total_iters']*1e6:.1f} us", flush=True)
