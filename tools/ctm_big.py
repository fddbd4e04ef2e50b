import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from spadot_b200 import ot_solvers
x, y = bench.synth(6100, 6000, 20)
cfg = dict(ot_solvers.default_config, lambda1=0.1, lambda2=5.0)
t0 = time.perf_counter()
T = ot_solvers.compute_transport_map(x, y, dict(cfg))
print("big plan", T.shape, T.dtype, float(T.sum()), time.perf_counter() - t0)
assert T.shape == (6100, 6000) and np.isfinite(T).all()
