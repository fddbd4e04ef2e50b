"""One compute_transport_map call at a ChickenHeart shape after a warm-up call (for an ncu launch list)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from spadot_b200 import ot_solvers
x, y = bench.synth(1966, 1916, 20)
cfg = dict(ot_solvers.default_config, lambda1=0.1, lambda2=5.0)
ot_solvers.compute_transport_map(x, y, dict(cfg))
torch.cuda.synchronize()
torch.cuda.profiler.start()
T = ot_solvers.compute_transport_map(x, y, dict(cfg))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(T.shape, float(T.sum()))
