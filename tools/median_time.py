"""Time the pieces of the exact streamed median (K5) — development aid."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from spadot_b200 import sinkhorn
from spadot_b200.cuda_ops import CudaOps

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
x, y = bench.synth(n, n, 32)
ops = CudaOps(x, y)
def t(fn, name):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter()-t0)*1e3:.1f} ms", flush=True); return r
med = 164.7
for tc in (True, False):
    ops.use_tc = tc
    t(lambda: ops.cost_histogram(med * 0.99, med * 1.01, 4096), f"hist tc={tc} bracket 1%")
    t(lambda: ops.cost_histogram(med * 0.999, med * 1.001, 4096), f"hist tc={tc} bracket 0.1%")
    t(lambda: ops.cost_collect(med * 0.99999, med * 1.00001, 1 << 26), f"collect tc={tc} 2e-5")
    t(lambda: ops.cost_collect(med * 2, med * 2.0000001, 1 << 20), f"collect tc={tc} empty")
ops.use_tc = True
g = ops.zeros(n)
t(lambda: ops.row_lse(g, 0.05), "lse pass")
info = {}
t(lambda: sinkhorn.median_cost(ops, info=info), "median_cost total")
print(info)
