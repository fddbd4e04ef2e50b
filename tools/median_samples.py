"""Exact streamed median (K5) against the size of its bracketing sample: total time, candidates, phase split."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from spadot_b200 import sinkhorn
from spadot_b200.cuda_ops import CudaOps
n = int(sys.argv[1])
x, y = bench.synth(n, n, 32)
ops = CudaOps(x, y)
sinkhorn.median_cost(ops)
for ns in (1 << 22, 1 << 24, 1 << 26):
    info = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    med = sinkhorn.median_cost(ops, n_samples=ns, info=info)
    torch.cuda.synchronize()
    print(n, ns, f"{time.perf_counter()-t0:.3f}s", repr(med), info["candidates"], {k: round(v, 4) for k, v in info["phases_s"].items()}, flush=True)
