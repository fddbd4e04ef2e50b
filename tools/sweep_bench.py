"""Full unbalanced-OT solves (exact median + 6-stage duality-gap solver to tolerance 1e-8) at scale:
BASELINE.json configs[4] (epsilon / lambda sweep at 250k x 250k) and configs[3] (1M x 1M).  Reports
iterations, wall time and the final duality gap per setting; one JSON line per solve.

    python tools/sweep_bench.py --n 250000 --m 250000 [--full-sweep]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from spadot_b200 import ot_solvers, sinkhorn  # noqa: E402
from spadot_b200.cuda_ops import CudaOps  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=250_000)
    ap.add_argument("--m", type=int, default=250_000)
    ap.add_argument("--d", type=int, default=32)
    ap.add_argument("--full-sweep", action="store_true")
    ap.add_argument("--max-iter", type=float, default=3000)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    x, y = bench.synth(a.n, a.m, a.d)
    t0 = time.perf_counter()
    ops = CudaOps(x, y)
    torch.cuda.synchronize()
    t_prep = time.perf_counter() - t0
    info = {}
    t0 = time.perf_counter()
    med = sinkhorn.median_cost(ops, info=info)
    torch.cuda.synchronize()
    t_med = time.perf_counter() - t0
    print(json.dumps(dict(stage="prep+median", n=a.n, m=a.m, d=a.d, prep_s=t_prep, median_s=t_med, median=med, **info)), flush=True)
    settings = [(0.05, 0.1, 5.0)]
    if a.full_sweep:
        settings += [(0.05, 1.0, 50.0), (0.01, 0.1, 5.0), (0.1, 1.0, 1.0)]
    for eps, l1, l2 in settings:
        cfg = dict(ot_solvers.default_config, epsilon=eps, lambda1=l1, lambda2=l2, tau=1000.0, max_iter=a.max_iter)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cp = ot_solvers.solve_coupling(x, y, cfg, median=med, ops=ops, dist=sinkhorn.Dist(enabled=False))
        mass = float(cp.row_mass().sum().item())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        it = cp.info["total_iters"]
        print(json.dumps(dict(stage="solve", n=a.n, m=a.m, epsilon=eps, lambda1=l1, lambda2=l2, seconds=dt,
                              iters_per_stage=cp.info["iters_per_stage"], total_iters=it, gap=cp.info["gap"],
                              iters_per_sec=it / dt, plan_mass=mass, tc=ops.use_tc)), flush=True)


if __name__ == "__main__":
    main()
