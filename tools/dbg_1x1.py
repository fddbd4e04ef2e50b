import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ot_dense
from spadot_b200 import ot_solvers, sinkhorn
from spadot_b200.cuda_ops import CudaOps
n, m, d = 1, 1, 3
a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=n + m)
G = np.exp(np.random.default_rng(n).normal(0, 0.4, n))
cfg = dict(ot_dense.DEFAULT_OT_CONFIG)
for name, resident, dotmax, fused, max_iter in [("fused-resident", True, 0.0, True, 1e7), ("fused-streamed-direct", False, 0.0, True, 1e7),
                                      ("fused-streamed-dot", False, 100.0, True, 1e7), ("host-direct", False, 0.0, False, 1e7), ("host-dot", False, 100.0, False, 1e7),
                                      ("fused-resident@10", True, 0.0, True, 10), ("host-direct@10", False, 0.0, False, 10)]:
    ops = CudaOps(a, b, tc="off")
    ops.RESIDENT_TILES, ops.RESIDENT_MAX_TILES_PER_CTA, ops.SIMT_DOT_MAX = resident, 6, dotmax
    if not fused:
        ops.fused_solve = None
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cp = ot_solvers.solve_coupling(a, b, dict(cfg, max_iter=max_iter), G=G, ops=ops, dist=sinkhorn.Dist(enabled=False), median=1.7 * d)
    print(name, cp.info["iters_per_stage"], ["%.3e" % v for v in cp.info["stage_criteria"]], float(cp.f[0]), float(cp.g[0]))
