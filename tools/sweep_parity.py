"""Diagnostic twin of tests/test_gpu_tc_sweep.py: prints the parity figures of every sweep case instead of asserting."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ot_dense  # noqa: E402  (checker only)
from spadot_b200 import ot_solvers, sinkhorn  # noqa: E402
from spadot_b200.cuda_ops import CudaOps  # noqa: E402

torch.cuda.set_device(0)
CFG = dict(ot_dense.DEFAULT_OT_CONFIG)
cases = [(eps, l1, l2, 3000, 2600) for eps in (0.01, 0.02, 0.1) for (l1, l2) in ((0.1, 5.0), (1.0, 1.0), (1.0, 50.0))] + \
        [(eps, 50.0, 50.0, 1000, 900) for eps in (0.01, 0.02, 0.1)]
if len(sys.argv) > 1:
    cases = [c for c in cases if str(c[0]) in sys.argv[1].split(",")]
tcs = (os.environ.get("SWEEP_TC", "on,off")).split(",")
for eps, lam1, lam2, n, m in cases:
    a, b, la, lb = ot_dense.synthetic_embeddings(n, m, 32, seed=17)
    cfg = dict(CFG, epsilon=eps, lambda1=lam1, lambda2=lam2)
    Cn, med = ot_dense.median_normalised_cost(a, b)
    info_ref = {}
    t0 = time.perf_counter()
    want = ot_dense.duality_gap_solve(Cn, np.ones(n), info=info_ref, **cfg)
    t_ref = time.perf_counter() - t0
    tab_ref = ot_dense.transition_table(want, la, lb, 10, 10)
    big = want > 1e-8 * want.max()
    for tc in tcs:
        ops = CudaOps(a, b, tc=tc)
        t0 = time.perf_counter()
        cp = ot_solvers.solve_coupling(a, b, cfg, ops=ops, dist=sinkhorn.Dist(enabled=False))
        got = cp.plan().cpu().numpy()
        t_gpu = time.perf_counter() - t0
        tab = cp.transition_table(la, lb, 10, 10).cpu().numpy()
        rel = (got - want)[big] / want[big]
        print(json.dumps(dict(eps=eps, lam=(lam1, lam2), n=n, m=m, tc=tc, iters=cp.info["iters_per_stage"], iters_ref=info_ref["iters_per_stage"],
                              rows=float(np.abs(got.sum(1) - want.sum(1)).max() / want.sum(1).max()),
                              cols=float(np.abs(got.sum(0) - want.sum(0)).max() / want.sum(0).max()),
                              entries_max=float(np.abs(rel).max()), entries_mean=float(rel.mean()),
                              table=float(np.abs(tab - tab_ref).max() / tab_ref.max()),
                              argmax=bool(np.array_equal(tab.argmax(1), tab_ref.argmax(1))), t_ref=t_ref, t_gpu=t_gpu)), flush=True)
