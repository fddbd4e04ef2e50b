"""Dense-cost entry points: the reference signatures that receive an explicit cost matrix.

`optimal_transport_duality_gap(C, G, ...)` and `transport_stablev2(C, ...)`
(SpaDOT/utils/OT_loss/ot_solvers.py:164,452) are what wot's `OTModel.solver` slot calls, and
`compute_transport_map(a, b, config, C=C)` (ot_solvers.py:95) accepts a precomputed C.  The same host
drivers as the streamed path run against `DenseOps`: fp64 log-sum-exp sweeps over C on the device, one
read of C per half-iteration, no K/_K/R matrices.  Returns R / J as an (N, M) float64 ndarray.
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np

from . import sinkhorn
from .cuda_ops import DenseOps


def _to_numpy(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def duality_gap_dense(C, G, lambda1, lambda2, epsilon, batch_size, tolerance, tau, epsilon0, max_iter, info=None):
    ops = DenseOps(_to_numpy(C))
    st, eps = sinkhorn.solve_duality_gap(ops, _to_numpy(G), lambda1, lambda2, epsilon, batch_size, tolerance, tau,
                                         epsilon0, max_iter, dist=sinkhorn.Dist(enabled=False), info=info)
    return ops.plan_dense(st.f, st.g, eps).cpu().numpy()


def stablev2_dense(C, lambda1, lambda2, epsilon, scaling_iter, G, tau, epsilon0, extra_iter, inner_iter_max, info=None):
    ops = DenseOps(_to_numpy(C))
    st, eps = sinkhorn.solve_stablev2(ops, _to_numpy(G), lambda1, lambda2, epsilon, scaling_iter, tau, epsilon0,
                                      extra_iter, inner_iter_max, dist=sinkhorn.Dist(enabled=False), info=info)
    return ops.plan_dense(st.f, st.g, eps).cpu().numpy()


def compute_transport_map_dense(C, config, G=None):
    """Growth loop of ot_solvers.py:105-121 for a caller-supplied C (config["C"], config["G"] mutated alike)."""
    C = _to_numpy(C)
    config["C"] = C
    config["G"] = np.ones(C.shape[0]) if G is None else G
    ops = DenseOps(C)
    single = sinkhorn.Dist(enabled=False)
    first, row_sums = None, config["G"]
    keys = ("lambda1", "lambda2", "epsilon", "batch_size", "tolerance", "tau", "epsilon0", "max_iter")
    for _ in range(int(config["growth_iters"])):
        config["G"] = row_sums
        st, eps = sinkhorn.solve_duality_gap(ops, _to_numpy(row_sums), dist=single, **{k: config[k] for k in keys if k in config})
        if first is None:
            first = ops.plan_dense(st.f, st.g, eps).cpu().numpy()
        row_sums = ops.row_mass(st.f, st.Lr, eps).cpu().numpy()
    return deepcopy(first)
