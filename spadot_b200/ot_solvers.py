"""Drop-in Python API of the reference's OT module, backed by the CUDA library.

Mirrors SpaDOT/utils/OT_loss/ot_solvers.py (names, keyword arguments, error behaviour):

* `compute_transport_map(a, b, config, C=None, G=None)`      ot_solvers.py:95-121
* `optimal_transport_duality_gap(C, G, lambda1, ...)`         ot_solvers.py:164-449   (dense C, see dense.py)
* `transport_stablev2(C, lambda1, ...)`                       ot_solvers.py:452-531   (dense C, see dense.py)
* `default_config`                                            ot_solvers.py:19-36

plus the streamed entry point the reference has no equivalent of:

* `solve_coupling(a, b, config, ...) -> Coupling` keeps everything on the device as
  potentials (f, g); the dense plan, its marginals and the domain transition table are
  produced on request, so 250k x 250k or 1M x 1M couplings never materialise N x M.
"""
from __future__ import annotations


import numpy as np
import torch

from . import sinkhorn
from .cuda_ops import CudaOps

default_config = {                      # ot_solvers.py:19-36
    "growth_iters": 3, "epsilon": 0.05, "lambda1": 1, "lambda2": 50, "epsilon0": 1, "tau": 1000,
    "scaling_iter": 3000, "inner_iter_max": 50, "tolerance": 1e-8, "max_iter": 1e7, "batch_size": 5,
    "extra_iter": 1000, "numItermax": 1000000, "use_Py": False, "use_C": True, "profiling": False,
}

_SOLVER_KEYS = ("lambda1", "lambda2", "epsilon", "batch_size", "tolerance", "tau", "epsilon0", "max_iter",
                "scaling_iter", "extra_iter", "inner_iter_max", "profiling")


class Coupling:
    """Result of one streamed solve: device potentials and everything derived from them."""

    def __init__(self, ops, state, eps, median, info, dist):
        self.ops, self.state, self.eps, self.median, self.info, self.dist = ops, state, eps, median, info, dist

    @property
    def f(self):
        return self.state.f

    @property
    def g(self):
        return self.state.g

    def row_mass(self):
        """gamma.sum(axis=1) for this rank's rows (ot_solvers.py:116), from the potentials."""
        return self.ops.row_mass(self.state.f, self.state.Lr, self.eps)

    def col_mass(self):
        """gamma.sum(axis=0): exp(g_j/eps + LSE_i[(f_i - C_ij)/eps]) / M (global over ranks)."""
        Lc = sinkhorn.combine_col_lse(self.ops.col_lse(self.state.f, self.eps), self.dist)
        return torch.exp(self.state.g / self.eps + Lc) / self.ops.m

    def plan(self):
        """Dense R / J (ot_solvers.py:449) for this rank's rows — only for sizes that fit."""
        return self.ops.plan_dense(self.state.f, self.state.g, self.eps)

    def transition_table(self, labels_x, labels_y, k0=None, k1=None):
        """P0^T . T . P1 (wot transition_table as used at utils/_analyze_utils.py:135-137)."""
        labels_x = np.asarray(labels_x)
        labels_y = np.asarray(labels_y)
        k0 = int(labels_x.max()) + 1 if k0 is None else int(k0)
        k1 = int(labels_y.max()) + 1 if k1 is None else int(k1)
        t = self.ops.transition_table(self.state.f, self.state.g, self.eps, labels_x, labels_y, k0, k1)
        return self.dist.sum_(t)


def _solver_kwargs(config):
    return {k: config[k] for k in _SOLVER_KEYS if k in config}


def solve_coupling(a, b, config, G=None, solver="duality_gap", median=None, ops=None, dist=None, device=None):
    """One unbalanced entropic OT solve between spot sets a (N,d) [this rank's rows] and b (M,d)."""
    dist = dist or sinkhorn.Dist()
    if ops is None:
        ops = CudaOps(a, b, device=device)
    info = {}
    if median is None:
        median = sinkhorn.median_cost(ops, dist, info=info)
    ops.set_median(median)
    G_local = np.ones(ops.n) if G is None else G
    fn = sinkhorn.solve_duality_gap if solver == "duality_gap" else sinkhorn.solve_stablev2
    st, eps = fn(ops, G_local, dist=dist, info=info, **_solver_kwargs(config))
    return Coupling(ops, st, eps, median, info, dist)


class _PlanDownload:
    """Device -> pinned host copy of a dense plan on a side stream; `result()` waits for it and returns the ndarray
    (a view of the pinned buffer, owned by the returned array)."""

    PINNED_MAX_BYTES = 1 << 28       # beyond this, page-locking the buffer costs more than it hides: plain copy at the end

    def __init__(self, plan_dev):
        self.dev = plan_dev
        self.done = None
        if plan_dev.numel() * plan_dev.element_size() > self.PINNED_MAX_BYTES:
            return
        self.host = torch.empty(plan_dev.shape, dtype=plan_dev.dtype, pin_memory=True)
        cur = torch.cuda.current_stream(plan_dev.device)
        self.stream = torch.cuda.Stream(plan_dev.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.host.copy_(plan_dev, non_blocking=True)
        self.done = self.stream.record_event()

    def result(self):
        if self.done is None:
            out = self.dev.cpu().numpy()
            self.dev = None
            return out
        self.done.synchronize()
        self.dev = None
        return self.host.numpy()


def compute_transport_map(a, b, config, C=None, G=None):
    """ot_solvers.py:95-121: cost = sqeuclidean / median, `growth_iters` solves with
    G <- gamma.sum(axis=1), returns gammas[0] as an (N, M) float64 ndarray.

    `config` is mutated like the reference mutates it (`config["G"]`), except that the dense cost is
    never stored under `config["C"]`."""
    if C is not None:
        from . import dense
        return dense.compute_transport_map_dense(C, config, G)
    # torch tensors (the train loop passes k-means centroids, utils/_train_utils.py:318) stay where they are:
    # CUDA tensors are used in place, CPU tensors / ndarrays are uploaded once
    a = a.detach() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach() if isinstance(b, torch.Tensor) else np.asarray(b)
    ops = CudaOps(a, b)
    dist = sinkhorn.Dist(enabled=False)
    median = sinkhorn.median_cost(ops, dist)
    config["G"] = np.ones(a.shape[0]) if G is None else G
    first, row_sums = None, config["G"]
    for i in range(int(config["growth_iters"])):
        config["G"] = row_sums                                    # ot_solvers.py:113-117
        cp = solve_coupling(a, b, config, G=row_sums, median=median, ops=ops, dist=dist)
        if first is None:
            first = _PlanDownload(cp.plan())     # gammas[0]: its device->host copy overlaps the remaining growth solves
        row_sums = cp.row_mass().cpu().numpy()
    return first.result()


def optimal_transport_duality_gap(C, G, lambda1, lambda2, epsilon, batch_size, tolerance, tau, epsilon0, max_iter,
                                  use_Py=False, use_C=True, profiling=False, **ignored):
    """ot_solvers.py:164-449 for a caller-supplied dense cost matrix (wot's `solver` slot)."""
    from . import dense
    return dense.duality_gap_dense(C, G, lambda1, lambda2, epsilon, batch_size, tolerance, tau, epsilon0, max_iter)


def transport_stablev2(C, lambda1, lambda2, epsilon, scaling_iter, G, tau, epsilon0, extra_iter, inner_iter_max,
                       **ignored):
    """ot_solvers.py:452-531 for a caller-supplied dense cost matrix."""
    from . import dense
    return dense.stablev2_dense(C, lambda1, lambda2, epsilon, scaling_iter, G, tau, epsilon0, extra_iter, inner_iter_max)
