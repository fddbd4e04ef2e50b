"""Host-side logic of spadot_b200/sinkhorn.py on CPU: the drivers run against an oracle-backed `ops`
stand-in (tests/_numpy_ops.py) so stage schedule, stopping rules, growth loop, exact-median selection
and the row-partitioned multi-rank path (gloo, world_size 2) are covered without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from _numpy_ops import NumpyOps
from oracle import ot_dense
from spadot_b200 import sinkhorn

CFG = dict(ot_dense.DEFAULT_OT_CONFIG)


def test_single_rank_driver_matches_dense_oracle():
    a, b, la, lb = ot_dense.synthetic_embeddings(90, 75, 7, seed=31)
    Cn, med = ot_dense.median_normalised_cost(a, b)
    ops = NumpyOps(a, b)
    assert sinkhorn.median_cost(ops, small_limit=100, n_samples=2048) == med
    ops.set_median(med)
    info, info_ref = {}, {}
    st, eps = sinkhorn.solve_duality_gap(ops, np.ones(90), info=info, **CFG)
    want = ot_dense.duality_gap_solve(Cn, np.ones(90), info=info_ref, **CFG)
    assert info["iters_per_stage"] == info_ref["iters_per_stage"]
    np.testing.assert_allclose(ops.plan_dense(st.f, st.g, eps).numpy(), want, rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(ops.row_mass(st.f, st.Lr, eps).numpy(), want.sum(1), rtol=1e-10)
    st2, eps2 = sinkhorn.solve_stablev2(ops, np.ones(90), **dict(CFG, scaling_iter=300, extra_iter=60, tau=3.0))
    want2 = ot_dense.transport_stablev2(C=Cn, G=np.ones(90), **dict(CFG, scaling_iter=300, extra_iter=60, tau=3.0))
    np.testing.assert_allclose(ops.plan_dense(st2.f, st2.g, eps2).numpy(), want2, rtol=1e-9, atol=1e-300)


@pytest.mark.parametrize("n,m", [(7, 6), (8, 8), (33, 30), (1, 1), (2, 1)])
def test_median_selection_even_and_odd(n, m):
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, 3, seed=n * 10 + m)
    want = float(np.median(ot_dense.sqeuclidean(a, b)))
    assert sinkhorn.median_cost(NumpyOps(a, b)) == want
    assert sinkhorn.median_cost(NumpyOps(a, b), small_limit=0, n_samples=1024, n_bins=16) == want


def test_median_with_ties_and_zero_costs():
    a = np.zeros((20, 3))
    b = np.zeros((30, 3))
    assert sinkhorn.median_cost(NumpyOps(a, b), small_limit=0, n_samples=1024) == 0.0
    b[:10] = 1.0
    want = float(np.median(ot_dense.sqeuclidean(a, b)))
    assert sinkhorn.median_cost(NumpyOps(a, b), small_limit=0, n_samples=1024) == want


def test_nan_gap_raises_like_reference():
    a, b, _, _ = ot_dense.synthetic_embeddings(10, 12, 3, seed=1)
    ops = NumpyOps(a, b)
    ops.set_median(float(np.median(ot_dense.sqeuclidean(a, b))))
    G = np.ones(10)
    G[3] = 0.0      # r log(r/p) with p = 0 -> NaN in the reference's primal (ot_func.cpp:309-322)
    with pytest.raises(RuntimeError, match="Overflow encountered in duality gap computation"):
        sinkhorn.solve_duality_gap(ops, G, **CFG)


# ---------------------------------------------------------------------------- world_size 2, gloo
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _CommonShiftOps(NumpyOps):
    """NumpyOps + the one-collective column step of the row-partitioned solve (CudaOps.col_shift_ready /
    col_partial_sums / col_update_from_sums / seed_col_shift), restated in fp64 numpy: every rank sums
    exp(local LSE - shift) against the same shift (the previous combined LSE + 1), one all-reduce(SUM) of M + 2 doubles
    carries the sums, the tau flag of the row update and the verification flag."""

    def __init__(self, x, y):
        super().__init__(x, y)
        self.shift, self.shift_key, self.merged_steps = None, None, 0

    def col_shift_ready(self, eps):
        return self.shift_key == (eps, self.inv_med)

    def seed_col_shift(self, st, eps):
        self.shift = torch.where(torch.isfinite(st.Lc), st.Lc + 1.0, torch.zeros_like(st.Lc))
        self.shift_key = (eps, self.inv_med)

    def col_partial_sums(self, st, eps, it):
        vec = torch.zeros(self.m + 2, dtype=torch.float64)
        vec[:self.m] = torch.exp(self.col_lse(st.f, eps) - self.shift)
        vec[self.m] = 1.0 if int(self.flag[0]) == it else 0.0
        return vec

    def col_update_from_sums(self, vec, st, eps, alpha2, it, log_tau, log_floor=float("-inf")):
        if float(vec[self.m]) > 0:
            self.flag[0] = max(int(self.flag[0]), it)
        st.Lc.copy_(self.shift + torch.log(vec[:self.m]))
        self.shift = torch.where(torch.isfinite(st.Lc), st.Lc + 1.0, torch.zeros_like(st.Lc))
        self.potential_update("col", st.Lc, st.logq, eps, alpha2, float(np.log(st.N)), st.g, st.v, st.lb_old, it, log_tau, log_floor)
        self.merged_steps += 1


def _worker(rank, world, port, a, b, G, la, lb, out, ops_cls=NumpyOps):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = a.shape[0]
        r0, r1 = (n * rank) // world, (n * (rank + 1)) // world
        dist = sinkhorn.Dist()
        ops = ops_cls(a[r0:r1], b)
        med = sinkhorn.median_cost(ops, dist, small_limit=0, n_samples=2048, n_bins=64)
        ops.set_median(med)
        info = {}
        st, eps = sinkhorn.solve_duality_gap(ops, G[r0:r1], dist=dist, info=info, **CFG)
        plan_rows = ops.plan_dense(st.f, st.g, eps).numpy()
        tab = dist.sum_(ops.transition_table(st.f, st.g, eps, la[r0:r1], lb, 10, 10)).numpy()
        st2, eps2 = sinkhorn.solve_stablev2(ops, G[r0:r1], dist=dist, **dict(CFG, scaling_iter=120, extra_iter=30))
        out[rank] = dict(med=med, rows=(r0, r1), plan=plan_rows, iters=info["iters_per_stage"], tab=tab,
                         plan2=ops.plan_dense(st2.f, st2.g, eps2).numpy(), collectives=dist.collectives,
                         merged_steps=getattr(ops, "merged_steps", 0))
    finally:
        td.destroy_process_group()


def test_row_partitioned_solve_world2_gloo():
    a, b, la, lb = ot_dense.synthetic_embeddings(61, 47, 6, seed=77)   # 61 rows: ragged split 30 / 31
    G = np.exp(np.random.default_rng(2).normal(0, 0.5, 61))
    Cn, med = ot_dense.median_normalised_cost(a, b)
    info_ref = {}
    want = ot_dense.duality_gap_solve(Cn, G, info=info_ref, **CFG)
    want2 = ot_dense.transport_stablev2(C=Cn, G=G, **dict(CFG, scaling_iter=120, extra_iter=30))
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), a, b, G, la, lb, out), nprocs=2, join=True)
    assert set(out.keys()) == {0, 1}
    plan = np.zeros_like(want)
    plan2 = np.zeros_like(want)
    for r in (0, 1):
        o = out[r]
        assert o["med"] == med
        assert o["iters"] == info_ref["iters_per_stage"]
        plan[o["rows"][0]:o["rows"][1]] = o["plan"]
        plan2[o["rows"][0]:o["rows"][1]] = o["plan2"]
        assert o["collectives"] > 0
    np.testing.assert_allclose(plan, want, rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(plan2, want2, rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(out[0]["tab"], ot_dense.transition_table(want, la, lb, 10, 10), rtol=1e-9)
    np.testing.assert_allclose(out[0]["tab"], out[1]["tab"], rtol=1e-14)


def test_row_partitioned_solve_one_collective_per_iteration_world2_gloo():
    """The common-shift column step: same plan / iterations as the oracle, and the collective count drops from 3 per
    iteration (max + sum of the M-vector, tau flag) to 1 for every iteration but the first of a stage."""
    a, b, la, lb = ot_dense.synthetic_embeddings(53, 41, 5, seed=78)
    G = np.exp(np.random.default_rng(3).normal(0, 0.5, 53))
    Cn, med = ot_dense.median_normalised_cost(a, b)
    info_ref = {}
    cfg_tau = dict(CFG)
    want = ot_dense.duality_gap_solve(Cn, G, info=info_ref, **cfg_tau)
    want2 = ot_dense.transport_stablev2(C=Cn, G=G, **dict(CFG, scaling_iter=120, extra_iter=30))
    mgr = mp.Manager()
    out, out_old = mgr.dict(), mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), a, b, G, la, lb, out, _CommonShiftOps), nprocs=2, join=True)
    mp.spawn(_worker, args=(2, _free_port(), a, b, G, la, lb, out_old, NumpyOps), nprocs=2, join=True)
    plan, plan2 = np.zeros_like(want), np.zeros_like(want)
    total = sum(info_ref["iters_per_stage"])
    for r in (0, 1):
        o = out[r]
        assert o["iters"] == info_ref["iters_per_stage"]
        plan[o["rows"][0]:o["rows"][1]] = o["plan"]
        plan2[o["rows"][0]:o["rows"][1]] = o["plan2"]
        assert o["merged_steps"] >= total - 6 + 150 - 4           # all but the first iteration of each stage / schedule step
        assert o["collectives"] < out_old[r]["collectives"] - 2 * (total - 6)
    np.testing.assert_allclose(plan, want, rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(plan2, want2, rtol=1e-9, atol=1e-300)


class _FaultyPredictingOps(NumpyOps):
    """NumpyOps that claims to predict, corrupts the potentials of its `fail_at`-th verified batch and reports it through
    settle(): the drivers must restore their snapshot and redo exactly that work (the protocol of CudaOps._Pred)."""

    def __init__(self, x, y, fail_at):
        super().__init__(x, y)
        self.fail_at, self.batches, self.pending_fault, self.ok = fail_at, 0, False, True
        self.redone = 0

    def begin_solve(self, dist=None):
        self.batches, self.pending_fault, self.ok = 0, False, True

    def predicting(self):
        return self.ok

    def snapshot(self, st):
        self.batches += 1
        if self.batches == self.fail_at:
            self.pending_fault = True
        vec = (st.f, st.g, st.u, st.v, st.la_old, st.lb_old, st.Lr, st.Lc)
        return [t.clone() for t in vec], self.flag.clone(), self._tick

    def restore(self, st, snap):
        vec = (st.f, st.g, st.u, st.v, st.la_old, st.lb_old, st.Lr, st.Lc)
        for t, c in zip(vec, snap[0]):
            t.copy_(c)
        self.flag.copy_(snap[1])
        self._tick = snap[2]
        self.redone += 1

    def potential_update(self, side, L, *args, **kw):
        if self.pending_fault and self.ok:
            L = L + 3.0                                   # a mispredicted pass: garbage LSE for this batch
        return super().potential_update(side, L, *args, **kw)

    def row_lse(self, g, eps, out=None, predict=False):
        return super().row_lse(g, eps, out)

    def col_lse(self, f, eps, out=None, predict=False):
        return super().col_lse(f, eps, out)

    def settle(self, dist=None):
        if self.pending_fault and self.ok:
            self.pending_fault, self.ok = False, False   # detected: predictions off for the rest of the solve
            return False
        return True


@pytest.mark.parametrize("fail_at", [1, 4, 17])
def test_mispredicted_batch_is_redone_from_the_snapshot(fail_at):
    a, b, _, _ = ot_dense.synthetic_embeddings(60, 50, 5, seed=4)
    Cn, med = ot_dense.median_normalised_cost(a, b)
    ref_ops = NumpyOps(a, b)
    ref_ops.set_median(med)
    info_ref = {}
    st_ref, _ = sinkhorn.solve_duality_gap(ref_ops, np.ones(60), info=info_ref, **CFG)
    ops = _FaultyPredictingOps(a, b, fail_at)
    ops.set_median(med)
    info = {}
    st, _ = sinkhorn.solve_duality_gap(ops, np.ones(60), info=info, **CFG)
    assert ops.redone == 1 and ops.ok is False
    assert info["iters_per_stage"] == info_ref["iters_per_stage"]
    np.testing.assert_allclose(st.f.numpy(), st_ref.f.numpy(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(st.g.numpy(), st_ref.g.numpy(), rtol=0, atol=1e-12)


def test_profiling_flag_prints_one_trace_line_per_stage(capsys):
    """config.yaml's ot_config.profiling (ot_solvers.py:244-245): per-stage trace, results unchanged."""
    a, b, _, _ = ot_dense.synthetic_embeddings(40, 30, 4, seed=1)
    _, med = ot_dense.median_normalised_cost(a, b)
    ops = NumpyOps(a, b)
    ops.set_median(med)
    info = {}
    sinkhorn.solve_duality_gap(ops, np.ones(40), info=info, **dict(CFG, profiling=True))
    out = capsys.readouterr().out
    assert out.count("Compute for epsilon scaling") == 6
    assert out.count("iterations,") == 6 and "stage 5: %d iterations" % info["iters_per_stage"][5] in out
    sinkhorn.solve_duality_gap(ops, np.ones(40), **CFG)
    assert capsys.readouterr().out == ""


def test_max_iter_warns_and_returns_like_the_reference():
    a, b, _, _ = ot_dense.synthetic_embeddings(30, 25, 4, seed=2)
    _, med = ot_dense.median_normalised_cost(a, b)
    ops = NumpyOps(a, b)
    ops.set_median(med)
    info = {}
    with pytest.warns(RuntimeWarning, match="Reached max_iter with duality gap still above threshold"):
        st, _ = sinkhorn.solve_duality_gap(ops, np.ones(30), info=info, **dict(CFG, max_iter=5))
    assert info["max_iter_reached"] and info["iters_per_stage"] == [5] * 6       # the budget is per stage (ot_solvers.py:288)
    assert torch.isfinite(st.f).all() and torch.isfinite(st.g).all()
