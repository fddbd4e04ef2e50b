"""fp64 log-domain restatement of the reference's two OT solvers.  TEST INFRASTRUCTURE ONLY.

Same algorithm as ot_dense.py (which is the literal restatement of
SpaDOT/utils/OT_loss/ot_solvers.py:164-531 + ot_func.cpp:831-930) but carried in
total potentials

    f = u + eps*log a,   g = v + eps*log b            
    f_i <- eps*alpha1*( log p_i - LSE_j[(g_j - C_ij)/eps + log dy_j] )
    g_j <- eps*alpha2*( log q_j - LSE_i[(f_i - C_ij)/eps + log dx_i] )

so the N x M matrices K, _K, R are never formed: every pass is a blocked
log-sum-exp over cost tiles recomputed from the embeddings (or sliced from a dense C).
It is the large-N oracle for the CUDA path and documents, in numpy, exactly what the
device driver in spadot_b200/sinkhorn.py does.  Absorption frames (u, v) are tracked
only because the reference's stage 0-4 stopping rule mixes them in
(ot_func.cpp:897-922); they do not affect the iterates.
"""
from __future__ import annotations

import numpy as np

from . import ot_dense


class CostOperator:
    """Blocked access to C = sqeuclidean(x, y) / median (ot_solvers.py:102-103) or a dense C."""

    def __init__(self, x=None, y=None, median=None, C=None, block=512):
        if C is not None:
            self.C = np.asarray(C, dtype=np.float64)
            self.n, self.m = self.C.shape
            self.x = self.y = None
        else:
            self.C = None
            self.x = np.asarray(x, dtype=np.float64)
            self.y = np.asarray(y, dtype=np.float64)
            self.n, self.m = self.x.shape[0], self.y.shape[0]
            self.inv_med = 1.0 / float(median)
        self.block = block

    def rows(self, s, e):
        if self.C is not None:
            return self.C[s:e]
        return ot_dense.sqeuclidean(self.x[s:e], self.y) * self.inv_med

    def row_lse(self, col_term, eps):
        """L_i = log sum_j exp(col_term_j - C_ij/eps)."""
        out = np.empty(self.n)
        for s in range(0, self.n, self.block):
            e = min(self.n, s + self.block)
            t = col_term[None, :] - self.rows(s, e) / eps
            mx = t.max(axis=1)
            mx = np.where(np.isfinite(mx), mx, 0.0)
            with np.errstate(divide="ignore"):
                out[s:e] = mx + np.log(np.exp(t - mx[:, None]).sum(axis=1))
        return out

    def col_lse(self, row_term, eps):
        """L_j = log sum_i exp(row_term_i - C_ij/eps)  (streamed over row blocks)."""
        mx = np.full(self.m, -np.inf)
        acc = np.zeros(self.m)
        for s in range(0, self.n, self.block):
            e = min(self.n, s + self.block)
            t = row_term[s:e, None] - self.rows(s, e) / eps
            bm = t.max(axis=0)
            new = np.maximum(mx, bm)
            safe = np.where(np.isfinite(new), new, 0.0)
            with np.errstate(invalid="ignore"):
                acc = acc * np.exp(np.where(np.isfinite(mx), mx - safe, -np.inf)) + np.exp(t - safe[None, :]).sum(axis=0)
            mx = new
        with np.errstate(divide="ignore"):
            return np.where(np.isfinite(mx), mx, 0.0) + np.log(acc)


def gap_from_marginals(f, g, Lr, Lc_at_f, sumK, p, q, eps, lam1, lam2):
    """Relative duality gap of ot_func.cpp:493-544 from O(N+M) quantities.

    Lr   = LSE_j[(g_j - C_ij)/eps]   at the current g (no weights)
    Lc   = LSE_i[(f_i - C_ij)/eps]   at the current f (no weights)
    sumK = sum_ij exp(-C_ij/eps)
    Uses eps*R log R + R*C = R (f_i + g_j)  (SURVEY.md §3.3).
    """
    I, J = len(f), len(g)
    dx, dy = 1.0 / I, 1.0 / J
    Rrow = np.exp(f / eps + Lr)            # sum_j R_ij
    Rcol = np.exp(g / eps + Lc_at_f)       # sum_i R_ij
    sumR = Rrow.sum()
    r = Rrow * dy
    c = Rcol * dx
    F1 = lam1 * np.sum(dx * (r * np.log(r / p) - r + p))
    F2 = lam2 * np.sum(dy * (c * np.log(c / q) - c + q))
    pri = F1 + F2 + (np.dot(f, Rrow) + np.dot(g, Rcol) - eps * sumR + eps * sumK) / (I * J)
    dua = (-lam1 * np.sum(p * dx * (np.exp(-f / lam1) - 1.0))
           - lam2 * np.sum(q * dy * (np.exp(-g / lam2) - 1.0))
           - eps * (sumR - sumK) / (I * J))
    return (pri - dua) / abs(pri), pri, dua


def duality_gap_solve(cost: CostOperator, G, lambda1, lambda2, epsilon, batch_size=5, tolerance=1e-8,
                      tau=1000.0, epsilon0=1.0, max_iter=1e7, info=None, **ignored):
    """Log-domain twin of ot_dense.duality_gap_solve.  Returns (f, g, eps_final, row_lse_at_g)."""
    I, J = cost.n, cost.m
    p = np.asarray(G, dtype=np.float64)
    q = np.ones(J) * np.average(G)
    logp, logq = np.log(p), np.log(q)
    log_dx, log_dy = -np.log(I), -np.log(J)
    scale_factor = np.exp(-np.log(epsilon) / ot_dense.EPSILON_SCALINGS)
    f, g = np.zeros(I), np.zeros(J)
    eps_i = epsilon0 * scale_factor
    log_tau = np.log(tau)
    iters_per_stage, total = [], 0
    gap, Lr = np.inf, None
    for e in range(ot_dense.EPSILON_SCALINGS + 1):
        u, v = f.copy(), g.copy()                 # absorb: ot_solvers.py:249-252
        eps_i = eps_i / scale_factor
        alpha1 = lambda1 / (lambda1 + eps_i)
        alpha2 = lambda2 / (lambda2 + eps_i)
        la_old, lb_old = np.zeros(I), np.zeros(J)
        final = e == ot_dense.EPSILON_SCALINGS
        threshold = tolerance if final else 1e-6
        n_inner = int(batch_size) if final else 5
        sumK = None
        gap, n_it = np.inf, 0
        Lr = None   # row LSE at the current g, when already known (re-used by the next update)
        while gap > threshold:
            for _ in range(n_inner):
                n_it += 1
                la_old, lb_old = (f - u) / eps_i, (g - v) / eps_i
                if Lr is None:
                    Lr = cost.row_lse(g / eps_i, eps_i)
                f = eps_i * alpha1 * (logp - (Lr + log_dy))
                Lc = cost.col_lse(f / eps_i, eps_i)
                g = eps_i * alpha2 * (logq - (Lc + log_dx))
                Lr = None
                if ((f - u) / eps_i).max() > log_tau or ((g - v) / eps_i).max() > log_tau:
                    u, v = f.copy(), g.copy()      # ot_func.cpp:778-819
            if final:
                if sumK is None:
                    sumK = np.exp(cost.row_lse(np.zeros(J), eps_i)).sum()
                Lr = cost.row_lse(g / eps_i, eps_i)
                gap, _, _ = gap_from_marginals(f, g, Lr, Lc, sumK, p, q, eps_i, lambda1, lambda2)
            else:
                with np.errstate(over="ignore", invalid="ignore"):
                    at = np.exp((f - u) / eps_i) * np.exp(u / eps_i)
                    bt = np.exp((g - v) / eps_i) * np.exp(v / eps_i)
                    va = np.linalg.norm(at - np.exp(la_old) * np.exp(u / eps_i)) / (1 + np.linalg.norm(at))
                    vb = np.linalg.norm(bt - np.exp(lb_old) * np.exp(v / eps_i)) / (1 + np.linalg.norm(bt))
                gap = max(va, vb)
            if n_it >= max_iter:
                break
        iters_per_stage.append(n_it)
        total += n_it
    if np.isnan(gap):
        raise RuntimeError("Overflow encountered in duality gap computation, please report this incident")
    if info is not None:
        info.update(iters_per_stage=iters_per_stage, total_iters=total, gap=float(gap))
    return f, g, eps_i, Lr


def stablev2_solve(cost: CostOperator, G, lambda1, lambda2, epsilon, scaling_iter=3000, tau=1000.0,
                   epsilon0=1.0, extra_iter=1000, inner_iter_max=50, info=None, **ignored):
    """Log-domain twin of ot_dense.transport_stablev2 (ot_solvers.py:452-531).

    The reference adds 1e-10 to the matvec result K(b*dy) where K is expressed in the
    current absorption frame: K(b dy)_i = exp(u_i/eps) * sum_j exp((g_j-C_ij)/eps) dy_j, so
    in total potentials the floor is  LSE <- logaddexp(LSE, log(1e-10) - u_i/eps)."""
    I, J = cost.n, cost.m
    p = np.asarray(G, dtype=np.float64)
    q = np.ones(J) * np.average(G)
    logp, logq = np.log(p), np.log(q)
    log_dx, log_dy = -np.log(I), -np.log(J)
    warm = tau is not None
    eps_i = epsilon0 if warm else epsilon
    f, g = np.zeros(I), np.zeros(J)
    u, v = np.zeros(I), np.zeros(J)
    log_floor = np.log(1e-10)

    def sweep(f, g):
        Lr = np.logaddexp(cost.row_lse(g / eps_i, eps_i) + log_dy, log_floor - u / eps_i)
        f = eps_i * alpha1 * (logp - Lr)
        Lc = np.logaddexp(cost.col_lse(f / eps_i, eps_i) + log_dx, log_floor - v / eps_i)
        g = eps_i * alpha2 * (logq - Lc)
        return f, g

    alpha1 = lambda1 / (lambda1 + eps_i)
    alpha2 = lambda2 / (lambda2 + eps_i)
    idx = since = n_absorb = 0
    log_tau = np.log(tau) if warm else np.inf
    for _ in range(int(scaling_iter)):
        f, g = sweep(f, g)
        since += 1
        if max(((f - u) / eps_i).max(), ((g - v) / eps_i).max()) > log_tau:
            # :506 tests max(|a|,|b|) > tau; a,b > 0 so |a| = a
            u, v = f.copy(), g.copy()
            n_absorb += 1
        if warm and since == inner_iter_max:
            idx += 1
            since = 0
            u, v = f.copy(), g.copy()
            eps_i = (epsilon0 - epsilon) * np.exp(-idx) + epsilon
            alpha1 = lambda1 / (lambda1 + eps_i)
            alpha2 = lambda2 / (lambda2 + eps_i)
    for _ in range(int(extra_iter)):
        f, g = sweep(f, g)
    if info is not None:
        info.update(n_absorb=n_absorb)
    return f, g, eps_i


def plan_from_potentials(cost: CostOperator, f, g, eps):
    """T_ij = exp((f_i + g_j - C_ij)/eps) / J  (ot_solvers.py:449)."""
    out = np.empty((cost.n, cost.m))
    for s in range(0, cost.n, cost.block):
        e = min(cost.n, s + cost.block)
        out[s:e] = np.exp((f[s:e, None] + g[None, :] - cost.rows(s, e)) / eps)
    return out / cost.m


def row_sums_from_potentials(f, Lr, eps, J):
    """gamma.sum(axis=1) = exp(f_i/eps + LSE_j[(g_j-C_ij)/eps]) / J  (growth update, ot_solvers.py:116)."""
    return np.exp(f / eps + Lr) / J
