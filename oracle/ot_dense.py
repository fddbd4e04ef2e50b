"""Dense fp64 numpy restatement of the reference's OT path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this
module; spadot_b200/ never does (see oracle/README.md).

Citations are relative to /root/reference/SpaDOT/utils/OT_loss/.
The restatement is intentionally literal (dense K, scaling vectors a/b, absorption at
tau) so that it reproduces the reference bit-for-bit up to summation order; the
log-domain twin lives in ot_logdomain.py.
"""
from __future__ import annotations

import numpy as np

EPSILON_SCALINGS = 5  # ot_solvers.py:217


# --------------------------------------------------------------------------- cost (a1)
def sqeuclidean(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """C_ij = sum_k (a_ik - b_jk)^2 by DIRECT differences.

    ot_solvers.py:102 calls sklearn pairwise_distances(metric="sqeuclidean", n_jobs=1),
    which dispatches to scipy cdist (direct differences, no |a|^2+|b|^2-2ab expansion).
    """
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float64)
    step = max(1, int(2e7 // max(1, b.shape[0] * a.shape[1])))
    for s in range(0, a.shape[0], step):
        diff = a[s:s + step, None, :] - b[None, :, :]
        out[s:s + step] = np.einsum("ijk,ijk->ij", diff, diff)
    return out


def median_normalised_cost(a, b):
    """ot_solvers.py:102-103: C / np.median(C) (median over all N*M entries)."""
    C = sqeuclidean(a, b)
    med = float(np.median(C))
    return C / med, med


# --------------------------------------------------------------------------- objective (a8)
def _fdiv(lam, x, p, dx):
    # ot_solvers.py:124-125 / ot_func.cpp:309-322
    return lam * np.sum(dx * (x * np.log(x / p) - x + p))


def _fdivstar(lam, u, p, dx):
    # ot_solvers.py:128-129 / ot_func.cpp:325-338
    return lam * np.sum((p * dx) * (np.exp(u / lam) - 1.0))


def primal(C, K_, R, dx, dy, p, q, eps, lam1, lam2):
    """ot_solvers.py:132-142 (K_ = exp(-C/eps))."""
    I, J = len(p), len(q)
    with np.errstate(divide="ignore", invalid="ignore"):
        ent = np.sum(R * np.nan_to_num(np.log(R)) - R + K_)
    return (_fdiv(lam1, R @ dy, p, dx) + _fdiv(lam2, R.T @ dx, q, dy)
            + (eps * ent + np.sum(R * C)) / (I * J))


def dual(K_, R, dx, dy, p, q, a_true, b_true, eps, lam1, lam2):
    """ot_solvers.py:145-158 with a_true=a*exp(u/eps), b_true=b*exp(v/eps)."""
    I, J = len(p), len(q)
    t1 = -_fdivstar(lam1, -eps * np.log(a_true), p, dx)
    t2 = -_fdivstar(lam2, -eps * np.log(b_true), q, dy)
    t3 = -eps * np.sum(R - K_) / (I * J)
    return t1 + t2 + t3


# --------------------------------------------------------------------------- solver (a3-a8)
def duality_gap_solve(C, G, lambda1, lambda2, epsilon, batch_size=5, tolerance=1e-8,
                      tau=1000.0, epsilon0=1.0, max_iter=1e7, info=None, **ignored):
    """optimal_transport_duality_gap, ot_solvers.py:164-449, with the inner loop of
    update_process<double> (ot_func.cpp:831-930) / step1_process (ot_func.cpp:690-828).

    Returns R / J (ot_solvers.py:449).  `info`, when a dict, receives iteration counts,
    the final gap and the total potentials f = u + eps*log a, g = v + eps*log b.
    """
    C = np.ascontiguousarray(C, dtype=np.float64)
    I, J = C.shape
    scale_factor = np.exp(-np.log(epsilon) / EPSILON_SCALINGS)  # :218
    dx, dy = np.ones(I) / I, np.ones(J) / J                     # :221
    p = np.asarray(G, dtype=np.float64)                          # :223
    q = np.ones(J) * np.average(G)                               # :224
    u, v = np.zeros(I), np.zeros(J)
    a, b = np.ones(I), np.ones(J)
    eps_i = epsilon0 * scale_factor                              # :240
    iters_per_stage, total_iters, gap = [], 0, np.inf
    R = None
    for e in range(EPSILON_SCALINGS + 1):
        u = u + eps_i * np.log(a)                                # :249 absorb
        v = v + eps_i * np.log(b)
        a, b = np.ones(I), np.ones(J)
        eps_i = eps_i / scale_factor                             # :254
        alpha1 = lambda1 / (lambda1 + eps_i)
        alpha2 = lambda2 / (lambda2 + eps_i)
        old_a, old_b = a.copy(), b.copy()
        threshold = tolerance if e == EPSILON_SCALINGS else 1e-6  # :262
        K_ = np.exp(-C / eps_i)                                  # :271 / ot_func.cpp:558
        K = np.exp((u[:, None] - C + v[None, :]) / eps_i)        # :272 / ot_func.cpp:563
        gap = np.inf
        stage_iters = 0   # the reference passes cur_iter by value: max_iter is per stage
        n_inner = int(batch_size) if e == EPSILON_SCALINGS else 5  # ot_func.cpp:867
        while gap > threshold:                                   # ot_func.cpp:866
            hit_max = False
            for _ in range(n_inner):                             # step1_process :726
                stage_iters += 1
                old_a, old_b = a, b
                a = (p / (K @ (b * dy))) ** alpha1 * np.exp(-u / (lambda1 + eps_i))   # ot_func.cpp:610-636
                b = (q / (K.T @ (a * dx))) ** alpha2 * np.exp(-v / (lambda2 + eps_i))  # ot_func.cpp:642-668
                if a.max() > tau or b.max() > tau:               # ot_func.cpp:778-790 (no abs)
                    u = u + eps_i * np.log(a)
                    v = v + eps_i * np.log(b)
                    K = np.exp((u[:, None] - C + v[None, :]) / eps_i)
                    a, b = np.ones(I), np.ones(J)
                if stage_iters >= max_iter:                      # ot_func.cpp:821-824
                    hit_max = True
                    break
            if hit_max:
                stage_iters = -1  # mirrors the C return value; the reference keeps looping
            a_true = a * np.exp(u / eps_i)                       # ot_func.cpp:878-884
            b_true = b * np.exp(v / eps_i)
            if e == EPSILON_SCALINGS:
                R = (K.T * a).T * b                              # update_R ot_func.cpp:571-584
                pri = primal(C, K_, R, dx, dy, p, q, eps_i, lambda1, lambda2)
                dua = dual(K_, R, dx, dy, p, q, a_true, b_true, eps_i, lambda1, lambda2)
                gap = (pri - dua) / abs(pri)                     # ot_func.cpp:543
            else:
                with np.errstate(over="ignore", invalid="ignore"):
                    va = np.linalg.norm(a_true - old_a * np.exp(u / eps_i)) / (1 + np.linalg.norm(a_true))
                    vb = np.linalg.norm(b_true - old_b * np.exp(v / eps_i)) / (1 + np.linalg.norm(b_true))
                gap = max(va, vb)                                # ot_func.cpp:897-922
            if hit_max:
                break
        iters_per_stage.append(stage_iters)
        total_iters += max(stage_iters, 0)
    if np.isnan(gap):                                            # :446-447
        raise RuntimeError("Overflow encountered in duality gap computation, please report this incident")
    if info is not None:
        info.update(iters_per_stage=iters_per_stage, total_iters=total_iters, gap=float(gap),
                    f=u + eps_i * np.log(a), g=v + eps_i * np.log(b), epsilon_final=eps_i)
    return R / J


def transport_stablev2(C, lambda1, lambda2, epsilon, scaling_iter, G, tau, epsilon0,
                       extra_iter, inner_iter_max, info=None, **ignored):
    """ot_solvers.py:452-531 (wot's fixed_iters solver)."""
    C = np.asarray(C, dtype=np.float64)
    I, J = C.shape
    warm_start = tau is not None
    eps_final = epsilon

    def get_reg(n):                                              # :478-480
        return (epsilon0 - eps_final) * np.exp(-n) + eps_final

    eps_i = epsilon0 if warm_start else epsilon
    dx, dy = np.ones(I) / I, np.ones(J) / J
    p = np.asarray(G, dtype=np.float64)
    q = np.ones(J) * np.average(G)
    u, v = np.zeros(I), np.zeros(J)
    b = np.ones(J)
    K = np.exp(-C / eps_i)
    alpha1 = lambda1 / (lambda1 + eps_i)
    alpha2 = lambda2 / (lambda2 + eps_i)
    eps_index, since = 0, 0
    floor = 1e-10                                                # :498
    n_absorb = 0
    for _ in range(int(scaling_iter)):
        a = (p / (K @ (b * dy) + floor)) ** alpha1 * np.exp(-u / (lambda1 + eps_i))
        b = (q / (K.T @ (a * dx) + floor)) ** alpha2 * np.exp(-v / (lambda2 + eps_i))
        since += 1
        if max(np.abs(a).max(), np.abs(b).max()) > tau:          # :506
            u = u + eps_i * np.log(a)
            v = v + eps_i * np.log(b)
            K = np.exp((u[:, None] - C + v[None, :]) / eps_i)
            a, b = np.ones(I), np.ones(J)
            n_absorb += 1
        if warm_start and since == inner_iter_max:               # :513
            eps_index += 1
            since = 0
            u = u + eps_i * np.log(a)
            v = v + eps_i * np.log(b)
            eps_i = get_reg(eps_index)
            alpha1 = lambda1 / (lambda1 + eps_i)
            alpha2 = lambda2 / (lambda2 + eps_i)
            K = np.exp((u[:, None] - C + v[None, :]) / eps_i)
            a, b = np.ones(I), np.ones(J)
    for _ in range(int(extra_iter)):                             # :525-527
        a = (p / (K @ (b * dy) + floor)) ** alpha1 * np.exp(-u / (lambda1 + eps_i))
        b = (q / (K.T @ (a * dx) + floor)) ** alpha2 * np.exp(-v / (lambda2 + eps_i))
    R = (K.T * a).T * b
    if info is not None:
        info.update(f=u + eps_i * np.log(a), g=v + eps_i * np.log(b), epsilon_final=eps_i,
                    n_absorb=n_absorb)
    return R / J


# --------------------------------------------------------------------------- growth loop (a2)
def compute_transport_map(a, b, config, C=None, G=None, return_all=False, solver=None):
    """ot_solvers.py:95-121.  Returns gammas[0] like SpaDOT (:121); return_all=True
    returns the list so the wot convention (last growth iteration) can be checked too."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if C is None:
        C, _ = median_normalised_cost(a, b)
    cfg = dict(config)
    cfg["C"] = C
    cfg["G"] = np.ones(C.shape[0]) if G is None else np.asarray(G, dtype=np.float64)
    solver = solver or duality_gap_solve
    gammas = []
    for i in range(int(cfg["growth_iters"])):
        if i > 0:
            cfg["G"] = gammas[-1].sum(axis=1)                     # :116
        gammas.append(solver(**cfg))
    return gammas if return_all else gammas[0]


# --------------------------------------------------------------------------- transition table (a10)
def transition_table(tmap, labels0, labels1, k0=None, k1=None):
    """wot TransportMapModel.transition_table as used at _analyze_utils.py:135-137:
    P0^T . T . P1 with 0/1 membership indicators (un-vendored; parity unpinned)."""
    labels0 = np.asarray(labels0)
    labels1 = np.asarray(labels1)
    k0 = int(labels0.max()) + 1 if k0 is None else k0
    k1 = int(labels1.max()) + 1 if k1 is None else k1
    P0 = np.zeros((tmap.shape[0], k0))
    P0[np.arange(tmap.shape[0]), labels0] = 1.0
    P1 = np.zeros((tmap.shape[1], k1))
    P1[np.arange(tmap.shape[1]), labels1] = 1.0
    return P0.T @ tmap @ P1


def plot_ot_normalisation(table):
    """_analyze_utils.py:185-193: min(column-normalised, row-normalised)."""
    col = table / table.sum(axis=0, keepdims=True)
    row = table / table.sum(axis=1, keepdims=True)
    return np.minimum(col, row)


# --------------------------------------------------------------------------- synthetic inputs
def synthetic_embeddings(n, m, d, seed=1993, n_clusters=10, centre_sd=1.5, sd=0.5, shift=0.1):
    """Mixture of Gaussians in R^d, target set shifted by +shift."""
    rng = np.random.default_rng(seed)
    centres = rng.normal(0.0, centre_sd, size=(n_clusters, d))
    la = rng.integers(0, n_clusters, size=n)
    lb = rng.integers(0, n_clusters, size=m)
    a = centres[la] + rng.normal(0.0, sd, size=(n, d))
    b = centres[lb] + rng.normal(0.0, sd, size=(m, d)) + shift
    return a, b, la, lb


DEFAULT_OT_CONFIG = dict(  # SpaDOT/config.yaml:39-57
    growth_iters=3, epsilon=0.05, epsilon0=1.0, lambda1=0.1, lambda2=5.0, tau=1000.0,
    scaling_iter=3000, inner_iter_max=50, tolerance=1e-8, max_iter=1e7, batch_size=5,
    extra_iter=1000, use_Py=False, use_C=True, profiling=False)
