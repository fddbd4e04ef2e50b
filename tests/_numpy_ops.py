"""Oracle-backed stand-in for spadot_b200.cuda_ops.CudaOps.  TEST INFRASTRUCTURE ONLY.

Implements the same `ops` interface with numpy fp64 on CPU torch tensors so that the host
drivers of spadot_b200/sinkhorn.py (stage schedule, stopping rules, growth loop, row
partition + collectives) can be exercised without a GPU, including world_size-2 gloo runs.
The product never imports this file.
"""
import numpy as np
import torch

from oracle import ot_dense
from oracle.ot_logdomain import CostOperator


class NumpyOps:
    def __init__(self, x_local, y):
        self.device = torch.device("cpu")
        self.x = np.asarray(x_local, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        self.n, self.d = self.x.shape
        self.m = self.y.shape[0]
        self.inv_med = 1.0
        self._tick = 0
        self.flag = torch.zeros(1, dtype=torch.int32)
        self.launches = 0

    def tick(self):
        self._tick += 1
        return self._tick

    def zeros(self, n, dtype=torch.float64):
        return torch.zeros(n, dtype=dtype)

    def tensor(self, a, dtype=torch.float64):
        return torch.as_tensor(np.asarray(a), dtype=dtype).clone()

    def set_median(self, median):
        self.inv_med = 1.0 / float(median)

    def _cost(self):
        return CostOperator(self.x, self.y, median=1.0 / self.inv_med)

    def row_lse(self, g, eps, out=None):
        gg = np.zeros(self.m) if g is None else g.numpy()
        L = torch.from_numpy(self._cost().row_lse(gg / eps, eps)) if self.n else torch.zeros(0, dtype=torch.float64)
        if out is None:
            return L
        out.copy_(L)
        return out

    def col_lse(self, f, eps, out=None):
        if self.n:
            L = torch.from_numpy(self._cost().col_lse(f.numpy() / eps, eps))
        else:
            L = torch.full((self.m,), float("-inf"), dtype=torch.float64)
        if out is None:
            return L
        out.copy_(L)
        return out

    def potential_update(self, side, L, logmarg, eps, alpha, log_n_other, pot, frame, la_old, it, log_tau,
                         log_floor=float("-inf")):
        la_old.copy_((pot - frame) / eps)
        LA = L - log_n_other
        if log_floor > float("-inf"):
            LA = torch.logaddexp(LA, log_floor - frame / eps)
        pot.copy_(eps * alpha * (logmarg - LA))
        if pot.numel() and float(((pot - frame) / eps).max()) > log_tau:
            self.flag[0] = max(int(self.flag[0]), it)

    def absorb_flag_tensor(self):
        return self.flag

    def absorb(self, it, f, g, u, v):
        if int(self.flag[0]) == it:
            u.copy_(f)
            v.copy_(g)

    def stage_criterion(self, f, u, la_old, g, v, lb_old, eps):
        with np.errstate(over="ignore", invalid="ignore"):
            eu, ev = np.exp(u.numpy() / eps), np.exp(v.numpy() / eps)
            at = np.exp((f.numpy() - u.numpy()) / eps) * eu
            bt = np.exp((g.numpy() - v.numpy()) / eps) * ev
            da = at - np.exp(la_old.numpy()) * eu
            db = bt - np.exp(lb_old.numpy()) * ev
            return torch.tensor([np.sum(da * da), np.sum(at * at), np.sum(db * db), np.sum(bt * bt)], dtype=torch.float64)

    def gap_terms(self, f, Lr, logp, g, Lc, logq, eps, lam1, lam2, dx, dy):
        f, Lr, logp, g, Lc, logq = (t.numpy() for t in (f, Lr, logp, g, Lc, logq))
        R, p = np.exp(f / eps + Lr), np.exp(logp)
        r = R * dy
        Rc, q = np.exp(g / eps + Lc), np.exp(logq)
        c = Rc * dx
        out = [R.sum(), (f * R).sum(), (dx * (r * (np.log(r) - logp) - r + p)).sum(), (p * dx * (np.exp(-f / lam1) - 1)).sum(),
               Rc.sum(), (g * Rc).sum(), (dy * (c * (np.log(c) - logq) - c + q)).sum(), (q * dy * (np.exp(-g / lam2) - 1)).sum(), 0, 0]
        return torch.tensor(out, dtype=torch.float64)

    def sum_exp(self, L):
        return torch.exp(L).sum().reshape(1)

    def row_mass(self, f, Lr, eps):
        return torch.exp(f / eps + Lr) / self.m

    def plan_dense(self, f, g, eps):
        C = ot_dense.sqeuclidean(self.x, self.y) * self.inv_med
        return torch.from_numpy(np.exp((f.numpy()[:, None] + g.numpy()[None, :] - C) / eps) / self.m)

    # median primitives (exact fp64; the "fp32 classification" of the device path is emulated in fp64)
    def pair_distances(self, ii, jj):
        d = self.x[np.asarray(ii)] - self.y[np.asarray(jj)]
        return torch.from_numpy(np.einsum("ij,ij->i", d, d))

    def cost_histogram(self, lo, hi, n_bins):
        C = ot_dense.sqeuclidean(self.x, self.y).ravel()
        below = int((C < lo).sum())
        inside = C[(C >= lo) & (C < hi)]
        bins = np.clip(((inside - lo) * (n_bins / (hi - lo))).astype(np.int64), 0, n_bins - 1)
        hist = np.bincount(bins, minlength=n_bins)
        return torch.from_numpy(hist.astype(np.int64)), torch.tensor([below, inside.size], dtype=torch.int64)

    def cost_collect(self, lo, hi, cap):
        C = ot_dense.sqeuclidean(self.x, self.y).ravel()
        below = int((C < lo).sum())
        inside = C[(C >= lo) & (C < hi)]
        cand = torch.zeros(cap, dtype=torch.float64)
        cand[:min(cap, inside.size)] = torch.from_numpy(inside[:cap])
        return cand, torch.tensor([below, inside.size], dtype=torch.int64)

    def radix_digit_hist(self, cand, n, shift, prefix):
        keys = cand[:n].numpy().view(np.uint64)
        match = np.ones(n, dtype=bool) if shift >= 56 else ((keys >> np.uint64(shift + 8)) == np.uint64(prefix))
        digits = ((keys[match] >> np.uint64(shift)) & np.uint64(255)).astype(np.int64)
        return torch.from_numpy(np.bincount(digits, minlength=256).astype(np.int64))

    def transition_table(self, f, g, eps, labels_x, labels_y, k0, k1):
        T = self.plan_dense(f, g, eps).numpy()
        return torch.from_numpy(ot_dense.transition_table(T, labels_x, labels_y, k0, k1))
