"""Drop-in for the `sklearn.cluster.KMeans(n_clusters, random_state, n_init=10)` calls of SpaDOT
(utils/_train_utils.py:255-269 every epoch, utils/_analyze_utils.py:10-39,42-105 in analyze).

The procedure is sklearn's, step for step, so the same `random_state` gives the same clustering:
centre the data, `tol * mean(var)`, then for each of the `n_init` runs draw the k-means++ seeds from ONE
RandomState stream (sklearn's own `kmeans_plusplus`, a few k x n distance evaluations on the host), run
Lloyd's iterations to strict label convergence or a squared centre shift below the tolerance, keep the run
with the smallest inertia.  Lloyd's iterations — the part that scales with n x k x iterations x n_init — run on
the device (csrc/sdb_kmeans.cu): for the sizes SpaDOT meets every epoch all n_init restarts are ONE launch, one CTA per
restart running its whole loop on chip; larger problems use per-iteration kernels (fused assignment + per-cluster
accumulation, centre update, inertia) driven from the host.
"""
from __future__ import annotations

import numpy as np
import torch
from sklearn.cluster import kmeans_plusplus
from sklearn.utils import check_random_state
from sklearn.utils.extmath import row_norms

from . import _lib


def _same_clustering(l1, l2, k):
    """sklearn's _is_same_clustering: every label of l1 maps to a single label of l2."""
    pairs = np.unique(l1.astype(np.int64) * k + l2.astype(np.int64))
    return pairs.size == np.unique(l1).size


class KMeans:
    def __init__(self, n_clusters=8, *, n_init=10, max_iter=300, tol=1e-4, random_state=None, device=None):
        self.n_clusters, self.n_init, self.max_iter, self.tol, self.random_state = n_clusters, n_init, max_iter, tol, random_state
        self.device = device

    # ------------------------------------------------------------------ one Lloyd run on the device
    def _lloyd(self, Xd, X_host, centers_init, tol):
        n, d = Xd.shape
        k = self.n_clusters
        dev = Xd.device
        st = torch.cuda.current_stream(dev).cuda_stream
        centers = torch.from_numpy(np.ascontiguousarray(centers_init)).to(dev)
        centers_new = torch.empty_like(centers)
        labels = torch.full((n,), -1, dtype=torch.int32, device=dev)
        work = torch.zeros(k * d + k + 2, dtype=torch.float64, device=dev)     # sums | counts | shift | (pad)
        sums, counts, shift = work[:k * d], work[k * d:k * d + k], work[k * d + k:k * d + k + 1]
        changed = torch.zeros(1, dtype=torch.int32, device=dev)
        strict, n_iter = False, 0
        for it in range(self.max_iter):
            n_iter = it + 1
            work.zero_()
            changed.zero_()
            _lib.call("sdb_kmeans_assign", Xd.data_ptr(), centers.data_ptr(), n, d, k, labels.data_ptr(), changed.data_ptr(),
                      sums.data_ptr(), counts.data_ptr(), st)
            cnt = counts.cpu().numpy()
            if (cnt == 0).any():
                self._relocate_empty(X_host, labels, centers, sums, counts, cnt)
            _lib.call("sdb_kmeans_update", sums.data_ptr(), counts.data_ptr(), centers.data_ptr(), centers_new.data_ptr(), d, k,
                      shift.data_ptr(), st)
            centers, centers_new = centers_new, centers
            if int(changed.item()) == 0:
                strict = True
                break
            if float(shift.item()) <= tol:
                break
        if not strict:
            # rerun the E-step so that the labels match the final centres (sklearn does the same)
            changed.zero_()
            _lib.call("sdb_kmeans_assign", Xd.data_ptr(), centers.data_ptr(), n, d, k, labels.data_ptr(), changed.data_ptr(), 0, 0, st)
        scratch = torch.zeros(10 * 1024 + 2, dtype=torch.float64, device=dev)
        out = torch.zeros(1, dtype=torch.float64, device=dev)
        _lib.call("sdb_kmeans_inertia", Xd.data_ptr(), centers.data_ptr(), labels.data_ptr(), n, d, out.data_ptr(), scratch.data_ptr(), st)
        return labels.cpu().numpy().astype(np.int32), float(out.item()), centers.cpu().numpy(), n_iter

    FUSED_MAX_ELEMS = 1 << 19     # n*d up to which all restarts run as one launch, one CTA per restart (measured crossover: 20k x 20 faster, 30k x 20 slower)

    def _lloyd_fused(self, Xd, inits, tol):
        """All restarts in one launch (sdb_kmeans_lloyd_runs); None for a restart that has to be redone on the host path."""
        n, d = Xd.shape
        k, r = self.n_clusters, len(inits)
        dev = Xd.device
        c0 = torch.from_numpy(np.ascontiguousarray(np.stack(inits))).to(dev)
        labels = torch.empty((r, n), dtype=torch.int32, device=dev)
        centers = torch.empty((r, k, d), dtype=torch.float64, device=dev)
        inertia = torch.empty(r, dtype=torch.float64, device=dev)
        meta = torch.empty((2, r), dtype=torch.int32, device=dev)
        _lib.call("sdb_kmeans_lloyd_runs", Xd.data_ptr(), n, d, k, c0.data_ptr(), r, int(self.max_iter), float(tol),
                  labels.data_ptr(), centers.data_ptr(), inertia.data_ptr(), meta[0].data_ptr(), meta[1].data_ptr(),
                  torch.cuda.current_stream(dev).cuda_stream)
        meta_h = meta.cpu().numpy()
        labels_h, centers_h, inertia_h = labels.cpu().numpy(), centers.cpu().numpy(), inertia.cpu().numpy()
        return [(labels_h[i].astype(np.int32), float(inertia_h[i]), centers_h[i], int(meta_h[0, i])) if meta_h[1, i] == 0 else None
                for i in range(r)]

    @staticmethod
    def _relocate_empty(X, labels, centers, sums, counts, cnt):
        """sklearn's _relocate_empty_clusters_dense: an empty cluster takes the sample farthest from its centre."""
        lab = labels.cpu().numpy()
        c_old = centers.cpu().numpy()
        s = sums.cpu().numpy().reshape(c_old.shape).copy()
        w = cnt.copy()
        empty = np.where(w == 0)[0]
        dist = ((X - c_old[lab]) ** 2).sum(axis=1)
        far = np.argpartition(dist, -empty.size)[:-empty.size - 1:-1]
        for new_id, far_idx in zip(empty, far):
            old_id = lab[far_idx]
            s[old_id] -= X[far_idx]
            s[new_id] = X[far_idx]
            w[new_id] = 1.0
            w[old_id] -= 1.0
        sums.copy_(torch.from_numpy(s.reshape(-1)).to(sums.device))
        counts.copy_(torch.from_numpy(w).to(counts.device))

    # ------------------------------------------------------------------ sklearn surface
    def fit(self, X, y=None):
        _lib.require_device()
        X = np.array(X, dtype=np.float64, order="C", copy=True)
        n, d = X.shape
        if n < self.n_clusters:
            raise ValueError(f"n_samples={n} should be >= n_clusters={self.n_clusters}.")
        tol = float(np.mean(np.var(X, axis=0)) * self.tol)
        rs = check_random_state(self.random_state)
        mean = X.mean(axis=0)
        X -= mean
        x_sq = row_norms(X, squared=True)
        dev = torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        Xd = torch.from_numpy(X).to(dev)
        # Lloyd's iterations draw no random numbers, so seeding every restart first consumes the RandomState stream
        # exactly as sklearn's seed / run / seed / run order does
        inits = [kmeans_plusplus(X, self.n_clusters, x_squared_norms=x_sq, random_state=rs)[0] for _ in range(self.n_init)]
        runs = self._lloyd_fused(Xd, inits, tol) if n * d <= self.FUSED_MAX_ELEMS else [None] * self.n_init
        best = None
        for centers_init, run in zip(inits, runs):
            # a fused run that met an empty cluster (or a problem too large for one CTA per run) goes through the
            # host-driven iterations, which implement sklearn's relocation
            labels, inertia, centers, n_iter = run if run is not None else self._lloyd(Xd, X, centers_init, tol)
            if best is None or (inertia < best[1] and not _same_clustering(labels, best[0], self.n_clusters)):
                best = (labels, inertia, centers, n_iter)
        self.labels_, self.inertia_, centers, self.n_iter_ = best
        self.cluster_centers_ = centers + mean
        return self

    def fit_predict(self, X, y=None):
        return self.fit(X).labels_


def select_n_clusters(wss, min_clusters=4, wss_threshold=0.1):
    """The elbow rule of `Adaptive_clustering` (utils/_analyze_utils.py:73-88) on the within-cluster sums of squares of
    k = min_clusters, min_clusters + 1, ...: among the k whose drop wss[k-1] - wss[k] exceeds `wss_threshold` of the WSS
    range, the one with the largest ratio between its drop and the next drop.  Returns (k, table) where table holds the
    reference's columns (clusters, wss, wss_diff, wss_diff_ratio; NaN where the reference has None)."""
    wss = np.asarray(wss, dtype=np.float64)
    K = wss.size
    if K < 3:
        raise ValueError("the elbow rule needs at least three cluster counts")
    diff = np.full(K, np.nan)
    diff[1:] = -(wss[1:] - wss[:-1])
    ratio = np.full(K, np.nan)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio[1:K - 1] = diff[1:K - 1] / diff[2:K]
    keep = diff > wss_threshold * (wss.max() - wss.min())            # NaN compares false, like pandas' filter
    cand = np.where(keep & ~np.isnan(ratio))[0]
    if cand.size == 0:
        raise ValueError("no cluster count passes the WSS threshold")      # pandas' idxmax raises on an all-NA column
    best = cand[np.argmax(ratio[cand])]                                       # first maximum, like idxmax
    table = dict(clusters=np.arange(min_clusters, min_clusters + K), wss=wss, wss_diff=diff, wss_diff_ratio=ratio)
    return int(min_clusters + best), table


def adaptive_clustering(X, min_clusters=4, max_clusters=20, wss_threshold=0.1, random_state=1993, n_init=10, device=None):
    """`Adaptive_clustering` for one timepoint's latent matrix (utils/_analyze_utils.py:42-105): k-means for every
    k in [min_clusters, max_clusters] on the device, the elbow rule above, then the labels of the selected k.
    Returns (labels, k, table).  (The reference re-fits the selected k with the same random_state, i.e. the same result.)"""
    fits = [KMeans(n_clusters=k, random_state=random_state, n_init=n_init, device=device).fit(X)
            for k in range(min_clusters, max_clusters + 1)]
    k, table = select_n_clusters([f.inertia_ for f in fits], min_clusters, wss_threshold)
    return fits[k - min_clusters].labels_, k, table
