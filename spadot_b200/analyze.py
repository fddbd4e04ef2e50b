"""`analyze`-side OT on arrays: the computation behind SpaDOT/utils/_analyze_utils.py:108-138 and :183-193.

The reference hands the latent AnnData to wot (`OTModel(..., epsilon=0.05, epsilon0=1, lambda1=0.1, lambda2=5,
growth_iters=3).compute_all_transport_maps()`, then `TransportMapModel.transition_table` over the k-means domain
labels).  wot is un-vendored; its solver text is the vendored `optimal_transport_duality_gap`, its conventions are
tau = 10000, batch_size = 5, tolerance = 1e-8, and the LAST growth iteration is the transport map.  Here every
adjacent pair of timepoints is one streamed coupling on the device: no N x M transport map is written unless asked.
"""
from __future__ import annotations

import numpy as np

from . import ot_solvers, sinkhorn
from .cuda_ops import CudaOps

WOT_CONFIG = dict(epsilon=0.05, epsilon0=1.0, lambda1=0.1, lambda2=5.0, growth_iters=3, batch_size=5,
                  tolerance=1e-8, tau=10000.0, max_iter=1e7)      # _analyze_utils.py:124 + wot defaults


def transport_between(X0, X1, config=None, growth_rates=None, dist=None):
    """wot's compute_transport_map for one pair: `growth_iters` solves with g <- row sums; returns the last
    Coupling and the per-iteration learned growth (the g0..gK columns of OT_g.txt)."""
    cfg = dict(WOT_CONFIG, **(config or {}))
    dist = dist or sinkhorn.Dist()
    ops = CudaOps(X0, X1)
    median = sinkhorn.median_cost(ops, dist)
    G = np.ones(ops.n) if growth_rates is None else np.asarray(growth_rates, dtype=np.float64)
    growth = [G]
    cp = None
    for _ in range(int(cfg["growth_iters"])):
        cp = ot_solvers.solve_coupling(X0, X1, cfg, G=growth[-1], median=median, ops=ops, dist=dist)
        growth.append(cp.row_mass().cpu().numpy())
    return cp, growth


def ot_analysis(embeddings, labels, config=None, n_domains=None):
    """Transition tables between the domains of adjacent timepoints (_analyze_utils.py:128-138).

    embeddings: list of (N_t, d) latent arrays in timepoint order; labels: list of integer domain labels.
    Returns a list of (k_t, k_{t+1}) float64 arrays, table[a, b] = sum of transported mass from domain a to b."""
    tables = []
    for t in range(len(embeddings) - 1):
        cp, _ = transport_between(embeddings[t], embeddings[t + 1], config)
        k0 = (n_domains[t] if n_domains else None)
        k1 = (n_domains[t + 1] if n_domains else None)
        tables.append(cp.transition_table(labels[t], labels[t + 1], k0, k1).cpu().numpy())
    return tables


def transition_probabilities(table):
    """plot_OT's normalisation (_analyze_utils.py:185-193): elementwise min of the column- and row-normalised table."""
    table = np.asarray(table, dtype=np.float64)
    col = table / table.sum(axis=0, keepdims=True)
    row = table / table.sum(axis=1, keepdims=True)
    return np.minimum(col, row)
