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
    # full (unpartitioned) embeddings come in here: a multi-rank solve needs an explicit `dist` and pre-partitioned rows
    dist = dist or sinkhorn.Dist(enabled=False)
    ops = CudaOps(X0, X1)
    median = sinkhorn.median_cost(ops, dist)
    G = np.ones(ops.n) if growth_rates is None else np.asarray(growth_rates, dtype=np.float64)
    growth = [G]
    cp = None
    for _ in range(int(cfg["growth_iters"])):
        cp = ot_solvers.solve_coupling(X0, X1, cfg, G=growth[-1], median=median, ops=ops, dist=dist)
        growth.append(cp.row_mass().cpu().numpy())
    return cp, growth


def label_codes(labels):
    """Domain labels -> (integer codes, sorted names).  The reference keys populations by the strings
    '<timepoint>_<kmeans label>' and wot orders them lexicographically (_analyze_utils.py:128-136)."""
    names, codes = np.unique(np.asarray(labels).astype(str), return_inverse=True)
    return codes.astype(np.int64), list(names)


def ot_analysis(embeddings, labels, config=None, n_domains=None, out_dir=None, ids=None, prefix="", write_tmaps=False):
    """Transition tables between the domains of adjacent timepoints (_analyze_utils.py:128-138).

    embeddings: list of (N_t, d) latent arrays in timepoint order; labels: list of domain labels (integers, or
    anything `label_codes` can order).  Returns a list of (k_t, k_{t+1}) float64 arrays, table[a, b] = transported
    mass from domain a to domain b.  With `out_dir`, also writes what wot leaves on disk in array form:
    `OT_g.txt` (TSV: id, g0..gK = learned growth per growth iteration, as examples/ChickenHeart_output/OT_g.txt) and
    `<prefix>transition_table_<t>_<t+1>.h5ad` (AnnData, as _analyze_utils.py:138 writes and :183 reads; `.npz` with the same
    content when anndata is not installed) and, with `write_tmaps`, wot's dense `OT_<t>_<t+1>.h5ad` transport maps."""
    import os
    tables, growth_rows = [], []
    coded = []
    for lab in labels:
        lab = np.asarray(lab)
        coded.append((lab.astype(np.int64), [str(v) for v in range(int(lab.max()) + 1)]) if np.issubdtype(lab.dtype, np.integer)
                     else label_codes(lab))
    for t in range(len(embeddings) - 1):
        cp, growth = transport_between(embeddings[t], embeddings[t + 1], config)
        k0 = n_domains[t] if n_domains else len(coded[t][1])
        k1 = n_domains[t + 1] if n_domains else len(coded[t + 1][1])
        tables.append(cp.transition_table(coded[t][0], coded[t + 1][0], k0, k1).cpu().numpy())
        growth_rows.append(np.stack(growth, axis=1))
        if out_dir is not None:
            os.makedirs(out_dir, exist_ok=True)
            # population names as the reference builds them: '<timepoint>_<kmeans label>' (_analyze_utils.py:128)
            # (integer k-means labels get the timepoint prefix; labels that already are population names stay as given)
            names = [[f"{tt}_{v}" for v in coded[tt][1]] if np.issubdtype(np.asarray(labels[tt]).dtype, np.integer) else coded[tt][1]
                     for tt in (t, t + 1)]
            write_transition_table(os.path.join(out_dir, f"{prefix}transition_table_{t}_{t + 1}"), tables[-1], names[0], names[1])
            if write_tmaps:
                write_transport_map(os.path.join(out_dir, f"OT_{t}_{t + 1}"), cp.plan().cpu().numpy(),
                                    ids[t] if ids is not None else [f"t{t}_{i}" for i in range(len(embeddings[t]))],
                                    ids[t + 1] if ids is not None else [f"t{t + 1}_{i}" for i in range(len(embeddings[t + 1]))],
                                    np.stack(growth, axis=1))
    if out_dir is not None and growth_rows:
        G = np.concatenate(growth_rows, axis=0)
        names = np.concatenate([np.asarray(ids[t]).astype(str) if ids is not None else
                                np.array([f"t{t}_{i}" for i in range(len(embeddings[t]))]) for t in range(len(embeddings) - 1)])
        with open(os.path.join(out_dir, "OT_g.txt"), "w") as fh:
            fh.write("id\t" + "\t".join(f"g{k}" for k in range(G.shape[1])) + "\n")
            for name, row in zip(names, G):
                fh.write(name + "\t" + "\t".join(repr(float(v)) for v in row) + "\n")
    return tables


def have_anndata():
    try:
        import anndata  # noqa: F401
        return True
    except Exception:
        return False


def write_transition_table(path_stem, table, row_names, col_names):
    """`transition_table.write_h5ad(<prefix>transition_table_<d0>_<d1>.h5ad)` of _analyze_utils.py:138 — an AnnData whose X is
    the table, obs_names the source populations and var_names the target populations (what `plot_OT` reads back at :183).
    anndata is an optional dependency (absent from this image): without it the same content goes to `<stem>.npz` and the
    path written is returned either way."""
    table = np.asarray(table, dtype=np.float64)
    if have_anndata():
        import anndata
        import pandas as pd
        ad = anndata.AnnData(X=table, obs=pd.DataFrame(index=pd.Index([str(r) for r in row_names])),
                             var=pd.DataFrame(index=pd.Index([str(c) for c in col_names])))
        ad.write_h5ad(path_stem + ".h5ad")
        return path_stem + ".h5ad"
    np.savez(path_stem + ".npz", table=table, rows=np.array([str(r) for r in row_names]), cols=np.array([str(c) for c in col_names]))
    return path_stem + ".npz"


def read_transition_table(path):
    """Reads what write_transition_table wrote (either format): (table, row names, column names)."""
    if path.endswith(".h5ad"):
        import anndata
        ad = anndata.read_h5ad(path)
        return np.asarray(ad.X, dtype=np.float64), list(ad.obs_names), list(ad.var_names)
    z = np.load(path)
    return z["table"], [str(v) for v in z["rows"]], [str(v) for v in z["cols"]]


def write_transport_map(path_stem, plan, source_ids, target_ids, growth):
    """wot's `OT_<t0>_<t1>.h5ad` (ot_model.compute_all_transport_maps(tmap_out=...), _analyze_utils.py:125): X = the dense
    transport map, obs = source cells with the learned growth columns g0..gK, var = target cells.  Dense N x M: only for
    sizes that fit in host memory, and only on request - the streamed solver never needs it."""
    plan = np.asarray(plan, dtype=np.float64)
    growth = np.asarray(growth, dtype=np.float64)
    if have_anndata():
        import anndata
        import pandas as pd
        obs = pd.DataFrame({f"g{k}": growth[:, k] for k in range(growth.shape[1])}, index=pd.Index([str(v) for v in source_ids]))
        ad = anndata.AnnData(X=plan, obs=obs, var=pd.DataFrame(index=pd.Index([str(v) for v in target_ids])))
        ad.write_h5ad(path_stem + ".h5ad")
        return path_stem + ".h5ad"
    np.savez(path_stem + ".npz", X=plan, obs_names=np.array([str(v) for v in source_ids]), var_names=np.array([str(v) for v in target_ids]),
             growth=growth)
    return path_stem + ".npz"


def transition_probabilities(table):
    """plot_OT's normalisation (_analyze_utils.py:185-193): elementwise min of the column- and row-normalised table."""
    table = np.asarray(table, dtype=np.float64)
    col = table / table.sum(axis=0, keepdims=True)
    row = table / table.sum(axis=1, keepdims=True)
    return np.minimum(col, row)
