"""numpy restatement of the reference's spatial graph feed.  TEST INFRASTRUCTURE ONLY.

SpaDOT/utils/_utils.py:52-100 (_Cal_Spatial_Net): sklearn NearestNeighbors(n_neighbors=max_neigh+1) on
the raw coordinates, neighbours 1..k_cutoff, edges Cell1 (query) -> Cell2 (neighbour), adjacency
G + I as a dense N x N matrix; SpaDOT/utils/_train_utils.py:69-72: k = min(30, 6*round(N/1000)),
dense_to_sparse(adj) => edge_index in row-major order of the non-zeros (row = source, col = target).
"""
import numpy as np
from sklearn.neighbors import NearestNeighbors


def knn_cutoff(n, max_neighbors=30, knn_cutoff_base=6):
    return int(min(max_neighbors, knn_cutoff_base * round(n / 1000)))      # _train_utils.py:69


def spatial_edge_index(coords, k_cutoff, max_neigh=30):
    n = coords.shape[0]
    nbrs = NearestNeighbors(n_neighbors=min(max_neigh + 1, n), algorithm="auto").fit(coords)
    _, idx = nbrs.kneighbors(coords)
    idx = idx[:, 1:k_cutoff + 1]
    adj = np.zeros((n, n), dtype=bool)
    adj[np.repeat(np.arange(n), idx.shape[1]), idx.ravel()] = True
    adj |= np.eye(n, dtype=bool)                                             # _utils.py:98-100
    src, dst = np.nonzero(adj)                                               # dense_to_sparse order
    return np.stack([src, dst]).astype(np.int64)


def two_hop_batches_ref(edge_index, num_nodes, batch_size=512, num_hops=2):
    """Plain-numpy restatement of the deterministic case of NeighborLoader(num_neighbors=[f, f], batch_size,
    subgraph_type="induced") as the reference builds it (utils/_train_utils.py:80-85, no shuffle): seeds in sequential
    order first, then the nodes newly reached by each hop over incoming edges (sorted), and EVERY edge of the full graph
    between sampled nodes.  Valid while fan-out >= every in-degree (no random sub-sampling)."""
    src, dst = np.asarray(edge_index[0]), np.asarray(edge_index[1])
    into = [[] for _ in range(num_nodes)]
    for s, t in zip(src, dst):
        into[t].append(s)
    for s0 in range(0, num_nodes, batch_size):
        seeds = list(range(s0, min(num_nodes, s0 + batch_size)))
        nodes, seen, frontier = list(seeds), set(seeds), seeds
        for _ in range(num_hops):
            new = sorted({s for t in frontier for s in into[t]} - seen)
            nodes += new
            seen |= set(new)
            frontier = new
        local = {g: i for i, g in enumerate(nodes)}
        es = [(local[s], local[t]) for t in nodes for s in sorted(into[t]) if s in local]
        lei = np.array(es, dtype=np.int64).T.reshape(2, -1)
        yield np.array(nodes, dtype=np.int64), lei, len(seeds)


def beta_cycle_linear(n_iter, start=0.0, stop=1.0, n_cycle=10, ratio=1):
    """utils/_train_utils.py:141-153 (_beta_cycle_linear): the beta1 schedule of the SVGP KL weight."""
    L = np.ones(n_iter) * stop
    period = n_iter / n_cycle
    step = (stop - start) / (period * ratio)
    for c in range(n_cycle):
        v, i = start, 0
        while v <= stop and (int(i + c * period) < n_iter):
            L[int(i + c * period)] = v
            v += step
            i += 1
    return L
