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
