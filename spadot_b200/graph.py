"""Spatial kNN graph and deterministic 2-hop mini-batches for the GAT encoder, without dense N x N
adjacency and without torch_geometric.

* `spatial_edge_index(coords, k)`  = `_Cal_Spatial_Net` + `dense_to_sparse`
  (SpaDOT/utils/_utils.py:52-100, utils/_train_utils.py:69-72): directed edges query -> neighbour for the
  k nearest other spots plus one self loop per spot, ordered like the non-zeros of the dense adjacency
  (by source, then target).
* `knn_cutoff(n)` = `min(30, 6 * round(n / 1000))`  (utils/_train_utils.py:69).
* `two_hop_batches(...)` stands in for `NeighborLoader(num_neighbors=[f, f], batch_size, input_nodes=None)`
  (utils/_train_utils.py:80-85, `subgraph_type="induced"`, no shuffling): seeds first, then newly reached nodes;
  edges = the induced sub-graph of the sampled nodes (see `two_hop_batches` for the fan-out deviation).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib


def knn_cutoff(n, max_neighbors=30, knn_cutoff_base=6):
    return int(min(max_neighbors, knn_cutoff_base * round(n / 1000)))


GRID_MIN_POINTS = 4096        # below this the brute-force kernel is a single wave of CTAs anyway
GRID_POINTS_PER_CELL = 4.0


def knn(coords, k, device=None, method="auto"):
    """(n,k) int64 indices of the k nearest other points, nearest first, ties by index (fp64 distances).
    method: "brute" = O(n^2) scan (sdb_knn_f64), "grid" = uniform-grid ring search for 2-D points (sdb_knn_grid_f64, O(n k),
    same result), "auto" = grid for 2-D sets of at least GRID_MIN_POINTS points."""
    _lib.require_device()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    pts = torch.as_tensor(np.asarray(coords) if not isinstance(coords, torch.Tensor) else coords).to(dev, torch.float64).contiguous()
    n, dim = pts.shape
    idx = torch.empty((n, k), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    use_grid = method == "grid" or (method == "auto" and dim == 2 and n >= GRID_MIN_POINTS)
    if use_grid and dim == 2 and n > k:
        lo, hi = pts.min(0).values, pts.max(0).values
        x0, y0 = float(lo[0]), float(lo[1])
        w, hgt = float(hi[0]) - x0, float(hi[1]) - y0
        if w > 0 and hgt > 0:
            h = math.sqrt(w * hgt * GRID_POINTS_PER_CELL / n)
            gx, gy = max(1, int(math.ceil(w / h))), max(1, int(math.ceil(hgt / h)))
            if gx * gy <= (1 << 26):
                cx = torch.clamp(((pts[:, 0] - x0) / h).floor().long(), 0, gx - 1)
                cy = torch.clamp(((pts[:, 1] - y0) / h).floor().long(), 0, gy - 1)
                cell = cy * gx + cx
                order = torch.argsort(cell, stable=True)
                cell_sorted = cell[order].to(torch.int32).contiguous()
                start = torch.zeros(gx * gy + 1, dtype=torch.int32, device=dev)
                start[1:] = torch.cumsum(torch.bincount(cell, minlength=gx * gy), 0).to(torch.int32)
                pts_sorted = pts[order].contiguous()
                order32 = order.to(torch.int32).contiguous()
                _lib.call("sdb_knn_grid_f64", pts_sorted.data_ptr(), order32.data_ptr(), cell_sorted.data_ptr(), start.data_ptr(), n,
                          gx, gy, x0, y0, h, int(k), idx.data_ptr(), 0, st)
                return idx.long()
    _lib.call("sdb_knn_f64", pts.data_ptr(), n, dim, int(k), idx.data_ptr(), 0, st)
    return idx.long()


def spatial_edge_index(coords, k_cutoff, device=None):
    nbr = knn(coords, k_cutoff, device)
    n = nbr.shape[0]
    src = torch.arange(n, device=nbr.device).repeat_interleave(k_cutoff + 1)
    dst = torch.cat([nbr, torch.arange(n, device=nbr.device)[:, None]], dim=1).reshape(-1)
    order = torch.argsort(src * n + dst)                      # dense_to_sparse order
    return torch.stack([src[order], dst[order]])


class InNeighbours:
    """CSR of in-neighbours (sources of the edges pointing to each node)."""

    def __init__(self, edge_index, num_nodes):
        src, dst = edge_index[0].long(), edge_index[1].long()
        order = torch.argsort(dst * num_nodes + src)
        self.src = src[order]
        self.rowptr = torch.zeros(num_nodes + 1, dtype=torch.long, device=edge_index.device)
        self.rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=num_nodes), 0)
        self.n = num_nodes

    def edges_into(self, nodes):
        """All edges (source, target) whose target is in `nodes` (1-D long tensor)."""
        start, end = self.rowptr[nodes], self.rowptr[nodes + 1]
        deg = end - start
        tgt = nodes.repeat_interleave(deg)
        offs = torch.arange(int(deg.sum()), device=nodes.device) - torch.cumsum(deg, 0).repeat_interleave(deg) + deg.repeat_interleave(deg)
        return self.src[start.repeat_interleave(deg) + offs], tgt


def two_hop_batches(edge_index, num_nodes, batch_size=512, num_hops=2):
    """Yields (node_ids, local_edge_index, n_seeds): seeds are node_ids[:n_seeds] in sequential order.

    Node set = seeds, then the nodes newly reached by each hop over incoming edges; edge set = the INDUCED sub-graph,
    i.e. every edge of the full graph whose source and target were both sampled (`subgraph_type="induced"`,
    utils/_train_utils.py:80-85) — the encoder has three GAT layers over a 2-hop batch, so the 2-hop nodes must
    aggregate from their sampled in-neighbours and not only from their self loop.
    Known deviation: the reference's fan-out is max(30, 6*round(N/1000)) per hop; a node whose in-degree (self loop
    included) exceeds it is randomly sub-sampled there (PyG's C++ RNG, not reproducible), here every in-neighbour is
    taken.  With k <= 12 (N < 2500, all ChickenHeart timepoints) in-degrees stay far below 30."""
    nb = InNeighbours(edge_index, num_nodes)
    dev = edge_index.device
    local = torch.full((num_nodes,), -1, dtype=torch.long, device=dev)     # one map, reset per batch where touched
    for s0 in range(0, num_nodes, batch_size):
        seeds = torch.arange(s0, min(num_nodes, s0 + batch_size), device=dev)
        nodes, frontier = seeds, seeds
        local[seeds] = torch.arange(seeds.numel(), device=dev)
        for _ in range(num_hops):
            s, _t = nb.edges_into(frontier)
            new = torch.unique(s[local[s] < 0])
            local[new] = torch.arange(nodes.numel(), nodes.numel() + new.numel(), device=dev)
            nodes = torch.cat([nodes, new])
            frontier = new
        s, t = nb.edges_into(nodes)
        keep = local[s] >= 0
        lei = torch.stack([local[s[keep]], local[t[keep]]])
        local[nodes] = -1
        yield nodes, lei, int(seeds.numel())


class TwoHopBatches:
    """The mini-batches of one timepoint as (node_ids, gat.CsrGraph of the induced sub-graph, n_seeds), sampled ONCE and kept.

    The reference builds its NeighborLoader without shuffling over a static graph (utils/_train_utils.py:80-85), so every
    epoch walks the same sequence of seed blocks and - with the fan-out above every in-degree - the same sub-graphs; PyG
    re-samples them anyway.  Here the first pass through the loader samples each batch (two_hop_batches) and prepares its graph
    for the GAT kernels (CSR by destination and by source, the Z-curve CTA order from `pos`, later the prefix plans the encoder
    asks for); from the second epoch on a step starts with everything in place.  Batches are kept while they fit `max_cache_bytes`
    of device memory (SYN-T: ~35 MB per 40k-node batch), otherwise every epoch re-samples like the reference.
    `spadot_b200.gat.GATEncoder` / `GATConv` take the CsrGraph wherever they take an edge_index."""

    def __init__(self, edge_index, num_nodes, batch_size=512, num_hops=2, pos=None, max_cache_bytes=8 << 30):
        self.edge_index, self.num_nodes, self.batch_size, self.num_hops = edge_index, int(num_nodes), int(batch_size), int(num_hops)
        self.pos = None if pos is None else torch.as_tensor(pos, device=edge_index.device)
        self.max_cache_bytes = int(max_cache_bytes)
        self.cached_bytes = 0
        self._cache = None

    def __len__(self):
        return -(-self.num_nodes // self.batch_size)

    @property
    def cached(self):
        return self._cache is not None

    def __iter__(self):
        if self._cache is not None:
            yield from self._cache
            return
        from . import gat
        kept, nbytes = [], 0
        for nodes, lei, ns in two_hop_batches(self.edge_index, self.num_nodes, self.batch_size, self.num_hops):
            g = gat.CsrGraph(lei, int(nodes.numel()), True, None if self.pos is None else self.pos[nodes])
            item = (nodes, g, ns)
            if kept is not None:
                nbytes += nodes.numel() * nodes.element_size() + g.nbytes()
                if nbytes > self.max_cache_bytes:
                    kept = None                      # does not fit: stream this and every later epoch
                else:
                    kept.append(item)
            yield item
        if kept is not None:                         # only a COMPLETE pass becomes the cache
            self._cache, self.cached_bytes = kept, nbytes
