"""Drop-in GATConv / GATEncoder (SpaDOT/model/encoder.py:37-61) without torch_geometric.

`GATConv(in_channels, out_channels, heads, concat)` keeps PyG's parameter names and shapes
(`lin.weight (H*C, in)`, `att_src`, `att_dst (1,H,C)`, `bias`) so state_dicts saved by the reference
(`train.py:39-41`) load unchanged.  The linear layer stays a cuBLAS GEMM; everything between it and the
bias add — gather a_src[j]+a_dst[i], LeakyReLU(0.2), segment softmax over incoming edges, weighted
aggregation, and the whole backward — runs in the fused CSR kernels of csrc/sdb_gat.cu (K2/K2b)
through one autograd Function.  Self-loop handling = PyG's remove_self_loops + add_self_loops.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


class CsrGraph:
    """Edges (source -> target) with self loops, grouped by destination (forward) and by source (backward)."""

    REORDER_MIN_NODES = 20000     # below this every gathered row stays in the 126 MB L2 anyway

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, add_self_loops: bool = True, pos=None):
        src, dst = edge_index[0].long(), edge_index[1].long()
        if add_self_loops:
            keep = src != dst
            loops = torch.arange(num_nodes, dtype=torch.long, device=edge_index.device)
            src, dst = torch.cat([src[keep], loops]), torch.cat([dst[keep], loops])
        order = torch.argsort(dst, stable=True)
        self.n = num_nodes
        self.E = int(src.numel())
        self.col = src[order].to(torch.int32).contiguous()                   # source of each edge, by-destination order
        dst_sorted = dst[order]
        self.rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=edge_index.device)
        self.rowptr[1:] = torch.cumsum(torch.bincount(dst_sorted, minlength=num_nodes), 0)
        by_src = torch.argsort(self.col.long(), stable=True)                 # positions (edge ids) grouped by source
        self.src_eid = by_src.to(torch.int32).contiguous()
        self.src_dst = dst_sorted[by_src].to(torch.int32).contiguous()
        self.src_rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=edge_index.device)
        self.src_rowptr[1:] = torch.cumsum(torch.bincount(self.col.long(), minlength=num_nodes), 0)
        self.order = None
        self._closures, self._orders = {}, {}
        self._base_tensors = (self.col, self.rowptr, self.src_eid, self.src_dst, self.src_rowptr)
        if num_nodes >= self.REORDER_MIN_NODES and pos is not None:
            # locality order for the CTAs from the spots' coordinates: a Z-curve sort on the device (tens of microseconds;
            # the sampled sub-graph of every training batch is a new graph, so this runs once per optimiser step)
            self.order = morton_order(pos)
        elif num_nodes >= self.REORDER_MIN_NODES:
            # no coordinates: reverse Cuthill-McKee of the symmetrised graph on the host (0.1 s at 40k nodes, once per graph)
            import scipy.sparse as sp
            from scipy.sparse.csgraph import reverse_cuthill_mckee
            s_np, d_np = src.cpu().numpy(), dst.cpu().numpy()
            A = sp.csr_matrix((np.ones(s_np.size, dtype=np.int8), (s_np, d_np)), shape=(num_nodes, num_nodes))
            perm = reverse_cuthill_mckee((A + A.T).tocsr(), symmetric_mode=True)
            self.order = torch.from_numpy(np.ascontiguousarray(perm).astype(np.int32)).to(edge_index.device)


def _graph_nbytes(self):
    """Device bytes held by the graph (CSR both ways, CTA order, cached prefix data)."""
    extra = [t for t in (self.order, getattr(self, "_col_cummax", None)) if t is not None] + list(self._orders.values())
    return int(sum(t.numel() * t.element_size() for t in list(self._base_tensors) + [t for t in extra if t is not None]))


def _graph_prefix_plan(self, n_out, n_layers):
    """Row counts a stack of `n_layers` GAT layers needs when only output rows [:n_out] of the last layer are used.

    Layer l (last first) has destinations [:d] and needs its input rows [:s], s = 1 + the largest source of an edge into [:d];
    the layer below then needs destinations [:s].  With hop-ordered batch nodes (seeds, then what each hop added;
    ref: utils/_train_utils.py:80-85) these prefixes ARE the hop sets, otherwise they are supersets of them - correct
    either way, because the edges of destinations [:d] are the first rowptr[d] entries of the by-destination CSR.
    Returns [(n_dst, n_src)] from the first layer to the last.  One host synchronisation per (graph, n_out)."""
    key = (int(n_out), int(n_layers))
    if key not in self._closures:
        if not hasattr(self, "_col_cummax"):
            self._col_cummax = torch.cummax(self.col, 0).values
        d = torch.as_tensor(int(n_out), device=self.col.device)
        need = []
        for _ in range(n_layers):
            s_ = self._col_cummax[(self.rowptr[d] - 1).clamp_min(0)].long() + 1
            need.append(torch.maximum(s_, d))           # self loops make this automatic; kept for graphs built without them
            d = need[-1]
        sizes = [int(v) for v in torch.stack(need).tolist()]
        dsts = [int(n_out)] + sizes[:-1]
        self._closures[key] = list(reversed(list(zip(dsts, sizes))))
    return self._closures[key]


def _graph_order_prefix(self, p):
    """CTA order over the nodes [:p] (cached)."""
    if self.order is None or p >= self.n:
        return self.order
    if p not in self._orders:
        self._orders[p] = _prefix_order(self.order, p)
    return self._orders[p]


CsrGraph.prefix_plan = _graph_prefix_plan
CsrGraph.nbytes = _graph_nbytes
CsrGraph.order_prefix = _graph_order_prefix


def morton_order(pos):
    """Node permutation (int32) along the Z-order curve of the first two coordinate columns, 16 bits per axis."""
    p = pos.detach()[:, :2].to(torch.float64)
    lo = p.min(dim=0).values
    span = (p.max(dim=0).values - lo).clamp_min(1e-300)
    q = ((p - lo) / span * 65535.0).to(torch.int64).clamp_(0, 65535)

    def spread(v):
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        return (v | (v << 1)) & 0x55555555

    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1)
    return torch.argsort(code).to(torch.int32).contiguous()


def _prefix_order(order, p):
    """The members < p of a permutation, in their original relative order (no host synchronisation)."""
    if order is None:
        return None
    first = torch.argsort((order >= p).to(torch.int8), stable=True)[:p]
    return order[first].contiguous()


_GRAPH_CACHE: dict = {}


def graph_for(edge_index, num_nodes, add_self_loops=True, pos=None):
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, num_nodes, add_self_loops, str(edge_index.device))
    hit = _GRAPH_CACHE.get(key)
    # the entry keeps the edge_index tensor alive, so its address cannot be recycled by the caching allocator for another
    # batch's edges while the entry exists; a hit must be the very same tensor object
    if hit is not None and hit[0] is edge_index:
        return hit[1]
    if len(_GRAPH_CACHE) > 64:
        _GRAPH_CACHE.clear()
    g = CsrGraph(edge_index, num_nodes, add_self_loops, pos)
    _GRAPH_CACHE[key] = (edge_index, g)
    return g


def _st(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class _EdgeSoftmaxAggregate(torch.autograd.Function):
    """feat (n_src,H,C), a_src (n_src,H), a_dst (n_dst,H) -> out (n_dst,H,C); n_dst < n_src is the prefix form (the layer's
    destinations are the first n_dst nodes of `graph`, see CsrGraph.prefix_plan).

    feat (n_src,C) - two dimensions - is the SHARED form: every head weighs the same row, out[i,h,:] = sum_j alpha^h_ij feat[j,:]
    (the aggregate-first layers of GATConv); its gradient comes back as (n_src,C), summed over the heads inside the kernel."""

    @staticmethod
    def forward(ctx, feat, a_src, a_dst, graph, slope):
        _lib.require_device()
        if not feat.is_cuda:
            raise RuntimeError("spadot_b200.gat needs CUDA tensors; there is no CPU fallback")
        shared = feat.dim() == 2
        n_src, H = a_src.shape
        C = feat.shape[-1]
        n_dst = a_dst.shape[0]
        if feat.shape[0] != n_src or (not shared and feat.shape[1] != H):
            raise ValueError(f"GAT features {tuple(feat.shape)} do not match attention scalars {tuple(a_src.shape)}")
        if not (0 < n_dst <= n_src <= graph.n):
            raise ValueError(f"GAT layer with {n_dst} destinations and {n_src} sources on a graph of {graph.n} nodes")
        feat, a_src, a_dst = feat.contiguous(), a_src.contiguous(), a_dst.contiguous()
        out = torch.empty((n_dst, H, C), dtype=feat.dtype, device=feat.device)
        alpha = torch.empty((graph.E, H), dtype=feat.dtype, device=feat.device)     # rows [:rowptr[n_dst]] are written
        is_double = int(feat.dtype == torch.float64)
        if feat.dtype not in (torch.float32, torch.float64):
            raise TypeError("GATConv supports float32 and float64")
        order = graph.order_prefix(n_dst)
        _lib.call("sdb_gat_forward_shared" if shared else "sdb_gat_forward", feat.data_ptr(), a_src.data_ptr(), a_dst.data_ptr(),
                  graph.rowptr.data_ptr(), graph.col.data_ptr(), 0 if order is None else order.data_ptr(), n_dst, H, C, float(slope),
                  is_double, out.data_ptr(), alpha.data_ptr(), _st(feat))
        ctx.save_for_backward(feat, a_src, a_dst, alpha)
        ctx.graph, ctx.slope, ctx.is_double, ctx.shared = graph, float(slope), is_double, shared
        return out

    @staticmethod
    def backward(ctx, grad_out):
        feat, a_src, a_dst, alpha = ctx.saved_tensors
        g = ctx.graph
        n_src, H = a_src.shape
        C = feat.shape[-1]
        n_dst = a_dst.shape[0]
        grad_out = grad_out.contiguous()
        dlogit = torch.empty_like(alpha)
        grad_feat = torch.empty_like(feat)
        grad_a_src = torch.empty_like(a_src)
        grad_a_dst = torch.empty_like(a_dst)
        o_dst, o_src = g.order_prefix(n_dst), g.order_prefix(n_src)
        _lib.call("sdb_gat_backward_shared" if ctx.shared else "sdb_gat_backward_prefix", feat.data_ptr(), a_src.data_ptr(),
                  a_dst.data_ptr(), g.rowptr.data_ptr(), g.col.data_ptr(), g.src_rowptr.data_ptr(), g.src_dst.data_ptr(),
                  g.src_eid.data_ptr(), 0 if o_dst is None else o_dst.data_ptr(), 0 if o_src is None else o_src.data_ptr(), n_dst, n_src,
                  H, C, ctx.slope, ctx.is_double, alpha.data_ptr(), grad_out.data_ptr(), dlogit.data_ptr(), grad_feat.data_ptr(),
                  grad_a_src.data_ptr(), grad_a_dst.data_ptr(), _st(feat))
        return grad_feat, grad_a_src, grad_a_dst, None, None


# GATConv runs a prefix layer (fewer destinations than sources) aggregate-first when the linear layer on all source rows would
# cost at least this many flops: below it the layer is launch-bound and the exchange buys nothing (ChickenHeart-sized batches:
# 1e10; SYN-T's second and third layers: 3.4e11 and 1.4e11).
AGGREGATE_FIRST_MIN_FLOPS = 4.0e10


def attention_scalars(h, att_src, att_dst, n_dst):
    """a_src[n,head] = <h[n,head,:], att_src[head,:]> for every node, a_dst likewise for the first n_dst nodes (PyG GATConv:
    `(x * att).sum(-1)`), both roles in ONE pass over h: a GEMM against the (H*C, 2H) block-diagonal arrangement of the two
    attention vectors (the zeros add exactly).  The elementwise form streams the (N, H*C) activations three times per role
    forward and eight times backward; at SYN-T's 40k-node batches those passes were most of the step's non-GEMM device time
    (8.7 -> 4.3 ms).  h (N,H,C); att_* (1,H,C).  Returns views (N,H), (n_dst,H)."""
    N, H, C = h.shape
    att = torch.stack([att_src[0], att_dst[0]], dim=-1)                                    # (H, C, 2)
    block = torch.einsum("hck,hg->hcgk", att, torch.eye(H, dtype=att.dtype, device=att.device))
    a = (h.reshape(N, H * C) @ block.reshape(H * C, 2 * H)).view(N, H, 2)
    return a[:, :, 0], a[:n_dst, :, 1]


class GATConv(nn.Module):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, add_self_loops=True, bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.negative_slope, self.add_self_loops = negative_slope, add_self_loops
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.lin.weight)
        nn.init.xavier_uniform_(self.att_src)
        nn.init.xavier_uniform_(self.att_dst)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, edge_index, pos=None, n_dst=None):
        """`pos` (optional, N x 2 spot coordinates) only orders the CTAs for cache locality; results do not depend on it.

        `n_dst` (optional): produce only output rows [:n_dst]; `x` then holds the rows [:n_src] those destinations read
        (CsrGraph.prefix_plan) and `edge_index` may be the CsrGraph of the whole batch."""
        H, C, N = self.heads, self.out_channels, x.shape[0]
        graph = edge_index if isinstance(edge_index, CsrGraph) else graph_for(edge_index, N, self.add_self_loops, pos)
        n_dst = N if n_dst is None else int(n_dst)
        # prefix layers of a large batch: aggregate first, transform the (fewer) aggregated rows afterwards
        if 2 * n_dst <= N and 2.0 * N * self.in_channels * H * C >= AGGREGATE_FIRST_MIN_FLOPS:
            out = self._aggregate_first(x, graph, n_dst)
        else:
            h = self.lin(x).view(N, H, C)
            a_src, a_dst = attention_scalars(h, self.att_src, self.att_dst, n_dst)
            out = _EdgeSoftmaxAggregate.apply(h, a_src, a_dst, graph, self.negative_slope)
        out = out.reshape(n_dst, H * C) if self.concat else out.mean(dim=1)
        return out + self.bias if self.bias is not None else out

    def _aggregate_first(self, x, graph, n_dst):
        """The same layer with the two linear steps exchanged:  out[i,h,:] = W_h (sum_j alpha^h_ij x_j)  instead of
        sum_j alpha^h_ij (W_h x_j).  The linear layer then runs on the n_dst aggregated rows per head instead of on all N source
        rows - n_dst / N of the GEMM work (forward, dX and dW alike) - and the attention scalars come from the projected
        attention vectors, a_src[j,h] = x_j . (W_h^T att_src[h]), so the (N, H*C) activations are never formed.  The price is
        an aggregation over in_channels instead of out_channels features per head.  Identical in exact arithmetic."""
        H, C, N, F_in = self.heads, self.out_channels, x.shape[0], self.in_channels
        W = self.lin.weight.view(H, C, F_in)
        att = torch.stack([self.att_src[0], self.att_dst[0]], dim=-1)                          # (H, C, 2)
        proj = torch.einsum("hcf,hck->fhk", W, att).reshape(F_in, 2 * H)
        a = (x @ proj).view(N, H, 2)
        z = _EdgeSoftmaxAggregate.apply(x, a[:, :, 0], a[:n_dst, :, 1], graph, self.negative_slope)   # shared rows: (n_dst, H, F_in)
        return torch.einsum("nhf,hcf->nhc", z, W)                                              # (n_dst, H, C)


class GATEncoder(nn.Module):
    """SpaDOT/model/encoder.py:37-61 with the fused GATConv."""

    def __init__(self, input_dim, GAT_z_dim, hidden_dim=512, num_heads=4):
        super().__init__()
        self.gat1 = GATConv(input_dim, hidden_dim, heads=num_heads, concat=True)
        self.gat2 = GATConv(hidden_dim * num_heads, hidden_dim, heads=num_heads, concat=True)
        self.gat3 = GATConv(hidden_dim * num_heads, hidden_dim, heads=num_heads, concat=False)
        self.GAT_fc = nn.Linear(hidden_dim, GAT_z_dim * 2)
        nn.init.xavier_uniform_(self.GAT_fc.weight)

    def forward(self, x, edge_index, pos=None, n_out=None):
        """`n_out`: only rows [:n_out] of the result are wanted (the seeds of a mini-batch, SpaDOT.py:83-84).  Each layer then
        runs on the rows the next one reads and no others: same values for those rows, less work (at SYN-T's 512-seed
        batches the third layer's GEMMs shrink from 40.7k to 16.4k rows and its aggregation from 40.7k to 512)."""
        N = x.shape[0]
        if isinstance(edge_index, CsrGraph) and edge_index.n != N:
            raise ValueError(f"CsrGraph of {edge_index.n} nodes for {N} feature rows")
        if n_out is not None and 0 < int(n_out) < N and x.is_cuda:
            # a prebuilt CsrGraph (graph.TwoHopBatches keeps one per mini-batch across epochs) carries its CTA order and prefix plans
            graph = edge_index if isinstance(edge_index, CsrGraph) else graph_for(edge_index, N, True, pos)
            (d1, s1), (d2, s2), (d3, s3) = graph.prefix_plan(int(n_out), 3)
            h = F.leaky_relu(self.gat1(x[:s1], graph, n_dst=d1))
            h = F.leaky_relu(self.gat2(h[:s2], graph, n_dst=d2))
            h = self.gat3(h[:s3], graph, n_dst=d3)
            GAT_z = self.GAT_fc(h)
            GAT_enc_mu, GAT_enc_logvar = torch.chunk(GAT_z, 2, dim=1)
            return GAT_enc_mu, torch.exp(GAT_enc_logvar)
        h = F.leaky_relu(self.gat1(x, edge_index, pos))
        h = F.leaky_relu(self.gat2(h, edge_index, pos))
        h = self.gat3(h, edge_index, pos)
        GAT_z = self.GAT_fc(h)
        GAT_enc_mu, GAT_enc_logvar = torch.chunk(GAT_z, 2, dim=1)
        return GAT_enc_mu, torch.exp(GAT_enc_logvar)
