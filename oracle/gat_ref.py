"""Plain-torch restatement of torch_geometric.nn.GATConv (PyG 2.6 semantics).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: torch_geometric is neither vendored under /root/reference nor installed in this
image, and no reference fixture pins GAT outputs.  The restatement follows the published PyG 2.6
GATConv.forward / edge_update / message and torch_geometric.utils.softmax as summarised in SURVEY.md
§8 a18, anchored on the reference's call sites SpaDOT/model/encoder.py:41-46 (constructor arguments,
`.lin.weight` re-initialisation => PyG >= 2.5 parameter naming) and :56-58 (forward).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def with_self_loops(edge_index, num_nodes):
    """remove_self_loops + add_self_loops (GATConv.forward with add_self_loops=True)."""
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    loops = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.stack([torch.cat([src[keep], loops]), torch.cat([dst[keep], loops])])


class GATConvRef(nn.Module):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, add_self_loops=True, bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.negative_slope, self.add_self_loops = negative_slope, add_self_loops
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels)) if bias else None
        nn.init.xavier_uniform_(self.lin.weight)          # PyG glorot
        nn.init.xavier_uniform_(self.att_src)
        nn.init.xavier_uniform_(self.att_dst)

    def forward(self, x, edge_index):
        H, C, N = self.heads, self.out_channels, x.shape[0]
        h = self.lin(x).view(N, H, C)
        a_src = (h * self.att_src).sum(-1)
        a_dst = (h * self.att_dst).sum(-1)
        ei = with_self_loops(edge_index, N) if self.add_self_loops else edge_index
        j, i = ei[0], ei[1]                                 # messages flow source j -> target i
        e = F.leaky_relu(a_src[j] + a_dst[i], self.negative_slope)
        mx = torch.full((N, H), float("-inf"), dtype=e.dtype, device=e.device).scatter_reduce(0, i[:, None].expand(-1, H), e, "amax")
        ex = torch.exp(e - mx[i])
        den = torch.zeros((N, H), dtype=e.dtype, device=e.device).index_add_(0, i, ex)
        alpha = ex / (den[i] + 1e-16)                       # torch_geometric.utils.softmax
        out = torch.zeros((N, H, C), dtype=h.dtype, device=h.device).index_add_(0, i, alpha.unsqueeze(-1) * h[j])
        out = out.view(N, H * C) if self.concat else out.mean(dim=1)
        return out + self.bias if self.bias is not None else out


class GATEncoderRef(nn.Module):
    """SpaDOT/model/encoder.py:37-61 on top of GATConvRef."""

    def __init__(self, input_dim, GAT_z_dim, hidden_dim=512, num_heads=4):
        super().__init__()
        self.gat1 = GATConvRef(input_dim, hidden_dim, heads=num_heads, concat=True)
        self.gat2 = GATConvRef(hidden_dim * num_heads, hidden_dim, heads=num_heads, concat=True)
        self.gat3 = GATConvRef(hidden_dim * num_heads, hidden_dim, heads=num_heads, concat=False)
        self.GAT_fc = nn.Linear(hidden_dim, GAT_z_dim * 2)
        nn.init.xavier_uniform_(self.GAT_fc.weight)

    def forward(self, x, edge_index):
        h = F.leaky_relu(self.gat1(x, edge_index))
        h = F.leaky_relu(self.gat2(h, edge_index))
        h = self.gat3(h, edge_index)
        mu, logvar = torch.chunk(self.GAT_fc(h), 2, dim=1)
        return mu, torch.exp(logvar)
