"""GAT edge-softmax message passing (K2/K2b).  Oracle: oracle/gat_ref.py (PyG semantics; parity
unpinned by the reference — torch_geometric is un-vendored).  Forward and every gradient are compared
in float64 (tight) and float32 (fp32 tolerance)."""
import numpy as np
import pytest
import torch

from oracle import gat_ref, graph_ref


def make_graph(n, seed=0):
    rng = np.random.default_rng(seed)
    coords = rng.uniform(0, 30, size=(n, 2))
    k = max(2, graph_ref.knn_cutoff(n)) if n >= 1000 else 6
    return torch.from_numpy(graph_ref.spatial_edge_index(coords, k))


def test_graph_oracle_shape_and_self_loops():
    ei = make_graph(200).numpy()
    assert ei.shape[0] == 2 and ei.shape[1] == 200 * 6 + 200
    assert np.all(np.bincount(ei[0], minlength=200) == 7)          # every node sends k edges + its self loop
    assert np.all((ei[0] == ei[1]).sum() == 200)
    # dense_to_sparse order: sorted by source, then target
    assert np.all(np.diff(ei[0]) >= 0)


def test_oracle_softmax_rows_sum_to_one_and_concat_mean_shapes():
    torch.manual_seed(0)
    ei = make_graph(50)
    conv = gat_ref.GATConvRef(7, 5, heads=3, concat=True).double()
    x = torch.randn(50, 7, dtype=torch.float64)
    assert conv(x, ei).shape == (50, 15)
    conv2 = gat_ref.GATConvRef(7, 5, heads=3, concat=False).double()
    assert conv2(x, ei).shape == (50, 5)
    assert set(conv.state_dict().keys()) == {"att_src", "att_dst", "bias", "lin.weight"}


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-11), (torch.float32, 2e-4)])
@pytest.mark.parametrize("n,fin,C,H,concat", [(300, 40, 16, 4, True), (257, 33, 24, 3, False), (1200, 64, 512, 4, True)])
def test_gatconv_forward_backward_match_oracle(dtype, tol, n, fin, C, H, concat):
    from spadot_b200 import gat
    dev = torch.device("cuda:0")
    torch.manual_seed(n)
    ei = make_graph(n, seed=n)
    ref = gat_ref.GATConvRef(fin, C, heads=H, concat=concat).to(dtype)
    mine = gat.GATConv(fin, C, heads=H, concat=concat).to(dtype).to(dev)
    assert set(mine.state_dict().keys()) == set(ref.state_dict().keys())
    mine.load_state_dict({k: v.to(dev) for k, v in ref.state_dict().items()})
    with torch.no_grad():
        ref.bias.normal_()
        mine.bias.copy_(ref.bias.to(dev))
    x = torch.randn(n, fin, dtype=dtype)
    xr = x.clone().requires_grad_(True)
    xm = x.clone().to(dev).requires_grad_(True)
    w = torch.randn(n, C * H if concat else C, dtype=dtype)
    out_r = ref(xr, ei)
    out_m = mine(xm, ei.to(dev))
    scale = float(out_r.abs().max())
    assert float((out_m.cpu() - out_r).abs().max()) < tol * max(scale, 1.0)
    (out_r * w).sum().backward()
    (out_m * w.to(dev)).sum().backward()
    for (name, pr), (_, pm) in zip(ref.named_parameters(), mine.named_parameters()):
        gs = float(pr.grad.abs().max())
        assert float((pm.grad.cpu() - pr.grad).abs().max()) < tol * 50 * max(gs, 1.0), name
    assert float((xm.grad.cpu() - xr.grad).abs().max()) < tol * 50 * max(float(xr.grad.abs().max()), 1.0)


@pytest.mark.gpu
def test_gat_encoder_matches_oracle_and_loads_reference_style_state_dict():
    from spadot_b200 import gat
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    n = 400
    ei = make_graph(n, seed=3)
    ref = gat_ref.GATEncoderRef(60, 10, hidden_dim=32, num_heads=4).double()
    mine = gat.GATEncoder(60, 10, hidden_dim=32, num_heads=4).double().to(dev)
    mine.load_state_dict(ref.state_dict())          # same keys: gatN.lin.weight, gatN.att_src, gatN.att_dst, gatN.bias, GAT_fc.*
    x = torch.randn(n, 60, dtype=torch.float64)
    mu_r, var_r = ref(x, ei)
    mu_m, var_m = mine(x.to(dev), ei.to(dev))
    np.testing.assert_allclose(mu_m.detach().cpu().numpy(), mu_r.detach().numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var_m.detach().cpu().numpy(), var_r.detach().numpy(), rtol=1e-9, atol=1e-11)


@pytest.mark.gpu
def test_gat_isolated_and_high_degree_nodes():
    """A hub with > 32 incoming edges (several warp strides) and a node whose only edge is its self loop."""
    from spadot_b200 import gat
    dev = torch.device("cuda:0")
    n = 120
    src = torch.arange(1, 100)
    ei = torch.stack([src, torch.zeros_like(src)])          # 99 edges into node 0; nodes 100..119 isolated
    torch.manual_seed(0)
    ref = gat_ref.GATConvRef(9, 8, heads=2).double()
    mine = gat.GATConv(9, 8, heads=2).double().to(dev)
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(n, 9, dtype=torch.float64)
    np.testing.assert_allclose(mine(x.to(dev), ei.to(dev)).detach().cpu().numpy(), ref(x, ei).detach().numpy(), rtol=1e-10, atol=1e-12)


@pytest.mark.gpu
def test_gat_large_graph_with_locality_order_matches_oracle():
    """n >= 20000 switches on the reverse Cuthill-McKee CTA order; results must not depend on it."""
    from spadot_b200 import gat
    dev = torch.device("cuda:0")
    n = 21000
    ei = make_graph(n, seed=4)
    torch.manual_seed(2)
    ref = gat_ref.GATConvRef(16, 8, heads=2).double()
    mine = gat.GATConv(16, 8, heads=2).double().to(dev)
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(n, 16, dtype=torch.float64)
    xr, xm = x.clone().requires_grad_(True), x.clone().to(dev).requires_grad_(True)
    out_r, out_m = ref(xr, ei), mine(xm, ei.to(dev))
    assert gat.graph_for(ei.to(dev), n).order is not None
    np.testing.assert_allclose(out_m.detach().cpu().numpy(), out_r.detach().numpy(), rtol=1e-10, atol=1e-12)
    out_r.sum().backward()
    out_m.sum().backward()
    np.testing.assert_allclose(xm.grad.cpu().numpy(), xr.grad.numpy(), rtol=1e-9, atol=1e-12)


def test_morton_order_is_a_permutation_with_locality():
    """The coordinate hint's Z-curve order: a permutation, and consecutive nodes are spatial neighbours."""
    from spadot_b200 import gat
    torch.manual_seed(0)
    pos = torch.rand(5000, 2, dtype=torch.float64)
    order = gat.morton_order(pos).long()
    assert sorted(order.tolist()) == list(range(5000))
    hop = (pos[order][1:] - pos[order][:-1]).norm(dim=1).mean()
    assert hop < 0.2 * (pos[1:] - pos[:-1]).norm(dim=1).mean()
    assert gat.morton_order(torch.zeros(7, 2)).numel() == 7                   # degenerate extent


@pytest.mark.gpu
def test_gat_coordinate_hint_orders_ctas_on_device_and_keeps_results():
    """With `pos` the CTA order is the device-side Z-curve (no host round trip); outputs and gradients equal the unhinted run."""
    from spadot_b200 import gat
    dev = torch.device("cuda:0")
    n = 21000
    ei = make_graph(n, seed=5).to(dev)
    torch.manual_seed(3)
    conv = gat.GATConv(16, 8, heads=2).double().to(dev)
    pos = torch.rand(n, 2, dtype=torch.float64, device=dev)
    x = torch.randn(n, 16, dtype=torch.float64, device=dev)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ei_a, ei_b = ei.clone(), ei.clone()
    out_a, out_b = conv(xa, ei_a), conv(xb, ei_b, pos)
    ga, gb = gat.graph_for(ei_a, n), gat.graph_for(ei_b, n)
    assert ga.order is not None and gb.order is not None and not torch.equal(ga.order, gb.order)
    assert torch.equal(gb.order, gat.morton_order(pos))
    # the tile form groups eight consecutive nodes of the order, so the order changes which rows are summed together:
    # equal to rounding, not bitwise
    np.testing.assert_allclose(out_b.detach().cpu().numpy(), out_a.detach().cpu().numpy(), rtol=1e-12, atol=1e-14)
    out_a.square().sum().backward()
    out_b.square().sum().backward()
    np.testing.assert_allclose(xb.grad.cpu().numpy(), xa.grad.cpu().numpy(), rtol=1e-12, atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("agg_first", [False, True])
@pytest.mark.parametrize("hop_ordered", [True, False])
def test_gat_encoder_seed_rows_only_equals_full_run(hop_ordered, agg_first, monkeypatch):
    """GATEncoder(n_out=seeds) runs every layer on the rows the next one reads (prefix form of the kernels): the seeds' outputs
    and every parameter gradient equal the full run's; with hop-ordered nodes the prefixes are the hop sets, with arbitrary
    numbering they degrade to (correct) supersets."""
    from spadot_b200 import gat, graph
    # agg_first: prefix layers with at most half as many destinations as sources run aggregate-first (GATConv._aggregate_first;
    # by default only for large batches) - same outputs and gradients as the full run
    monkeypatch.setattr(gat, "AGGREGATE_FIRST_MIN_FLOPS", 0.0 if agg_first else float("inf"))
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    n_all, bs = 6000, 64
    coords = rng.uniform(0, 100, size=(n_all, 2))
    ei_all = graph.spatial_edge_index(coords, 8, device=dev)
    nodes, lei, ns = next(iter(graph.two_hop_batches(ei_all, n_all, batch_size=bs)))
    n = nodes.numel()
    if not hop_ordered:                                  # scramble the non-seed numbering
        perm = torch.cat([torch.arange(ns), ns + torch.from_numpy(rng.permutation(n - ns))]).to(dev)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(n, device=dev)
        lei = inv[lei]
    torch.manual_seed(5)
    enc = gat.GATEncoder(40, 6, hidden_dim=16, num_heads=4).double().to(dev)
    x = torch.randn(n, 40, dtype=torch.float64, device=dev)
    g = gat.graph_for(lei, n, True)
    plan = g.prefix_plan(ns, 3)
    assert plan[-1][0] == ns and all(d <= s for d, s in plan) and plan[0][1] <= n
    if hop_ordered:
        assert plan[2][1] < n // 2 and plan[1][1] == n          # 1-hop prefix for the last layer, everything below it
    res = []
    for n_out in (None, ns):
        enc.zero_grad()
        mu, var = enc(x, lei, n_out=n_out)
        ((mu[:ns] ** 2).sum() + var[:ns].sum()).backward()
        res.append((mu[:ns].detach().clone(), var[:ns].detach().clone(), [p.grad.clone() for p in enc.parameters()]))
    (mu_a, var_a, ga), (mu_b, var_b, gb) = res
    assert mu_b.shape == (ns, 6)
    np.testing.assert_allclose(mu_b.cpu().numpy(), mu_a.cpu().numpy(), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(var_b.cpu().numpy(), var_a.cpu().numpy(), rtol=1e-11, atol=1e-13)
    for a, b in zip(ga, gb):
        scale = float(a.abs().max()) + 1e-30
        assert float((a - b).abs().max()) <= 1e-11 * scale


def test_prefix_plan_is_the_hop_closure():
    """CsrGraph.prefix_plan: for every layer, the source prefix contains every in-neighbour of the destination prefix (brute
    force over the edge list), the last layer's destinations are the seeds, and each layer's destinations are the next layer's
    sources; on a hop-ordered batch the prefixes are exactly the hop sets."""
    from spadot_b200 import gat, graph
    rng = np.random.default_rng(3)
    coords = rng.uniform(0, 50, size=(1500, 2))
    from oracle import graph_ref
    ei_all = torch.as_tensor(np.asarray(graph_ref.spatial_edge_index(coords, 6)), dtype=torch.long)     # (the product's kNN needs the GPU)
    nodes, lei, ns = next(iter(graph.two_hop_batches(ei_all, 1500, batch_size=32)))
    n = int(nodes.numel())
    g = gat.CsrGraph(lei, n, add_self_loops=True)
    plan = g.prefix_plan(ns, 3)
    assert plan[-1][0] == ns
    src, dst = lei[0].numpy(), lei[1].numpy()
    for (d, s_), nxt in zip(plan, plan[1:] + [None]):
        need = int(max(src[dst < d].max() + 1, d))
        assert s_ == need                                           # the smallest prefix that holds every source
        if nxt is not None:
            assert nxt[1] == d                                      # this layer produces exactly what the next one reads
    hop1 = int(np.unique(np.concatenate([src[dst < ns], np.arange(ns)])).size)
    assert plan[2] == (ns, hop1) and plan[1] == (hop1, n) and plan[0] == (n, n)
    assert g.prefix_plan(ns, 3) is plan                             # cached


def _random_graph(kind, n, rng):
    """Edge lists that stress the tile kernels: (a) spatial kNN (shared neighbours: the fast path), (b) uniformly random pairs
    (no sharing: tiles overflow their distinct-row budget and fall back, next to tiles that fit), (c) hubs, isolated nodes and
    parallel edges."""
    if kind == "knn":
        coords = rng.uniform(0, 30, size=(n, 2))
        return torch.from_numpy(graph_ref.spatial_edge_index(coords, 12))
    if kind == "random":
        return torch.from_numpy(rng.integers(0, n, size=(2, 14 * n)))
    src = np.concatenate([np.arange(1, 100), np.arange(0, 700), rng.integers(0, n, 300), rng.integers(0, n, 300)])
    dst = np.concatenate([np.zeros(99, dtype=np.int64), np.full(700, 5), rng.integers(0, n // 2, 300), rng.integers(0, n // 2, 300)])
    src, dst = np.concatenate([src, src[-200:]]), np.concatenate([dst, dst[-200:]])          # parallel edges
    return torch.from_numpy(np.stack([src, dst]))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
@pytest.mark.parametrize("kind,n,H,C,n_dst", [("knn", 1203, 4, 512, None), ("knn", 1203, 4, 512, 300), ("knn", 500, 3, 24, None),
                                              ("knn", 500, 2, 25, 123), ("random", 900, 4, 128, None), ("random", 900, 4, 64, 400),
                                              ("hubs", 800, 4, 256, None), ("hubs", 800, 1, 7, None), ("knn", 40, 5, 16, None),
                                              ("knn", 500, 2, 2048, None), ("knn", 300, 1, 4096, 100)])   # heads wider than a CTA sweep
def test_gat_tile_kernels_equal_the_per_node_kernels(dtype, tol, kind, n, H, C, n_dst):
    """The tile form (8 nodes per CTA, de-duplicated gathers, dense weights; by-destination backward on the butterfly) against
    the per-node form of the same library (SDB_GAT_TILES=0), forward and all three gradients, full and prefix layers."""
    import os
    from spadot_b200 import gat
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(n + H + C)
    ei = _random_graph(kind, n, rng).to(dev)
    g = gat.CsrGraph(ei, n, add_self_loops=(kind != "hubs"))
    if n_dst is not None:                                    # prefix layer: sources = everything the first n_dst destinations read
        n_src = int(g.col[: int(g.rowptr[n_dst])].max()) + 1 if int(g.rowptr[n_dst]) else n_dst
        n_src = max(n_src, n_dst)
    else:
        n_dst = n_src = n
    torch.manual_seed(0)
    feat = torch.randn(n_src, H, C, dtype=dtype, device=dev)
    a_s = torch.randn(n_src, H, dtype=dtype, device=dev)
    a_d = torch.randn(n_dst, H, dtype=dtype, device=dev)
    go = torch.randn(n_dst, H, C, dtype=dtype, device=dev)
    res = {}
    try:
        for mode in ("0", "1"):
            os.environ["SDB_GAT_TILES"] = mode
            f, s, d = (t.clone().requires_grad_(True) for t in (feat, a_s, a_d))
            out = gat._EdgeSoftmaxAggregate.apply(f, s, d, g, 0.2)
            out.backward(go)
            torch.cuda.synchronize()
            res[mode] = dict(out=out.detach(), grad_feat=f.grad, grad_a_src=s.grad, grad_a_dst=d.grad)
    finally:
        os.environ.pop("SDB_GAT_TILES", None)
    for key in res["0"]:
        a, b = res["0"][key], res["1"][key]
        assert torch.isfinite(b).all(), key
        scale = max(float(a.abs().max()), 1.0)
        assert float((a - b).abs().max()) <= tol * scale, (key, float((a - b).abs().max()), scale)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-13), (torch.float32, 1e-5)])
def test_attention_scalars_one_gemm_equals_the_elementwise_form(dtype, tol):
    """gat.attention_scalars (block-diagonal GEMM, one pass over the activations) against PyG's (x * att).sum(-1), values
    and every gradient, with fewer destinations than sources (prefix layers)."""
    from spadot_b200 import gat
    torch.manual_seed(0)
    N, H, C, n_dst = 57, 4, 24, 31
    h = torch.randn(N, H, C, dtype=dtype, requires_grad=True)
    a_s = torch.randn(1, H, C, dtype=dtype, requires_grad=True)
    a_d = torch.randn(1, H, C, dtype=dtype, requires_grad=True)
    w1, w2 = torch.randn(N, H, dtype=dtype), torch.randn(n_dst, H, dtype=dtype)
    want = ((h * a_s).sum(-1), (h[:n_dst] * a_d).sum(-1))
    got = gat.attention_scalars(h, a_s, a_d, n_dst)
    assert got[0].shape == (N, H) and got[1].shape == (n_dst, H)
    for g, w in zip(got, want):
        assert float((g - w).abs().max()) <= tol * max(1.0, float(w.abs().max()))
    gw = torch.autograd.grad((want[0] * w1).sum() + (want[1] * w2).sum(), [h, a_s, a_d])
    gg = torch.autograd.grad((got[0] * w1).sum() + (got[1] * w2).sum(), [h, a_s, a_d])
    for g, w in zip(gg, gw):
        assert float((g - w).abs().max()) <= 10 * tol * max(1.0, float(w.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 5e-5)])
@pytest.mark.parametrize("concat", [True, False])
def test_gatconv_aggregate_first_equals_transform_first(dtype, tol, concat, monkeypatch):
    """A prefix layer (300 destinations of 1203 sources) computed as W_h (sum_j alpha x_j) and as sum_j alpha (W_h x_j): same
    output, same gradients of the input and of every parameter."""
    from spadot_b200 import gat
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(7)
    n, n_dst, f_in, C, H = 1203, 300, 48, 32, 4
    ei = _random_graph("knn", n, rng).to(dev)
    g = gat.CsrGraph(ei, n, add_self_loops=True)
    n_src = max(int(g.col[: int(g.rowptr[n_dst])].max()) + 1, n_dst)
    torch.manual_seed(1)
    conv = gat.GATConv(f_in, C, heads=H, concat=concat).to(dtype).to(dev)
    with torch.no_grad():
        conv.bias.normal_()
    x = torch.randn(n_src, f_in, dtype=dtype, device=dev)
    w = torch.randn(n_dst, H * C if concat else C, dtype=dtype, device=dev)
    res = []
    for min_flops in (float("inf"), 0.0):
        monkeypatch.setattr(gat, "AGGREGATE_FIRST_MIN_FLOPS", min_flops)
        conv.zero_grad()
        xi = x.clone().requires_grad_(True)
        out = conv(xi, g, n_dst=n_dst)
        (out * w).sum().backward()
        res.append([out.detach(), xi.grad] + [p.grad.clone() for p in conv.parameters()])
    assert 2 * n_dst <= n_src
    for a, b in zip(*res):
        scale = max(float(a.abs().max()), 1.0)
        assert float((a - b).abs().max()) <= tol * scale


@pytest.mark.gpu
@pytest.mark.parametrize("tiles", ["0", "1"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
@pytest.mark.parametrize("kind,n,H,C,n_dst", [("knn", 1203, 4, 512, 300), ("knn", 1203, 4, 2048, 200), ("knn", 500, 3, 24, None),
                                              ("knn", 500, 2, 25, 123), ("random", 900, 4, 128, 400), ("hubs", 800, 4, 256, None),
                                              ("hubs", 800, 1, 7, None), ("knn", 300, 5, 16, 100)])
def test_gat_shared_feature_form_equals_the_expanded_features(tiles, dtype, tol, kind, n, H, C, n_dst):
    """sdb_gat_forward_shared / sdb_gat_backward_shared (rows (n, C) read by every head, head sum of the feature gradient inside
    the kernel) against the per-head kernels fed with the same row copied H times - in the tile and in the per-node form."""
    import os
    from spadot_b200 import gat
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(n + H + C + 1)
    ei = _random_graph(kind, n, rng).to(dev)
    g = gat.CsrGraph(ei, n, add_self_loops=(kind != "hubs"))
    if n_dst is not None:
        n_src = max(int(g.col[: int(g.rowptr[n_dst])].max()) + 1 if int(g.rowptr[n_dst]) else n_dst, n_dst)
    else:
        n_dst = n_src = n
    torch.manual_seed(0)
    x = torch.randn(n_src, C, dtype=dtype, device=dev)
    a_s = torch.randn(n_src, H, dtype=dtype, device=dev)
    a_d = torch.randn(n_dst, H, dtype=dtype, device=dev)
    go = torch.randn(n_dst, H, C, dtype=dtype, device=dev)
    res = []
    os.environ["SDB_GAT_TILES"] = tiles
    try:
        for shared in (False, True):
            xi, s, d = (t.clone().requires_grad_(True) for t in (x, a_s, a_d))
            feat = xi if shared else xi.unsqueeze(1).expand(n_src, H, C)
            out = gat._EdgeSoftmaxAggregate.apply(feat, s, d, g, 0.2)
            out.backward(go)
            torch.cuda.synchronize()
            res.append(dict(out=out.detach(), grad_x=xi.grad, grad_a_src=s.grad, grad_a_dst=d.grad))
    finally:
        os.environ.pop("SDB_GAT_TILES", None)
    assert res[1]["grad_x"].shape == (n_src, C)
    for key in res[0]:
        a, b = res[0][key], res[1][key]
        assert torch.isfinite(b).all(), key
        scale = max(float(a.abs().max()), 1.0)
        assert float((a - b).abs().max()) <= tol * scale, (key, float((a - b).abs().max()), scale)


class _DenseAggregateStandIn:
    """CPU stand-in for gat._EdgeSoftmaxAggregate (plain torch, autograd through it): softmax over each destination's incoming
    edges of leaky_relu(a_src[j] + a_dst[i]), then the weighted sum of per-head (n,H,C) or shared (n,C) rows."""

    @staticmethod
    def apply(feat, a_src, a_dst, graph, slope):
        n_src, n_dst, H = a_src.shape[0], a_dst.shape[0], a_src.shape[1]
        E = int(graph.rowptr[n_dst])
        src = graph.col[:E].long()
        dst = torch.repeat_interleave(torch.arange(n_dst), (graph.rowptr[1:n_dst + 1] - graph.rowptr[:n_dst]))
        logit = torch.nn.functional.leaky_relu(a_src[src] + a_dst[dst], slope)                       # (E, H)
        mask = torch.zeros(n_dst, n_src, dtype=torch.bool)
        mask[dst, src] = True
        dense = torch.full((n_dst, n_src, H), float("-inf"), dtype=feat.dtype).index_put((dst, src), logit)
        alpha = torch.softmax(dense, dim=1).masked_fill(~mask[:, :, None], 0.0)
        rows = feat if feat.dim() == 3 else feat.unsqueeze(1).expand(n_src, H, feat.shape[-1])
        return torch.einsum("inh,nhc->ihc", alpha, rows)


@pytest.mark.parametrize("concat", [True, False])
def test_gatconv_host_logic_aggregate_first_on_cpu(concat, monkeypatch):
    """The host side of GATConv without the CUDA kernels: with a dense torch stand-in for the aggregation, the aggregate-first
    form (projected attention vectors, shared rows, per-head linear map afterwards) equals the transform-first form in output and
    in every gradient, and the rule that picks it looks at destinations / sources and at the GEMM size."""
    from spadot_b200 import gat
    monkeypatch.setattr(gat, "_EdgeSoftmaxAggregate", _DenseAggregateStandIn)
    rng = np.random.default_rng(2)
    n, n_dst, f_in, C, H = 90, 30, 11, 6, 3
    coords = rng.uniform(0, 10, size=(n, 2))
    g = gat.CsrGraph(torch.from_numpy(graph_ref.spatial_edge_index(coords, 5)), n, add_self_loops=True)
    n_src = max(int(g.col[: int(g.rowptr[n_dst])].max()) + 1, n_dst)
    assert 2 * n_dst <= n_src
    torch.manual_seed(3)
    conv = gat.GATConv(f_in, C, heads=H, concat=concat).double()
    with torch.no_grad():
        conv.bias.normal_()
    x = torch.randn(n_src, f_in, dtype=torch.float64)
    w = torch.randn(n_dst, H * C if concat else C, dtype=torch.float64)
    res, used = [], []
    real = gat.GATConv._aggregate_first
    monkeypatch.setattr(gat.GATConv, "_aggregate_first", lambda self, *a: (used.append(True), real(self, *a))[1])
    for min_flops in (float("inf"), 0.0):
        monkeypatch.setattr(gat, "AGGREGATE_FIRST_MIN_FLOPS", min_flops)
        conv.zero_grad()
        xi = x.clone().requires_grad_(True)
        out = conv(xi, g, n_dst=n_dst)
        (out * w).sum().backward()
        res.append([out.detach(), xi.grad] + [p.grad.clone() for p in conv.parameters()])
    assert used == [True]                                   # only the second run took the exchanged form
    for a, b in zip(*res):
        assert float((a - b).abs().max()) <= 1e-12 * max(1.0, float(a.abs().max()))
    used.clear()
    conv(torch.randn(n, f_in, dtype=torch.float64), g)       # a full layer (as many destinations as sources): never exchanged
    assert used == []


def test_two_hop_batches_loader_samples_once_and_keeps_the_graphs():
    """graph.TwoHopBatches: same node sets and induced edges as two_hop_batches, a prepared CsrGraph per batch, the very same
    objects on the second epoch; an interrupted first pass or a budget that is too small leaves nothing cached."""
    from spadot_b200 import gat, graph
    rng = np.random.default_rng(8)
    n = 700
    coords = rng.uniform(0, 40, size=(n, 2))
    ei = torch.as_tensor(np.asarray(graph_ref.spatial_edge_index(coords, 6)), dtype=torch.long)
    plain = list(graph.two_hop_batches(ei, n, batch_size=128))
    loader = graph.TwoHopBatches(ei, n, batch_size=128, pos=torch.from_numpy(coords))
    assert len(loader) == len(plain) == 6 and not loader.cached
    first = list(loader)
    assert loader.cached and loader.cached_bytes > 0
    for (nodes_a, lei, ns_a), (nodes_b, g, ns_b) in zip(plain, first):
        assert ns_a == ns_b and torch.equal(nodes_a, nodes_b) and isinstance(g, gat.CsrGraph) and g.n == nodes_a.numel()
        # the CSR holds the induced edges plus one self loop per node (existing self loops replaced)
        src, dst = lei[0], lei[1]
        keep = src != dst
        want = set(zip(src[keep].tolist(), dst[keep].tolist())) | {(i, i) for i in range(g.n)}
        dst_of_edge = torch.repeat_interleave(torch.arange(g.n), g.rowptr[1:] - g.rowptr[:-1])
        assert set(zip(g.col.tolist(), dst_of_edge.tolist())) == want
    second = list(loader)
    assert all(a[1] is b[1] and a[0] is b[0] for a, b in zip(first, second))          # nothing re-sampled, nothing rebuilt
    partial = graph.TwoHopBatches(ei, n, batch_size=128)
    next(iter(partial))
    assert not partial.cached
    small = graph.TwoHopBatches(ei, n, batch_size=128, max_cache_bytes=1000)
    assert len(list(small)) == 6 and not small.cached and len(list(small)) == 6


def test_gat_encoder_takes_a_prebuilt_graph_on_cpu(monkeypatch):
    """GATEncoder / GATConv accept a gat.CsrGraph wherever they accept an edge_index (what graph.TwoHopBatches hands out); with
    the dense stand-in for the kernels the two calls agree, and a graph of the wrong size is refused."""
    from spadot_b200 import gat
    monkeypatch.setattr(gat, "_EdgeSoftmaxAggregate", _DenseAggregateStandIn)
    rng = np.random.default_rng(4)
    n = 80
    ei = torch.from_numpy(graph_ref.spatial_edge_index(rng.uniform(0, 10, size=(n, 2)), 5))
    torch.manual_seed(9)
    enc = gat.GATEncoder(13, 4, hidden_dim=8, num_heads=2).double()
    x = torch.randn(n, 13, dtype=torch.float64)
    mu_a, var_a = enc(x, ei)
    mu_b, var_b = enc(x, gat.CsrGraph(ei, n, add_self_loops=True))
    np.testing.assert_allclose(mu_b.detach().numpy(), mu_a.detach().numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(var_b.detach().numpy(), var_a.detach().numpy(), rtol=1e-12, atol=1e-14)
    with pytest.raises(ValueError):
        enc(x[:50], gat.CsrGraph(ei, n, add_self_loops=True))


@pytest.mark.gpu
def test_loader_batches_with_prebuilt_graphs_train_like_plain_batches():
    """graph.TwoHopBatches on the device: the encoder fed with a kept (nodes, CsrGraph, n_seeds) batch returns what it returns for
    the plain (nodes, edge_index, n_seeds) batch, forward and parameter gradients, on the first and on the second epoch."""
    from spadot_b200 import gat, graph
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(21)
    n_all = 3000
    coords = torch.from_numpy(rng.uniform(0, 60, size=(n_all, 2))).to(dev)
    ei_all = graph.spatial_edge_index(coords.cpu().numpy(), 8, device=dev)
    loader = graph.TwoHopBatches(ei_all, n_all, batch_size=256, pos=coords)
    plain = list(graph.two_hop_batches(ei_all, n_all, batch_size=256))
    torch.manual_seed(2)
    enc = gat.GATEncoder(20, 4, hidden_dim=16, num_heads=4).double().to(dev)
    feats = torch.randn(n_all, 20, dtype=torch.float64, device=dev)

    def run(x, g, ns):
        enc.zero_grad()
        mu, var = enc(x, g, pos=None, n_out=ns)
        ((mu ** 2).sum() + var.sum()).backward()
        return mu.detach().clone(), [p.grad.clone() for p in enc.parameters()]

    for epoch in range(2):
        for (nodes_a, lei, ns_a), (nodes_b, g, ns_b) in zip(plain[:3], loader):
            assert torch.equal(nodes_a, nodes_b) and ns_a == ns_b
            mu_a, ga = run(feats[nodes_a], lei, ns_a)
            mu_b, gb = run(feats[nodes_b], g, ns_b)
            np.testing.assert_allclose(mu_b.cpu().numpy(), mu_a.cpu().numpy(), rtol=1e-11, atol=1e-13)
            for a, b in zip(ga, gb):
                assert float((a - b).abs().max()) <= 1e-11 * (float(a.abs().max()) + 1e-30)
    assert not loader.cached                  # both epochs were cut short after three batches: nothing is kept
    assert len(list(loader)) == len(plain) and loader.cached
