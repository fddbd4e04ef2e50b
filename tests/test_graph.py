"""Spatial graph feed: CUDA kNN + edge ordering vs the sklearn-based oracle, and the
deterministic 2-hop batches vs a plain-Python breadth-first restatement."""
import numpy as np
import pytest
import torch

from oracle import graph_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,k", [(300, 6), (1500, 12), (2000, 30), (40, 3)])
def test_spatial_edge_index_matches_sklearn_oracle(n, k):
    from spadot_b200 import graph
    coords = np.random.default_rng(n).uniform(0, 5000, size=(n, 2))
    want = graph_ref.spatial_edge_index(coords, k)
    got = graph.spatial_edge_index(coords, k).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_hex_grid_rings_are_complete():
    """Visium-like hexagonal lattice: k=6 must pick exactly the first ring of an interior spot."""
    from spadot_b200 import graph
    pts = np.array([[i + 0.5 * (j % 2), j * np.sqrt(3) / 2] for j in range(20) for i in range(20)], dtype=np.float64) * 100
    nbr = graph.knn(pts, 6).cpu().numpy()
    centre = 10 * 20 + 10
    d = np.linalg.norm(pts[nbr[centre]] - pts[centre], axis=1)
    assert np.allclose(d, 100.0)


def test_knn_cutoff_rule():
    from spadot_b200 import graph
    assert [graph.knn_cutoff(n) for n in (747, 1966, 1916, 1967, 100000)] == [6, 12, 12, 12, 30]


def test_two_hop_batches_cover_every_needed_edge():
    from spadot_b200 import graph
    n, k = 700, 6
    coords = np.random.default_rng(1).uniform(0, 100, size=(n, 2))
    ei = graph.spatial_edge_index(coords, k)
    src, dst = ei.cpu().numpy()
    into = {i: set(src[dst == i]) for i in range(n)}
    seen_seeds = 0
    for nodes, lei, n_seeds in graph.two_hop_batches(ei, n, batch_size=256):
        nodes = nodes.cpu().numpy()
        assert np.array_equal(nodes[:n_seeds], np.arange(seen_seeds, seen_seeds + n_seeds))
        seen_seeds += n_seeds
        assert len(set(nodes)) == len(nodes)
        hop1 = set(nodes[:n_seeds])
        for s in nodes[:n_seeds]:
            hop1 |= into[s]
        hop2 = set(hop1)
        for s in hop1:
            hop2 |= into[s]
        assert set(nodes) == hop2
        ls, lt = lei.cpu().numpy()
        edges = set(zip(nodes[ls], nodes[lt]))
        want = {(j, i) for i in hop1 for j in into[i]}
        assert edges == want
    assert seen_seeds == n


def test_gat_on_batch_equals_gat_on_full_graph_for_one_layer():
    """One GATConv on a 2-hop batch gives the seeds the same output as on the full graph."""
    from spadot_b200 import gat, graph
    dev = torch.device("cuda:0")
    n, k = 900, 6
    coords = np.random.default_rng(2).uniform(0, 100, size=(n, 2))
    ei = graph.spatial_edge_index(coords, k)
    torch.manual_seed(0)
    conv = gat.GATConv(12, 8, heads=2).double().to(dev)
    x = torch.randn(n, 12, dtype=torch.float64, device=dev)
    full = conv(x, ei)
    for nodes, lei, n_seeds in graph.two_hop_batches(ei, n, batch_size=300):
        sub = conv(x[nodes], lei)
        assert float((sub[:n_seeds] - full[nodes[:n_seeds]]).abs().max()) < 1e-12
