"""K7: device k-means must reproduce sklearn.cluster.KMeans (the reference's call,
utils/_train_utils.py:262-266 / utils/_analyze_utils.py:33-35: n_init=10, random_state=1993)."""
import numpy as np
import pytest
from sklearn.cluster import KMeans as SkKMeans

pytestmark = pytest.mark.gpu


def latent(n, d, k, seed, sd=0.5):
    rng = np.random.default_rng(seed)
    centres = rng.normal(0, 1.5, size=(k, d))
    return centres[rng.integers(0, k, n)] + rng.normal(0, sd, size=(n, d))


@pytest.mark.parametrize("n,d,k,seed", [(747, 20, 10, 1), (1966, 20, 10, 2), (3000, 32, 7, 3), (500, 20, 12, 4)])
def test_matches_sklearn_clustering(n, d, k, seed):
    from spadot_b200.kmeans import KMeans
    X = latent(n, d, 10, seed, sd=0.9)
    want = SkKMeans(n_clusters=k, random_state=1993, n_init=10).fit(X)
    got = KMeans(n_clusters=k, random_state=1993, n_init=10).fit(X)
    assert got.inertia_ == pytest.approx(want.inertia_, rel=1e-10)
    assert np.array_equal(got.labels_, want.labels_)
    np.testing.assert_allclose(got.cluster_centers_, want.cluster_centers_, rtol=1e-9, atol=1e-11)


def test_overlapping_blobs_same_inertia():
    """Heavily overlapping data (many near-ties): the optimum found must be sklearn's (inertia), labels may
    differ only on exact ties."""
    from spadot_b200.kmeans import KMeans
    X = np.random.default_rng(5).normal(0, 1, size=(4000, 16))
    want = SkKMeans(n_clusters=9, random_state=7, n_init=5).fit(X)
    got = KMeans(n_clusters=9, random_state=7, n_init=5).fit(X)
    assert got.inertia_ == pytest.approx(want.inertia_, rel=1e-9)
    assert (got.labels_ == want.labels_).mean() > 0.999


def test_duplicate_points_trigger_empty_cluster_relocation():
    from spadot_b200.kmeans import KMeans
    X = np.repeat(np.random.default_rng(0).normal(0, 1, size=(6, 4)), 30, axis=0)
    X += np.random.default_rng(1).normal(0, 1e-3, size=X.shape)
    got = KMeans(n_clusters=6, random_state=0, n_init=3).fit(X)
    want = SkKMeans(n_clusters=6, random_state=0, n_init=3).fit(X)
    assert len(set(got.labels_)) == 6
    assert got.inertia_ == pytest.approx(want.inertia_, rel=1e-8)


@pytest.mark.parametrize("n,d,k,seed", [(747, 20, 10, 1), (2500, 20, 15, 6), (64, 3, 5, 9)])
def test_one_launch_restarts_match_host_driven_iterations(n, d, k, seed):
    """sdb_kmeans_lloyd_runs (all n_init restarts in one launch, whole Lloyd loop on chip) against the per-iteration
    kernels driven from the host and against sklearn: same labels, inertia and iteration count."""
    from spadot_b200.kmeans import KMeans
    X = latent(n, d, 8, seed, sd=0.8)
    fused = KMeans(n_clusters=k, random_state=1993, n_init=10).fit(X)
    host = KMeans(n_clusters=k, random_state=1993, n_init=10)
    host.FUSED_MAX_ELEMS = 0
    host.fit(X)
    want = SkKMeans(n_clusters=k, random_state=1993, n_init=10).fit(X)
    assert np.array_equal(fused.labels_, host.labels_) and np.array_equal(fused.labels_, want.labels_)
    assert fused.inertia_ == pytest.approx(host.inertia_, rel=1e-12)
    assert fused.inertia_ == pytest.approx(want.inertia_, rel=1e-10)
    assert fused.n_iter_ == host.n_iter_ == want.n_iter_
