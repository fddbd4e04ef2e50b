"""`Adaptive_clustering`'s elbow rule (SpaDOT/utils/_analyze_utils.py:73-88): the product's numpy form against the
literal pandas restatement, on random WSS curves; on the GPU the whole driver against sklearn + that rule."""
import numpy as np
import pytest

from oracle import cluster_ref


def test_elbow_rule_matches_pandas_restatement():
    from spadot_b200.kmeans import select_n_clusters
    rng = np.random.default_rng(0)
    for trial in range(300):
        K = int(rng.integers(3, 18))
        drops = np.abs(rng.normal(0, 1, K)) * np.exp(-0.3 * np.arange(K) * rng.uniform(0, 2))
        if trial % 7 == 0:
            drops[rng.integers(0, K)] *= -0.2                    # a non-monotone WSS curve (k-means is not exact)
        wss = 100.0 + np.cumsum(drops[::-1])[::-1]
        try:
            want = cluster_ref.select_n_clusters_ref(wss, 4, 0.1)
        except Exception:
            with pytest.raises(ValueError):
                select_n_clusters(wss, 4, 0.1)
            continue
        got, table = select_n_clusters(wss, 4, 0.1)
        assert got == want, (trial, wss)
        assert table["clusters"][0] == 4 and table["wss"].size == K


@pytest.mark.gpu
def test_adaptive_clustering_matches_sklearn_driver():
    from sklearn.cluster import KMeans as SkKMeans
    from spadot_b200.kmeans import adaptive_clustering
    rng = np.random.default_rng(3)
    centres = rng.normal(0, 2.0, size=(7, 10))
    X = centres[rng.integers(0, 7, 900)] + rng.normal(0, 0.5, size=(900, 10))
    wss = [SkKMeans(n_clusters=k, random_state=1993, n_init=10).fit(X).inertia_ for k in range(4, 13)]
    k_ref = cluster_ref.select_n_clusters_ref(wss, 4, 0.1)
    labels, k, table = adaptive_clustering(X, 4, 12)
    assert k == k_ref
    np.testing.assert_allclose(table["wss"], wss, rtol=1e-9)
    assert np.array_equal(labels, SkKMeans(n_clusters=k_ref, random_state=1993, n_init=10).fit(X).labels_)
