"""k-means (n_init=10, k=10, random_state=1993) wall time: sklearn on the host cores vs the device Lloyd iterations."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.cluster import KMeans as SkKMeans
from spadot_b200.kmeans import KMeans

out = {}
for n, d in ((1966, 20), (20000, 20), (100000, 32)):
    rng = np.random.default_rng(n)
    centres = rng.normal(0, 1.5, size=(10, d))
    X = centres[rng.integers(0, 10, n)] + rng.normal(0, 0.9, size=(n, d))
    KMeans(n_clusters=10, random_state=1993, n_init=1).fit(X[:500])
    t0 = time.perf_counter(); a = SkKMeans(n_clusters=10, random_state=1993, n_init=10).fit(X); t1 = time.perf_counter()
    b = KMeans(n_clusters=10, random_state=1993, n_init=10).fit(X); t2 = time.perf_counter()
    out[f"{n}x{d}"] = dict(sklearn_s=t1 - t0, spadot_b200_s=t2 - t1, same_labels=bool((a.labels_ == b.labels_).all()),
                           inertia_rel=abs(a.inertia_ - b.inertia_) / a.inertia_)
print(json.dumps(out))
