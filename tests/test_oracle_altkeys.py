"""oracle/altkeys.py against scikit-learn's NearestNeighbors with the notebook's own arguments (the CPU stand-in for cuML,
SURVEY.md 8(c)): n_neighbors = 11, algorithm='brute', metric='euclidean', entry [0] dropped."""
import numpy as np
import pytest

from oracle import altkeys


@pytest.mark.parametrize("n,d,seed", [(400, 16, 0), (257, 36, 1), (90, 64, 2)])
def test_bruteforce_knn_equals_sklearn(n, d, seed):
    from sklearn.neighbors import NearestNeighbors
    x = np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)
    nbrs = NearestNeighbors(n_neighbors=11, algorithm="brute", metric="euclidean").fit(x)
    want = nbrs.kneighbors(x, return_distance=False)[:, 1:]
    got, dist = altkeys.knn_bruteforce(x, k=10)
    assert (got == want).mean() > 0.999                 # sklearn's fp32 GEMM may swap an exact near-tie
    assert (np.diff(dist, axis=1) >= 0).all()


def test_most_popular_is_the_first_maximum_and_the_key_format():
    nbr = np.array([[3, 1, 2], [0, 2, 3], [1, 0, -1]])
    freq = np.array([5, 9, 9, 1])
    assert altkeys.most_popular(nbr, freq).tolist() == [1, 2, 1]
    assert altkeys.most_popular(nbr, None).tolist() == [3, 0, 1]
    rng = np.random.default_rng(3)
    tables = [rng.standard_normal((30, 8)).astype(np.float32), rng.standard_normal((50, 8)).astype(np.float32)]
    keys = altkeys.alt_keys(tables)
    assert [k.dtype for k in keys] == [np.uint32, np.uint32] and [len(k) for k in keys] == [30, 50]
    t = np.concatenate(keys) % 100
    r = np.concatenate(keys) // 100
    assert set(t.tolist()) <= {1, 2} and (r[t == 1] < 30).all() and (r[t == 2] < 50).all()
