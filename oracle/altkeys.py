"""Alt-key generation oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Restates the reference's two notebooks (script/approximate_embedding/phase2_similarity_analysis/):
``get_neighbors_GPU.ipynb`` -- all rows of all tables in one matrix, NearestNeighbors(n_neighbors=11, algorithm='brute',
metric='euclidean'), entry [0] dropped -- and ``most_popular_neighbor.ipynb`` -- of the 10 neighbours the one with the highest
workload frequency, ``freqArr.index(max_value)`` = the first maximum, absent rows counting 0.  cuML is not installed here;
scikit-learn's NearestNeighbors with the same arguments is, and tests/test_oracle_altkeys.py checks this
restatement against it.  Distances in float64, ties by row index."""
from __future__ import annotations

import numpy as np


def knn_bruteforce(x: np.ndarray, q: np.ndarray | None = None, k: int = 10):
    """(neighbours int64 [nq, k], squared distances float64 [nq, k]): the k + 1 nearest rows of x with entry [0] dropped."""
    x = np.asarray(x, dtype=np.float64)
    q = x if q is None else np.asarray(q, dtype=np.float64)
    n = x.shape[0]
    xn = (x * x).sum(axis=1)
    nbr = np.full((q.shape[0], k), -1, dtype=np.int64)
    dist = np.full((q.shape[0], k), np.inf)
    for i in range(q.shape[0]):
        d = xn - 2.0 * (x @ q[i]) + (q[i] * q[i]).sum()
        order = np.lexsort((np.arange(n), d))[:k + 1]            # by distance, then by index
        take = order[1:]
        nbr[i, :len(take)] = take
        dist[i, :len(take)] = d[take]
    return nbr, dist


def most_popular(nbr: np.ndarray, freq: np.ndarray | None):
    """Per row the neighbour with the highest frequency, the first one on ties (most_popular_neighbor.ipynb)."""
    out = np.full(nbr.shape[0], -1, dtype=np.int64)
    for i, row in enumerate(nbr):
        row = row[row >= 0]
        if len(row) == 0:
            continue
        f = np.zeros(len(row), dtype=np.int64) if freq is None else np.asarray(freq)[row].astype(np.int64)
        out[i] = row[int(np.argmax(f))]                           # argmax returns the first maximum
    return out


def alt_keys(tables, freq=None, k: int = 10):
    """One uint32 array per table: alt_row * 100 + alt_table (1-based), convert_altkeys_to_binary.py:50."""
    rows = [t.shape[0] for t in tables]
    off = np.concatenate([[0], np.cumsum(rows)])
    x = np.concatenate([np.asarray(t, dtype=np.float32) for t in tables])
    nbr, _ = knn_bruteforce(x, k=k)
    pick = most_popular(nbr, None if freq is None else np.concatenate([np.asarray(f) for f in freq]))
    t_of = np.searchsorted(off, pick, side="right") - 1
    key = np.where(pick >= 0, (pick - off[np.maximum(t_of, 0)]) * 100 + (t_of + 1), 0).astype(np.uint32)
    return [key[off[t]:off[t + 1]] for t in range(len(rows))]
