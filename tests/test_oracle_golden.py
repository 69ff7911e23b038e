"""The oracle against the reference's own outputs (tests/golden/*.npz, made by
tests/golden/make_golden.py from /root/reference/cache_algo/EvLFU_C1.py)."""
import os
from collections import OrderedDict

import numpy as np
import pytest

from oracle.evlfu import BatchEvLFU, SeqEvLFU

T = 26
CASES = ["evlfu_c1_small", "evlfu_c1_skew", "evlfu_c1_flush", "evlfu_c1_approx"]


def _load(golden_dir, name):
    with np.load(os.path.join(golden_dir, name + ".npz")) as z:
        g = {k: z[k] for k in z.files}          # NpzFile re-inflates on every access
    hits = np.unpackbits(g["hits"], axis=1)[:, :T].astype(bool)
    return g, hits


@pytest.mark.parametrize("name", CASES)
def test_seq_oracle_equals_reference(golden_dir, name):
    """Hit vector, eviction order, flush order, answering rows, final FIFO state: bit-exact."""
    g, hits = _load(golden_dir, name)
    o = SeqEvLFU(int(g["cap"]))
    thres = int(g["approx_thres"])
    for i, req in enumerate(g["trace"]):
        h, src, _agg = o.request(req, thres)
        assert list(h) == list(hits[i]), f"hit vector, request {i}"
        assert o.evicted == list(g["ev_keys"][g["ev_off"][i]:g["ev_off"][i + 1]]), f"evictions, request {i}"
        assert o.flushed == list(g["fl_keys"][g["fl_off"][i]:g["fl_off"][i + 1]]), f"flush, request {i}"
        for t in range(T):
            if src[t] is None:
                assert g["src_t"][i, t] == -1
            else:
                assert (g["src_t"][i, t], g["src_r"][i, t]) == src[t], f"value source, request {i} table {t}"
    want = [list(g["state_keys"][g["state_off"][b]:g["state_off"][b + 1]]) for b in range(T + 1)]
    assert o.state() == want
    assert o.n_perfect == int(g["n_perfect"]) and o.min == int(g["min_bucket"])


def _clone_to_batch(o):
    p = BatchEvLFU(o.cap)
    p.lists = [OrderedDict((k, None) for k in l) for l in o.lists]
    p.entries = dict(o.vals)
    p.n_perfect = o.n_perfect
    return p


@pytest.mark.parametrize("name", CASES)
def test_batch_policy_at_b1_equals_sequential(golden_dir, name):
    """From the same pre-state, one sample through the batch policy gives the sequential
    policy's hits, eviction/flush sets and post-state on every request that does not take
    the same-request re-fetch corner (EvLFU_C1.py:84-95)."""
    g, _ = _load(golden_dir, name)
    o = SeqEvLFU(int(g["cap"]))
    thres = int(g["approx_thres"])
    corners = 0
    for i, req in enumerate(g["trace"]):
        p = _clone_to_batch(o)
        h, src, agg = o.request(req, thres)
        ph, pt, pr, pagg = p.lookup_batch(np.asarray(req).reshape(T, 1), thres)
        if o.refetched:
            corners += 1
            continue
        assert list(ph[0]) == list(h) and int(pagg[0]) == agg, i
        assert sorted(p.evicted) == sorted(o.evicted), i
        assert sorted(p.flushed) == sorted(o.flushed), i
        assert p.state() == o.state() and p.n_perfect == o.n_perfect, i
        for t in range(T):
            if src[t] is not None:
                assert (pt[0, t], pr[0, t]) == src[t]
    assert corners < len(g["trace"]) // 4


def test_batch_policy_duplicates_and_capacity():
    """Same-batch duplicate misses insert once; size never exceeds cap after a batch."""
    rng = np.random.default_rng(0)
    p = BatchEvLFU(64, n_tables=4)
    for _ in range(50):
        idx = rng.integers(0, 40, size=(4, 32))
        idx[:, 1] = idx[:, 0]
        hit, st, sr, agg = p.lookup_batch(idx)
        assert len(p.entries) <= 64
        assert len(set(p.inserted)) == len(p.inserted)
        assert sum(len(l) for l in p.lists) == len(p.entries)
        assert (hit[0] == hit[1]).all()


def test_batch_policy_empty_batch():
    p = BatchEvLFU(64, n_tables=4)
    hit, st, sr, agg = p.lookup_batch(np.zeros((4, 0), dtype=np.int64))
    assert hit.shape == (0, 4) and not p.evicted and not p.inserted


def test_slices_of_bags_in_the_batch_oracle():
    """oracle.evlfu.expand_bags / pool_bags and the absent-key handling of BatchEvLFU: with one index per bag the slices ARE
    the batch (identical streams), absent positions are never probed or inserted, pooling sums in ascending j."""
    from oracle.evlfu import BatchEvLFU, expand_bags, pool_bags
    rng = np.random.default_rng(0)
    rows = [40, 7, 300, 5]
    a, b = BatchEvLFU(60, n_tables=4), BatchEvLFU(60, n_tables=4)
    for _ in range(20):
        idx = np.stack([rng.integers(0, n, size=16) for n in rows])
        ha = a.lookup_batch(idx)[0]
        v = expand_bags([idx[t] for t in range(4)], [np.arange(16) for _ in range(4)], 16, 1)
        hb = b.lookup_batch(v)[0]
        assert (ha == hb).all() and a.evicted == b.evicted and a.state() == b.state()
    c = BatchEvLFU(60, n_tables=4)
    idx_lists = [np.array([1, 2, 3]), np.array([], dtype=np.int64), np.array([5]), np.array([0, 0])]
    off_lists = [np.array([0, 2]), np.array([0, 0]), np.array([0, 0]), np.array([0, 1])]
    v = expand_bags(idx_lists, off_lists, 2, 2)
    assert v.tolist() == [[1, 2, 3, -1], [-1, -1, -1, -1], [-1, -1, 5, -1], [0, -1, 0, -1]]
    hit, st, sr, agg = c.lookup_batch(v)
    assert not hit.any() and (st[v.T < 0] == -1).all() and len(c.entries) == 5
    rows_v = np.arange(4 * 4 * 2, dtype=np.float32).reshape(4, 4, 2)
    pooled = pool_bags(rows_v, v, 2, 2)
    assert (pooled[0, 0] == rows_v[0, 0] + rows_v[1, 0]).all() and (pooled[1, 0] == rows_v[2, 0]).all() and (pooled[:, 1] == 0).all()
