"""The LRU oracles (oracle/lru.py) against the reference's own outputs (tests/golden/lru_*.npz, made by
tests/golden/make_golden.py from /root/reference/cache_algo/LRU.py)."""
import os
from collections import OrderedDict

import numpy as np
import pytest

from oracle.lru import BatchLRU, SeqLRU

T = 26
CASES = ["lru_small", "lru_skew"]


def _load(golden_dir, name):
    with np.load(os.path.join(golden_dir, name + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    return g, np.unpackbits(g["hits"], axis=1)[:, :T].astype(bool)


@pytest.mark.parametrize("name", CASES)
def test_seq_lru_equals_reference(golden_dir, name):
    """Hit vector and eviction order of every request, final recency order: bit-exact."""
    g, hits = _load(golden_dir, name)
    o = SeqLRU(int(g["cap"]))
    for i, req in enumerate(g["trace"]):
        assert o.request(req) == list(hits[i]), f"hit vector, request {i}"
        assert o.evicted == list(g["ev_keys"][g["ev_off"][i]:g["ev_off"][i + 1]]), f"evictions, request {i}"
    assert o.state() == list(g["state_keys"])


@pytest.mark.parametrize("name", CASES)
def test_batch_lru_at_b1_equals_sequential(golden_dir, name):
    """From the same pre-state one sample through the batch policy gives the sequential policy's hits,
    victims and recency order, on every request that does not take the same-request corner."""
    g, _ = _load(golden_dir, name)
    o = SeqLRU(int(g["cap"]))
    corners = 0
    for i, req in enumerate(g["trace"]):
        p = BatchLRU(o.cap)
        p.entries = OrderedDict((k, None) for k in o.od)
        h = o.request(req)
        ph, _st, _sr, _agg = p.lookup_batch(np.asarray(req).reshape(T, 1))
        if o.corner:
            corners += 1
            continue
        assert list(ph[0]) == h, i
        assert p.evicted == o.evicted, i
        assert p.state()[0] == o.state(), i
    assert corners < len(g["trace"]) // 3


def test_batch_lru_duplicates_capacity_and_protection():
    rng = np.random.default_rng(1)
    p = BatchLRU(48, n_tables=4)
    for _ in range(60):
        idx = rng.integers(0, 60, size=(4, 40))
        idx[:, 1] = idx[:, 0]
        hit, _st, _sr, _agg = p.lookup_batch(idx)
        assert len(p.entries) <= 48 and len(set(p.inserted)) == len(p.inserted)
        assert (hit[0] == hit[1]).all()
        if p.inserted:
            assert p.inserted[-1] in p.entries          # the last insert survives (LRU.py evicts before it inserts)
    # a batch with more new keys than the cache holds keeps the most recent ones
    p = BatchLRU(8, n_tables=1)
    p.lookup_batch(np.arange(100, 130).reshape(1, 30))
    assert p.state()[0] == list(range(100 + 22, 130))


@pytest.mark.parametrize("name", ["lfu_small", "lfu_skew"])
def test_seq_lfu_equals_reference(golden_dir, name):
    """cache_algo/LFU.py: hit vector and evicted keys of every request, final frequency lists and least_freq."""
    from oracle.lru import SeqLFU
    g, hits = _load(golden_dir, name)
    o = SeqLFU(int(g["cap"]))
    for i, req in enumerate(g["trace"]):
        before = set(o.key)
        assert o.request(req) == list(hits[i]), f"hit vector, request {i}"
        # the fixture lists the keys resident before the request that are gone after it (a key inserted and evicted
        # again by the same request -- LFU evicts among the newest keys -- never shows, nor does one that came back)
        gone = sorted((set(o.evicted) & before) - set(o.key))
        assert gone == list(g["ev_keys"][g["ev_off"][i]:g["ev_off"][i + 1]]), f"evictions, request {i}"
    least, lists = o.state()
    want = [list(g["state_keys"][g["state_off"][f]:g["state_off"][f + 1]]) for f in range(len(g["state_off"]) - 1)]
    assert least == int(g["least_freq"]) and lists == want


def _clone_lfu(o, n_tables=T):
    from oracle.lru import BatchLFU
    p = BatchLFU(o.cap, n_tables=n_tables)
    for f in range(1, len(o.freq)):
        for k in o.freq[f]:
            p.lists[f - 1][k] = None
            p.entries[k] = f - 1
    return p


@pytest.mark.parametrize("name", ["lfu_small", "lfu_skew"])
def test_batch_lfu_at_b1_equals_capped_sequential(golden_dir, name):
    """BatchLFU (what policy="lfu" runs on the GPU) from the pre-state of SeqLFU(freq_cap = 27): the same hits, the same set
    of victims and the same frequency lists after every request that does not take the same-request corner.  (The ORDER of a
    request's victims may differ: the sequential policy alternates insert / evict and can evict a key it inserted a moment
    ago before an older, more frequent one; the batch evicts lowest frequency first.)  And the capped sequential policy is the
    reference's as long as no frequency reaches the cap."""
    from oracle.lru import SeqLFU
    g, hits = _load(golden_dir, name)
    o = SeqLFU(int(g["cap"]), freq_cap=T + 1)
    ref = SeqLFU(int(g["cap"]))
    same_as_ref = True
    corners = checked = 0
    for i, req in enumerate(g["trace"]):
        p = _clone_lfu(o)
        h = o.request(req)
        if same_as_ref:
            hr = ref.request(req)
            if len(ref.freq) - 1 > T + 1:
                same_as_ref = False                     # a key went past the cap: from here on the two may part
            else:
                assert hr == h and ref.evicted == o.evicted, i
        ph, _st, _sr, _agg = p.lookup_batch(np.asarray(req).reshape(T, 1))
        if o.corner:
            corners += 1
            continue
        assert list(ph[0]) == h, i
        assert sorted(p.evicted) == sorted(o.evicted), i
        _least, lists = o.state()
        want = [l for l in lists] + [[] for _ in range(T + 1 - len(lists))]
        assert p.state() == want[:T + 1], i
        checked += 1
    assert checked > len(g["trace"]) // 2


def test_batch_lfu_frequency_counts_batches_and_saturates():
    from oracle.lru import BatchLFU
    p = BatchLFU(100, n_tables=3)
    idx = np.array([[5, 5, 5], [1, 1, 2], [7, 8, 9]])          # table 0: the same key three times in one batch
    for n in range(6):
        hit, *_ = p.lookup_batch(idx)
        assert p.entries[(0 << 40) | 5] == min(n, 3)            # one step per batch, saturating at bucket n_tables
        assert hit[:, 0].all() == (n > 0)
    # eviction: lowest frequency first, oldest first, the last insert survives
    q = BatchLFU(4, n_tables=1)
    q.lookup_batch(np.array([[1, 2, 3, 4]]))
    q.lookup_batch(np.array([[1, 2]]))                          # 1, 2 -> frequency 2
    q.lookup_batch(np.array([[9, 10]]))
    assert q.evicted == [3, 4] and q.state()[0] == [9, 10] and q.state()[1] == [1, 2]
    q.lookup_batch(np.array([[11, 12, 13, 14, 15]]))            # more new keys than the cache holds
    assert 15 in q.entries and len(q.entries) == 4


@pytest.mark.parametrize("seed", range(6))
def test_batch_lfu_random_configurations_against_the_sequential_policy(seed):
    """Random table counts, capacities and skews: BatchLFU fed one request at a time tracks SeqLFU(freq_cap = T + 1) -- and, while no
    frequency reaches the cap, the reference-exact SeqLFU -- request by request (state re-cloned after a same-request corner)."""
    from oracle.lru import BatchLFU, SeqLFU
    rng = np.random.default_rng(100 + seed)
    Tn = int(rng.integers(2, 9))
    rows = rng.integers(3, 60, size=Tn)
    cap = int(rng.integers(6, 50))
    o = SeqLFU(cap, n_tables=Tn, freq_cap=Tn + 1)
    checked = 0
    for i in range(400):
        req = [int(rng.integers(0, rows[t]) if rng.random() < 0.7 else rng.integers(0, min(3, rows[t]))) for t in range(Tn)]
        p = _clone_lfu(o, n_tables=Tn)
        h = o.request(req)
        ph, *_ = p.lookup_batch(np.asarray(req).reshape(Tn, 1))
        if o.corner:
            continue
        assert list(ph[0]) == h and sorted(p.evicted) == sorted(o.evicted), (seed, i)
        _least, lists = o.state()
        assert p.state() == (lists + [[] for _ in range(Tn + 1 - len(lists))])[:Tn + 1], (seed, i)
        checked += 1
    assert checked > 150
