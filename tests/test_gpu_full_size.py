"""BASELINE configs[1] at full size (Kaggle-shape tables, 33.76 M rows, cache 13 %, batch 2048), where the
Python oracle would take hours: size-independent properties of the path.

* every answered row IS the backing-store row of its key (the cache only ever holds copies) -- checked on
  every 40th batch through 5 000 batches, i.e. with the slab being recycled by ~1 400 evictions per batch;
* hit + miss == lookups, inserts - evictions - flushed == resident size <= capacity at every checkpoint;
* a batch looked up twice in a row is answered entirely from the cache the second time, except keys the
  first pass's own evictions removed (bounded by that batch's eviction count);
* the resident structure is self-consistent (evs_dump_state re-derives every bucket's live records from the
  rings and compares them with the counters) and holds no duplicate key.
"""
import numpy as np
import pytest

from helpers import pkg

pytestmark = pytest.mark.gpu


def test_kaggle_shape_full_size_properties():
    import torch
    p = pkg()
    rows, dim, B = p.workload.KAGGLE_ROWS, 16, 2048
    cap = p.workload.KAGGLE_CACHE_ROWS
    tables = p.workload.make_tables(rows, dim)
    n = 5000
    idx = p.workload.ZipfTrace(rows, seed=42).batches(n, B)
    store = p.EvStore(tables, p.CacheConfig(total_size=cap, max_batch=B))
    d_idx = torch.from_numpy(idx).cuda()
    out = torch.empty((B, 26, dim), dtype=torch.float32, device="cuda")
    hit = torch.empty((B, 26), dtype=torch.uint8, device="cuda")
    checked = 0
    for k in range(n):
        store.lookup(d_idx[k], out=out, hit=hit)
        if k % 40 == 0 or k == n - 1:
            torch.cuda.synchronize()
            o = out.cpu().numpy()
            for t in range(26):
                assert np.array_equal(o[:, t], tables[t][idx[k, t]]), f"batch {k} table {t}: row differs from the backing store"
            checked += 1
        if k % 1000 == 999:
            s = store.stats()
            assert s["hits"][0] + s["misses"] == s["lookups"] == (k + 1) * B * 26
            assert s["inserts"][0] - s["evictions"][0] - s["flushed"][0] == s["size"][0] <= cap
    store.sync()
    s = store.stats()
    assert s["size"][0] == cap and s["evictions"][0] > 100000, s        # the cache filled up and has been evicting
    # the same batch twice: the second pass misses only what the first pass's own evictions removed
    h1 = hit.cpu().numpy().copy()
    store.lookup(d_idx[n - 1], out=out, hit=hit)
    torch.cuda.synchronize()
    h2 = hit.cpu().numpy()
    first_pass_misses = int((h1 == 0).sum())
    # (pass 1 evicts as many keys as it inserts; a victim may occupy more than one position of the batch)
    assert int((h2 == 0).sum()) <= 2 * first_pass_misses + 16, (int((h2 == 0).sum()), first_pass_misses)
    # structure
    state, _n_perfect = store.dump_state()
    keys = np.concatenate([np.asarray(b, dtype=np.int64) for b in state if len(b)])
    assert len(keys) == cap and len(np.unique(keys)) == cap
    assert checked >= 125
    store.close()


def _follow_full_size(layers_cfg, n_compare, B=2048, prefetch=True):
    """The CUDA path next to the C restatement of the batch-granular oracle (oracle/evlfu_batch.c, pinned to
    oracle/evlfu.py by tests/test_oracle_c_batch.py) at the BASELINE sizes: the oracle follows every batch of the
    warm-up (the cache fills to its 4 389 135 entries and the rings wrap), then hit maps, agg_hit, eviction streams
    (order included) and n_perfect are compared batch by batch, and the whole FIFO state at the end."""
    import torch
    from oracle.evlfu_c import CBatchEvLFU
    p = pkg()
    rows, dim = p.workload.KAGGLE_ROWS, 16
    cap = p.workload.KAGGLE_CACHE_ROWS
    tables = p.workload.make_tables(rows, dim)
    # the fill point of the trace (bench.py uses the same rule), then evictions for a while, then the compared batches
    import bench
    n_search = 4200
    idx = p.workload.ZipfTrace(rows, seed=42).batches(n_search + n_compare + 2, B)
    warm = min(n_search, bench.batches_until_full(idx[:n_search], rows, cap) + 150)
    store = p.EvStore(tables, p.CacheConfig(total_size=cap, max_batch=B, record_events=True, **layers_cfg))
    oracle = CBatchEvLFU(cap, n_tables=26, max_keys_per_batch=B * 26)
    d_idx = torch.from_numpy(idx[:warm + n_compare + 1]).cuda()
    out = torch.empty((B, 26, dim), dtype=torch.float32, device="cuda")
    hit = torch.empty((B, 26), dtype=torch.uint8, device="cuda")
    for k in range(warm):
        store.lookup(d_idx[k], out=out, hit=hit)
        if prefetch:
            store.prefetch(d_idx[k + 1])
        oracle.lookup_batch(idx[k])
    torch.cuda.synchronize()
    assert oracle.size == cap and store.stats()["size"][0] == cap
    n_ev = 0
    for k in range(warm, warm + n_compare):
        store.lookup(d_idx[k], out=out, hit=hit)
        if prefetch:
            store.prefetch(d_idx[k + 1])
        o_hit, _st, _sr, _agg = oracle.lookup_batch(idx[k])
        torch.cuda.synchronize()
        assert np.array_equal(hit.cpu().numpy().astype(bool), o_hit), f"hit map, batch {k}"
        ev, fl = store.last_events()
        assert ev.tolist() == oracle.evicted, f"eviction stream, batch {k}: {len(ev)} vs {len(oracle.evicted)} keys"
        assert fl.tolist() == oracle.flushed, f"flush stream, batch {k}"
        n_ev += len(oracle.evicted)
        if (k - warm) % 50 == 0:
            o = out.cpu().numpy()
            for t in range(26):
                assert np.array_equal(o[:, t], tables[t][idx[k, t]]), f"batch {k} table {t}: row differs from the backing store"
    state, n_perfect = store.dump_state()
    assert n_perfect == oracle.n_perfect
    assert state == oracle.state(), "resident FIFO state after the compared batches"
    store.close()
    oracle.close()
    return n_ev, warm


def test_kaggle_shape_full_size_eviction_order_equals_the_oracle():
    """configs[1] at full size: 33.76 M rows, cache 4 389 135, batch 2048 -- bit-exact against the oracle for 200 batches
    after the cache has filled and evicted for 150 batches, with the look-ahead on."""
    n_ev, warm = _follow_full_size({}, 200)
    assert n_ev > 200 * 800 and warm > 1000, (n_ev, warm)
