"""The LRU policy of the CUDA path (CacheConfig(policy="lru"); the reference's comparison policy
cache_algo/LRU.py, selected there by --cache-algo) vs oracle.lru.BatchLRU, through the C-ABI: hit stream,
fp32 rows, eviction stream and the resident recency order must be bit-exact, every batch."""
import numpy as np
import pytest

from helpers import SKEW_ROWS, SMALL_ROWS, pkg, run_single_tier_parity

pytestmark = pytest.mark.gpu


def test_lru_small_batches():
    t = run_single_tier_parity(SMALL_ROWS, 16, 32, 600, [64], 40, policy="lru")
    assert t["evicted"] > 0


def test_lru_batch_of_one_and_ragged_sizes():
    """B = 1 is the reference's own granularity (one request of 26 keys per call)."""
    run_single_tier_parity(SMALL_ROWS, 16, 32, 500, [1, 3, 8, 9, 31, 64, 2, 100, 1, 1], 80, policy="lru")


def test_lru_skewed_tables_large_batch_ring_compaction():
    """Every hit appends a recency record, so the ring wraps within a few batches and k_compact runs."""
    t = run_single_tier_parity(SKEW_ROWS, 64, 32, 6000, [512], 40, check_state_every=4, policy="lru")
    assert t["evicted"] > 0


@pytest.mark.parametrize("prec", [16, 8])
def test_lru_quantised(prec):
    run_single_tier_parity(SMALL_ROWS, 36, prec, 150, [33, 64], 24, check_state_every=3, policy="lru")


def test_lru_more_new_keys_than_capacity():
    """A cold batch larger than the cache: the most recent keys survive, the last insert is never the victim."""
    run_single_tier_parity(SMALL_ROWS, 16, 32, 40, [64, 7], 12, policy="lru", alpha=0.2)


def test_lru_rejects_multi_layer_configs():
    p = pkg()
    tables = p.workload.make_tables(SMALL_ROWS, 16)
    with pytest.raises(Exception):
        p.EvStore(tables, p.CacheConfig(n_layers=2, main_precision=8, secondary_precision=4, total_size=200, policy="lru"))
    with pytest.raises(ValueError):
        p.EvStore(tables, p.CacheConfig(total_size=200, policy="mru"))


def test_lru_vs_evlfu_hit_rate_on_the_same_trace():
    """The paper's headline comparison (experiments.md:694-802): on a Zipf trace with few perfect-hit samples
    EvLFU keeps more of the hot keys than LRU at the same capacity; both answers are exact copies of the rows."""
    import torch
    p = pkg()
    rows = SKEW_ROWS
    tables = p.workload.make_tables(rows, 16)
    rates = {}
    for policy in ("evlfu", "lru"):
        tr = p.workload.ZipfTrace(rows, seed=5)
        store = p.EvStore(tables, p.CacheConfig(total_size=2500, max_batch=256, policy=policy))
        for it in range(120):
            idx = tr.batch(256)
            out, hit = store.lookup(torch.from_numpy(idx).cuda())
            if it % 30 == 0:
                o = out.cpu().numpy()
                for t in range(26):
                    assert np.array_equal(o[:, t], tables[t][idx[t]])
        store.sync()
        s = store.stats()
        rates[policy] = s["hits"][0] / s["lookups"]
        assert s["size"][0] <= 2500
        store.close()
    assert 0.3 < rates["lru"] < 1.0 and 0.3 < rates["evlfu"] < 1.0, rates
