"""The LFU policy of the CUDA path (CacheConfig(policy="lfu"); the reference's cache_algo/LFU.py, the third simulator its
driver initialises, dlrm_s_pytorch_C1_C2_C3.py:1294) vs oracle.lru.BatchLFU, through the C-ABI: hit stream, fp32 rows,
eviction stream and the per-frequency FIFO lists must be bit-exact, every batch.  Frequencies saturate at n_tables + 1
(one bucket ring per frequency); BatchLFU at B = 1 is pinned to the reference's LFU.py in tests/test_oracle_lru.py."""
import numpy as np
import pytest

from helpers import SKEW_ROWS, SMALL_ROWS, TINY_ROWS, pkg, run_single_tier_parity

pytestmark = pytest.mark.gpu


def test_lfu_small_batches():
    t = run_single_tier_parity(SMALL_ROWS, 16, 32, 600, [64], 40, policy="lfu")
    assert t["evicted"] > 0


def test_lfu_batch_of_one_and_ragged_sizes():
    """B = 1 is the reference's own granularity (one request of 26 keys per call)."""
    run_single_tier_parity(SMALL_ROWS, 16, 32, 500, [1, 3, 8, 9, 31, 64, 2, 100, 1, 1], 80, policy="lfu")


def test_lfu_saturating_frequencies():
    """Tiny tables: every key is requested in almost every batch, the frequencies run into the cap (bucket 26) and the keys
    there are re-appended like in an LRU list."""
    run_single_tier_parity(TINY_ROWS, 16, 32, 110, [4, 16], 120, check_state_every=5, policy="lfu")


def test_lfu_skewed_tables_large_batch():
    t = run_single_tier_parity(SKEW_ROWS, 64, 32, 6000, [512], 30, check_state_every=3, policy="lfu")
    assert t["evicted"] > 0


@pytest.mark.parametrize("prec", [16, 8])
def test_lfu_quantised(prec):
    run_single_tier_parity(SMALL_ROWS, 36, prec, 150, [33, 64], 20, check_state_every=3, policy="lfu")


def test_lfu_more_new_keys_than_capacity_and_look_ahead():
    run_single_tier_parity(SMALL_ROWS, 16, 32, 40, [64, 7], 12, policy="lfu", alpha=0.2)
    run_single_tier_parity(SMALL_ROWS, 16, 32, 600, [64], 20, policy="lfu", prefetch=True)


def test_lfu_few_tables_packed_warps():
    """A handle of 5 tables packs four samples into a warp: the per-bucket ranks inside a warp cross sample boundaries."""
    run_single_tier_parity([400, 300, 50, 2500, 9], 16, 32, 300, [64, 33], 30, policy="lfu")


def test_lfu_rejects_multi_layer_configs():
    p = pkg()
    tables = p.workload.make_tables(SMALL_ROWS, 16)
    with pytest.raises(Exception):
        p.EvStore(tables, p.CacheConfig(n_layers=2, main_precision=8, secondary_precision=4, total_size=200, policy="lfu"))
