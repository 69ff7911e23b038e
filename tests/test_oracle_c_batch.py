"""oracle/evlfu_batch.c (the C restatement used at the BASELINE sizes) against oracle/evlfu.py: BatchEvLFU, which is
itself pinned to the reference's EvLFU_C1.py by the golden traces: hit maps, agg_hit, eviction and flush streams,
n_perfect and the FIFO state of every bucket after every batch."""
import numpy as np
import pytest

from helpers import SKEW_ROWS, SMALL_ROWS, TINY_ROWS, pkg
from oracle.evlfu import BatchEvLFU
from oracle.evlfu_c import CBatchEvLFU


def _follow(rows, cap, B_list, n, seed, alpha=1.05, table_ids=None, agg_shift=False):
    p = pkg()
    Tl = len(rows)
    trace = p.workload.ZipfTrace(rows, alpha=alpha, seed=seed)
    py, c = BatchEvLFU(cap, n_tables=26), CBatchEvLFU(cap, n_tables=26, max_keys_per_batch=max(B_list) * Tl)
    rng = np.random.default_rng(seed)
    tot = dict(ev=0, fl=0)
    for it in range(n):
        idx = trace.batch(B_list[it % len(B_list)])
        agg = None
        if agg_shift:          # an externally supplied agg_hit (table-wise sharding: the other ranks' counts are added)
            local = np.array([sum(((table_ids[t] if table_ids else t) << 40 | int(idx[t, s])) in py.entries for t in range(Tl))
                              for s in range(idx.shape[1])])
            agg = np.minimum(26, local + rng.integers(0, 26 - Tl + 1, size=idx.shape[1]))
        a = py.lookup_batch(idx, agg=agg, table_ids=table_ids)
        b = c.lookup_batch(idx, agg=agg, table_ids=table_ids)
        assert np.array_equal(a[0], b[0]), f"hit map, batch {it}"
        assert np.array_equal(a[3], b[3]), f"agg_hit, batch {it}"
        assert py.evicted == c.evicted, f"eviction stream, batch {it}"
        assert py.flushed == c.flushed, f"flush stream, batch {it}"
        assert len(py.inserted) == c.n_inserted
        assert py.n_perfect == c.n_perfect and len(py.entries) == c.size
        if it % 3 == 0 or it == n - 1:
            assert py.state() == c.state(), f"FIFO state, batch {it}"
        tot["ev"] += len(py.evicted)
        tot["fl"] += len(py.flushed)
    return tot


def test_c_batch_oracle_equals_python_on_zipf_traces():
    t = _follow(SMALL_ROWS, 600, [64, 5, 1, 33], 60, seed=1)
    assert t["ev"] > 0
    t = _follow(SKEW_ROWS, 2500, [300, 64, 700], 25, seed=2)
    assert t["ev"] > 0


def test_c_batch_oracle_flush_rule():
    t = _follow(TINY_ROWS, 60, [16, 4], 120, seed=3, alpha=2.5)
    assert t["fl"] > 0, t


@pytest.mark.parametrize("ids", [[3, 4, 5, 6, 7], [0, 2, 11, 20, 25]])
def test_c_batch_oracle_sharded_inputs(ids):
    rows = [SKEW_ROWS[t] for t in ids]
    t = _follow(rows, 250, [128, 31], 30, seed=4, table_ids=ids, agg_shift=True)
    assert t["ev"] > 0
