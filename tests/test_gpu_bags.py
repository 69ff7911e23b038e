"""Bags on the cached path (evs_lookup_bags): pooling factor > 1, ragged and empty bags.  The policy is the batch-granular
EvLFU over the batch's SLICES (oracle.evlfu.expand_bags), the pooled rows are nn.EmbeddingBag(mode="sum") over the rows the
cache serves.  Bit-exact: hit code per index, pooled fp32 rows, eviction stream, FIFO state."""
import numpy as np
import pytest

from helpers import SKEW_ROWS, SMALL_ROWS, decoded_tables, pkg
from oracle.evlfu import BatchEvLFU, expand_bags, gather_rows, pool_bags

pytestmark = pytest.mark.gpu


def random_bags(rng, trace_rows, B, P, alpha_pick=1.3):
    """Per table: ragged bags of 0..P Zipf-ish indices."""
    idx_lists, off_lists = [], []
    for n in trace_rows:
        lens = rng.integers(0, P + 1, size=B)
        off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
        r = np.minimum((n * rng.random(int(lens.sum())) ** alpha_pick * rng.random(int(lens.sum()))).astype(np.int64), n - 1)
        idx_lists.append(r)
        off_lists.append(off)
    return idx_lists, off_lists


def drive(rows, dim, prec, total, B, P, n_batches, layers=1, sec=0):
    import torch
    p = pkg()
    tables = p.workload.make_tables(rows, dim)
    T = len(rows)
    rng = np.random.default_rng(3)
    if layers == 1:
        dec = decoded_tables(tables, prec)
        oracle = BatchEvLFU(total * (32 // prec), n_tables=T)
    else:
        from oracle import tiers as otiers
        dec = [decoded_tables(tables, prec), decoded_tables(tables, sec)]
        caps = otiers.capacities(layers, prec, sec, total, "", dim)
        oracle = otiers.BatchTiers(caps, n_layers=layers, T=T)
    cfg = p.CacheConfig(n_layers=layers, main_precision=prec, secondary_precision=sec, total_size=total, max_batch=B * P,
                        record_events=True)
    store = p.EvStore(tables, cfg)
    n_ev = 0
    try:
        for it in range(n_batches):
            idx_lists, off_lists = random_bags(rng, rows, B, P)
            lS_i = [torch.from_numpy(x).cuda() for x in idx_lists]
            lS_o = [torch.from_numpy(x).cuda() for x in off_lists]
            out, hit = store.lookup_bags(lS_o, lS_i, max_per_bag=P)
            torch.cuda.synchronize()
            store.check()
            idx_v = expand_bags(idx_lists, off_lists, B, P)
            if layers == 1:
                o_hit, st, sr, _ = oracle.lookup_batch(idx_v)
                rows_v = gather_rows(dec, st, sr)
                code_v = o_hit.astype(np.uint8)
                got_code = lambda h: (h != 0).astype(np.uint8)
                evs = [(0, oracle)]
            else:
                code_v, val_tier, st, sr, _ = oracle.lookup_batch(idx_v)
                rows_v = otiers.gather_tier_rows(dec, val_tier, st, sr)
                got_code = lambda h: h
                evs = [(0, oracle.c1), (1, oracle.c2)]
            want = pool_bags(rows_v, idx_v, B, P)
            assert (out.cpu().numpy() == want).all(), f"pooled rows, batch {it}"
            for t in range(T):
                off = list(off_lists[t]) + [len(idx_lists[t])]
                w = np.concatenate([code_v[s * P:s * P + (off[s + 1] - off[s]), t] for s in range(B)]) if len(idx_lists[t]) else np.zeros(0, np.uint8)
                assert (got_code(hit[t].cpu().numpy()) == w).all(), f"hit codes of table {t}, batch {it}"
            for ti, ot in evs:
                ev, fl = store.last_events(ti)
                assert ev.tolist() == ot.evicted and fl.tolist() == ot.flushed, f"eviction stream of tier {ti}, batch {it}"
                state, n_perfect = store.dump_state(ti)
                assert state == ot.state() and n_perfect == ot.n_perfect, f"FIFO state of tier {ti}, batch {it}"
                n_ev += len(ot.evicted)
        s = store.stats()
        assert s["lookups"] > 0
    finally:
        store.close()
    return n_ev


def test_bags_fp32_ragged():
    assert drive(SKEW_ROWS, 16, 32, 2500, 24, 5, 14) > 0


def test_bags_pooling_factor_10_dim64():
    assert drive(SKEW_ROWS, 64, 32, 900, 16, 10, 10) > 0


@pytest.mark.parametrize("prec", [16, 8, 4])
def test_bags_quantised_tier(prec):
    drive(SMALL_ROWS, 16, prec, 200, 12, 4, 8)


def test_bags_two_tiers():
    drive(SKEW_ROWS, 16, 32, 1500, 16, 4, 10, layers=2, sec=8)


def test_bags_equal_embedding_bag_sum():
    """fp32 tier: the pooled rows equal nn.EmbeddingBag(mode='sum') over the tables within one rounding per add of the
    other summation order (bit-exact for bags of up to 2 rows), and empty bags give zeros."""
    import torch
    p = pkg()
    rows, dim, B, P = SMALL_ROWS, 16, 20, 6
    tables = p.workload.make_tables(rows, dim)
    store = p.EvStore(tables, p.CacheConfig(total_size=400, max_batch=B * P))
    rng = np.random.default_rng(9)
    try:
        idx_lists, off_lists = random_bags(rng, rows, B, P)
        out, _ = store.lookup_bags([torch.from_numpy(x).cuda() for x in off_lists], [torch.from_numpy(x).cuda() for x in idx_lists], max_per_bag=P)
        torch.cuda.synchronize()
        for t in range(len(rows)):
            want = torch.nn.functional.embedding_bag(torch.from_numpy(idx_lists[t]), torch.from_numpy(tables[t]), torch.from_numpy(off_lists[t]), mode="sum")
            assert torch.allclose(out[:, t, :].cpu(), want, atol=1e-6, rtol=1e-6), f"table {t}"
            off = list(off_lists[t]) + [len(idx_lists[t])]
            empty = [s for s in range(B) if off[s + 1] == off[s]]
            assert (out[empty, t, :].cpu() == 0).all()
    finally:
        store.close()


def test_bags_errors():
    import torch
    p = pkg()
    tables = p.workload.make_tables(SMALL_ROWS, 16)
    store = p.EvStore(tables, p.CacheConfig(total_size=300, max_batch=64))
    try:
        B = 8
        lS_i = [torch.zeros(3 * B, dtype=torch.int64, device="cuda") for _ in SMALL_ROWS]
        lS_o = [torch.arange(B, dtype=torch.int64, device="cuda") * 3 for _ in SMALL_ROWS]
        with pytest.raises(p.EvsError):
            store.lookup_bags(lS_o, lS_i, max_per_bag=10)       # 8 * 10 slices > max_batch
        store.lookup_bags(lS_o, lS_i, max_per_bag=2)           # bags of 3 with max_per_bag 2
        with pytest.raises(p.EvsError):
            store.sync()
    finally:
        store.close()
