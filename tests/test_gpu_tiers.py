"""C1+C2 mixed-precision tiers and C3 approximate embeddings: CUDA path vs the batch-granular
oracle (oracle/tiers.py, whose sequential form is pinned to the compiled reference by
tests/test_oracle_tiers.py).  Integer streams are bit-exact; the fp32 rows are bit-exact with the
reference's own dequantisers (evlfu_16.cpp:332, evlfu_8.cpp:370, evlfu_4.cpp:319)."""
import pytest

from helpers import SKEW_ROWS, SMALL_ROWS, TINY_ROWS, run_tier_parity

pytestmark = pytest.mark.gpu


def test_two_tier_8_4_reference_dimension():
    t = run_tier_parity(SMALL_ROWS, 36, 2, 8, 4, 200, [64, 17], 40)
    assert t["c2"] > 0 and t["ev1"] > 0 and t["ev2"] > 0, t


@pytest.mark.parametrize("main,sec,total", [(32, 16, 300), (32, 8, 100), (32, 4, 200), (16, 8, 100), (16, 4, 200)])
def test_two_tier_precision_pairs(main, sec, total):
    t = run_tier_parity(SMALL_ROWS, 16, 2, main, sec, total, [48, 5, 64], 36, check_state_every=3)
    assert t["c2"] > 0 and t["ev1"] > 0, t


def test_two_tier_batch_of_one_and_ragged():
    run_tier_parity(SMALL_ROWS, 16, 2, 8, 4, 160, [1, 2, 31, 64, 7, 100], 90, check_state_every=5)


def test_two_tier_larger_batches_skewed_tables():
    t = run_tier_parity(SKEW_ROWS, 64, 2, 32, 8, 3000, [512], 20, check_state_every=4)
    assert t["ev1"] > 0, t


def test_two_tier_flush_rule():
    """A hot trace fills bucket 26 of C1 to 95 % of its capacity; with high_agghit_threshold above 26
    odd tables keep inserting into C1, so the flush rule fires there.  C1 then falls below capacity
    and the not-full routing (everything to C1, C2 untouched) resumes until it is full again."""
    t = run_tier_parity(SMALL_ROWS, 16, 2, 8, 4, 30, [4, 16], 260, check_state_every=5, alpha=2.5, high_thres=27)
    assert t["fl1"] > 0 and t["ev1"] > 0 and t["ev2"] > 0, t


def test_three_tier_c3_substitution():
    t = run_tier_parity(SMALL_ROWS, 16, 3, 8, 4, 200, [64, 9], 60, prop="45-45-10", check_state_every=2)
    assert t["c3"] > 0 and t["ev1"] > 0 and t["ev2"] > 0, t


def test_three_tier_even_split_dim36():
    t = run_tier_parity(SMALL_ROWS, 36, 3, 8, 4, 150, [32], 50, check_state_every=5)
    assert t["c3"] > 0, t


def test_three_tier_zero_c3_share_is_two_tier():
    """SIZE_PROPORTION with 0 % for C3 disables the third layer (evlfu_8.cpp:81-84)."""
    run_tier_parity(SMALL_ROWS, 16, 3, 8, 4, 200, [64], 20, prop="50-50-0", check_state_every=4)


def test_two_tier_very_large_batch_scan_path():
    """More than 2048 serve-CTAs (B > 16384): ring positions come from the k_scan prefix pass."""
    t = run_tier_parity(SKEW_ROWS, 16, 2, 8, 4, 2000, [17000, 16400], 4, check_state_every=1)
    assert t["ev1"] > 0 and t["c2"] > 0, t


@pytest.mark.parametrize("n_tables", [3, 4, 7])
def test_two_and_three_tier_with_packed_warps(n_tables):
    rows = SMALL_ROWS[2:2 + n_tables]
    t = run_tier_parity(rows, 16, 2, 8, 4, 60, [200, 33], 30, check_state_every=3)
    assert t["ev1"] > 0, t
    run_tier_parity(rows, 16, 3, 8, 4, 90, [150], 30, prop="40-40-20", check_state_every=3)


def test_three_tier_cache_over_a_stored_model_directory(tmp_path):
    """The reference's stored model, unchanged on disk (ev-table-8/binary, ev-table-4/binary, training_config.txt,
    alt-key binaries) -> storage_manager.open_model_dir -> EvStore.from_raw_stores:
    same streams and rows as the oracle, without an fp32 copy of the tables."""
    import numpy as np
    import torch
    from helpers import decoded_tables, pkg
    from oracle import tiers as otiers
    p = pkg()
    sm = p.storage_manager
    rows, dim = SMALL_ROWS, 36
    tables = p.workload.make_tables(rows, dim)
    alt = p.workload.make_alt_keys(rows)
    sm.ev_dimension, sm.n_tables = dim, 26
    try:
        for prec in (8, 4):
            sm.load_tables(tables, precisions=(prec,))
            sm.save_ev_tables(str(tmp_path / sm.PRECISION_DIRS[prec]), prec)
        sm.store_training_config(str(tmp_path / "training_config.txt"), {i: i for i in range(26)}, 1, 1, np.array(rows), 13)
        sm.save_alt_keys(str(tmp_path / "alt"), alt)
        sm.close_any_db_conn()
        rows_back, stores, alt_back = sm.open_model_dir(str(tmp_path), (8, 4), dim=dim, alt_path=str(tmp_path / "alt"))
    finally:
        sm.ev_precs, sm.ev_dimension, sm.n_tables = 32, 36, 26
    total, prop, B = 150, "45-45-10", 48
    cfg = p.CacheConfig(n_layers=3, main_precision=8, secondary_precision=4, total_size=total, size_proportion=prop,
                        max_batch=B, record_events=True)
    store = p.EvStore.from_raw_stores(rows_back, dim, cfg, stores, alt_keys=alt_back)
    caps = otiers.capacities(3, 8, 4, total, prop, dim)
    oracle = otiers.BatchTiers(caps, n_layers=3, T=26, alt_keys=alt)
    dec = [decoded_tables(tables, 8), decoded_tables(tables, 4)]
    trace = p.workload.ZipfTrace(rows, seed=21)
    c3 = 0
    try:
        for it in range(40):
            idx = trace.batch(B)
            o, h = store.lookup(torch.from_numpy(idx).cuda())
            torch.cuda.synchronize()
            code, val_tier, st, sr, _agg = oracle.lookup_batch(idx)
            assert (h.cpu().numpy() == code).all(), it
            assert (o.cpu().numpy() == otiers.gather_tier_rows(dec, val_tier, st, sr)).all(), it
            for ti, ot in ((0, oracle.c1), (1, oracle.c2)):
                ev, _fl = store.last_events(ti)
                assert ev.tolist() == ot.evicted, (it, ti)
            c3 += int((code == 3).sum())
        assert c3 > 0
    finally:
        store.close()


@pytest.mark.parametrize("dim,main,sec,total", [(64, 32, 8, 400), (32, 16, 8, 300), (32, 8, 4, 300)])
def test_two_tier_look_ahead_prefetch(dim, main, sec, total):
    """evs_prefetch with two tiers: the staged row is tagged with the tier the look-ahead guessed (odd tables C1, even
    tables C2 once C1 is full); a wrong guess is fetched the usual way.  Everything stays bit-exact."""
    t = run_tier_parity(SKEW_ROWS, dim, 2, main, sec, total, [256, 64, 700], 30, check_state_every=3, prefetch=True)
    assert t["c2"] > 0 and t["ev1"] > 0, t


def test_three_tier_look_ahead_prefetch():
    t = run_tier_parity(SMALL_ROWS, 32, 3, 8, 4, 300, [64, 9, 128], 40, prop="45-45-10", check_state_every=2, prefetch=True)
    assert t["c3"] > 0 and t["ev1"] > 0, t
