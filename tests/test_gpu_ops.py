"""The operators either side of the cache, through the reference-named host API: sum pooling
(nn.EmbeddingBag), the no-cache storage path, apply_emb_evstore, interact_features, and the legacy
one-sample ev_lookup surface of libcachemanager.so."""
import ctypes as C

import numpy as np
import pytest

from helpers import SMALL_ROWS, pkg
from oracle import codecs as ocodecs
from oracle.evlfu import BatchEvLFU, gather_rows

pytestmark = pytest.mark.gpu


def _bag_sum(dec, idx, off, w=None):
    """Sequential fp32 sum, j ascending -- the definition of nn.EmbeddingBag(mode='sum')."""
    B = len(off)
    out = np.zeros((B, dec.shape[1]), dtype=np.float32)
    ends = list(off[1:]) + [len(idx)]
    for b in range(B):
        acc = np.zeros(dec.shape[1], dtype=np.float32)
        for j in range(off[b], ends[b]):
            x = dec[idx[j]]
            acc = acc + (np.float32(w[j]) * x if w is not None else x)
        out[b] = acc
    return out


@pytest.mark.parametrize("prec,dim", [(32, 16), (32, 36), (32, 64), (16, 16), (8, 36), (8, 64), (4, 16), (4, 36), (16, 128)])
def test_embedding_bag_matches_sequential_sum(prec, dim):
    import torch
    p = pkg()
    rng = np.random.default_rng(prec * 1000 + dim)
    rows, B = 5000, 257
    table = p.workload.make_table(3, rows, dim)
    raw = ocodecs.quantize_table(table, prec)
    dec = ocodecs.dequantize_rows(raw, prec)
    lens = rng.integers(0, 11, size=B)                       # ragged bags, some empty (--num-indices-per-lookup=10)
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    idx = rng.integers(0, rows, size=int(lens.sum())).astype(np.int64)
    w = rng.random(len(idx), dtype=np.float32)
    t_raw = torch.from_numpy(raw.view(np.uint8) if prec != 32 else raw).cuda()
    d_idx, d_off = torch.from_numpy(idx).cuda(), torch.from_numpy(off).cuda()
    got = p.dlrm_ops.embedding_bag(t_raw, d_idx, d_off, precision=prec, dim=dim)
    assert np.array_equal(got.cpu().numpy(), _bag_sum(dec, idx, off)), "unweighted sum must be bit-exact"
    got_w = p.dlrm_ops.embedding_bag(t_raw, d_idx, d_off, torch.from_numpy(w).cuda(), precision=prec, dim=dim)
    assert np.array_equal(got_w.cpu().numpy(), _bag_sum(dec, idx, off, w)), "weighted sum (mul then add) must be bit-exact"
    if prec == 32:
        ref = torch.nn.functional.embedding_bag(d_idx, torch.from_numpy(table).cuda(), d_off, mode="sum")
        assert torch.allclose(got, ref, rtol=1e-6, atol=1e-6)       # ATen may sum in another order
    assert p.load_library().evs_embedding_bag_status() == 0


def test_embedding_bag_flags_bad_index():
    import torch
    p = pkg()
    table = torch.zeros((10, 16), device="cuda")
    idx = torch.tensor([1, 99], device="cuda")
    off = torch.tensor([0, 1], device="cuda")
    p.dlrm_ops.embedding_bag(table, idx, off)
    assert p.load_library().evs_embedding_bag_status() == -3         # EVS_ERR_INDEX
    assert p.load_library().evs_embedding_bag_status() == 0


def test_apply_emb_evstore_cache_and_storage_paths_and_interaction():
    import torch
    p = pkg()
    dim, B = 16, 96
    tables = p.workload.make_tables(SMALL_ROWS, dim)
    store = p.EvStore(tables, p.CacheConfig(total_size=500, max_batch=B))
    oracle = BatchEvLFU(500)
    p.dlrm_ops.set_store(store)
    trace = p.workload.ZipfTrace(SMALL_ROWS, seed=31)
    for it in range(6):
        idx = trace.batch(B)
        lS_i = torch.from_numpy(idx).cuda()
        lS_o = torch.arange(B, device="cuda").repeat(26, 1)
        ly = p.dlrm_ops.apply_emb_evstore(lS_o, lS_i, use_emb_cache=True)
        o_hit, st, sr, _ = oracle.lookup_batch(idx)
        want = gather_rows(tables, st, sr)
        assert len(ly) == 26 and tuple(ly[0].shape) == (B, dim)
        for k in range(26):
            assert np.array_equal(ly[k].cpu().numpy(), want[:, k]), (it, k)
        assert np.array_equal(p.dlrm_ops.last_hit.cpu().numpy().astype(bool), o_hit)
        # storage path: same rows, the cache is not touched
        ly2 = p.dlrm_ops.apply_emb_evstore(lS_o, lS_i, use_emb_cache=False)
        for k in range(26):
            assert np.array_equal(ly2[k].cpu().numpy(), tables[k][idx[k]]), (it, k)
        # interaction on the list of views (no repack) == torch
        x = torch.randn((B, dim), device="cuda")
        r = p.dlrm_ops.interact_features(x, ly)
        Tm = torch.cat([x.unsqueeze(1), torch.stack(ly, dim=1)], dim=1)
        Z = torch.bmm(Tm, Tm.transpose(1, 2))
        li = torch.tensor([i for i in range(27) for j in range(i)])
        lj = torch.tensor([j for i in range(27) for j in range(i)])
        assert torch.allclose(r, torch.cat([x, Z[:, li, lj]], dim=1), rtol=1e-5, atol=1e-4)
    store.sync()
    assert store.stats()["lookups"] == 6 * B * 26           # the storage path did not count as cache lookups
    with pytest.raises(RuntimeError):
        p.dlrm_ops.apply_emb_evstore(None, torch.zeros((26, 4), dtype=torch.int64))      # CPU indices: no fallback
    store.close()


def test_legacy_ev_lookup_surface():
    """cache_algo/cpp_socket_client.py's calls on our library: one sample per ev_lookup call, the
    answer in a library-owned buffer; equals the batch policy at B = 1; print_perfect_hit resets."""
    p = pkg()
    cl = p.cpp_socket_client
    dim = 36
    tables = p.workload.make_tables(SMALL_ROWS, dim)
    cl.init_ctypes_lib(tables, p.CacheConfig(total_size=400, max_batch=64))
    assert (cl.N_EVTable, cl.EV_DIMENSION) == (26, 36)
    oracle = BatchEvLFU(400)
    trace = p.workload.ZipfTrace(SMALL_ROWS, seed=33).batches(1, 300)[0].T       # [300, 26]
    perfect = 0
    for i, req in enumerate(trace):
        ly = cl.request_to_cpp_cache([int(x) for x in req])
        _h, st, sr, agg = oracle.lookup_batch(req.reshape(26, 1))
        want = gather_rows(tables, st, sr)[0]
        perfect += int(agg[0] == 26)
        assert len(ly) == 26 and tuple(ly[0].shape) == (1, dim)
        for t in range(26):
            assert np.array_equal(ly[t].numpy()[0], want[t]), (i, t)
    lib = cl.cache_manager_cpp
    st = cl.legacy_store().stats()
    assert st["perfect_hits"] == perfect and st["samples"] == 300
    lib.print_perfect_hit()
    lib.test_arr((C.c_int * 5)(1, 2, 3, 4, 5))
    assert lib.ev_lookup_based_on_list_keys((C.c_int * 26)()) == -1
    # batched call on the same process-global cache
    import torch
    idx = p.workload.ZipfTrace(SMALL_ROWS, seed=34).batch(64)
    out, hit = cl.request_batch_to_cpp_cache(torch.from_numpy(idx).cuda())
    o_hit, st_, sr_, _ = oracle.lookup_batch(idx)
    assert np.array_equal(out.cpu().numpy(), gather_rows(tables, st_, sr_))
    assert np.array_equal(hit.cpu().numpy().astype(bool), o_hit)
