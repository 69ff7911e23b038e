"""Table-wise sharding on the GPU: table_base / n_tables_total / probe / agg_in of the C-ABI against
the oracle (two shards on one device), and the full NCCL path when the box has two GPUs."""
import os

import numpy as np
import pytest

from helpers import SKEW_ROWS, SMALL_ROWS, pkg
from oracle.evlfu import BatchEvLFU, gather_rows

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 8])
def test_shards_on_one_device_exact_groupability(world):
    import torch
    p = pkg()
    dim, B, cap = 16, 64, (260 if world == 2 else 70)
    tables = p.workload.make_tables(SMALL_ROWS, dim)
    trace = p.workload.ZipfTrace(SMALL_ROWS, seed=41)
    stores, oracles, slices = [], [], []
    for r in range(world):
        sl = p.sharded.get_my_slice(26, r, world)
        cfg = p.CacheConfig(total_size=cap, max_batch=B, n_tables_total=26, table_base=sl.start, record_events=True)
        stores.append(p.EvStore(tables[sl], cfg))
        oracles.append(BatchEvLFU(cap, n_tables=26))
        slices.append(sl)
    n_ev = 0
    for it in range(30):
        idx = trace.batch(B)
        aggs = [st.probe(torch.from_numpy(np.ascontiguousarray(idx[sl])).cuda()) for st, sl in zip(stores, slices)]
        agg = torch.stack([x.to(torch.int32) for x in aggs]).sum(dim=0).to(torch.uint8)      # the all-reduce
        # the probe equals the oracle's view of the state
        o_agg = sum(np.array([[((sl.start + t) << 40 | int(idx[sl][t, s])) in o.entries for t in range(sl.stop - sl.start)]
                              for s in range(B)]).sum(axis=1) for o, sl in zip(oracles, slices))
        assert np.array_equal(agg.cpu().numpy(), o_agg)
        for st, o, sl in zip(stores, oracles, slices):
            li = torch.from_numpy(np.ascontiguousarray(idx[sl])).cuda()
            out, hit = st.lookup(li, agg_in=agg)
            torch.cuda.synchronize()
            o_hit, s_t, s_r, _ = o.lookup_batch(idx[sl], agg=o_agg, table_base=sl.start)
            assert np.array_equal(hit.cpu().numpy().astype(bool), o_hit), (it, sl)
            assert np.array_equal(out.cpu().numpy(), gather_rows(tables[sl], s_t - sl.start, s_r)), (it, sl)
            ev, fl = st.last_events()
            assert ev.tolist() == o.evicted and fl.tolist() == o.flushed, (it, sl)
            n_ev += len(o.evicted)
            state, n_perfect = st.dump_state()
            assert state == o.state() and n_perfect == o.n_perfect
    assert n_ev > 0
    for st in stores:
        st.close()


CASES = {"small": (SMALL_ROWS, 16, 64, 260, 8), "skew": (SKEW_ROWS, 64, 256, 2000, 10)}


def _worker(rank, world, port, q, backend, transport, same_device, case):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    devno = 0 if same_device else rank
    torch.cuda.set_device(devno)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", devno))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    p = pkg()
    rows, dim, B, cap, n = CASES[case]
    tables = p.workload.make_tables(rows, dim)
    batches = p.workload.ZipfTrace(rows, seed=43).batches(n, B)
    sl = p.sharded.get_my_slice(26, rank, world)
    cfg = p.CacheConfig(total_size=cap, max_batch=B, n_tables_total=26, table_base=sl.start, device=devno)
    store = p.EvStore(tables[sl], cfg)
    sh = p.sharded.ShardedLookup(store, 26, dim, rank, world, transport=transport, batch_max=B)
    res = []
    err = None
    try:
        for idx in batches:
            ly, hit = sh.lookup(torch.from_numpy(np.ascontiguousarray(idx[sl])).cuda())
            torch.cuda.synchronize()
            res.append((ly.cpu().numpy().copy(), hit.cpu().numpy().copy()))
        store.sync()
    except Exception as e:                               # report instead of hanging the peer
        err = repr(e)
    q.put((rank, res, err))
    dist.barrier()
    store.close()
    dist.destroy_process_group()


def _run_sharded(world, backend, transport, same_device, case):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + hash((backend, transport, case))) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, backend, transport, same_device, case)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = {}
    for _ in range(world):
        r, res, err = q.get(timeout=300)
        assert err is None, f"rank {r}: {err}"
        got[r] = res
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = pkg()
    rows, dim, B, cap, n = CASES[case]
    tables = p.workload.make_tables(rows, dim)
    batches = p.workload.ZipfTrace(rows, seed=43).batches(n, B)
    oracles = [BatchEvLFU(cap, n_tables=26) for _ in range(world)]
    slices = [p.sharded.get_my_slice(26, r, world) for r in range(world)]
    Bl = B // world
    n_ev = 0
    for k, idx in enumerate(batches):
        agg = sum(np.array([[((sl.start + t) << 40 | int(idx[sl][t, s])) in o.entries for t in range(sl.stop - sl.start)]
                            for s in range(B)]).sum(axis=1) for o, sl in zip(oracles, slices))
        full = np.empty((B, 26, dim), dtype=np.float32)
        for r, (o, sl) in enumerate(zip(oracles, slices)):
            o_hit, s_t, s_r, _ = o.lookup_batch(idx[sl], agg=agg, table_base=sl.start)
            full[:, sl] = gather_rows(tables[sl], s_t - sl.start, s_r)
            n_ev += len(o.evicted)
            assert np.array_equal(got[r][k][1].astype(bool), o_hit), (k, r)
        for r in range(world):
            assert np.array_equal(got[r][k][0], full[r * Bl:(r + 1) * Bl]), (k, r)
    return n_ev


def test_sharded_lookup_p2p_two_processes_on_one_gpu():
    """The fused exchange (peer-memory stores + epoch flags, evs_shard_*) with both ranks on cuda:0: CUDA IPC
    maps the peer's block just the same; the two processes' kernels time-slice, so this is slow but exact."""
    n_ev = _run_sharded(2, "gloo", "p2p", True, "small")
    assert n_ev > 0


def test_sharded_lookup_nccl_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run_sharded(2, "nccl", "nccl", False, "skew")


def test_sharded_lookup_p2p_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run_sharded(2, "nccl", "p2p", False, "skew")
