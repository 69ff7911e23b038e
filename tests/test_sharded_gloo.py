"""Host logic of the table-wise sharded lookup (ev-store-dlrm_b200/sharded.py) on CPU: two processes,
gloo.  Each rank drives a CPU stand-in for its EvStore (the oracle's batch policy on its own tables);
the test checks the plumbing the product adds around the store: table split, all-reduce of the
per-sample hit counts, all-to-all of the pooled rows and the reassembly into [B/size, 26, d]."""
import os
import sys

import numpy as np
import pytest

from helpers import SMALL_ROWS, pkg

WORLD = 2
DIM, B, CAP, N_BATCH = 16, 32, 300, 12


class OracleStore:
    """CPU test double with EvStore's probe / lookup signatures."""

    def __init__(self, tables_all, cap, ids):
        from oracle.evlfu import BatchEvLFU
        self.o = BatchEvLFU(cap, n_tables=26)
        self.tables, self.ids = tables_all, list(ids)

    def probe(self, lS_i, agg_out=None):
        import torch
        from oracle.evlfu import make_key
        idx = lS_i.numpy()
        cnt = np.array([sum(make_key(self.ids[t], idx[t, s]) in self.o.entries for t in range(idx.shape[0]))
                        for s in range(idx.shape[1])], dtype=np.uint8)
        agg_out.copy_(torch.from_numpy(cnt))
        return agg_out

    def lookup(self, lS_i, out=None, hit=None, agg_in=None):
        import torch
        from oracle.evlfu import gather_rows
        idx = lS_i.numpy()
        h, st, sr, _ = self.o.lookup_batch(idx, agg=None if agg_in is None else agg_in.numpy(), table_ids=self.ids)
        rows = gather_rows(self.tables, st, sr)
        out.copy_(torch.from_numpy(rows))
        hit.copy_(torch.from_numpy(h.astype(np.uint8)))
        return out, hit


def _placement(kind):
    p = pkg()
    return p.sharded.balanced_placement(SMALL_ROWS, WORLD) if kind == "balanced" else p.sharded.contiguous_placement(26, WORLD)


def _simulate(tables, batches, placement):
    """All ranks in one process: what the distributed run must reproduce."""
    stores = [OracleStore(tables, CAP, placement[r]) for r in range(WORLD)]
    outs, hits = [], []
    import torch
    for idx in batches:
        aggs = []
        for r, st in enumerate(stores):
            sl = placement[r]
            a = torch.empty(B, dtype=torch.uint8)
            st.probe(torch.from_numpy(idx[sl]), agg_out=a)
            aggs.append(a.numpy().astype(np.int64))
        agg = torch.from_numpy(sum(aggs).astype(np.uint8))
        full = np.empty((B, 26, DIM), dtype=np.float32)
        hh = np.empty((B, 26), dtype=np.uint8)
        for r, st in enumerate(stores):
            sl = placement[r]
            o = torch.empty((B, len(sl), DIM))
            h = torch.empty((B, len(sl)), dtype=torch.uint8)
            st.lookup(torch.from_numpy(idx[sl]), out=o, hit=h, agg_in=agg)
            full[:, sl] = o.numpy()
            hh[:, sl] = h.numpy()
        outs.append(full)
        hits.append(hh)
    return outs, hits


def _worker(rank, port, q, kind):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    p = pkg()
    tables = p.workload.make_tables(SMALL_ROWS, DIM)
    batches = p.workload.ZipfTrace(SMALL_ROWS, seed=21).batches(N_BATCH, B)
    placement = _placement(kind)
    sl = placement[rank]
    store = OracleStore(tables, CAP, sl)
    sh = p.sharded.ShardedLookup(store, 26, DIM, rank, WORLD, placement=placement)
    res = []
    local = [torch.from_numpy(np.ascontiguousarray(idx[sl])) for idx in batches]
    if kind == "balanced":
        # a queue of batches per call (ShardedLookup.lookup_many; with this transport the batches run one by one)
        for k in range(0, len(local), 5):
            res += [(ly.numpy().copy(), hit.numpy().copy()) for ly, hit in sh.lookup_many(local[k:k + 5])]
    else:
        for li in local:
            ly, hit = sh.lookup(li)
            res.append((ly.numpy().copy(), hit.numpy().copy()))
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["contiguous", "balanced"])
def test_sharded_lookup_two_ranks_gloo(kind):
    """contiguous = the reference's get_my_slice split; balanced = tables spread by row count (non-contiguous ids:
    the reassembly after the all-to-all scatters each rank's block to its tables' columns)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + (7 if kind == "balanced" else 0)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, port, q, kind)) for r in range(WORLD)]
    for pr in procs:
        pr.start()
    got = dict(q.get(timeout=180) for _ in range(WORLD))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = pkg()
    tables = p.workload.make_tables(SMALL_ROWS, DIM)
    batches = p.workload.ZipfTrace(SMALL_ROWS, seed=21).batches(N_BATCH, B)
    placement = _placement(kind)
    if kind == "balanced":
        assert placement != p.sharded.contiguous_placement(26, WORLD)
    outs, hits = _simulate(tables, batches, placement)
    Bl = B // WORLD
    n_hits = 0
    for k in range(N_BATCH):
        for r in range(WORLD):
            ly, hit = got[r][k]
            sl = placement[r]
            assert ly.shape == (Bl, 26, DIM)
            assert np.array_equal(ly, outs[k][r * Bl:(r + 1) * Bl]), f"batch {k} rank {r}: pooled rows after the all-to-all"
            assert np.array_equal(hit, hits[k][:, sl]), f"batch {k} rank {r}: hit map"
            n_hits += int(hit.sum())
    assert n_hits > 0
