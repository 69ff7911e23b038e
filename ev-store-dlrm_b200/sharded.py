"""Table-wise sharding of the lookup path over the GPUs of one box (BASELINE configs[4]).

Semantics of the reference's distributed forward (dlrm_s_pytorch.py:529-586 with
extend_distributed.py): rank r owns a contiguous slice of the tables (``get_split_lengths`` /
``get_my_slice``, extend_distributed.py:47-62), looks the FULL batch up in its tables, and one
all-to-all turns the table-sharded pooled rows ``[B, T_local*d]`` into batch-sharded ones
``[B/size, 26*d]`` (``ext_dist.alltoall``, extend_distributed.py:541-576; ``all_to_all_single``
at :414).  One process per GPU, ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU
tests of this host logic).

EvLFU's only cross-table coupling is agg_hit, the number of a sample's keys that hit.  The
reference never ran its cache sharded; here the ranks probe (``evs_probe_batch``), all-reduce the
per-sample hit counts (uint8[B], 2-16 KB) and pass the sum to ``evs_lookup_batch`` as ``agg_in``,
so the union of the ranks' hit / eviction streams equals the single-GPU policy's.
"""
from __future__ import annotations


def get_split_lengths(n: int, size: int):
    """extend_distributed.py:54-62: the first n % size ranks get one extra table."""
    k, m = divmod(n, size)
    return [k + 1] * m + [k] * (size - m) if m else [k] * size


def get_my_slice(n: int, rank: int, size: int) -> slice:
    """extend_distributed.py:47-51."""
    k, m = divmod(n, size)
    return slice(rank * k + min(rank, m), (rank + 1) * k + min(rank + 1, m), 1)


class ShardedLookup:
    """One rank's part of the sharded lookup.  ``store`` serves this rank's tables
    (``EvStore`` built with table_base / n_tables_total; any object with ``probe`` and ``lookup``
    of the same signatures works, which is how the CPU tests drive the host logic)."""

    def __init__(self, store, n_tables_total: int, dim: int, rank: int, world: int, group=None, exact_agg: bool = True,
                 transport: str = "nccl", batch_max: int = 0):
        """transport "nccl": torch.distributed all-reduce + all_to_all_single (any backend; what the reference does).
        transport "p2p": the exchange is fused into the kernels over NVLink peer memory (evs_shard_*); needs
        ``batch_max`` (the global batch) and a store with shard_create / shard_connect / shard_lookup."""
        self.store, self.T, self.dim = store, n_tables_total, dim
        self.rank, self.world, self.group = rank, world, group
        self.exact_agg = exact_agg
        self.transport = transport if world > 1 else "nccl"
        if self.transport == "p2p":
            import torch.distributed as dist
            mine = store.shard_create(rank, world, batch_max)
            handles = [None] * world
            dist.all_gather_object(handles, mine, group=group)
            store.shard_connect(handles)
            dist.barrier(group=group)
        self.splits = get_split_lengths(n_tables_total, world)
        self.my = get_my_slice(n_tables_total, rank, world)
        self.T_local = self.splits[rank]
        self._bufs = {}

    def _buffers(self, B, device):
        import torch
        key = (B, str(device))
        if key not in self._bufs:
            Bl = B // self.world
            send = torch.empty((B, self.T_local, self.dim), dtype=torch.float32, device=device)
            recv = torch.empty((Bl * self.T * self.dim,), dtype=torch.float32, device=device)
            out = torch.empty((Bl, self.T, self.dim), dtype=torch.float32, device=device)
            hit = torch.empty((B, self.T_local), dtype=torch.uint8, device=device)
            agg = torch.empty((B,), dtype=torch.uint8, device=device)
            self._bufs[key] = (send, recv, out, hit, agg)
        return self._bufs[key]

    def lookup(self, lS_i_local):
        """lS_i_local: int64 [T_local, B], this rank's tables for the whole batch (B % world == 0).
        Returns (ly [B/world, 26, dim] for this rank's slice of the batch, hit [B, T_local])."""
        import torch
        import torch.distributed as dist
        T_local, B = lS_i_local.shape
        assert T_local == self.T_local and B % self.world == 0
        if self.transport == "p2p":
            return self.store.shard_lookup(lS_i_local)
        send, recv, out, hit, agg = self._buffers(B, lS_i_local.device)
        agg_in = None
        if self.exact_agg and self.world > 1:
            self.store.probe(lS_i_local, agg_out=agg)
            dist.all_reduce(agg, op=dist.ReduceOp.SUM, group=self.group)
            agg_in = agg
        self.store.lookup(lS_i_local, out=send, hit=hit, agg_in=agg_in)
        if self.world == 1:
            return send, hit
        Bl = B // self.world
        # send: [B, T_local*d] split along the batch; recv: world blocks [Bl, T_src, d], source-rank major
        in_splits = [Bl * self.T_local * self.dim] * self.world
        out_splits = [Bl * t * self.dim for t in self.splits]
        dist.all_to_all_single(recv, send.view(-1), out_splits, in_splits, group=self.group)
        o = 0
        t0 = 0
        for t in self.splits:
            n = Bl * t * self.dim
            out[:, t0:t0 + t, :] = recv[o:o + n].view(Bl, t, self.dim)
            o += n
            t0 += t
        return out, hit

    def alltoall_bytes(self, B: int) -> int:
        """Bytes this rank sends to its peers per batch."""
        return B * self.T_local * self.dim * 4 * (self.world - 1) // self.world
