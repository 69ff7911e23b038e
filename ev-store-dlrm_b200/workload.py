"""Synthetic Criteo-shaped tables and Zipf index traces.

Tables follow DLRM's create_emb initialisation (dlrm_s_pytorch_C1_C2_C3.py:446-449):
U(-1/sqrt(rows), 1/sqrt(rows)), clipped into the dense range of the 16-bit codec.
Indices: per table, rank r drawn with p ~ r^-alpha, mapped to a row by a fixed permutation;
layout lS_i[n_tables, B] int64 (collate_wrapper_criteo_offset, dlrm_data_pytorch.py:397).
"""
from __future__ import annotations

import os

import numpy as np

# logs/sample-inference-criteo_kaggle_all.txt:31
KAGGLE_ROWS = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
               5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]
# the reference's "13 %" operating point (cache_manager.cpp:16) applied to all Kaggle rows
KAGGLE_CACHE_ROWS = int(sum(KAGGLE_ROWS) * 0.13)
# MLPerf DLRM Criteo-Terabyte cardinalities capped at 40M (not in the reference repo)
TERABYTE_ROWS = [39884406, 39043, 17289, 7420, 20263, 3, 7120, 1543, 63, 38532951, 2953546, 403346, 10, 2208,
                 11938, 155, 4, 976, 14, 39979771, 25641295, 39664984, 585935, 12972, 108, 36]


def scaled_rows(rows, scale: float, floor: int = 3):
    """Shrink a cardinality list for tests / small hosts, keeping the skew."""
    return [max(floor, int(r * scale)) if r > 1000 else r for r in rows]


def make_table(t: int, rows: int, dim: int, seed: int = 1234) -> np.ndarray:
    rng = np.random.default_rng(seed + t)
    bound = np.float32(min(1.0 / np.sqrt(rows), 0.6499))
    w = rng.random(size=(rows, dim), dtype=np.float32)
    w *= np.float32(2) * bound
    w -= bound
    return w


def make_tables(rows, dim: int, seed: int = 1234):
    return [make_table(t, r, dim, seed) for t, r in enumerate(rows)]


class ZipfTrace:
    """Reproducible stream of index batches lS_i[n_tables, B]."""

    def __init__(self, rows, alpha: float = 1.05, seed: int = 42, perm_seed: int = 7):
        self.rows = list(rows)
        self.rng = np.random.default_rng(seed)
        self.cdfs, self.perms = [], []
        for t, n in enumerate(self.rows):
            w = np.arange(1, n + 1, dtype=np.float64) ** (-alpha)
            c = np.cumsum(w)
            self.cdfs.append(c / c[-1])
            self.perms.append(np.random.default_rng(perm_seed + t).permutation(n).astype(np.int64))

    def batches(self, n: int, B: int) -> np.ndarray:
        """n consecutive batches at once: int64 [n, n_tables, B]."""
        out = np.empty((n, len(self.rows), B), dtype=np.int64)
        # the uniform draws are taken table by table from the one generator (the sequence is part of the
        # golden fixtures); the inverse-CDF searches are independent and run on a thread pool
        u = [self.rng.random(n * B) for _ in self.rows]

        def one(t):
            r = np.searchsorted(self.cdfs[t], u[t], side="left")
            out[:, t, :] = self.perms[t][np.minimum(r, self.rows[t] - 1)].reshape(n, B)

        if n * B >= (1 << 16):
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=min(len(self.rows), os.cpu_count() or 1)) as ex:
                list(ex.map(one, range(len(self.rows))))
        else:
            for t in range(len(self.rows)):
                one(t)
        return out

    def batch(self, B: int) -> np.ndarray:
        out = np.empty((len(self.rows), B), dtype=np.int64)
        for t, n in enumerate(self.rows):
            r = np.searchsorted(self.cdfs[t], self.rng.random(B), side="left")
            out[t] = self.perms[t][np.minimum(r, n - 1)]
        return out


def make_alt_keys(rows, seed: int = 11):
    """A stand-in for the offline kNN alt-key tables (script/approximate_embedding/...):
    alt_key = alt_row*100 + alt_table(1-based) (convert_altkeys_to_binary.py:50).  Each row's
    alternative is a more popular row (lower Zipf rank) of the same table under ZipfTrace's
    permutation -- the property the reference's "most popular neighbour" pick provides."""
    out = []
    for t, n in enumerate(rows):
        perm = np.random.default_rng(7 + t).permutation(n).astype(np.int64)
        rank_of_row = np.empty(n, dtype=np.int64)
        rank_of_row[perm] = np.arange(n)
        alt_rank = rank_of_row // 2
        alt_row = perm[alt_rank]
        out.append((alt_row * 100 + (t + 1)).astype(np.uint32))
    return out


# ---- closed-form tables ------------------------------------------------------------------------------------
# value(table, row, col) is a pure function, so ANY process can say what a backing row holds without owning the
# table: a rank of a table-wise sharded run checks the rows it RECEIVED from its peers against it
# (bench_sharded.py "verified"), and a 40 M x 64 table is generated on the GPU in a second instead of by a host RNG.
# Integer mixing in int64 with 32-bit masks (no overflow anywhere), then k / 2^24 -> U(-bound, bound) in float32:
# every step is exact or a single IEEE rounding, so numpy, torch-CPU and torch-CUDA give identical bits.
def _synth_bound(rows: int) -> float:
    return float(np.float32(min(1.0 / np.sqrt(rows), 0.6499)))


def synth_rows(table_id: int, row_ids, dim: int, rows_in_table: int):
    """fp32 [n, dim] rows of the closed-form table `table_id`; row_ids: int64 numpy array or torch tensor (any device)."""
    is_np = isinstance(row_ids, np.ndarray)
    if is_np:
        r = row_ids.astype(np.int64).reshape(-1, 1)
        c = np.arange(dim, dtype=np.int64).reshape(1, -1)
    else:
        import torch
        r = row_ids.to(torch.int64).reshape(-1, 1)
        c = torch.arange(dim, dtype=torch.int64, device=row_ids.device).reshape(1, -1)
    x = ((r * 73856093) ^ (c * 19349663) ^ (int(table_id) * 83492791 + 12345)) & 0xFFFFFFFF
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & 0xFFFFFFFF
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & 0xFFFFFFFF
    x = (x ^ (x >> 16)) >> 8                                   # 24 bits
    b = _synth_bound(rows_in_table)
    if is_np:
        u = x.astype(np.float32) * np.float32(1.0 / 16777216.0)
        return (u * np.float32(2.0) - np.float32(1.0)) * np.float32(b)
    import torch
    u = x.to(torch.float32) * (1.0 / 16777216.0)
    return (u * 2.0 - 1.0) * b


def synth_table_pinned(table_id: int, rows: int, dim: int, device, chunk_rows: int = 1 << 21):
    """The whole closed-form table as a pinned host tensor [rows, dim] fp32, generated on `device` chunk by chunk."""
    import torch
    out = torch.empty((rows, dim), dtype=torch.float32, pin_memory=True)
    for r0 in range(0, rows, chunk_rows):
        r1 = min(rows, r0 + chunk_rows)
        ids = torch.arange(r0, r1, dtype=torch.int64, device=device)
        out[r0:r1].copy_(synth_rows(table_id, ids, dim, rows), non_blocking=True)
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)
    return out


def synth_table_mapped(table_id: int, rows: int, dim: int, device, chunk_rows: int = 1 << 21):
    """The same table in ``host_rows`` memory (evs_host_alloc: host memory the device maps with large pages); numpy [rows, dim]."""
    import torch
    from .cache_manager import host_rows
    dev = torch.device(device)
    out = host_rows((rows, dim), np.float32, dev.index or 0)
    view = torch.from_numpy(out)
    for r0 in range(0, rows, chunk_rows):
        r1 = min(rows, r0 + chunk_rows)
        ids = torch.arange(r0, r1, dtype=torch.int64, device=dev)
        view[r0:r1].copy_(synth_rows(table_id, ids, dim, rows))
    return out
