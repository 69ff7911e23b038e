"""ctypes front of oracle/evlfu_batch.c (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

``CBatchEvLFU`` has the interface of ``oracle.evlfu.BatchEvLFU`` (without the approximate substitution) and is pinned to
it by tests/test_oracle_c_batch.py; it exists so that the CUDA path can be followed at the BASELINE sizes
(tests/test_gpu_full_size.py), where the Python oracle would take hours."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "evlfu_batch.c")
LIB = os.path.join(HERE, "_ref", "libevlfu_batch.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", SRC, "-o", LIB], check=True)
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = C.CDLL(LIB)
        lib.evb_create.argtypes = [C.c_int64, C.c_int, C.c_int64, C.c_double, C.c_double]
        lib.evb_create.restype = C.c_void_p
        lib.evb_destroy.argtypes = [C.c_void_p]
        lib.evb_lookup_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        lib.evb_lookup_batch.restype = C.c_int
        lib.evb_size.argtypes = [C.c_void_p]
        lib.evb_size.restype = C.c_int64
        lib.evb_n_perfect.argtypes = [C.c_void_p]
        lib.evb_n_perfect.restype = C.c_int64
        lib.evb_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


class CBatchEvLFU:
    def __init__(self, capacity: int, n_tables: int = 26, max_keys_per_batch: int = 1 << 16, flush_rate: float = 0.3,
                 perfect_item_cap: float = 0.95):
        self.lib = _load()
        self.cap, self.T, self.max_keys = int(capacity), int(n_tables), int(max_keys_per_batch)
        self.h = self.lib.evb_create(self.cap, self.T, self.max_keys, flush_rate, perfect_item_cap)
        self._ev = np.empty(self.max_keys + 8, dtype=np.int64)
        self._fl = np.empty(int(flush_rate * self.cap) + 8, dtype=np.int64)
        self.evicted, self.flushed, self.n_inserted = [], [], 0

    def lookup_batch(self, idx, agg=None, table_base: int = 0, table_ids=None):
        """idx int64 [Tl, B].  Returns (hit [B, Tl] bool, src_t [B, Tl] int32, src_r [B, Tl] int64, agg [B])."""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        Tl, B = idx.shape
        gid = np.asarray([table_base + t for t in range(Tl)] if table_ids is None else list(table_ids), dtype=np.int32)
        hit = np.empty((B, Tl), dtype=np.uint8)
        agg_out = np.empty(B, dtype=np.int32)
        agg_in = None if agg is None else np.ascontiguousarray(agg, dtype=np.int32)
        n = np.zeros(3, dtype=np.int64)
        rc = self.lib.evb_lookup_batch(self.h, idx.ctypes.data, Tl, B, gid.ctypes.data, None if agg_in is None else agg_in.ctypes.data,
                                       hit.ctypes.data, agg_out.ctypes.data, self._ev.ctypes.data, self._fl.ctypes.data, n.ctypes.data)
        if rc != 0:
            raise RuntimeError("evb_lookup_batch: batch larger than max_keys_per_batch")
        self.evicted = self._ev[:n[0]].tolist()
        self.flushed = self._fl[:n[1]].tolist()
        self.n_inserted = int(n[2])
        src_t = np.tile(gid, (B, 1))
        return hit.astype(bool), src_t, np.ascontiguousarray(idx.T), agg_out.astype(np.int64)

    @property
    def size(self) -> int:
        return int(self.lib.evb_size(self.h))

    @property
    def n_perfect(self) -> int:
        return int(self.lib.evb_n_perfect(self.h))

    def state(self):
        keys = np.empty(self.size + 8, dtype=np.int64)
        off = np.zeros(self.T + 2, dtype=np.int64)
        self.lib.evb_state(self.h, keys.ctypes.data, off.ctypes.data)
        return [keys[off[b]:off[b + 1]].tolist() for b in range(self.T + 1)]

    def close(self):
        if self.h:
            self.lib.evb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
