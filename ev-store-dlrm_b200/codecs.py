"""Host-side codecs for the backing store (numpy).

Produce the per-precision row files the reference reads (binary/ev-table-N.bin of
ev-table, ev-table-16, ev-table-8, ev-table-4) from fp32 tables, i.e. the job of
script/reduce_precision.py + script/convert_ev_to_binary.py, vectorised.  The GPU kernels
decode these formats (csrc/evs_codec.cuh); ``decode_rows`` is the same map on the host for
``storage_manager.get_val_from_storage``.
"""
from __future__ import annotations

import numpy as np

PRECISIONS = (32, 16, 8, 4)
# evlfu_4.hpp:46 value_mapping; code 15 cannot be produced by the quantiser and decodes to -1
_LUT4 = np.array([1, 0.8, 0.6, 0.4, 0.0625, 0.00390625, 0.0000153, 0, -0.0000153, -0.00390625,
                  -0.0625, -0.4, -0.6, -0.8, -1, -1], dtype=np.float32)
_POS_EDGES = np.array([0.00025, 0.015, 0.25, 0.4, 0.6, 0.8])      # codes 5,4,3,2,1,0 from the top
_NEG_EDGES = np.array([-1.0, -0.8, -0.6, -0.4, -0.25, -0.015, -0.00025])


def row_bytes(dim: int, prec: int) -> int:
    if (dim * prec) % 8:
        raise ValueError("dim*prec must be a whole number of bytes")
    return dim * prec // 8


def encode_table(table: np.ndarray, prec: int) -> np.ndarray:
    """fp32 [rows, dim] -> C-contiguous raw rows (float32 / uint16 / uint8 / packed nibbles)."""
    t = np.ascontiguousarray(table, dtype=np.float32)
    if prec == 32:
        return t
    v = t.astype(np.float64)
    if prec == 16:    # reduce_precision.py:26-51
        below = np.trunc(-100.0 * (0.65 + v)).astype(np.int64)
        below += (below % 2 == 0)
        above = np.trunc(100.0 * (v - 0.65)).astype(np.int64)
        above -= (above % 2 == 1)
        mid = np.trunc((v + 0.65) / 1.3 * 65000).astype(np.int64)
        q = np.where(v < -0.65, 65000 + below, np.where(v > 0.65, 65000 + above, mid))
        return np.ascontiguousarray(q.astype(np.uint16))
    if prec == 8:     # reduce_precision.py:270 (round half to even)
        return np.ascontiguousarray(np.rint((v + 1.0) / 2.0 * 254.0).astype(np.uint8))
    if prec == 4:     # reduce_precision.py:140-172, packed high nibble first (:316-322)
        pos = 6 - np.searchsorted(_POS_EDGES, v, side="right")           # v >= edge counts
        neg = 15 - np.searchsorted(_NEG_EDGES, v, side="right")          # edges > v count
        neg = np.where(v >= -0.00025, 8, neg)
        code = np.where(v == 0, 7, np.where(v > 0, pos, neg)).astype(np.uint8)
        if code.shape[1] % 2:
            raise ValueError("4-bit rows need an even dim")
        return np.ascontiguousarray(code[:, 0::2] * 16 + code[:, 1::2])
    raise ValueError(f"precision {prec}")


def decode_rows(raw: np.ndarray, prec: int) -> np.ndarray:
    """Raw rows -> fp32, bit-exact with the C++ decoders (evlfu_16.cpp:332, evlfu_8.cpp:370, evlfu_4.cpp:319)."""
    if prec == 32:
        return np.asarray(raw, dtype=np.float32)
    if prec == 16:
        q = np.asarray(raw).astype(np.int64)
        f = q.astype(np.float32)
        mid = (f.astype(np.float64) * 0.00002 - 0.65).astype(np.float32)
        far = 0.65 + ((q - 65000).astype(np.float32) / np.float32(100)).astype(np.float64)
        far = np.where(q & 1, -far, far).astype(np.float32)
        return np.where(q > 65000, far, mid)
    if prec == 8:
        f = np.asarray(raw).astype(np.float32)
        return (f / np.float32(254)) * np.float32(2) - np.float32(1)
    if prec == 4:
        b = np.asarray(raw, dtype=np.uint8)
        out = np.empty(b.shape[:-1] + (2 * b.shape[-1],), dtype=np.float32)
        out[..., 0::2] = _LUT4[b >> 4]
        out[..., 1::2] = _LUT4[b & 15]
        return out
    raise ValueError(f"precision {prec}")
