"""Storage codecs of the reference (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Quantisers restate script/reduce_precision.py (the offline tool that defines the formats);
dequantisers restate the C++ ``chars_buffer_to_floats`` of each tier, which is what the
reference's lookup path actually returns:

  16-bit  quantise  reduce_precision.py:26-51     decode  evlfu_16.cpp:332-356
  8-bit   quantise  reduce_precision.py:270       decode  evlfu_8.cpp:370-378
  4-bit   quantise  reduce_precision.py:140-172   decode  evlfu_4.cpp:319-341 + evlfu_4.hpp:46
          packing   reduce_precision.py:316-322 (first value in the high nibble)

The quantisers are evaluated in float64 on float32 inputs: the reference applies Python
lambdas to a float32 pandas column, which under its NumPy 1.x promotes to float64.
"""
import numpy as np

LUT4 = np.array([1, 0.8, 0.6, 0.4, 0.0625, 0.00390625, 0.0000153, 0, -0.0000153, -0.00390625,
                 -0.0625, -0.4, -0.6, -0.8, -1, -1], dtype=np.float32)   # entry 15: see DESIGN.md


def quant16(x):
    v = np.asarray(x, dtype=np.float32).astype(np.float64)
    dense = ((v + 0.65) / 1.3 * 65000).astype(np.int64)             # int() truncates
    lo = (-100 * (0.65 + v)).astype(np.int64)
    lo = np.where(lo % 2 == 0, lo + 1, lo)
    hi = (100 * (v - 0.65)).astype(np.int64)
    hi = np.where(hi % 2 == 1, hi - 1, hi)
    q = np.where(v < -0.65, 65000 + lo, np.where(v > 0.65, 65000 + hi, dense))
    return q.astype(np.uint16)


def dequant16(q):
    q = np.asarray(q).astype(np.int64)
    dense = (q.astype(np.float32).astype(np.float64) * 0.00002 - 0.65).astype(np.float32)
    diff = (q - 65000).astype(np.float32) / np.float32(100)
    big = 0.65 + diff.astype(np.float64)
    big = np.where(q % 2 == 1, -big, big).astype(np.float32)
    return np.where(q > 65000, big, dense).astype(np.float32)


def quant8(x):
    v = np.asarray(x, dtype=np.float32).astype(np.float64)
    return np.rint(((v + 1) / 2) * 254).astype(np.uint8)             # Python round == half-to-even


def dequant8(q):
    q = np.asarray(q).astype(np.float32)
    return ((q / np.float32(254)) * np.float32(2) - np.float32(1)).astype(np.float32)


def quant4_codes(x):
    v = np.asarray(x, dtype=np.float32).astype(np.float64)
    pos = np.select([v >= 0.8, v >= 0.6, v >= 0.4, v >= 0.25, v >= 0.015, v >= 0.00025], [0, 1, 2, 3, 4, 5], 6)
    neg = np.select([v >= -0.00025, v < -1, v < -0.8, v < -0.6, v < -0.4, v < -0.25, v < -0.015],
                    [8, 15, 14, 13, 12, 11, 10], 9)
    return np.where(v == 0, 7, np.where(v > 0, pos, neg)).astype(np.uint8)


def pack4(codes):
    codes = np.asarray(codes, dtype=np.uint8)
    assert codes.shape[-1] % 2 == 0
    return (codes[..., 0::2] * 16 + codes[..., 1::2]).astype(np.uint8)


def dequant4_packed(b):
    b = np.asarray(b, dtype=np.uint8)
    out = np.empty(b.shape[:-1] + (b.shape[-1] * 2,), dtype=np.float32)
    out[..., 0::2] = LUT4[b >> 4]
    out[..., 1::2] = LUT4[b & 15]
    return out


def quantize_table(table, prec):
    """fp32 [rows, d] -> raw rows in the layout of binary/ev-table-N.bin at `prec` bits."""
    if prec == 32:
        return np.ascontiguousarray(table, dtype=np.float32)
    if prec == 16:
        return quant16(table)
    if prec == 8:
        return quant8(table)
    if prec == 4:
        return pack4(quant4_codes(table))
    raise ValueError(prec)


def dequantize_rows(raw, prec):
    if prec == 32:
        return np.asarray(raw, dtype=np.float32)
    if prec == 16:
        return dequant16(raw)
    if prec == 8:
        return dequant8(raw)
    if prec == 4:
        return dequant4_packed(raw)
    raise ValueError(prec)
