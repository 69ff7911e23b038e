"""Backing store of the embedding rows: the counterpart of the reference's
emb_storage/storage_manager.py, same module-level API (storage_type, ev_precs,
load_ev_table_into_emb_stor, get_val_from_storage, get_arr_val_from_storage,
request_to_emb_storage, close_any_db_conn; emb_storage/storage_manager.py:20-186).

The reference reads one row per call from RocksDB / SQLite / a file (`seek` + `read` of
binary/ev-table-N.bin, emb_storage/file_read.py:27-33).  Here the store is one row-major array per
table and precision in the same on-disk layout (script/convert_ev_to_binary.py), kept in pinned
host memory; the GPU cache reads missing rows from it zero-copy (csrc/evs_kernels.cuh k_fetch) and
``request_batch_to_emb_storage`` gathers whole index batches from it on the GPU
(evs_embedding_bag).  The storage *engines* of the reference (RocksDB, SQLite) are out of scope
(DESIGN.md section 8); their enum values are accepted and served from the same arrays.
"""
from __future__ import annotations

import os

import numpy as np

from . import codecs


class EmbStorage:                      # emb_storage/storage_manager.py:20-27
    DUMMY = 1
    ROCKSDB = 2
    FILEPY = 3
    MMAPFILEPY = 4
    SQLITE = 5
    FILEC = 6
    CPP_CACHING_LAYER = 7


BINARY_DIR_NAME = "binary/"            # emb_storage/file_read.py:5
PRECISION_DIRS = {32: "ev-table", 16: "ev-table-16", 8: "ev-table-8", 4: "ev-table-4"}

storage_type = EmbStorage.DUMMY
ev_precs = 32
ev_dimension = 36                      # EV_DIMENSION (emb_storage/file_read.py:8); set before loading files
n_tables = 26
training_config_path = "/should/point/to/training_config_path"

_raw: dict[int, list[np.ndarray]] = {}     # precision -> per table raw rows
_rows: list[int] = []


class StorageError(RuntimeError):
    pass


def load_tables(tables_fp32, precisions=(32,)):
    """In-memory load (the reference's storage_dummy.load reads CSVs into python lists,
    emb_storage/storage_dummy.py:22-95): quantise the fp32 tables with the reference's quantisers."""
    global _rows
    _raw.clear()
    _rows = [int(t.shape[0]) for t in tables_fp32]
    for p in precisions:
        _raw[p] = [codecs.encode_table(t, p) for t in tables_fp32]
    return _raw


def load_ev_table_into_emb_stor(ev_path_c1, overwrite_db=True, rows=None):
    """Open <ev_path>/binary/ev-table-{1..n}.bin at ``ev_precs`` bits (file_read.open_files_as_binary,
    emb_storage/file_read.py:11-25).  Row counts come from the file sizes."""
    global _rows
    rb = codecs.row_bytes(ev_dimension, ev_precs)
    tabs = []
    for t in range(n_tables):
        path = os.path.join(ev_path_c1, BINARY_DIR_NAME, f"ev-table-{t + 1}.bin")
        if not os.path.exists(path):
            raise StorageError(f"ERROR: cannot open {path}")
        size = os.path.getsize(path)
        if size % rb:
            raise StorageError(f"{path}: size {size} is not a multiple of the row size {rb} (dim {ev_dimension}, {ev_precs} bits)")
        a = np.fromfile(path, dtype=np.uint8).reshape(size // rb, rb)
        if ev_precs == 32:
            a = a.view(np.float32)
        elif ev_precs == 16:
            a = a.view(np.uint16)
        tabs.append(np.ascontiguousarray(a))
    _raw[ev_precs] = tabs
    _rows = [int(a.shape[0]) for a in tabs]
    if rows is not None and list(rows) != _rows:
        raise StorageError(f"row counts {_rows} differ from the training configuration {list(rows)}")
    return tabs


def save_ev_tables(ev_path, precision=None):
    """Write the loaded tables as binary/ev-table-N.bin (script/convert_ev_to_binary.py:31-56)."""
    p = precision or ev_precs
    d = os.path.join(ev_path, BINARY_DIR_NAME)
    os.makedirs(d, exist_ok=True)
    for t, a in enumerate(_raw[p]):
        np.ascontiguousarray(a).tofile(os.path.join(d, f"ev-table-{t + 1}.bin"))


def raw_tables(precision=None):
    p = precision or ev_precs
    if p not in _raw:
        raise StorageError(f"no table loaded at {p} bits")
    return _raw[p]


def get_val_from_storage(tableId, rowId):
    """tableId is 1-based, rowId 0-based (emb_storage/storage_manager.py:73-94); returns the row as
    a tuple of python floats like file_read.get (struct.unpack)."""
    tabs = raw_tables()
    if not (1 <= tableId <= len(tabs)) or not (0 <= rowId < tabs[tableId - 1].shape[0]):
        raise StorageError(f"ERROR: key {tableId}-{rowId} is outside the embedding storage")
    return tuple(codecs.decode_rows(tabs[tableId - 1][rowId], ev_precs).tolist())


def get_arr_val_from_storage(keys):
    return [get_val_from_storage(t, r) for t, r in keys]


def request_to_emb_storage(group_rowIds, use_gpu=False):
    """One sample straight from the store, bypassing the cache (emb_storage/storage_manager.py:125-139):
    returns (-1, list of n_tables FloatTensor[1, dim])."""
    import torch
    emb_weights = []
    for i, rowId in enumerate(group_rowIds):
        t = torch.FloatTensor([get_val_from_storage(i + 1, int(rowId))])
        if use_gpu:
            t = t.to(torch.device("cuda:0"))
        emb_weights.append(t)
    return -1, emb_weights


def close_any_db_conn():
    _raw.clear()
    print("All db connections are closed!")
