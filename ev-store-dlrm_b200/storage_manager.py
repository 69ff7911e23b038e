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


# ---- alternative keys (C3) ---------------------------------------------------------------------
def load_alt_keys(alt_path, rows=None):
    """Read <alt_path>/binary/ev-table-{1..n}.bin: one big-endian uint32 per row,
    alt_key = alt_row * 100 + alt_table (1-based) -- written by script/convert_altkeys_to_binary.py:27-57
    (struct.pack('>I')), read by APRX_EV::get_from_file_as_uint (aprx_embedding.cpp:222-250).  Returns per
    table a host-order uint32 array, the form evs_config.alt_keys takes."""
    out = []
    for t in range(n_tables):
        path = os.path.join(alt_path, BINARY_DIR_NAME, f"ev-table-{t + 1}.bin")
        if not os.path.exists(path):
            raise StorageError(f"ERROR: cannot open {path}")
        if os.path.getsize(path) % 4:
            raise StorageError(f"{path}: size is not a multiple of 4 bytes")
        a = np.fromfile(path, dtype=">u4").astype(np.uint32)
        if rows is not None and len(a) != int(rows[t]):
            raise StorageError(f"{path}: {len(a)} alternative keys for a table of {int(rows[t])} rows")
        tab = a % 100
        if len(a) and (tab.min() < 1 or tab.max() > n_tables):
            raise StorageError(f"{path}: alternative key with table id outside 1..{n_tables}")
        out.append(a)
    return out


def save_alt_keys(alt_path, alt_keys):
    """The inverse: per table uint32 alt keys -> big-endian binary/ev-table-N.bin."""
    d = os.path.join(alt_path, BINARY_DIR_NAME)
    os.makedirs(d, exist_ok=True)
    for t, a in enumerate(alt_keys):
        np.asarray(a, dtype=">u4").tofile(os.path.join(d, f"ev-table-{t + 1}.bin"))


def alt_keys_from_text(lines):
    """One "<tableId>-<rowId>" per row (the CSV the kNN notebooks emit) -> uint32 alt keys, as
    convert_altkeys_to_binary (script/convert_altkeys_to_binary.py:41-52) computes them."""
    out = np.empty(len(lines), dtype=np.uint32)
    for i, ln in enumerate(lines):
        t, r = ln.strip().split("-")
        out[i] = int(t) + 100 * int(r)
    return out


# ---- training_config.txt -------------------------------------------------------------------------
def read_training_config(file_path):
    """evstore_utils.read_training_config (evstore_utils.py:43-53): line 0 is a caption, then
    table_feature_map, nbatches, nbatches_test, ln_emb (table cardinalities), m_den."""
    import ast
    with open(file_path) as f:
        lines = [line.rstrip() for line in f]
    if len(lines) < 6:
        raise StorageError(f"{file_path}: expected 6 lines, found {len(lines)}")
    table_feature_map = ast.literal_eval(lines[1])
    nbatches = int(lines[2])
    nbatches_test = int(lines[3])
    ln_emb = np.array(ast.literal_eval(lines[4]))
    m_den = int(lines[5])
    return table_feature_map, nbatches, nbatches_test, ln_emb, m_den


def store_training_config(file_path, table_feature_map, nbatches, nbatches_test, ln_emb, m_den):
    """evstore_utils.store_training_config (evstore_utils.py:31-41), same text."""
    with open(file_path, "w") as f:
        f.write("The order of the arguments: table_feature_map, nbatches, nbatches_test, ln_emb, m_den\n")
        f.write(str(table_feature_map) + "\n")
        f.write(str(nbatches) + "\n")
        f.write(str(nbatches_test) + "\n")
        f.write(str(np.asarray(ln_emb).tolist()) + "\n")
        f.write(str(m_den) + "\n")


def open_model_dir(model_dir, precisions, dim=None, alt_path=None, mapped_device=None):
    """Everything the GPU cache needs from the reference's stored model, unchanged on disk:
    <model_dir>/training_config.txt (cardinalities), <model_dir>/<ev-table[-16|-8|-4]>/binary/ev-table-N.bin per
    precision (evlfu_32.hpp:61, evlfu_16.hpp, evlfu_8.hpp:58, evlfu_4.hpp:61) and, for three layers, the alt-key
    directory.  Returns (rows, {precision: [raw tables]}, alt_keys | None) -- pass them to
    ``EvStore.from_raw_stores``.  mapped_device: a CUDA device index -> the rows are moved into ``evs_host_alloc`` memory
    (host memory that device maps with large pages: about twice the zero-copy miss-fetch rate over multi-GB tables)."""
    global ev_precs, ev_dimension, n_tables
    cfg = os.path.join(model_dir, "training_config.txt")
    rows = None
    if os.path.exists(cfg):
        rows = [int(x) for x in read_training_config(cfg)[3]]
        n_tables = len(rows)
    if dim is not None:
        ev_dimension = int(dim)
    stores = {}
    keep = ev_precs
    try:
        for p in precisions:
            ev_precs = int(p)
            stores[int(p)] = load_ev_table_into_emb_stor(os.path.join(model_dir, PRECISION_DIRS[int(p)]), rows=rows)
    finally:
        ev_precs = keep
    rows = rows or list(_rows)
    alt = load_alt_keys(alt_path, rows) if alt_path else None
    if mapped_device is not None:
        from .cache_manager import to_host_rows
        stores = {p: [to_host_rows(t, int(mapped_device)) for t in tabs] for p, tabs in stores.items()}
        if alt is not None:
            alt = [to_host_rows(a, int(mapped_device)) for a in alt]
    return rows, stores, alt


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
