"""The reference's cache client (cache_algo/cpp_socket_client.py) over libevstore_b200.so.

Same entry points: ``init_ctypes_lib`` (:63), ``cache_lookup_via_ctypes`` (:119),
``request_to_cpp_cache`` (:129-157), ``print_n_reset_perfect_hit`` (:85).  The reference's library
is configured at compile time and builds its cache when it is dlopen'ed; ours is configured at run
time, so ``init_ctypes_lib`` takes the tables and a ``CacheConfig`` (once) and calls
``evs_legacy_configure``.  After that the four legacy C symbols behave like the reference's:
one sample (26 int32 row ids) per ``ev_lookup`` call, answer in a library-owned float buffer.

``request_batch_to_cpp_cache`` is the batched call the B200 path is built for.  The TCP transport
(use_socket=True; "50 % of the total latency", :132) is out of scope and raises.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native
from .cache_manager import CacheConfig, EvStore

N_EVTable = 26
EV_DIMENSION = 36

cache_manager_cpp = None
_store: EvStore | None = None
emb_weights_in_tensor: list = []        # module-global list reused across calls (:18,154)


def init_ctypes_lib(tables_fp32=None, cfg: CacheConfig | None = None, stores=None, alt_keys=None):
    """dlopen the CUDA library, declare the reference's prototypes (:66-83) and build the
    process-global cache.  Raises if the library is missing: there is no CPU fallback."""
    global cache_manager_cpp, _store, N_EVTable, EV_DIMENSION, emb_weights_in_tensor
    print("Initiating ctypes cache_manager_cpp library (libevstore_b200.so) ...")
    cache_manager_cpp = _native.load_library()
    if tables_fp32 is None:
        return cache_manager_cpp
    cfg = cfg or CacheConfig()
    _store = EvStore.__new__(EvStore)
    _store._init_config(tables_fp32, cfg, stores, alt_keys)
    _native.check(cache_manager_cpp.evs_legacy_configure(ctypes.byref(_store._c)), "evs_legacy_configure")
    _store.handle = ctypes.c_void_p(cache_manager_cpp.evs_legacy_handle())
    _store._owns_handle = False
    N_EVTable, EV_DIMENSION = _store.n_tables, _store.dim
    emb_weights_in_tensor = [None] * N_EVTable
    return cache_manager_cpp


def legacy_store() -> EvStore:
    if _store is None:
        raise RuntimeError("init_ctypes_lib(tables, cfg) has not been called")
    return _store


def print_n_reset_perfect_hit():
    if cache_manager_cpp is not None:
        cache_manager_cpp.print_perfect_hit()


def establish_socket_conn():
    raise NotImplementedError("the TCP transport of cache_manager.cpp is out of scope; use the ctypes path")


def cache_lookup_via_ctypes(group_rowIds):
    p = cache_manager_cpp.ev_lookup((ctypes.c_int * N_EVTable)(*[int(x) for x in group_rowIds]))
    if not p:
        raise _native.EvsError("ev_lookup failed: " + (cache_manager_cpp.evs_last_error() or b"").decode())
    return p


def request_to_cpp_cache(group_rowIds, use_gpu=False, use_socket=False, evstore_gpu_id=0):
    """One sample: returns the module-global list of N_EVTable FloatTensor[1, EV_DIMENSION] (:129-157)."""
    import torch
    if use_socket:
        establish_socket_conn()
    clean = np.ctypeslib.as_array(cache_lookup_via_ctypes(group_rowIds), shape=(N_EVTable, EV_DIMENSION))
    for t in range(N_EVTable):
        ev = torch.from_numpy(clean[t:t + 1].copy())          # the library buffer is overwritten by the next call
        if use_gpu:
            ev = ev.to(torch.device("cuda:" + str(evstore_gpu_id)))
        emb_weights_in_tensor[t] = ev
    return emb_weights_in_tensor


def request_batch_to_cpp_cache(lS_i, out=None, hit=None):
    """The whole index batch lS_i [N_EVTable, B] (int64, CUDA) in one call -> fp32 [B, N_EVTable, dim]."""
    return legacy_store().lookup(lS_i, out=out, hit=hit)
