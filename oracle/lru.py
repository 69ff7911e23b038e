"""LRU oracles (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

The reference's comparison policy ``/root/reference/cache_algo/LRU.py``:
an ``OrderedDict`` in recency order, a hit moves the key to the MRU end (:30), a miss evicts the LRU
key when ``len >= cap`` (:17-19) and inserts.

* ``SeqLRU``   -- statement-for-statement restatement, one request of n_tables keys at a time, each key
                  probed when its turn comes (LRU.py:40-65).  Pinned by tests/golden/lru_*.npz, generated
                  from the reference module itself (tests/golden/make_golden.py).
* ``BatchLRU`` -- the batch-granular generalisation the CUDA path implements (``policy="lru"``): all keys
                  of a batch are probed against the state at batch start; per key the LAST occurrence
                  (highest position, sample-major / table-minor) decides its place in the recency order;
                  occurrences are applied in position order; then the cache is evicted back down to
                  capacity from the LRU end, never the last inserted key (the reference evicts before it
                  inserts).  With one sample per batch it equals ``SeqLRU`` on every request that does not
                  take the same-request corner (a key resident at request start that an earlier insert of
                  the same request evicted: the reference then misses and re-inserts it).
Keys are ``(table0 << 40) | row`` as in oracle/evlfu.py.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from .evlfu import make_key


class SeqLRU:
    def __init__(self, capacity: int, n_tables: int = 26):
        self.cap, self.T = int(capacity), int(n_tables)
        self.od: OrderedDict[int, None] = OrderedDict()        # LRU.py:8
        self.evicted: list[int] = []
        self.corner = False

    def request(self, row_ids):
        """Returns hit[T].  self.evicted: keys evicted by this request, in order."""
        self.evicted = []
        self.corner = False
        start = set(self.od) if self.cap <= 4096 else None
        hit = []
        for i, r in enumerate(row_ids):
            k = make_key(i, r)
            if k in self.od:                                   # :27-31
                self.od.move_to_end(k, last=True)
                hit.append(True)
            else:                                              # :32-36 -> set() :15-22
                if start is not None and k in start:
                    self.corner = True
                if len(self.od) >= self.cap:
                    ek, _ = self.od.popitem(last=False)
                    self.evicted.append(ek)
                self.od[k] = None
                hit.append(False)
        return hit

    def state(self):
        return list(self.od.keys())


class BatchLRU:
    """Same interface as oracle.evlfu.BatchEvLFU (the GPU parity helpers drive either)."""

    def __init__(self, capacity: int, n_tables: int = 26):
        self.cap, self.T = int(capacity), int(n_tables)
        self.entries: OrderedDict[int, None] = OrderedDict()
        self.n_perfect = 0
        self.evicted: list[int] = []
        self.flushed: list[int] = []
        self.inserted: list[int] = []

    def lookup_batch(self, idx, approx_emb_thres: int = -1, agg=None, table_base: int = 0):
        idx = np.asarray(idx)
        Tl, B = idx.shape
        ent = self.entries
        hit = np.zeros((B, Tl), dtype=bool)
        last: dict[int, int] = {}
        for s in range(B):
            for t in range(Tl):
                k = make_key(table_base + t, idx[t, s])
                hit[s, t] = k in ent
                last[k] = s * Tl + t
        self.inserted = []
        prot = None
        for k, _p in sorted(last.items(), key=lambda kv: kv[1]):
            if k in ent:
                ent.move_to_end(k, last=True)
            else:
                ent[k] = None
                self.inserted.append(k)
                prot = k
        self.evicted = []
        need = len(ent) - self.cap
        if need > 0:
            for k in ent:
                if k == prot:
                    continue
                self.evicted.append(k)
                if len(self.evicted) == need:
                    break
            for k in self.evicted:
                del ent[k]
        src_t = np.tile(np.arange(Tl, dtype=np.int32) + table_base, (B, 1))
        src_r = np.ascontiguousarray(idx.T).astype(np.int64)
        return hit, src_t, src_r, hit.sum(axis=1).astype(np.int64)

    def state(self):
        """Per-bucket FIFO lists in the layout of the CUDA path's dump: everything lives in bucket 0."""
        return [list(self.entries.keys())] + [[] for _ in range(self.T)]


class SeqLFU:
    """``/root/reference/cache_algo/LFU.py`` restated (the third policy the reference's driver initialises,
    dlrm_s_pytorch_C1_C2_C3.py:1294): per-frequency FIFO lists ``node_for_freq[f]`` (:9,33-34), ``node_for_key``
    key -> frequency (:11,31), a lazily maintained ``least_freq`` pointer (:25-28 advances it by ONE when its list
    empties, :50 resets it to 1 on every insert), eviction = the oldest key of ``node_for_freq[least_freq]``
    (:40-44).  Pinned by tests/golden/lfu_*.npz.  The CUDA path (``policy="lfu"``) implements ``BatchLFU`` below, whose
    frequencies saturate at n_tables + 1 (``freq_cap``)."""

    def __init__(self, capacity: int, n_tables: int = 26, freq_cap: int = 0):
        """freq_cap = 0: the reference (unbounded frequencies).  freq_cap = F: the frequency saturates at F -- a hit on a key
        of frequency F re-appends it to list F -- which is the variant the CUDA path implements (F = n_tables + 1, one bucket
        ring per frequency); identical to the reference as long as no key is hit more than F - 1 times."""
        self.cap, self.T = int(capacity), int(n_tables)
        self.fcap = int(freq_cap)
        self.least = 1                                         # :7
        self.freq: list = [0, OrderedDict()]                   # :16-17 (index 0 unused)
        self.key: dict[int, int] = {}
        self.evicted: list[int] = []
        self.corner = False

    def request(self, row_ids):
        self.evicted = []
        self.corner = False
        start = set(self.key) if self.cap <= 4096 else None
        hit = []
        for i, r in enumerate(row_ids):
            k = make_key(i, r)
            f = self.key.get(k)
            if f is not None:                                  # :56-61 -> _update :19-34
                del self.freq[f][k]
                if len(self.freq[self.least]) == 0:
                    self.least += 1
                nf = f + 1 if (self.fcap == 0 or f < self.fcap) else f
                if nf == f and self.least > f:                 # saturated key re-appended to the list least_freq just left
                    self.least = f
                self.key[k] = nf
                if nf == len(self.freq):
                    self.freq.append(OrderedDict())
                self.freq[nf][k] = None
                hit.append(True)
            else:                                              # :62-66 -> set :36-50
                if start is not None and k in start:
                    self.corner = True
                if len(self.key) >= self.cap:
                    ek, _ = self.freq[self.least].popitem(last=False)
                    del self.key[ek]
                    self.evicted.append(ek)
                self.key[k] = 1
                self.freq[1][k] = None
                self.least = 1
                hit.append(False)
        return hit

    def state(self):
        """(least_freq, per-frequency key lists from frequency 1 up)."""
        return self.least, [list(d.keys()) for d in self.freq[1:]]


class BatchLFU:
    """The batch-granular LFU the CUDA path implements (``policy="lfu"``; same interface as BatchEvLFU / BatchLRU).

    One FIFO list per frequency 1 .. F (F = n_tables + 1: the tier's bucket rings), bucket b = frequency - 1.  All keys of a
    batch are probed against the state at batch start.  A key that hit moves to the end of the next list (a key already in
    list F is re-appended to list F), a key that missed enters list 1; a key requested several times in a batch moves ONCE
    -- its frequency counts the batches that asked for it -- and takes the place of its last occurrence (highest position,
    sample-major / table-minor); moves are applied in position order.  Then the cache is evicted back down to capacity from
    the lowest frequency up, oldest first, never the last key inserted (the reference evicts before it inserts,
    LFU.py:40-50).  With one sample per batch this is ``SeqLFU(freq_cap=F)`` on every request that does not take the
    same-request corner (a key resident at request start that an earlier insert of the same request evicted)."""

    def __init__(self, capacity: int, n_tables: int = 26):
        self.cap, self.T = int(capacity), int(n_tables)
        self.entries: dict[int, int] = {}                      # key -> bucket
        self.lists = [OrderedDict() for _ in range(self.T + 1)]
        self.n_perfect = 0
        self.evicted: list[int] = []
        self.flushed: list[int] = []
        self.inserted: list[int] = []

    def lookup_batch(self, idx, approx_emb_thres: int = -1, agg=None, table_base: int = 0):
        idx = np.asarray(idx)
        Tl, B = idx.shape
        ent = self.entries
        top = self.T
        hit = np.zeros((B, Tl), dtype=bool)
        last: dict[int, int] = {}
        for s in range(B):
            for t in range(Tl):
                k = make_key(table_base + t, idx[t, s])
                hit[s, t] = k in ent
                last[k] = s * Tl + t
        self.inserted = []
        prot = None
        for k, _p in sorted(last.items(), key=lambda kv: kv[1]):
            old = ent.get(k)
            if old is not None:
                del self.lists[old][k]
                nb = min(old + 1, top)
            else:
                nb = 0
                self.inserted.append(k)
                prot = k
            self.lists[nb][k] = None
            ent[k] = nb
        self.evicted = []
        need = len(ent) - self.cap
        b = 0
        while need > 0 and b <= top:
            victims = []
            for k in self.lists[b]:
                if k == prot:
                    continue
                victims.append(k)
                if len(victims) == need:
                    break
            for k in victims:
                del self.lists[b][k]
                del ent[k]
            self.evicted.extend(victims)
            need -= len(victims)
            b += 1
        src_t = np.tile(np.arange(Tl, dtype=np.int32) + table_base, (B, 1))
        src_r = np.ascontiguousarray(idx.T).astype(np.int64)
        return hit, src_t, src_r, hit.sum(axis=1).astype(np.int64)

    def state(self):
        return [list(l.keys()) for l in self.lists]
