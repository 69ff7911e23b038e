"""EvLFU oracles (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Two policies, both single-tier (C1):

* ``SeqEvLFU``   -- a restatement of the reference's canonical sequential EvLFU,
                    ``/root/reference/cache_algo/EvLFU_C1.py`` (one request of
                    ``n_tables`` keys at a time).  Pinned against the reference
                    itself by ``tests/golden/`` (see ``make_golden.py``).
* ``BatchEvLFU`` -- the batch-granular generalisation the CUDA path implements
                    (DESIGN.md "Batch-granular EvLFU").  With a batch of one
                    sample it produces the same hit vector, eviction set and
                    final state as ``SeqEvLFU`` on every request that does not
                    take the reference's same-request re-fetch corner
                    (EvLFU_C1.py:84-95, a hit key evicted/flushed by an earlier
                    insert of the same request).

Keys are integers ``(table0 << 40) | row`` where ``table0`` is the 0-based table
index (the reference's string key is ``f"{table0+1}-{row}"``, EvLFU_C1.py:113).

The oracles track *which* backing-store row answers each position, not the
float payload: a cached value is by construction a copy of the backing row, so
``src`` (table0, row) determines the fp32 output bit-exactly.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

KEY_SHIFT = 40


def make_key(table0: int, row: int) -> int:
    return (int(table0) << KEY_SHIFT) | int(row)


def split_key(key: int):
    return key >> KEY_SHIFT, key & ((1 << KEY_SHIFT) - 1)


class SeqEvLFU:
    """Sequential EvLFU, statement-for-statement semantics of EvLFU_C1.py.

    State (EvLFU_C1.py:7-19): ``vals`` key -> agg_hit, ``lists[0..T]`` FIFO per
    agg_hit bucket, ``min`` bucket pointer, ``n_perfect``, ``max_perfect =
    int(cap*0.95)``, ``flush_rate = 0.3``.
    """

    def __init__(self, capacity: int, n_tables: int = 26,
                 flush_rate: float = 0.3, perfect_item_cap: float = 0.95):
        self.cap = int(capacity)
        self.T = int(n_tables)
        self.min = 0                                      # EvLFU_C1.py:9
        self.vals: dict[int, int] = {}                    # EvLFU_C1.py:10 (agg only)
        self.lists = [OrderedDict() for _ in range(self.T + 1)]   # :26-28
        self.n_perfect = 0                                # :13
        self.max_perfect = int(self.cap * perfect_item_cap)       # :29
        self.flush_n = int(flush_rate * self.cap) + 1     # :39
        # per-request logs (cleared by request())
        self.evicted: list[int] = []
        self.flushed: list[int] = []
        self.refetched: list[int] = []

    # EvLFU_C1.py:32-63
    def _set(self, key: int, agg: int) -> None:
        if self.n_perfect >= self.max_perfect:
            for _ in range(self.flush_n):
                k, _v = self.lists[self.T].popitem(last=False)   # KeyError == ref IndexError
                del self.vals[k]
                self.flushed.append(k)
            self.n_perfect = len(self.lists[self.T])
        elif len(self.vals) >= self.cap:
            while not self.lists[self.min]:
                self.min += 1
                if self.min > self.T:
                    self.min = 1
            k, _v = self.lists[self.min].popitem(last=False)
            del self.vals[k]
            self.evicted.append(k)
        self.vals[key] = agg
        self.lists[agg][key] = None
        if agg < self.min:
            self.min = agg

    # EvLFU_C1.py:65-78
    def _update_agg_hit(self, key: int, agg: int) -> bool:
        old = self.vals.get(key)
        if old is None:
            return False
        if old < agg:
            del self.lists[old][key]
            self.lists[agg][key] = None
            self.vals[key] = agg
        return True

    # EvLFU_C1.py:97-166
    def request(self, row_ids, approx_emb_thres: int = -1):
        """One request.  Returns (hit[T] bool, src[T] list of (table0,row)|None, agg)."""
        self.evicted, self.flushed, self.refetched = [], [], []
        keys = [make_key(i, r) for i, r in enumerate(row_ids)]
        hit = [k in self.vals for k in keys]
        agg = sum(hit)
        pick_random = approx_emb_thres > 0 and agg >= approx_emb_thres   # :122
        last_src = None              # the reference starts from random floats (:104-107)
        src = [None] * len(keys)
        for i, k in enumerate(keys):
            if hit[i]:
                if not self._update_agg_hit(k, agg):       # evicted meanwhile (:84-95)
                    self.refetched.append(k)
                    self._set(k, agg)
                src[i] = (i, int(row_ids[i]))
                last_src = src[i]                          # :139
            elif pick_random:
                src[i] = last_src                          # :150
                hit[i] = True                              # :151
            else:
                # update() with the pre-fetched value -> update_agg_hit misses -> set()
                if not self._update_agg_hit(k, agg):
                    self._set(k, agg)
                src[i] = (i, int(row_ids[i]))
        if agg == self.T:                                  # :163-165
            self.n_perfect = len(self.lists[self.T])
        return hit, src, agg

    def state(self):
        """Per-bucket FIFO key lists (eviction order)."""
        return [list(l.keys()) for l in self.lists]


class BatchEvLFU:
    """Batch-granular EvLFU (the policy of the CUDA path).

    For a batch of B samples x Tl keys, positions are numbered sample-major,
    table-minor (``p = s*Tl + t``):

    1. probe every key against the state at batch start; ``agg[s]`` = hits of
       sample s (or the caller's ``agg`` for the table-sharded exact mode);
    2. an occurrence is *flagged* when it is a miss, or a hit whose stored
       bucket is below ``agg[s]``; per key the winning occurrence is the
       maximum of ``(agg, p)``;
    3. winners are applied in position order: the key moves to (or is inserted
       at) the tail of bucket ``agg`` -- new keys may push the size above cap;
    4. if there is at least one new key and ``n_perfect >= max_perfect`` the
       oldest ``int(0.3*cap)+1`` keys of bucket T are flushed and ``n_perfect``
       is refreshed (EvLFU_C1.py:36-44);
    5. the cache is evicted back down to ``cap`` in (bucket, FIFO) order, never
       evicting the highest-ranked new key (the reference evicts *before* each
       insert, so its last insert always survives, EvLFU_C1.py:46-60);
    6. ``n_perfect = len(bucket T)`` if any sample had ``agg == T`` (:163-165).
    """

    def __init__(self, capacity: int, n_tables: int = 26,
                 flush_rate: float = 0.3, perfect_item_cap: float = 0.95):
        self.cap = int(capacity)
        self.T = int(n_tables)
        self.entries: dict[int, int] = {}
        self.lists = [OrderedDict() for _ in range(self.T + 1)]
        self.n_perfect = 0
        self.max_perfect = int(self.cap * perfect_item_cap)
        self.flush_n = int(flush_rate * self.cap) + 1
        self.evicted: list[int] = []
        self.flushed: list[int] = []
        self.inserted: list[int] = []

    def lookup_batch(self, idx, approx_emb_thres: int = -1, agg=None, table_base: int = 0, table_ids=None):
        """idx: int array [Tl, B] (the reference's lS_i layout).  Local table t has the global id
        ``table_base + t`` (a rank's contiguous slice) or ``table_ids[t]`` (any placement).

        Returns (hit [B,Tl] bool, src_t [B,Tl] int32, src_r [B,Tl] int64, agg [B]).
        ``src_t < 0`` marks a substituted position with no earlier hit (the
        reference answers those with random floats).
        """
        idx = np.asarray(idx)
        Tl, B = idx.shape
        T = self.T
        ent = self.entries
        gid = [table_base + t for t in range(Tl)] if table_ids is None else [int(g) for g in table_ids]
        # a negative index = no key at this position (a slice of ragged bags, ``expand_bags``): it is not probed, counts
        # for nothing and is reported as a miss with src_t = -1
        keys = [[make_key(gid[t], idx[t, s]) if idx[t, s] >= 0 else None for t in range(Tl)] for s in range(B)]
        hit0 = np.array([[k is not None and k in ent for k in ks] for ks in keys], dtype=bool).reshape(B, Tl)
        if agg is None:
            agg = hit0.sum(axis=1).astype(np.int64)
        agg = np.asarray(agg, dtype=np.int64)
        hit = hit0.copy()
        src_t = np.tile(np.asarray(gid, dtype=np.int32), (B, 1))
        src_r = np.ascontiguousarray(idx.T).astype(np.int64)
        src_t[src_r < 0] = -1

        winners: dict[int, tuple[int, int]] = {}
        for s in range(B):
            a = int(agg[s])
            approx = approx_emb_thres > 0 and a >= approx_emb_thres
            ks = keys[s]
            last = -1
            first = -1
            if approx:
                hs = np.flatnonzero(hit0[s])
                first = int(hs[0]) if len(hs) else -1
            for t in range(Tl):
                k = ks[t]
                p = s * Tl + t
                if k is None:
                    continue
                if hit0[s, t]:
                    last = t
                    if ent[k] >= a:
                        continue
                elif approx:
                    j = last if last >= 0 else first
                    hit[s, t] = True
                    if j >= 0:
                        src_t[s, t] = gid[j]
                        src_r[s, t] = idx[j, s]
                    else:
                        src_t[s, t] = -1
                    continue
                w = winners.get(k)
                if w is None or (a, p) > w:
                    winners[k] = (a, p)

        self.inserted = []
        prot, prot_rank = None, None
        for k, (a, p) in sorted(winners.items(), key=lambda kv: kv[1][1]):
            old = ent.get(k)
            if old is not None:
                del self.lists[old][k]
            else:
                self.inserted.append(k)
                if prot_rank is None or (a, p) > prot_rank:
                    prot, prot_rank = k, (a, p)
            self.lists[a][k] = None
            ent[k] = a

        self.flushed = []
        if self.inserted and self.n_perfect >= self.max_perfect:
            lt = self.lists[T]
            for _ in range(min(self.flush_n, len(lt))):
                k, _v = lt.popitem(last=False)
                del ent[k]
                self.flushed.append(k)
            self.n_perfect = len(lt)

        self.evicted = []
        need = len(ent) - self.cap
        b = 0
        while need > 0 and b <= T:
            lst = self.lists[b]
            victims = []
            for k in lst:
                if k == prot:
                    continue
                victims.append(k)
                if len(victims) == need:
                    break
            for k in victims:
                del lst[k]
                del ent[k]
            self.evicted.extend(victims)
            need -= len(victims)
            b += 1

        if (agg == T).any():
            self.n_perfect = len(self.lists[T])
        return hit, src_t, src_r, agg

    def state(self):
        return [list(l.keys()) for l in self.lists]


def gather_rows(tables, src_t, src_r, fill=0.0):
    """fp32 rows for (src_t, src_r) [B,Tl] from ``tables`` (list of [rows,d] float32)."""
    B, Tl = src_t.shape
    d = tables[0].shape[1]
    out = np.full((B, Tl, d), fill, dtype=np.float32)
    for t in np.unique(src_t):
        if t < 0:
            continue
        m = src_t == t
        out[m] = tables[int(t)][src_r[m]]
    return out


def expand_bags(idx_lists, off_lists, B: int, P: int):
    """Ragged bags -> slices.  idx_lists[t]: the indices of table t (all bags concatenated), off_lists[t]: int [B] start of
    each bag (nn.EmbeddingBag's offsets; bag s ends where bag s+1 starts).  Returns int64 [T, B*P]: virtual sample
    s*P + j holds the j-th index of every table's bag of sample s, -1 where the bag is shorter.  A batch of bags is
    looked up as these B*P groups (each one a request group in the reference's sense: its agg_hit counts the hits among
    its keys) and pooled per (sample, table) in ascending j -- the sum order of nn.EmbeddingBag(mode="sum")."""
    T = len(idx_lists)
    out = np.full((T, B * P), -1, dtype=np.int64)
    for t in range(T):
        off = list(off_lists[t]) + [len(idx_lists[t])]
        for s in range(B):
            n = off[s + 1] - off[s]
            assert 0 <= n <= P, "bag larger than P"
            out[t, s * P:s * P + n] = idx_lists[t][off[s]:off[s + 1]]
    return out


def pool_bags(rows_v, idx_v, B: int, P: int):
    """rows_v fp32 [B*P, T, d] (the rows of the slices), idx_v [T, B*P] -> pooled [B, T, d]: sum over j ascending, fp32."""
    BP, T, d = rows_v.shape
    out = np.zeros((B, T, d), dtype=np.float32)
    for j in range(P):
        r = rows_v.reshape(B, P, T, d)[:, j]
        m = (idx_v.reshape(T, B, P)[:, :, j] >= 0).T            # [B, T]
        out = np.where(m[:, :, None], out + r, out).astype(np.float32)
    return out
