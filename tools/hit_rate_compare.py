#!/usr/bin/env python
"""Hit rate of the batch-granular EvLFU policy (what the CUDA path implements, oracle.evlfu.BatchEvLFU)
next to the reference's sequential EvLFU simulator (cache_algo/EvLFU_C1.py, restated and pinned as
oracle.evlfu.SeqEvLFU) on the same Zipf(1.05) trace.  CPU only; writes profiles/<round>_hit_rate.md.

    python tools/hit_rate_compare.py r1 [--scale 0.01] [--samples 65536]
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.evlfu import BatchEvLFU, SeqEvLFU  # noqa: E402
from oracle.lru import BatchLFU, BatchLRU, SeqLFU, SeqLRU  # noqa: E402


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
    scale, n = 0.01, 65536
    for i, a in enumerate(sys.argv):
        if a == "--scale":
            scale = float(sys.argv[i + 1])
        if a == "--samples":
            n = int(sys.argv[i + 1])
    pkg = importlib.import_module("ev-store-dlrm_b200")
    rows = pkg.workload.scaled_rows(pkg.workload.KAGGLE_ROWS, scale)
    cap = int(sum(rows) * 0.13)
    trace = pkg.workload.ZipfTrace(rows, seed=42).batches(1, n)[0]            # [26, n]
    out = [f"# {rnd}: hit rate, batch-granular EvLFU (and LRU) vs the reference's sequential EvLFU_C1.py, LRU.py and LFU.py\n",
           f"Kaggle-shape tables scaled by {scale} ({sum(rows)} rows), cache {cap} rows (13 %), Zipf(1.05), {n} samples; "
           f"the second half of the trace is measured (the first half warms the cache).\n",
           "| policy | per-lookup hit rate | perfect-hit samples | evictions |", "|---|---|---|---|"]
    half = n // 2
    t0 = time.time()
    seq = SeqEvLFU(cap)
    hits = perfect = ev = 0
    for s in range(n):
        h, _src, agg = seq.request(trace[:, s])
        if s >= half:
            hits += sum(h)
            perfect += int(agg == 26)
            ev += len(seq.evicted)
    out.append(f"| sequential (EvLFU_C1.py, 1 sample per request) | {hits / (26 * (n - half)):.4f} | {perfect} | {ev} |")
    print(out[-1], f"({time.time() - t0:.0f}s)")
    for B in (1, 128, 2048, 16384):
        if B > half:
            continue
        t0 = time.time()
        o = BatchEvLFU(cap)
        hits = perfect = ev = 0
        for k in range(0, n, B):
            h, _st, _sr, agg = o.lookup_batch(trace[:, k:k + B])
            if k >= half:
                hits += int(h.sum())
                perfect += int((agg == 26).sum())
                ev += len(o.evicted)
        out.append(f"| batch-granular, B = {B} | {hits / (26 * (n - half)):.4f} | {perfect} | {ev} |")
        print(out[-1], f"({time.time() - t0:.0f}s)")
    # the reference's comparison policy (cache_algo/LRU.py) on the same trace
    t0 = time.time()
    lru = SeqLRU(cap)
    hits = perfect = ev = 0
    for s in range(n):
        h = lru.request(trace[:, s])
        if s >= half:
            hits += sum(h)
            perfect += int(all(h))
            ev += len(lru.evicted)
    out.append(f"| sequential LRU (LRU.py, 1 sample per request) | {hits / (26 * (n - half)):.4f} | {perfect} | {ev} |")
    print(out[-1], f"({time.time() - t0:.0f}s)")
    for B in (1, 2048):
        t0 = time.time()
        o = BatchLRU(cap)
        hits = perfect = ev = 0
        for k in range(0, n, B):
            h, _st, _sr, agg = o.lookup_batch(trace[:, k:k + B])
            if k >= half:
                hits += int(h.sum())
                perfect += int((agg == 26).sum())
                ev += len(o.evicted)
        out.append(f"| batch-granular LRU, B = {B} | {hits / (26 * (n - half)):.4f} | {perfect} | {ev} |")
        print(out[-1], f"({time.time() - t0:.0f}s)")
    # and its plain LFU (cache_algo/LFU.py), then the batch-granular LFU the CUDA path runs (frequencies saturate at 27)
    t0 = time.time()
    lfu = SeqLFU(cap)
    hits = perfect = ev = 0
    for s in range(n):
        h = lfu.request(trace[:, s])
        if s >= half:
            hits += sum(h)
            perfect += int(all(h))
            ev += len(lfu.evicted)
    out.append(f"| sequential LFU (LFU.py, 1 sample per request) | {hits / (26 * (n - half)):.4f} | {perfect} | {ev} |")
    print(out[-1], f"({time.time() - t0:.0f}s)")
    for B in (1, 2048):
        t0 = time.time()
        o = BatchLFU(cap)
        hits = perfect = ev = 0
        for k in range(0, n, B):
            h, _st, _sr, agg = o.lookup_batch(trace[:, k:k + B])
            if k >= half:
                hits += int(h.sum())
                perfect += int((agg == 26).sum())
                ev += len(o.evicted)
        out.append(f"| batch-granular LFU (policy=\"lfu\"), B = {B} | {hits / (26 * (n - half)):.4f} | {perfect} | {ev} |")
        print(out[-1], f"({time.time() - t0:.0f}s)")
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"{rnd}_hit_rate.md"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
