#!/usr/bin/env python
"""A handle like a rank of an 8-GPU run that owns only small tables: 3 tables, the 8-GPU global batch (16384), a cache
of ~150k rows.  Its bucket rings are small, so a pessimistic host-side occupancy bound forces a stream synchronise every
few batches -- and in a sharded run every other rank waits for this one.  Prints us per step (CUDA events, 200 steps)."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    pkg = importlib.import_module("ev-store-dlrm_b200")
    rows = [93145, 5683, 8351593 // 64]           # ~0.23 M rows
    dim, B, steps = 16, 16384, 200
    tables = pkg.workload.make_tables(rows, dim)
    stores = {32: [pkg.to_host_rows(t) for t in tables]}
    idx = pkg.workload.ZipfTrace(rows).batches(steps * 2 + 60, B)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    idx_dev = torch.from_numpy(idx).to(dev)
    for cap in (150000, 30000):
        cfg = pkg.CacheConfig(n_layers=1, main_precision=32, total_size=cap, max_batch=B, device=0, n_tables_total=26, table_ids=(3, 9, 20))
        store = pkg.EvStore(tables, cfg, stores=stores)
        out = torch.empty((B, 3, dim), dtype=torch.float32, device=dev)
        hit = torch.empty((B, 3), dtype=torch.uint8, device=dev)
        for k in range(50):
            store.lookup(idx_dev[k], out=out, hit=hit)
            store.prefetch(idx_dev[k + 1])
        res = {}
        for mode in ("single", "many"):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            base = 50 + (0 if mode == "single" else steps)
            t0 = time.perf_counter()
            e0.record()
            if mode == "single":
                for k in range(steps):
                    store.lookup(idx_dev[base + k], out=out, hit=hit)
                    store.prefetch(idx_dev[base + k + 1])
            else:
                for k in range(0, steps, 4):
                    store.lookup_many([idx_dev[base + k + j] for j in range(4)], outs=out, hits=hit)
                    store.prefetch(idx_dev[base + k + 4])
            e1.record()
            host = time.perf_counter() - t0
            torch.cuda.synchronize()
            res[mode] = {"us_per_step": 1e3 * e0.elapsed_time(e1) / steps, "host_us_per_step": 1e6 * host / steps}
        st = store.stats()
        print(json.dumps({"cache_rows": cap, "hit_rate": st["hits"][0] / max(1, st["lookups"]), **res}), flush=True)
        store.close()


if __name__ == "__main__":
    main()
