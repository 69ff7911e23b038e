#!/usr/bin/env python
"""sequential_forward (dlrm_s_pytorch_C1_C2_C3.py:742-768) eager vs captured in one CUDA graph (DLRMInference.capture):
Kaggle architecture (13 dense, 26 tables, dim 16, bot 13-512-256-64-16, top 367-512-256-1), batch 2048, random weights,
Kaggle-shape tables scaled down (the cache kernels' cost does not depend on the table size; the point here is the launch
structure).  Prints us per forward (CUDA events over 200 forwards) for both."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    pkg = importlib.import_module("ev-store-dlrm_b200")
    B, dim, n = 2048, 16, 200
    rows = pkg.workload.scaled_rows(pkg.workload.KAGGLE_ROWS, 0.05)
    tables = pkg.workload.make_tables(rows, dim)
    rng = np.random.default_rng(0)
    sizes_b, sizes_t = [13, 512, 256, 64, 16], [16 + 27 * 26 // 2, 512, 256, 1]
    mk = lambda sz: [((rng.standard_normal((o, i)) / np.sqrt(i)).astype(np.float32), np.zeros(o, np.float32)) for i, o in zip(sz[:-1], sz[1:])]
    idx = pkg.workload.ZipfTrace(rows, seed=3).batches(n + 60, B)
    dev = torch.device("cuda", 0)
    idx_dev = torch.from_numpy(idx).to(dev)
    dense = torch.rand((B, 13), device=dev)
    res = {}
    for mode in ("eager", "graph"):
        store = pkg.EvStore(tables, pkg.CacheConfig(total_size=int(sum(rows) * 0.13), max_batch=B),
                            stores={32: [pkg.to_host_rows(t) for t in tables]})
        net = pkg.dlrm_ops.DLRMInference(mk(sizes_b), mk(sizes_t), store)
        if mode == "graph":
            net.capture(B)
        step = (lambda k: net.replay(dense, idx_dev[k])) if mode == "graph" else (lambda k: net.sequential_forward(dense, None, idx_dev[k]))
        for k in range(50):
            step(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(n):
            step(50 + k)
        e1.record()
        torch.cuda.synchronize()
        res[mode] = 1e3 * e0.elapsed_time(e1) / n
        store.close()
    print(json.dumps({"batch": B, "us_per_forward_eager": res["eager"], "us_per_forward_one_graph": res["graph"],
                      "samples_per_s_one_graph": B / (res["graph"] * 1e-6)}))


if __name__ == "__main__":
    main()
