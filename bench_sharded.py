"""bench.py --gpus N (N > 1): the lookup path table-wise sharded over the GPUs of one box.

One process per GPU (torchrun), NCCL.  Rank r owns the tables get_my_slice(26, r, N)
(extend_distributed.py:47-62), looks the whole global batch up in them (probe -> all-reduce of the
per-sample hit counts -> batch-granular EvLFU with the exact agg_hit), and an all-to-all turns the
pooled rows [B, T_local*d] into [B/N, 26*d] (dlrm_s_pytorch.py:564-570).

Weak scaling: the global batch is 2048*N samples, so every rank handles (26/N tables) x (2048*N
samples) = 53 248 lookups per step whatever N is, and ends with the 2048 samples of its batch slice.
Default shape: the Kaggle tables of configs[1] (so the N = 1 line of bench.py is the same workload,
unsharded, and the reference arm is comparable); --shape terabyte selects configs[4] (dim 64,
MLPerf cardinalities capped at 40 M).
"""
from __future__ import annotations

import importlib
import json
import os
import time

import numpy as np


def main_sharded(args):
    import torch
    import torch.distributed as dist

    from bench import ClockSampler, bytes_per_lookup, log, measured_peak_hbm

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("ev-store-dlrm_b200")

    shape = getattr(args, "shape", "kaggle")
    all_rows = pkg.workload.TERABYTE_ROWS if shape == "terabyte" else pkg.workload.KAGGLE_ROWS
    if args.scale != 1.0:
        all_rows = pkg.workload.scaled_rows(all_rows, args.scale)
    dim = args.dim or (64 if shape == "terabyte" else 16)
    prec = args.precision
    T = len(all_rows)
    Bl = args.batch or 2048
    B = Bl * world
    K, W = args.steps, max(args.warmup, 3)
    sl = pkg.sharded.get_my_slice(T, rank, world)
    rows = all_rows[sl]
    T_local = len(rows)
    # the reference's 13 % operating point (cache_manager.cpp:16), each rank caching 13 % of its own rows
    cache_rows = max(1024, int(sum(rows) * 0.13))

    t0 = time.time()
    tables = [pkg.workload.make_table(sl.start + t, r, dim) for t, r in enumerate(rows)]
    raw = [pkg.codecs.encode_table(t, prec) for t in tables]
    pinned = [torch.from_numpy(r).pin_memory() for r in raw]
    stores = {prec: [q.numpy() for q in pinned]}
    del raw
    trace = pkg.workload.ZipfTrace(rows, alpha=1.05, seed=42 + sl.start, perm_seed=7 + sl.start)
    log(f"[rank {rank}] tables {sl.start}..{sl.stop - 1}: {sum(rows) / 1e6:.2f} M rows, {sum(t.nbytes for t in tables) / 1e9:.2f} GB, "
        f"cache {cache_rows} rows, built in {time.time() - t0:.1f}s")
    cfg = pkg.CacheConfig(n_layers=1, main_precision=prec, total_size=cache_rows * prec // 32, max_batch=B, device=local_rank,
                          n_tables_total=T, table_base=sl.start)
    store = pkg.EvStore(tables, cfg, stores=stores)
    del tables
    transport = getattr(args, "transport", "p2p")
    sh = pkg.sharded.ShardedLookup(store, T, dim, rank, world, transport=transport, batch_max=B)

    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    # ---- fill the cache ------------------------------------------------------------------------
    t0 = time.time()
    warm_max = args.cache_warm if args.cache_warm >= 0 else 6000
    done = 0
    full = torch.zeros(1, dtype=torch.int32, device=dev)
    while done < warm_max:
        n = min(200, warm_max - done)
        idx = torch.from_numpy(trace.batches(n, B)).to(dev)
        for k in range(n):
            sh.lookup(idx[k])
        done += n
        st = store.stats()
        full[0] = 1 if st["size"][0] >= st["capacity"][0] else 0
        dist.all_reduce(full, op=dist.ReduceOp.MIN)
        if int(full.item()) == 1:
            break
    st = store.stats(reset=True)
    log(f"[rank {rank}] cache warm: {done} batches in {time.time() - t0:.1f}s, resident {st['size'][0]}/{st['capacity'][0]}, "
        f"hit rate so far {st['hits'][0] / max(1, st['lookups']):.3f}")

    n_batches = 3 * (W + K) + 8
    idx_host = torch.from_numpy(trace.batches(n_batches, B)).pin_memory()        # [n, T_local, B]
    idx_dev = idx_host.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0 and not args.no_clocks:
        sampler.start()
        time.sleep(0.3)

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: index batches resident in HBM ------------------------------------------------------
    base = 0
    for k in range(W):
        sh.lookup(idx_dev[base + k])
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    l0 = store.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        sh.lookup(idx_dev[base + W + k])
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    launches = store.launch_count() - l0
    st = store.stats(reset=True)
    lookups = K * B * T                                   # whole job: every rank's tables
    value = lookups / (ms_dev * 1e-3)
    hits = torch.tensor([st["hits"][0], st["lookups"], st["misses"]], dtype=torch.float64, device=dev)
    dist.all_reduce(hits)
    hit_rate = float(hits[0] / max(1.0, float(hits[1])))

    # ---- e2e: pinned host indices in, this rank's pooled rows out to pinned host memory -----------------
    base += W + K
    out_host = [torch.empty((Bl, T, dim), dtype=torch.float32).pin_memory() for _ in range(2)]
    idx_stage = [torch.empty((T_local, B), dtype=torch.int64, device=dev) for _ in range(2)]
    for k in range(W):
        idx_stage[k % 2].copy_(idx_host[base + k], non_blocking=True)
        ly, _ = sh.lookup(idx_stage[k % 2])
        out_host[k % 2].copy_(ly, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for k in range(K):
        idx_stage[k % 2].copy_(idx_host[base + W + k], non_blocking=True)
        ly, _ = sh.lookup(idx_stage[k % 2])
        out_host[k % 2].copy_(ly, non_blocking=True)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    dist.barrier()
    e2e_value = lookups / e2e_s

    clocks = sampler.stop() if (rank == 0 and not args.no_clocks) else {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}

    # ---- the all-to-all alone (NVLink share of the step) ---------------------------------------------
    send = torch.empty((B, T_local, dim), dtype=torch.float32, device=dev)
    recv = torch.empty((Bl * T * dim,), dtype=torch.float32, device=dev)
    in_splits = [Bl * T_local * dim] * world
    out_splits = [Bl * t * dim for t in sh.splits]
    for _ in range(5):
        dist.all_to_all_single(recv, send.view(-1), out_splits, in_splits)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(50):
        dist.all_to_all_single(recv, send.view(-1), out_splits, in_splits)
    e1.record()
    torch.cuda.synchronize()
    a2a_ms = max_over_ranks(e0.elapsed_time(e1)) / 50
    a2a_bytes = sh.alltoall_bytes(B)

    # ---- roofline: per-kernel CUDA-event times on rank 0's shard ----------------------------------------
    base += W + K
    store.kernel_times(reset=True)
    store.set_profiling(True)
    for k in range(K):
        sh.lookup(idx_dev[base + (k % (W + K))])
    torch.cuda.synchronize()
    kt = store.kernel_times(reset=True)
    store.set_profiling(False)
    dist.barrier()
    per_kernel = {n: {"avg_us": 1e3 * ms / max(1, timed), "launches": timed} for n, (ms, timed, _l) in kt.items() if timed}
    tot_us = sum(v["avg_us"] * v["launches"] for v in per_kernel.values())
    for v in per_kernel.values():
        v["share"] = v["avg_us"] * v["launches"] / max(tot_us, 1e-9)
    dom = max(per_kernel, key=lambda n: per_kernel[n]["share"])
    peak, peak_src = measured_peak_hbm()
    bpl = bytes_per_lookup(dim, prec)
    alg_bytes = B * T_local * bpl                                   # what one launch of a rank's kernel covers
    # as in bench.py: the HBM-bound kernel of the step is k_serve (rank 0's launch; under p2p it also waits for the
    # peers' hit counts and stores its rows over NVLink); the kernel with the largest share is named beside it
    dom_us = per_kernel["k_serve"]["avg_us"] if "k_serve" in per_kernel else per_kernel[dom]["avg_us"]
    roofline = {"bound": "hbm", "kernel": "k_serve" if "k_serve" in per_kernel else dom,
                "achieved": alg_bytes / (dom_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg_bytes / (dom_us * 1e-6) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_lookup": bpl, "kernel_avg_us": dom_us,
                "dominant_by_time": {"kernel": dom, "share": per_kernel[dom]["share"], "avg_us": per_kernel[dom]["avg_us"]},
                "per_kernel_rank0": per_kernel,
                "alltoall_nccl_alone": {"bytes_sent_per_rank": a2a_bytes, "ms": a2a_ms, "achieved_GBps": a2a_bytes / (a2a_ms * 1e-3) / 1e9,
                             "peak_GBps": 770.0, "peak_source": "B200_PROFILING.md measured peer copy per direction",
                             "frac": a2a_bytes / (a2a_ms * 1e-3) / 1e9 / 770.0}}

    if rank == 0:
        line = {
            "metric": "ev_lookups_per_s", "value": value, "unit": "lookups/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if prec == 32 else f"u{prec}->f32", "data": "synthetic", "samples_per_s": value / T,
            "hit_rate": hit_rate,
            "config": {"workload": "%s-shape 26 tables (%.2fM rows), dim %d, C1 EvLFU fp%d tier, Zipf(1.05), table-wise sharded over %d "
                                   "GPUs (13 %% of each rank's rows cached), global batch %d = %d per GPU, exact agg_hit; exchange: %s"
                                   % (shape, sum(all_rows) / 1e6, dim, prec, world, B, Bl,
                                      "fused into the kernels (peer-memory stores over NVLink + epoch flags, evs_shard_*)"
                                      if transport == "p2p" else "NCCL all-reduce of the hit counts + all_to_all_single of the pooled rows"),
                       "batch": B, "dim": dim, "precision": prec, "parallelism": "table-wise x%d + all-to-all" % world,
                       "transport": transport,
                       "cache_warm_batches": done,
                       "l2": "no flush: index+slab working set exceeds the 126 MB L2 and every step reads a distinct index batch"},
            "e2e": {"value": e2e_value, "unit": "lookups/s", "h2d_bytes_per_step": T_local * B * 8 * world,
                    "d2h_bytes_per_step": Bl * T * dim * 4 * world, "ms_per_step": 1e3 * e2e_s / K,
                    "api": "ShardedLookup.lookup (%s) on pinned host indices, pooled rows copied back to pinned host memory" % transport},
            "gpu_launches": int(launches) * world, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": {"value": None, "unit": "lookups/s", "cores": 0, "kind": "reference", "sample": "reported at N = 1 only"},
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    store.close()
    dist.destroy_process_group()
    return 0
