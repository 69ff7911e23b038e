"""bench.py --gpus N (N > 1): the lookup path table-wise sharded over the GPUs of one box.

One process per GPU (torchrun).  Rank r owns a set of tables, looks the whole global batch up in them (batch-granular
EvLFU with the EXACT agg_hit: the sum of the ranks' per-sample hit counts) and the pooled rows [B, T_local*d] end up
batch-sharded, [B/N, 26*d] (DLRM_Net.distributed_forward, dlrm_s_pytorch.py:529-586; ext_dist.alltoall,
extend_distributed.py:541-576).  The exchange is part of the kernels (peer-memory stores over NVLink + epoch words,
evs_shard_*); --transport nccl selects all-reduce + all_to_all_single instead.

Legs of one run (every rank 53 248 lookups per step whatever N is -- weak scaling):
  main        Kaggle-shape tables of configs[1] (so the N = 1 line of bench.py is the same workload unsharded), global
              batch 2048*N, tables placed by ROW COUNT (sharded.balanced_placement; departs from get_my_slice)  -> `value`
  contiguous  the same with the reference's contiguous get_my_slice placement                               -> "placement_contiguous"
  configs4    BASELINE configs[4]: Criteo-Terabyte shape (dim 64, cardinalities capped at 40 M, 48 GB of backing
              rows in pinned host memory), same placement rule                                              -> "configs4"
Every leg is VERIFIED on untimed steps: each rank all-gathers the index batch and checks every row it received
against the closed-form backing table of its key (workload.synth_rows); one more leg replays a scaled-down shard
against the batch-granular oracle (hit maps, eviction counts).  A mismatch makes the run exit non-zero.
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import time

import numpy as np


def _max_over_ranks(x: float, dev, world) -> float:
    import torch
    import torch.distributed as dist
    if world == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(vals, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    return [float(v) for v in t.tolist()]


def _gather_over_ranks(vals, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    if world <= 1:
        return [t.tolist()]
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [x.tolist() for x in out]


def _barrier(world):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


GROUP = max(1, int(os.environ.get("EVS_BENCH_GROUP", "4")))      # batches per library call in the timed regions


class Leg:
    """One sharded cache over one table shape: build, fill, verify, time."""

    def __init__(self, pkg, log, rank, world, local_rank, all_rows, dim, prec, Bl, placement_kind, transport, label):
        import torch
        self.pkg, self.log, self.rank, self.world, self.label = pkg, log, rank, world, label
        self.dev = torch.device("cuda", local_rank)
        self.all_rows, self.dim, self.prec = list(all_rows), dim, prec
        self.T = len(all_rows)
        self.Bl, self.B = Bl, Bl * world
        self.placement = (pkg.sharded.balanced_placement(all_rows, world) if placement_kind == "balanced"
                          else pkg.sharded.contiguous_placement(self.T, world))
        self.ids = self.placement[rank]
        self.rows = [all_rows[t] for t in self.ids]
        self.T_local = len(self.ids)
        # the reference's 13 % operating point (cache_manager.cpp:16), each rank caching 13 % of its own rows
        # the reference's 13 % operating point (cache_manager.cpp:16) is a budget for the whole cache: 13 % of ALL rows, split
        # evenly over the GPUs, a rank with less than its share passing the rest on (sharded.split_cache_budget)
        per_rank_rows = [sum(all_rows[t] for t in x) for x in self.placement]
        self.cache_split = pkg.sharded.split_cache_budget(per_rank_rows, int(sum(all_rows) * 0.13))
        self.cache_rows = self.cache_split[rank]
        t0 = time.time()
        assert prec == 32, "the sharded bench serves the fp32 tier"
        # backing rows in evs_host_alloc memory (host memory mapped into this rank's device with large pages)
        if os.environ.get("EVS_BENCH_STORE", "mapped") == "mapped":
            self.pinned = [pkg.workload.synth_table_mapped(t, all_rows[t], dim, self.dev) for t in self.ids]
            stores = {32: self.pinned}
        else:
            self.pinned = [pkg.workload.synth_table_pinned(t, all_rows[t], dim, self.dev) for t in self.ids]
            stores = {32: [q.numpy() for q in self.pinned]}
        self.trace = pkg.workload.ZipfTrace(self.rows, alpha=1.05, seed=42 + self.ids[0], perm_seed=7 + self.ids[0])
        cfg = pkg.CacheConfig(n_layers=1, main_precision=prec, total_size=self.cache_rows, max_batch=self.B, device=local_rank,
                              n_tables_total=self.T, table_ids=tuple(self.ids))
        self.store = pkg.EvStore.from_raw_stores(self.rows, dim, cfg, stores)
        self.sh = pkg.sharded.ShardedLookup(self.store, self.T, dim, rank, world, transport=transport, batch_max=self.B,
                                            placement=self.placement)
        log(f"[{label} rank {rank}] tables {self.ids}: {sum(self.rows) / 1e6:.2f} M rows, {sum(self.rows) * dim * 4 / 1e9:.2f} GB of host rows, "
            f"cache {self.cache_rows} rows, HBM {self.store.memory_footprint() / 1e9:.2f} GB, built in {time.time() - t0:.1f}s")

    def batches(self, n):
        import torch
        return torch.from_numpy(self.trace.batches(n, self.B))

    def fill(self, warm_max):
        import torch
        import torch.distributed as dist
        t0 = time.time()
        done = 0
        full = torch.zeros(1, dtype=torch.int32, device=self.dev)
        while done < warm_max:
            n = min(200, warm_max - done)
            idx = self.batches(n + 1).to(self.dev)
            for k in range(n):
                self.sh.lookup(idx[k], next_idx=idx[k + 1])
            done += n
            st = self.store.stats()
            # (a rank whose budget covers all its rows never evicts: it counts as full from the start)
            full[0] = 1 if (st["size"][0] >= st["capacity"][0] or self.cache_rows >= sum(self.rows)) else 0
            if self.world > 1:
                dist.all_reduce(full, op=dist.ReduceOp.MIN)
            if int(full.item()) == 1:
                break
        # evictions running on every rank before anything is timed
        idx = self.batches(101).to(self.dev)
        for k in range(100):
            self.sh.lookup(idx[k], next_idx=idx[k + 1])
        done += 100
        st = self.store.stats(reset=True)
        self.store.check()
        self.log(f"[{self.label} rank {self.rank}] cache warm: {done} batches in {time.time() - t0:.1f}s, resident {st['size'][0]}/{st['capacity'][0]}, "
                 f"hit rate so far {st['hits'][0] / max(1, st['lookups']):.3f}")
        return done

    def verify(self, idx_dev, n_steps=2):
        """Untimed: every row this rank received == the closed-form backing row of its key (bit-exact), and the hit map
        agrees with it (a hit and a miss both deliver the backing row).  Returns (ok, rows checked)."""
        import torch
        import torch.distributed as dist
        ok, checked = True, 0
        tmax = max(len(x) for x in self.placement)
        for k in range(n_steps):
            ly, hit = self.sh.lookup(idx_dev[k], next_idx=idx_dev[k + 1])
            mine = torch.zeros((tmax, self.B), dtype=torch.int64, device=self.dev)
            mine[:self.T_local] = idx_dev[k]
            if self.world > 1:
                allidx = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(allidx, mine)
            else:
                allidx = [mine]
            s0 = self.rank * self.Bl
            for r, ids in enumerate(self.placement):
                for j, t in enumerate(ids):
                    want = self.pkg.workload.synth_rows(t, allidx[r][j, s0:s0 + self.Bl], self.dim, self.all_rows[t])
                    good = bool(torch.equal(ly[:, t, :], want))
                    ok = ok and good
                    checked += self.Bl
                    if not good:
                        self.log(f"[{self.label} rank {self.rank}] VERIFY FAILED: table {t} from rank {r}, step {k}: "
                                 f"{int((ly[:, t, :] != want).any(dim=1).sum())} of {self.Bl} rows differ")
            torch.cuda.synchronize()
        self.store.check()
        return ok, checked

    def timed(self, idx_dev, base, W, K, reps=3):
        """`reps` regions of exactly K steps (barrier + synchronize on both sides, device time, max over ranks)."""
        import torch
        for k in range(W):
            self.sh.lookup(idx_dev[base + k], next_idx=idx_dev[base + k + 1])
        base += W
        _barrier(self.world)
        self.store.phase_times()                  # the phase accumulators cover the timed regions only
        regions = []
        launches = 0
        for _ in range(reps):
            _barrier(self.world)
            l0 = self.store.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # the serving loop hands the library GROUP queued batches per call (groups of 4 = one captured graph on every
            # rank; each batch is still a full, strictly ordered pass) and announces the first batch of the next call
            hit = torch.empty((self.B, self.T_local), dtype=torch.uint8, device=self.dev)
            calls = [[idx_dev[base + k + j] for j in range(min(GROUP, K - k))] + [idx_dev[base + min(k + GROUP, K)]] for k in range(0, K, GROUP)]
            e0.record()
            for c in calls:
                self.sh.lookup_many(c[:-1], next_idx=c[-1], hits=hit)
            e1.record()
            _barrier(self.world)
            regions.append(_max_over_ranks(e0.elapsed_time(e1), self.dev, self.world))
            launches = self.store.launch_count() - l0
            base += K
        return regions, base, launches

    def close(self):
        self.store.close()
        self.pinned = None


def verify_policy_small(pkg, log, rank, world, local_rank, transport):
    """A scaled-down shard against the batch-granular oracle: per-rank hit maps, eviction counts and resident sizes
    (the full-size caches are checked for delivered VALUES; this checks the policy under the sharded exchange)."""
    import torch
    import torch.distributed as dist
    from oracle.evlfu import BatchEvLFU, make_key
    dev = torch.device("cuda", local_rank)
    rows = [3, 5, 4000, 2500, 7, 4, 60, 9, 3, 300, 50, 3500, 40, 4, 80, 3000, 5, 45, 30, 4, 3800, 6, 5, 600, 11, 400]
    dim, Bl, cap, n = 16, 64, 500, 14
    B = Bl * world
    placement = pkg.sharded.balanced_placement(rows, world)
    ids = placement[rank]
    lrows = [rows[t] for t in ids]
    tables = [pkg.workload.synth_rows(t, np.arange(rows[t]), dim, rows[t]) for t in ids]
    cfg = pkg.CacheConfig(total_size=cap, max_batch=B, n_tables_total=26, table_ids=tuple(ids), device=local_rank, record_events=True)
    store = pkg.EvStore(tables, cfg)
    sh = pkg.sharded.ShardedLookup(store, 26, dim, rank, world, transport=transport, batch_max=B, placement=placement)
    oracle = BatchEvLFU(cap, n_tables=26)
    trace = pkg.workload.ZipfTrace(lrows, seed=91 + ids[0], perm_seed=3 + ids[0])
    batches = trace.batches(n + 1, B)
    dev_idx = [torch.from_numpy(np.ascontiguousarray(b)).to(dev) for b in batches]
    ok = True
    n_ev = 0
    for k in range(n):
        ly, hit = sh.lookup(dev_idx[k], next_idx=dev_idx[k + 1])
        idx = batches[k]
        local = np.array([sum(make_key(ids[t], idx[t, s]) in oracle.entries for t in range(len(ids))) for s in range(B)], dtype=np.int32)
        agg = torch.from_numpy(local).to(dev)
        if world > 1:
            dist.all_reduce(agg)
        o_hit, _st, _sr, _ = oracle.lookup_batch(idx, agg=agg.cpu().numpy(), table_ids=ids)
        torch.cuda.synchronize()
        ev, _fl = store.last_events()
        good = bool(np.array_equal(hit.cpu().numpy().astype(bool), o_hit)) and ev.tolist() == oracle.evicted
        if not good:
            log(f"[policy check rank {rank}] batch {k}: hit map / eviction stream differs from the oracle")
        ok = ok and good
        n_ev += len(oracle.evicted)
    st = store.stats()
    ok = ok and st["size"][0] == len(oracle.entries)
    store.close()
    return ok, n_ev


def main_sharded(args):
    import torch
    import torch.distributed as dist

    from bench import log

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("ev-store-dlrm_b200")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    K = args.steps
    transport = getattr(args, "transport", "p2p")

    def shape_rows(shape):
        r = pkg.workload.TERABYTE_ROWS if shape == "terabyte" else pkg.workload.KAGGLE_ROWS
        return pkg.workload.scaled_rows(r, args.scale) if args.scale != 1.0 else r

    verified = {}
    # ---- policy under the sharded exchange, against the oracle (scaled-down shard) -----------------------------------
    okp, n_ev = verify_policy_small(pkg, log, rank, world, local_rank, transport)
    bad, ev_all = _sum_over_ranks([0.0 if okp else 1.0, float(n_ev)], dev, world)        # a rank that owns only tiny tables never evicts
    verified["policy_vs_oracle_scaled_shard"] = bool(bad == 0.0 and ev_all > 0)

    def run_leg(shape, placement_kind, label, K_leg, full_report):
        return run_one_leg(pkg, log, args, rank, world, local_rank, shape, placement_kind, label, K_leg, full_report, verified)

    shape = getattr(args, "shape", "kaggle")
    main = run_leg(shape, "balanced", "main", K, True)
    legs = {}
    if not getattr(args, "only_main", False):
        legs["placement_contiguous"] = run_leg(shape, "contiguous", "contiguous", min(K, 50), False)
        if shape == "kaggle" and args.scale == 1.0:
            legs["configs4"] = run_leg("terabyte", "balanced", "configs4", min(K, 50), False)
    return finish_sharded(pkg, log, args, rank, world, dev, main, legs, verified, shape_rows(shape))


def configs4_text(world, B):
    return ("configs[4]: Criteo-Terabyte shape, 26 tables (187.77M rows, capped at 40M), dim 64, 48 GB of fp32 backing rows in pinned host "
            "memory, C1 EvLFU fp32 tier caching 13 %% of each rank's rows, Zipf(1.05), table-wise sharded over %d GPU%s (tables placed by "
            "row count), global batch %d, exact agg_hit" % (world, "" if world == 1 else "s", B))


def run_one_leg(pkg, log, args, rank, world, local_rank, shape, placement_kind, label, K_leg, full_report, verified):
    """Build, fill, verify and time one sharded cache; returns the leg's report (identical on every rank where it matters)."""
    import torch
    from bench import ClockSampler, bytes_per_lookup, measured_peak_hbm
    dev = torch.device("cuda", local_rank)
    prec = args.precision
    Bl = args.batch or 2048
    B = Bl * world
    W = max(args.warmup, 3)
    transport = getattr(args, "transport", "p2p")
    peak, peak_src = measured_peak_hbm()

    def shape_rows(shape):
        r = pkg.workload.TERABYTE_ROWS if shape == "terabyte" else pkg.workload.KAGGLE_ROWS
        return pkg.workload.scaled_rows(r, args.scale) if args.scale != 1.0 else r

    dim = args.dim or (64 if shape == "terabyte" else 16)
    leg = Leg(pkg, log, rank, world, local_rank, shape_rows(shape), dim, prec, Bl, placement_kind, transport, label)
    warm_done = leg.fill(args.cache_warm if args.cache_warm >= 0 else 24000)
    n_batches = (4 if full_report else 3) * K_leg + 2 * W + 8
    idx_host = leg.batches(n_batches).pin_memory()                   # [n, T_local, B]
    idx_dev = idx_host.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    ok, checked = leg.verify(idx_dev, 2)
    ok_all = bool(_sum_over_ranks([0.0 if ok else 1.0], dev, world)[0] == 0.0)
    verified[label] = ok_all
    leg.store.phase_times()
    sampler = None
    if full_report and rank == 0 and not args.no_clocks:
        sampler = ClockSampler(local_rank)
        sampler.start()
        time.sleep(0.3)
    regions, base, launches = leg.timed(idx_dev, 2, W, K_leg)
    ms_dev = sorted(regions)[1]
    ph = leg.store.phase_times()
    st = leg.store.stats(reset=True)
    tot = _sum_over_ranks([st["hits"][0], st["lookups"], st["misses"]], dev, world)
    T = leg.T
    lookups = K_leg * B * T
    res = {"value": lookups / (ms_dev * 1e-3), "samples_per_s": K_leg * B / (ms_dev * 1e-3), "ms_per_step": ms_dev / K_leg,
           "value_regions_ms": regions, "hit_rate": tot[0] / max(1.0, tot[1]), "misses_per_step_all_ranks": tot[2] / (3 * K_leg + W),
           "cache_warm_batches": warm_done, "rows_verified_per_rank": checked, "verified": ok_all, "dim": dim,
           "placement": placement_kind, "tables_per_rank": [len(x) for x in leg.placement],
           "rows_per_rank_M": [round(sum(leg.all_rows[t] for t in x) / 1e6, 2) for x in leg.placement],
           "cache_rows_per_rank": leg.cache_split,
           # where rank 0 waits for its peers (in-kernel %globaltimer, averaged over the timed steps): inside k_serve for their
           # hit counts (part of avg_serve) and at the end of k_evict for their rows
           "rank0_phases_us": {"serve_incl_wait_for_counts": ph["avg_serve"], "update": ph["avg_update"],
                               "evict_incl_fetch_and_wait_for_rows": ph["avg_evict"], "wait_for_peer_rows": ph["avg_peer_wait"],
                                   "wait_for_peer_counts_cta0": ph["avg_count_wait_cta0"],
                               "fetch_role_since_evict_start": ph["avg_fetch_since_evict_start"]},
           "nvlink_bytes_per_rank_per_step": leg.sh.alltoall_bytes(B)}
    # every rank's in-kernel phase spans and misses per step: which rank the others wait for, and in which phase
    mine = [ph["avg_serve"], ph["avg_count_wait_cta0"], ph["avg_update"], ph["avg_evict"], ph["avg_peer_wait"],
            ph["avg_fetch_since_evict_start"], st["misses"] / (3 * K_leg + W)]
    allp = _gather_over_ranks(mine, dev, world)
    res["phases_us_by_rank"] = {"columns": ["serve", "of_it_wait_counts_cta0", "update", "evict", "of_it_wait_rows", "fetch_since_evict_start",
                                            "misses_per_step"], "rows": [[round(v, 2) for v in r] for r in allp]}
    res["nvlink_GBps_per_rank"] = res["nvlink_bytes_per_rank_per_step"] / (res["ms_per_step"] * 1e-3) / 1e9
    res["nvlink_frac_of_peer_copy_peak"] = res["nvlink_GBps_per_rank"] / 770.0
    clocks = None
    if full_report:
        # ---- e2e: pinned host indices in, this rank's pooled rows out to pinned host memory -----------------
        out_host = [torch.empty((Bl, T, dim), dtype=torch.float32).pin_memory() for _ in range(2)]
        idx_stage = [torch.empty((leg.T_local, B), dtype=torch.int64, device=dev) for _ in range(2)]

        def e2e_step(k):
            # the look-ahead for step k + 1 is announced as soon as its indices are on their way (ready event)
            idx_stage[k % 2].copy_(idx_host[base + k], non_blocking=True)
            ly, _ = leg.sh.lookup(idx_stage[k % 2])
            out_host[k % 2].copy_(ly, non_blocking=True)

        for k in range(W):
            e2e_step(k)
        _barrier(world)
        t0 = time.perf_counter()
        for k in range(W, W + K_leg):
            e2e_step(k)
        torch.cuda.synchronize()
        e2e_s = _max_over_ranks(time.perf_counter() - t0, dev, world)
        _barrier(world)
        res["e2e"] = {"value": lookups / e2e_s, "unit": "lookups/s", "h2d_bytes_per_step": leg.T_local * B * 8 * world,
                      "d2h_bytes_per_step": Bl * T * dim * 4 * world, "ms_per_step": 1e3 * e2e_s / K_leg,
                      "api": "ShardedLookup.lookup (%s) on pinned host indices, pooled rows copied back to pinned host memory" % transport}
        clocks = sampler.stop() if sampler is not None else {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        # ---- per-kernel CUDA-event times on every rank's shard (plain launches), rank 0 reported ----------------
        leg.store.kernel_times(reset=True)
        leg.store.set_profiling(True)
        for k in range(K_leg):
            leg.sh.lookup(idx_dev[2 + W + (k % (3 * K_leg))])
        torch.cuda.synchronize()
        kt = leg.store.kernel_times(reset=True)
        leg.store.set_profiling(False)
        _barrier(world)
        per_kernel = {n: {"avg_us": 1e3 * ms / max(1, timed), "launches": timed} for n, (ms, timed, _l) in kt.items() if timed}
        tot_us = sum(v["avg_us"] * v["launches"] for v in per_kernel.values())
        for v in per_kernel.values():
            v["share"] = v["avg_us"] * v["launches"] / max(tot_us, 1e-9)
        dom = max(per_kernel, key=lambda n: per_kernel[n]["share"])
        bpl = bytes_per_lookup(dim, prec)
        alg_bytes = B * leg.T_local * bpl
        sv = per_kernel["k_serve"]["avg_us"]
        res["roofline"] = {"bound": "hbm", "kernel": "k_serve", "achieved": alg_bytes / (sv * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": alg_bytes / (sv * 1e-6) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                           "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_lookup": bpl, "kernel_avg_us": sv,
                           "note": "rank 0's launch; under the fused exchange k_serve also waits for the peers' hit counts and stores its rows over NVLink",
                           "dominant_by_time": {"kernel": dom, "share": per_kernel[dom]["share"], "avg_us": per_kernel[dom]["avg_us"]},
                           "per_kernel_rank0": per_kernel}
        res["gpu_launches"] = int(launches) * world
        res["clocks"] = clocks
    if "configs4" == label:
        res["workload"] = configs4_text(world, B)
    leg.close()
    del idx_dev, idx_host
    torch.cuda.empty_cache()
    return res


def finish_sharded(pkg, log, args, rank, world, dev, main, legs, verified, rows_main):
    import torch
    import torch.distributed as dist
    from bench import sharded_config
    Bl = args.batch or 2048
    B = Bl * world
    K, W = args.steps, max(args.warmup, 3)
    transport = getattr(args, "transport", "p2p")
    if "configs4" in legs:
        legs["configs4"]["workload"] = configs4_text(world, B)
    # ---- the NCCL all-to-all alone, for scale (what the fused exchange replaces) ----------------------------------------
    a2a = None
    try:
        dim = main["dim"]
        T = 26
        pl = pkg.sharded.balanced_placement(rows_main, world)
        T_local = len(pl[rank])
        send = torch.empty((B, T_local, dim), dtype=torch.float32, device=dev)
        recv = torch.empty((Bl * T * dim,), dtype=torch.float32, device=dev)
        in_splits = [Bl * T_local * dim] * world
        out_splits = [Bl * len(x) * dim for x in pl]
        for _ in range(5):
            dist.all_to_all_single(recv, send.view(-1), out_splits, in_splits)
        _barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            dist.all_to_all_single(recv, send.view(-1), out_splits, in_splits)
        e1.record()
        torch.cuda.synchronize()
        a2a_ms = _max_over_ranks(e0.elapsed_time(e1), dev, world) / 50
        a2a_bytes = B * T_local * dim * 4 * (world - 1) // world
        a2a = {"bytes_sent_per_rank": a2a_bytes, "ms": a2a_ms, "achieved_GBps": a2a_bytes / (a2a_ms * 1e-3) / 1e9, "peak_GBps": 770.0,
               "peak_source": "B200_PROFILING.md measured peer copy per direction", "frac": a2a_bytes / (a2a_ms * 1e-3) / 1e9 / 770.0}
    except Exception as e:
        log("all-to-all leg failed:", repr(e))

    all_ok = all(verified.values())
    if rank == 0:
        roof = main.pop("roofline")
        roof["alltoall_nccl_alone"] = a2a
        cfg = sharded_config(world, B, main["dim"], transport)
        line = {
            "metric": "ev_lookups_per_s", "value": main["value"], "unit": "lookups/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "samples_per_s": main["samples_per_s"], "hit_rate": main["hit_rate"],
            "config": cfg, "e2e": main.pop("e2e"), "gpu_launches": main.pop("gpu_launches"), "clocks": main.pop("clocks"),
            "roofline": roof, "verified": all_ok, "verified_legs": verified, "main": main,
            "cpu_baseline": {"value": None, "unit": "lookups/s", "cores": 0, "kind": "reference", "sample": "reported at N = 1 only"},
        }
        line.update(legs)
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if all_ok else 3
