// Kernels of one batch of the embedding-lookup hot path (single tier shown first; the
// two/three-tier lookup lives in evs_tiers.cuh and reuses everything below).
//
//   k_lookup      probe + per-sample agg_hit + gather/dequantise hits + flag promotions/misses
//   k_miss        claim index slots for missing keys (dedup), fetch rows from the backing
//                 store (zero-copy), fill slab + output
//   k_hist_scan   per-bucket exclusive scan of the per-CTA append counts -> ring positions
//   k_append      append promoted / inserted rows to their bucket's FIFO ring in position order
//   k_evict       flush rule, then evict in (bucket, FIFO) order down to capacity
//   k_compact     squeeze dead records out of one bucket ring (rare, host-triggered)
#pragma once
#include "evs_codec.cuh"
#include "evs_types.cuh"

namespace evs {

constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint4 ldg16(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

// ---- index probe ---------------------------------------------------------------------
// Linear probing; a probe ends at the key or at the first slot no resident key's path
// crosses (pass == 0), so deletions need no tombstones.
__device__ __forceinline__ bool probe(const Slot *__restrict__ slots, unsigned mask, unsigned long long key,
                                      unsigned &slot_out, unsigned &rowword_out) {
    unsigned i = hash_key(key, mask);
    while (true) {
        uint4 v = ldg16(slots + i);
        unsigned long long k = (static_cast<unsigned long long>(v.y) << 32) | v.x;
        if (k == key) {
            slot_out = i;
            rowword_out = v.z;
            return true;
        }
        if (v.w == 0u) return false;
        i = (i + 1) & mask;
    }
}

// ---- k_lookup --------------------------------------------------------------------------
// One warp per sample, lane t < T holds the key of table t.  A CTA covers kSamplesPerCta
// consecutive samples so the int64 index tile is read as T runs of 64 contiguous bytes.
template <int PREC>
__global__ void __launch_bounds__(kLookupThreads) k_lookup(TierDev tier, LookupArgs a) {
    __shared__ long long s_idx[kSamplesPerCta][kMaxTables];
    __shared__ unsigned s_hist[kMaxBuckets];
    __shared__ unsigned s_stat[4];           // hits, subst, perfect, misses
    __shared__ CodecLut s_lut;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = a.T, B = a.B, D = a.D;
    const int s0 = blockIdx.x * kSamplesPerCta;

    for (int i = threadIdx.x; i < kSamplesPerCta * T; i += kLookupThreads) {
        int t = i / kSamplesPerCta, j = i - t * kSamplesPerCta;
        int s = s0 + j;
        s_idx[j][t] = (s < B) ? __ldg(a.idx + static_cast<size_t>(t) * B + s) : 0;
    }
    if (threadIdx.x < kMaxBuckets) s_hist[threadIdx.x] = 0;
    if (threadIdx.x < 4) s_stat[threadIdx.x] = 0;
    codec_lut_init<PREC>(&s_lut);
    __syncthreads();

    const int s = s0 + warp;
    const bool act = (s < B) && (lane < T);
    bool ishit = false;
    unsigned slot = 0, rowword = 0;
    if (act) {
        long long r = s_idx[warp][lane];
        if (r < 0 || r >= __ldg(a.rows + lane)) {
            a.g->error = 1u;               // EVS_ERR_INDEX; answer from row 0
            r = 0;
            s_idx[warp][lane] = 0;
        }
        ishit = probe(tier.slots, tier.hash_mask, make_key(a.table_base + lane, r), slot, rowword);
    }
    const unsigned hitmask = __ballot_sync(kFull, ishit);
    int agg = __popc(hitmask);
    if (a.agg_in != nullptr && s < B) agg = a.agg_in[s];
    const bool approx = (a.approx_thres > 0) && (agg >= a.approx_thres) && (hitmask != 0u);

    if (s < B) {
        // EvLFU bookkeeping requests for the later kernels
        uint8_t f = 0;
        const int p = s * T + lane;
        if (act) {
            const int bucket = static_cast<int>(rowword >> kRowBits) - 1;
            if (ishit) {
                if (bucket < agg) {
                    f = static_cast<uint8_t>(agg + 1);
                    tier.pos_slot[p] = slot;
                }
            } else if (!approx) {
                f = static_cast<uint8_t>(kFlagMiss | (agg + 1));
            }
            tier.flags[p] = f;
            if (a.hit != nullptr) a.hit[p] = (ishit || approx) ? 1 : 0;
            if (f) atomicAdd(&s_hist[(f & 0x3F) - 1], 1u);
        }
        const unsigned missmask = __ballot_sync(kFull, (f & kFlagMiss) != 0);
        if (missmask) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&tier.ctl->miss_count, static_cast<unsigned>(__popc(missmask)));
            base = __shfl_sync(kFull, base, 0);
            if (f & kFlagMiss) tier.miss_list[base + __popc(missmask & ((1u << lane) - 1u))] = p;
        }
        if (lane == 0) {
            a.agg_out[s] = static_cast<uint8_t>(agg);
            const int nh = __popc(hitmask);
            atomicAdd(&s_stat[0], static_cast<unsigned>(nh));
            if (approx) atomicAdd(&s_stat[1], static_cast<unsigned>(T - nh));
            else atomicAdd(&s_stat[3], static_cast<unsigned>(T - nh));
            if (agg == a.n_perfect_agg) {
                atomicAdd(&s_stat[2], 1u);
                tier.ctl->any_perfect = 1u;
            }
        }

        // gather + dequantise: lanes walk the sample's T*CPR 16-byte chunks, so consecutive
        // lanes read consecutive 16 B of a row and write consecutive floats of the output
        const int cpr = static_cast<int>(tier.row_stride >> 4);
        const int total = T * cpr;
        const unsigned row = rowword & kRowMask;
        float *orow = a.out + static_cast<size_t>(s) * a.out_stride;
        const bool vec = ((reinterpret_cast<uintptr_t>(a.out) & 15u) == 0) && ((a.out_stride & 3) == 0) && ((D & 3) == 0);
        for (int c0 = 0; c0 < total; c0 += 32) {
            const int c = c0 + lane;
            const bool inb = c < total;
            const int tt = inb ? c / cpr : 0;
            const int part = c - tt * cpr;
            bool srv = (hitmask >> tt) & 1u;
            int src = tt;
            if (!srv && approx) {           // EvLFU_C1.py:140-152: value of the latest earlier hit
                const unsigned lower = hitmask & ((1u << tt) - 1u);
                src = lower ? (31 - __clz(lower)) : (__ffs(hitmask) - 1);
                srv = true;
            }
            const unsigned r = __shfl_sync(kFull, row, src);
            if (inb && srv) {
                uint4 v = ldg16(tier.slab + static_cast<size_t>(r) * tier.row_stride + (part << 4));
                decode_store<PREC>(v, orow + tt * D, part, D, vec, &s_lut);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < kMaxBuckets) tier.hist[blockIdx.x * kMaxBuckets + threadIdx.x] = s_hist[threadIdx.x];
    if (threadIdx.x == 0) {
        GlobalCtl *g = a.g;
        const int ns = min(kSamplesPerCta, B - s0);
        atomicAdd(&g->lookups, static_cast<unsigned long long>(ns) * T);
        atomicAdd(&g->samples, static_cast<unsigned long long>(ns));
        if (s_stat[0]) atomicAdd(&g->hits[0], static_cast<unsigned long long>(s_stat[0]));
        if (s_stat[1]) atomicAdd(&g->approx_subst, static_cast<unsigned long long>(s_stat[1]));
        if (s_stat[2]) atomicAdd(&g->perfect_hits, static_cast<unsigned long long>(s_stat[2]));
        if (s_stat[3]) atomicAdd(&g->misses, static_cast<unsigned long long>(s_stat[3]));
    }
}

// Probe only: local hit count per sample (table-sharded exact groupability).
__global__ void __launch_bounds__(kLookupThreads) k_probe(TierDev tier, LookupArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * kSamplesPerCta + warp;
    if (s >= a.B) return;
    bool ishit = false;
    if (lane < a.T) {
        long long r = __ldg(a.idx + static_cast<size_t>(lane) * a.B + s);
        if (r < 0 || r >= __ldg(a.rows + lane)) r = 0;
        unsigned slot, rowword;
        ishit = probe(tier.slots, tier.hash_mask, make_key(a.table_base + lane, r), slot, rowword);
    }
    const unsigned hitmask = __ballot_sync(kFull, ishit);
    if (lane == 0) a.agg_out[s] = static_cast<uint8_t>(__popc(hitmask));
}

// ---- k_miss ------------------------------------------------------------------------------
// Claim a slot for `key` (or find the slot a same-batch duplicate already claimed).
// Only inserts run concurrently here, so slots go EMPTY -> occupied monotonically and every
// thread inserting the same key converges on one slot.
__device__ __forceinline__ unsigned claim_slot(const TierDev &tier, unsigned long long key, bool &claimed) {
    const unsigned mask = tier.hash_mask;
    const unsigned home = hash_key(key, mask);
    unsigned i = home;
    claimed = false;
    while (true) {
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&tier.slots[i].key);
        if (cur == key) break;
        if (cur == kEmptyKey) {
            unsigned long long old = atomicCAS(&tier.slots[i].key, kEmptyKey, key);
            if (old == kEmptyKey) {
                claimed = true;
                break;
            }
            if (old == key) break;
        }
        i = (i + 1) & mask;
    }
    if (claimed) {
        TierCtl *c = tier.ctl;
        const unsigned ft = atomicSub(&c->free_top, 1u);
        if (ft == 0u || ft > tier.rows_total) {
            c->error = 4u;                       // EVS_ERR_CAPACITY
            return i;
        }
        const unsigned row = tier.free_rows[ft - 1];
        tier.slots[i].rowword = row;             // bucket field 0 == not in any bucket yet
        tier.row_key[row] = key;
        tier.row_slot[row] = i;
        tier.row_meta[row] = 0ull;
        for (unsigned j = home; j != i; j = (j + 1) & mask) atomicAdd(&tier.slots[j].pass, 1u);
        atomicAdd(&c->n_new, 1u);
    }
    return i;
}

// Copy one backing-store row (row_bytes, arbitrary alignment) into a 16-byte aligned staging
// row of `stride` bytes, zero padded.  Zero-copy reads when the store lives in pinned host memory.
__device__ __forceinline__ void fetch_row(const unsigned char *__restrict__ src, unsigned row_bytes,
                                          unsigned char *stage, unsigned stride, int lane) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
    if (((a | row_bytes) & 15u) == 0) {
        for (unsigned o = lane * 16u; o < row_bytes; o += 512u)
            *reinterpret_cast<uint4 *>(stage + o) = ldg16(src + o);
    } else if (((a | row_bytes) & 3u) == 0) {
        for (unsigned o = lane * 4u; o < row_bytes; o += 128u)
            *reinterpret_cast<unsigned *>(stage + o) = __ldg(reinterpret_cast<const unsigned *>(src + o));
    } else if (((a | row_bytes) & 1u) == 0) {
        for (unsigned o = lane * 2u; o < row_bytes; o += 64u)
            *reinterpret_cast<unsigned short *>(stage + o) = __ldg(reinterpret_cast<const unsigned short *>(src + o));
    } else {
        for (unsigned o = lane; o < row_bytes; o += 32u) stage[o] = __ldg(src + o);
    }
    for (unsigned o = row_bytes + lane; o < stride; o += 32u) stage[o] = 0;
}

// One warp per missing position.  Dynamic shared memory: warps * row_stride bytes.
template <int PREC>
__global__ void __launch_bounds__(256) k_miss(TierDev tier, LookupArgs a) {
    extern __shared__ __align__(16) unsigned char s_stage[];
    __shared__ CodecLut s_lut;
    codec_lut_init<PREC>(&s_lut);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    unsigned char *stage = s_stage + static_cast<size_t>(warp) * tier.row_stride;
    const unsigned n = tier.ctl->miss_count;
    const int T = a.T, D = a.D;
    const int cpr = static_cast<int>(tier.row_stride >> 4);
    const bool vec = ((reinterpret_cast<uintptr_t>(a.out) & 15u) == 0) && ((a.out_stride & 3) == 0) && ((D & 3) == 0);

    for (unsigned m = blockIdx.x * warps_per_cta + warp; m < n; m += gridDim.x * warps_per_cta) {
        const unsigned p = tier.miss_list[m];
        const int s = p / T, t = p - s * T;
        long long r = __ldg(a.idx + static_cast<size_t>(t) * a.B + s);
        if (r < 0 || r >= __ldg(a.rows + t)) r = 0;
        const unsigned long long key = make_key(a.table_base + t, r);
        unsigned slot = 0, row = 0;
        int claimed_i = 0;
        if (lane == 0) {
            bool claimed;
            slot = claim_slot(tier, key, claimed);
            claimed_i = claimed ? 1 : 0;
            if (claimed) row = tier.slots[slot].rowword & kRowMask;
            tier.pos_slot[p] = slot;
            const unsigned f = tier.flags[p];
            atomicMax(&tier.ctl->prot, (static_cast<unsigned long long>(f & 0x3Fu) << 32) | p);
        }
        claimed_i = __shfl_sync(kFull, claimed_i, 0);
        row = __shfl_sync(kFull, row, 0);

        fetch_row(tier.store[t] + static_cast<size_t>(r) * tier.row_bytes, tier.row_bytes, stage, tier.row_stride, lane);
        __syncwarp();
        float *orow = a.out + static_cast<size_t>(s) * a.out_stride + t * D;
        for (int c = lane; c < cpr; c += 32) {
            uint4 v = *reinterpret_cast<const uint4 *>(stage + (c << 4));
            if (claimed_i) *reinterpret_cast<uint4 *>(tier.slab + static_cast<size_t>(row) * tier.row_stride + (c << 4)) = v;
            decode_store<PREC>(v, orow, c, D, vec, &s_lut);
        }
        __syncwarp();
    }
}

// ---- k_hist_scan -----------------------------------------------------------------------
// hist[chunk][b] (append requests of lookup-CTA `chunk` for bucket b) -> offset of that chunk's
// first record past the old tail of bucket b; tails advance by the bucket totals.  One warp per bucket.
__global__ void __launch_bounds__(1024) k_hist_scan(TierDev tier, int n_chunks) {
    const int lane = threadIdx.x & 31, b = threadIdx.x >> 5;
    if (b >= tier.n_buckets) return;
    TierCtl *c = tier.ctl;
    const unsigned long long t0 = c->tail[b];
    unsigned running = 0;
    for (int c0 = 0; c0 < n_chunks; c0 += 32) {
        const int ch = c0 + lane;
        unsigned v = (ch < n_chunks) ? tier.hist[ch * kMaxBuckets + b] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned n = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += n;
        }
        if (ch < n_chunks) tier.hist[ch * kMaxBuckets + b] = running + (incl - v);
        running += __shfl_sync(kFull, incl, 31);
    }
    if (lane == 0) {
        if (t0 + running - c->head[b] > tier.ring_cap) c->error = 4u;
        c->tail_prev[b] = t0;
        c->tail[b] = t0 + running;
    }
}

// ---- k_append --------------------------------------------------------------------------
// Same thread <-> position mapping as k_lookup, so ring order inside a bucket is position
// order: by warp (sample) then lane (table).
__global__ void __launch_bounds__(kLookupThreads) k_append(TierDev tier, int B, int T) {
    __shared__ unsigned s_cnt[kSamplesPerCta][kMaxBuckets];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * kSamplesPerCta + warp;
    const bool act = (s < B) && (lane < T);
    const int p = s * T + lane;
    const unsigned f = act ? tier.flags[p] : 0u;
    if (__syncthreads_or(f != 0u) == 0) return;

    for (int i = threadIdx.x; i < kSamplesPerCta * kMaxBuckets; i += kLookupThreads) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int b = static_cast<int>(f & 0x3Fu) - 1;
    const unsigned grp = __match_any_sync(kFull, f ? b : 255);
    const unsigned rank = __popc(grp & ((1u << lane) - 1u));
    if (f && rank == 0) s_cnt[warp][b] = __popc(grp);
    __syncthreads();
    if (!f) return;

    unsigned pre = 0;
    for (int w = 0; w < warp; ++w) pre += s_cnt[w][b];
    const unsigned long long q = tier.ctl->tail_prev[b] + tier.hist[blockIdx.x * kMaxBuckets + b] + pre + rank;

    const unsigned slot = tier.pos_slot[p];
    const unsigned row = tier.slots[slot].rowword & kRowMask;
    tier.ring[static_cast<size_t>(b) * tier.ring_cap + (q & (tier.ring_cap - 1))] = row;
    const unsigned long long mine = pack_meta(b, q);
    const unsigned long long old = atomicMax(&tier.row_meta[row], mine);
    if (mine > old) {
        const int ob = meta_bucket(old);
        if (ob != b) {
            atomicAdd(&tier.ctl->count[b], 1u);
            if (ob >= 0) atomicSub(&tier.ctl->count[ob], 1u);
            else atomicAdd(&tier.ctl->stat_inserts, 1ull);
        }
    }
    atomicMax(&tier.slots[slot].rowword, (static_cast<unsigned>(b + 1) << kRowBits) | row);
}

// ---- k_evict ---------------------------------------------------------------------------
__device__ __forceinline__ void evict_row(const TierDev &tier, unsigned row, unsigned long long *out_keys, unsigned out_idx) {
    const unsigned long long key = tier.row_key[row];
    const unsigned slot = tier.row_slot[row];
    const unsigned mask = tier.hash_mask;
    for (unsigned j = hash_key(key, mask); j != slot; j = (j + 1) & mask) atomicSub(&tier.slots[j].pass, 1u);
    tier.slots[slot].rowword = 0u;
    tier.slots[slot].key = kEmptyKey;
    tier.row_meta[row] = 0ull;
    const unsigned fi = atomicAdd(&tier.ctl->free_top, 1u);
    tier.free_rows[fi] = row;
    if (out_keys != nullptr) out_keys[out_idx] = key;
}

// Pop up to `want` live records from the head of bucket b (whole CTA).  The protected row is
// skipped but keeps its place.  Returns the number popped (uniform across the CTA).
__device__ unsigned pop_bucket(const TierDev &tier, int b, unsigned want, unsigned prot_row,
                               unsigned long long *out_keys, unsigned out_base) {
    __shared__ unsigned s_wsum[32];
    __shared__ unsigned s_total;
    __shared__ unsigned long long s_first_kept;
    TierCtl *c = tier.ctl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const unsigned long long h = c->head[b], tl = c->tail[b];
    __syncthreads();                       // everyone has read head/tail before thread 0 rewrites them
    if (threadIdx.x == 0) s_first_kept = ~0ull;
    __syncthreads();
    unsigned got = 0;
    unsigned long long base = h;
    for (; base < tl && got < want; base += blockDim.x) {
        const unsigned long long q = base + threadIdx.x;
        bool valid = false;
        unsigned row = kNoRow;
        if (q < tl) {
            row = tier.ring[static_cast<size_t>(b) * tier.ring_cap + (q & (tier.ring_cap - 1))];
            if (row < tier.rows_total) valid = (tier.row_meta[row] == pack_meta(b, q));
        }
        const bool cand = valid && (row != prot_row);
        const unsigned bal = __ballot_sync(kFull, cand);
        if (lane == 0) s_wsum[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            unsigned v = (lane < nwarp) ? s_wsum[lane] : 0u;
            unsigned incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned n = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += n;
            }
            s_wsum[lane] = incl - v;
            if (lane == 31) s_total = incl;
        }
        __syncthreads();
        const unsigned idx = s_wsum[warp] + __popc(bal & ((1u << lane) - 1u));
        const bool take = cand && (got + idx < want);
        if (take) evict_row(tier, row, out_keys, out_base + got + idx);
        if (valid && !take) atomicMin(&s_first_kept, q);
        got += min(s_total, want - got);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned long long fk = s_first_kept;
        c->head[b] = (fk != ~0ull) ? fk : (base < tl ? base : tl);
        c->count[b] -= got;
    }
    __syncthreads();
    return got;
}

__global__ void __launch_bounds__(1024) k_evict(TierDev tier, GlobalCtl *g) {
    __shared__ unsigned s_size;
    TierCtl *c = tier.ctl;
    const int top = tier.n_buckets - 1;
    if (threadIdx.x == 0) {
        unsigned sz = 0;
        for (int b = 0; b < tier.n_buckets; ++b) sz += c->count[b];
        s_size = sz;
    }
    __syncthreads();
    unsigned size = s_size;
    const unsigned n_new = c->n_new;
    const unsigned long long prot = c->prot;
    const unsigned n_perfect = c->n_perfect;
    const unsigned any_perfect = c->any_perfect;
    __syncthreads();

    // flush rule (EvLFU_C1.py:36-44, evlfu_32.cpp:209-218): bucket `top` is trimmed FIFO
    unsigned flushed = 0;
    if (n_new > 0 && n_perfect >= tier.max_perfect) {
        const unsigned want = min(tier.flush_n, c->count[top]);
        flushed = pop_bucket(tier, top, want, kNoRow, tier.flushed, 0);
        size -= flushed;
        if (threadIdx.x == 0) c->n_perfect = c->count[top];
        __syncthreads();
    }

    // evict down to capacity, lowest bucket first, FIFO inside a bucket; the highest ranked
    // new key is never the victim (the reference evicts before it inserts)
    unsigned prot_row = kNoRow;
    if (n_new > 0) {
        const unsigned pp = static_cast<unsigned>(prot & 0xFFFFFFFFull);
        prot_row = tier.slots[tier.pos_slot[pp]].rowword & kRowMask;
    }
    unsigned need = size > tier.cap ? size - tier.cap : 0u;
    unsigned ev = 0;
    for (int b = 0; b <= top && need > 0; ++b) {
        const unsigned got = pop_bucket(tier, b, need, prot_row, tier.evicted, ev);
        ev += got;
        need -= got;
    }
    size -= ev;

    if (threadIdx.x == 0) {
        if (any_perfect) c->n_perfect = c->count[top];      // EvLFU_C1.py:163-165
        c->stat_evictions += ev;
        c->stat_flushed += flushed;
        c->n_evicted_last = ev;
        c->n_flushed_last = flushed;
        c->full_at_start = (size >= tier.cap) ? 1u : 0u;
        c->miss_count = 0;
        c->n_new = 0;
        c->prot = 0ull;
        c->any_perfect = 0;
        if (need > 0) c->error = 4u;
        if (g != nullptr) atomicAdd(&g->batches, 1ull);
    }
}

// ---- k_compact ---------------------------------------------------------------------------
// Rewrites bucket b's ring so that [head, tail) holds only live records, order preserved.
__global__ void __launch_bounds__(1024) k_compact(TierDev tier, int b) {
    __shared__ unsigned s_wsum[32];
    __shared__ unsigned s_total;
    TierCtl *c = tier.ctl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const unsigned long long h = c->head[b], tl = c->tail[b];
    unsigned long long dst = h;
    for (unsigned long long base = h; base < tl; base += blockDim.x) {
        const unsigned long long q = base + threadIdx.x;
        bool valid = false;
        unsigned row = kNoRow;
        if (q < tl) {
            row = tier.ring[static_cast<size_t>(b) * tier.ring_cap + (q & (tier.ring_cap - 1))];
            if (row < tier.rows_total) valid = (tier.row_meta[row] == pack_meta(b, q));
        }
        const unsigned bal = __ballot_sync(kFull, valid);
        if (lane == 0) s_wsum[warp] = __popc(bal);
        __syncthreads();                    // also: every record of this window has been read
        if (warp == 0) {
            unsigned v = (lane < nwarp) ? s_wsum[lane] : 0u;
            unsigned incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned n = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += n;
            }
            s_wsum[lane] = incl - v;
            if (lane == 31) s_total = incl;
        }
        __syncthreads();
        if (valid) {
            const unsigned long long nq = dst + s_wsum[warp] + __popc(bal & ((1u << lane) - 1u));
            tier.ring[static_cast<size_t>(b) * tier.ring_cap + (nq & (tier.ring_cap - 1))] = row;
            tier.row_meta[row] = pack_meta(b, nq);
        }
        dst += s_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) c->tail[b] = dst;
}

// ---- one-time initialisation --------------------------------------------------------------
__global__ void k_init_tier(TierDev tier) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t j = i; j <= tier.hash_mask; j += stride) {
        tier.slots[j].key = kEmptyKey;
        tier.slots[j].rowword = 0;
        tier.slots[j].pass = 0;
    }
    for (size_t j = i; j < tier.rows_total; j += stride) {
        tier.free_rows[j] = tier.rows_total - 1 - static_cast<unsigned>(j);   // row 0 is handed out first
        tier.row_meta[j] = 0ull;
        tier.row_key[j] = kEmptyKey;
        tier.row_slot[j] = 0;
    }
}

}  // namespace evs
