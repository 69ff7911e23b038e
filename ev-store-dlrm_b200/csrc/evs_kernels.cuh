// Kernels of one batch of the embedding-lookup hot path, for one or two cache tiers (+ C3).
//
//   k_serve    probe both tiers -> per-sample agg_hit -> route (EvLFU promote / insert per tier) ->
//              gather + dequantise every resident row into the fp32 output.  Read-only on the cache.
//   k_scan     per (tier, bucket): exclusive scan of the per-CTA append counts -> ring positions
//   k_update   claim index slots for missing keys (dedup by CAS) and append promoted / inserted
//              entries to their bucket's FIFO ring in position order
//   k_fetch    rows of the missing keys from the host-pinned backing store (zero-copy) -> slab +
//              output; runs on a side stream next to k_evict
//   k_evict    flush rule, then evict in (bucket, FIFO) order down to capacity; one CTA per tier
//   k_compact  squeeze dead records out of one bucket ring (rare, host-triggered)
#pragma once
#include "evs_codec.cuh"
#include "evs_types.cuh"

namespace evs {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kEvictThreads = 1024;
constexpr int kEvictPerThread = 4;           // ring records one thread examines per window

__device__ __forceinline__ uint4 ldg16(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ unsigned long long u64_of(unsigned lo, unsigned hi) {
    return (static_cast<unsigned long long>(hi) << 32) | lo;
}

// ---- index probe ---------------------------------------------------------------------
// Linear probing.  The walk ends at the key or at the first slot no resident key's path crosses
// (pass == 0).  `v` is the already loaded content of slot `i` (lets two tiers' first loads overlap).
__device__ __forceinline__ bool probe_from(const Slot *__restrict__ slots, unsigned mask, unsigned long long key,
                                           unsigned i, uint4 v, unsigned &slot_out, unsigned long long &meta_out) {
    while (true) {
        const unsigned long long kw = u64_of(v.x, v.y);
        if ((kw & kKeyMask) == key) {
            slot_out = i;
            meta_out = u64_of(v.z, v.w);
            return true;
        }
        if ((kw >> 48) == 0ull) return false;
        i = (i + 1) & mask;
        v = ldg16(slots + i);
    }
}

__device__ __forceinline__ bool probe(const TierDev &t, unsigned long long key, unsigned &slot, unsigned long long &meta) {
    const unsigned i = hash_key(key, t.hash_mask);
    return probe_from(t.slots, t.hash_mask, key, i, ldg16(t.slots + i), slot, meta);
}

// C3: key -> alt key (aprx_embedding.cpp:344 get_altkey_str).  Plain loads: the recency flag of
// these slots is written by this same kernel.
__device__ __forceinline__ bool c3_find(const C3Dev &c, unsigned long long key, unsigned &slot_out, unsigned &alt_out) {
    unsigned i = hash_key(key, c.hash_mask);
    while (true) {
        const unsigned long long kw = c.slots[i].kw;
        if ((kw & kKeyMask) == key) {
            slot_out = i;
            alt_out = c.slots[i].alt;
            return true;
        }
        if ((kw >> 48) == 0ull) return false;
        i = (i + 1) & c.hash_mask;
    }
}

// ---- k_serve ---------------------------------------------------------------------------
// Gather the rows tier `k` serves for this warp's sample: lanes walk the sample's T*CPR 16-byte
// chunks so consecutive lanes read consecutive 16 B of a row and write consecutive floats of
// the output; all loads of an unrolled group are issued before the first decode/store.
template <int PREC>
__device__ __forceinline__ void gather_tier(const TierDev &t, int k, int src_t, unsigned src_s, float *orow, int T, int D,
                                            bool vec, int lane, const CodecLut *lut) {
    if (__ballot_sync(kFull, src_t == k) == 0u) return;
    constexpr int U = 4;
    const int cpr = static_cast<int>(t.row_stride >> 4);
    const int total = T * cpr;
    for (int c0 = 0; c0 < total; c0 += 32 * U) {
        uint4 v[U];
        int tt[U], part[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + u * 32 + lane;
            const bool inb = c < total;
            tt[u] = inb ? c / cpr : 0;
            part[u] = c - tt[u] * cpr;
            const int st = __shfl_sync(kFull, src_t, tt[u]);
            const unsigned sl = __shfl_sync(kFull, src_s, tt[u]);
            ok[u] = inb && (st == k);
            if (ok[u]) v[u] = ldg16(t.slab + static_cast<size_t>(sl) * t.row_stride + (part[u] << 4));
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (ok[u]) decode_store<PREC>(v[u], orow + tt[u] * D, part[u], D, vec, lut);
    }
}

// One warp per sample, lane t < T holds the key of table t.  A CTA covers kSamplesPerCta
// consecutive samples so the int64 index tile is read as T runs of 64 contiguous bytes.
// P1 == 0: single tier.  The cache state is only read here (C3 recency flags excepted).
template <int P0, int P1>
__global__ void __launch_bounds__(kLookupThreads) k_serve(const __grid_constant__ Params p) {
    __shared__ long long s_idx[kSamplesPerCta][kMaxTables];
    __shared__ unsigned s_hist[kSeqs];
    __shared__ unsigned s_stat[8];           // hits C1, hits C2, C3, approx, misses, perfect
    __shared__ CodecLut s_lut;

    const BatchArgs a = *p.args;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = p.T, B = a.B, D = p.D;
    const int s0 = blockIdx.x * kSamplesPerCta;
    if (s0 >= B) return;

    for (int i = threadIdx.x; i < kSamplesPerCta * T; i += kLookupThreads) {
        const int t = i / kSamplesPerCta, j = i - t * kSamplesPerCta;
        const int s = s0 + j;
        s_idx[j][t] = (s < B) ? __ldg(a.idx + static_cast<size_t>(t) * B + s) : 0;
    }
    if (threadIdx.x < kSeqs) s_hist[threadIdx.x] = 0;
    if (threadIdx.x < 8) s_stat[threadIdx.x] = 0;
    codec_lut_init<P0, P1>(&s_lut);
    const TierDev &t0 = p.tier[0];
    const TierDev &t1 = p.tier[1];
    const int full = (P1 != 0) ? static_cast<int>(t0.ctl->full_at_start) : 0;
    __syncthreads();

    const int s = s0 + warp;
    const bool wact = s < B;                   // a warp past the batch end only joins the barriers
    const bool act = wact && lane < T;
    unsigned long long key = 0, m0 = 0, m1 = 0;
    unsigned slot0 = 0, slot1 = 0;
    bool h0 = false, h1 = false;
    if (act) {
        long long r = s_idx[warp][lane];
        if (r < 0 || r >= __ldg(p.rows + lane)) {
            p.g->error = 1u;                   // EVS_ERR_INDEX; answer from row 0
            r = 0;
        }
        key = make_key(p.table_base + lane, r);
        const unsigned i0 = hash_key(key, t0.hash_mask);
        const uint4 v0 = ldg16(t0.slots + i0);
        if (P1 != 0) {
            const unsigned i1 = hash_key(key, t1.hash_mask);
            const uint4 v1 = ldg16(t1.slots + i1);
            h1 = probe_from(t1.slots, t1.hash_mask, key, i1, v1, slot1, m1);
        }
        h0 = probe_from(t0.slots, t0.hash_mask, key, i0, v0, slot0, m0);
    }

    // C3 (evlfu_8.cpp:474-489, 528-558): a double miss whose alternative key is resident in C1,
    // else C2, is answered with that row; the C3 entry gets its recency flag.
    bool c3hit = false;
    int c3_tier = 0;
    unsigned c3_slot = 0;
    if (P1 != 0 && p.c3.active) {
        if (act && !h0 && !h1) {
            unsigned cs, alt;
            if (c3_find(p.c3, key, cs, alt)) {
                const unsigned long long akey = make_key(static_cast<int>(alt % 100u) - 1, static_cast<long long>(alt / 100u));
                unsigned long long am;
                if (probe(t0, akey, c3_slot, am)) {
                    c3hit = true;
                    c3_tier = 0;
                } else if (probe(t1, akey, c3_slot, am)) {
                    c3hit = true;
                    c3_tier = 1;
                }
                if (c3hit) p.c3.slots[cs].flag = 1u;       // set_recency_flag_c3 (aprx_embedding.cpp:402)
            }
        }
    }

    const unsigned m_h0 = __ballot_sync(kFull, h0);
    const unsigned m_h1 = __ballot_sync(kFull, h1);
    const unsigned m_c3 = __ballot_sync(kFull, c3hit);
    int agg;
    if (P1 == 0) agg = __popc(m_h0);                                   // evlfu_32.cpp:477-492
    else if (full) agg = __popc(m_h0 | m_h1) + __popc(m_c3);           // evlfu_8.cpp:512-541
    else agg = __popc(m_h0);                                           // evlfu_8.cpp:601 (C1 not full)
    const int local_agg = agg;
    if (a.agg_in != nullptr && wact) agg = a.agg_in[s];
    if (a.probe_only) {
        if (lane == 0 && wact) a.agg_out[s] = static_cast<uint8_t>(local_agg);
        return;
    }
    const bool approx = (P1 == 0) && (p.approx_thres > 0) && (agg >= p.approx_thres) && (m_h0 != 0u);

    uint8_t f = 0, hc = kHitMiss;
    int src_t = -1;
    unsigned src_s = 0;
    const int pos = s * T + lane;
    if (act) {
        if (h0) {                                                      // C1 serves (and overrides C2)
            hc = kHitC1;
            src_t = 0;
            src_s = slot0;
            if (meta_bucket(m0) < agg) {
                f = static_cast<uint8_t>(agg + 1);
                p.pos_slot[pos] = slot0;
            }
        } else if (c3hit) {
            hc = kHitC3;
            src_t = c3_tier;
            src_s = c3_slot;
        } else if (P1 != 0 && full && h1) {                            // C2 serves
            hc = kHitC2;
            src_t = 1;
            src_s = slot1;
            if (meta_bucket(m1) < agg) {
                f = static_cast<uint8_t>(kFlagTier | (agg + 1));
                p.pos_slot[pos] = slot1;
            }
        } else if (approx) {                                           // EvLFU_C1.py:140-152
            hc = kHitApprox;
            src_t = 0;
        } else {
            // fetch and insert.  C1 not full: everything goes to C1 (evlfu_8.cpp:590-602).  C1 full:
            // odd tables to C1, even to C2 while agg < high_agghit_threshold, else all to C2 (:573-588).
            int tier_ins = 0;
            if (P1 != 0 && full) tier_ins = (agg < p.high_thres && ((p.table_base + lane) & 1)) ? 0 : 1;
            f = static_cast<uint8_t>(kFlagMiss | (tier_ins ? kFlagTier : 0) | (agg + 1));
        }
        p.flags[pos] = f;
        if (a.hit != nullptr) a.hit[pos] = hc;
    }
    if (P1 == 0 && p.approx_thres > 0) {
        // value of the latest earlier hit in table order, else of the first hit
        const unsigned lower = m_h0 & ((1u << lane) - 1u);
        const int j = lower ? (31 - __clz(lower)) : (m_h0 ? __ffs(m_h0) - 1 : 0);
        const unsigned subst = __shfl_sync(kFull, slot0, j);
        if (hc == kHitApprox) src_s = subst;
    }

    const unsigned m_f0 = __ballot_sync(kFull, f != 0 && !(f & kFlagTier));
    const unsigned m_f1 = __ballot_sync(kFull, (f & kFlagTier) != 0);
    const unsigned m_miss = __ballot_sync(kFull, (f & kFlagMiss) != 0);
    const unsigned m_c2 = __ballot_sync(kFull, hc == kHitC2);
    const unsigned m_ap = __ballot_sync(kFull, hc == kHitApprox);
    if (lane == 0 && wact) {
        a.agg_out[s] = static_cast<uint8_t>(agg);
        if (m_f0) atomicAdd(&s_hist[agg], static_cast<unsigned>(__popc(m_f0)));
        if (m_f1) atomicAdd(&s_hist[kMaxBuckets + agg], static_cast<unsigned>(__popc(m_f1)));
        if (m_h0) atomicAdd(&s_stat[0], static_cast<unsigned>(__popc(m_h0)));
        if (m_c2) atomicAdd(&s_stat[1], static_cast<unsigned>(__popc(m_c2)));
        if (m_c3) atomicAdd(&s_stat[2], static_cast<unsigned>(__popc(m_c3)));
        if (m_ap) atomicAdd(&s_stat[3], static_cast<unsigned>(__popc(m_ap)));
        if (m_miss) atomicAdd(&s_stat[4], static_cast<unsigned>(__popc(m_miss)));
        if (agg == p.n_perfect_agg) atomicAdd(&s_stat[5], 1u);
    }

    if (wact) {
        float *orow = a.out + static_cast<size_t>(s) * a.out_stride;
        const bool vec = ((reinterpret_cast<uintptr_t>(a.out) & 15u) == 0) && ((a.out_stride & 3) == 0) && ((D & 3) == 0);
        gather_tier<P0>(t0, 0, src_t, src_s, orow, T, D, vec, lane, &s_lut);
        if (P1 != 0) gather_tier<(P1 != 0 ? P1 : 32)>(t1, 1, src_t, src_s, orow, T, D, vec, lane, &s_lut);
    }

    __syncthreads();
    const int n_seq = p.n_tiers * kMaxBuckets;
    if (threadIdx.x < n_seq) p.hist[static_cast<size_t>(threadIdx.x) * p.n_chunks_max + blockIdx.x] = s_hist[threadIdx.x];
    if (threadIdx.x == 0) {
        GlobalCtl *g = p.g;
        const int ns = min(kSamplesPerCta, B - s0);
        atomicAdd(&g->lookups, static_cast<unsigned long long>(ns) * T);
        atomicAdd(&g->samples, static_cast<unsigned long long>(ns));
        if (s_stat[0]) atomicAdd(&g->hits[0], static_cast<unsigned long long>(s_stat[0]));
        if (s_stat[1]) atomicAdd(&g->hits[1], static_cast<unsigned long long>(s_stat[1]));
        if (s_stat[2]) atomicAdd(&g->c3_hits, static_cast<unsigned long long>(s_stat[2]));
        if (s_stat[3]) atomicAdd(&g->approx_subst, static_cast<unsigned long long>(s_stat[3]));
        if (s_stat[4]) atomicAdd(&g->misses, static_cast<unsigned long long>(s_stat[4]));
        if (s_stat[5]) {
            atomicAdd(&g->perfect_hits, static_cast<unsigned long long>(s_stat[5]));
            t0.ctl->any_perfect = 1u;                              // EvLFU_C1.py:163-165
            if (P1 != 0 && full) t1.ctl->any_perfect = 1u;         // evlfu_8.cpp:439-441 (only when C2 is updated)
        }
    }
}

// ---- k_scan ------------------------------------------------------------------------------
// hist[seq][chunk] (append requests of serve-CTA `chunk` for (tier, bucket) = seq) -> offset of
// that chunk's first record past the old tail; tails advance by the totals.  One CTA per seq.
__global__ void __launch_bounds__(256) k_scan(const __grid_constant__ Params p) {
    __shared__ unsigned s_w[8];
    const int B = p.args->B;
    const int n_chunks = (B + kSamplesPerCta - 1) / kSamplesPerCta;
    const int nb = p.tier[0].n_buckets;
    const int tier = blockIdx.x / nb, b = blockIdx.x - tier * nb;
    unsigned *h = p.hist + static_cast<size_t>(tier * kMaxBuckets + b) * p.n_chunks_max;
    TierCtl *c = p.tier[tier].ctl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned running = 0;
    for (int base = 0; base < n_chunks; base += 256) {
        const int i = base + threadIdx.x;
        const unsigned v = (i < n_chunks) ? h[i] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned n = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += n;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        unsigned woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const unsigned x = s_w[w];
            if (w < warp) woff += x;
            tot += x;
        }
        if (i < n_chunks) h[i] = running + woff + incl - v;
        running += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned long long tl = c->tail[b];
        if (tl + running - c->head[b] > p.tier[tier].ring_cap) c->error = 4u;
        c->tail_prev[b] = tl;
        c->tail[b] = tl + running;
    }
}

// ---- k_update ----------------------------------------------------------------------------
// Claim a slot for `key` (or find the slot a same-batch duplicate already claimed).  Only inserts
// run concurrently here, so the key field goes EMPTY -> occupied monotonically and every thread
// inserting the same key converges on one slot.  pass bits of the word may change under us
// (other claimers crossing), hence the CAS retry on the same slot.
__device__ __forceinline__ unsigned claim_slot(const TierDev &tier, unsigned long long key, bool &claimed) {
    const unsigned mask = tier.hash_mask;
    const unsigned home = hash_key(key, mask);
    unsigned i = home;
    claimed = false;
    while (true) {
        const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&tier.slots[i].kw);
        const unsigned long long k = cur & kKeyMask;
        if (k == key) break;
        if (k == kEmptyKey) {
            const unsigned long long old = atomicCAS(&tier.slots[i].kw, cur, (cur & ~kKeyMask) | key);
            if (old == cur) {
                claimed = true;
                break;
            }
            continue;                            // look at the same slot again
        }
        i = (i + 1) & mask;
    }
    if (claimed) {
        for (unsigned j = home; j != i; j = (j + 1) & mask) {
            const unsigned long long old = atomicAdd(&tier.slots[j].kw, kPassOne);
            if ((old >> 48) == 0xFFFFull) tier.ctl->error = 5u;      // pass counter overflow
        }
        atomicAdd(&tier.ctl->n_new, 1u);
    }
    return i;
}

// Same thread <-> position mapping as k_serve, so ring order inside a bucket is position
// order: by warp (sample) then lane (table).  All flagged positions of a sample share the
// bucket agg_hit(sample).
__global__ void __launch_bounds__(kLookupThreads) k_update(const __grid_constant__ Params p) {
    __shared__ unsigned s_cnt[kSamplesPerCta][kMaxTiers];
    __shared__ int s_b[kSamplesPerCta];
    const BatchArgs a = *p.args;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = p.T, B = a.B;
    const int s = blockIdx.x * kSamplesPerCta + warp;
    if (blockIdx.x * kSamplesPerCta >= B) return;
    const bool act = (s < B) && (lane < T);
    const int pos = s * T + lane;
    const unsigned f = act ? p.flags[pos] : 0u;
    if (__syncthreads_or(f != 0u) == 0) return;

    const int tr = (f & kFlagTier) ? 1 : 0;
    const int b = static_cast<int>(f & 0x3Fu) - 1;
    const unsigned m0 = __ballot_sync(kFull, f != 0u && tr == 0);
    const unsigned m1 = __ballot_sync(kFull, f != 0u && tr == 1);
    if (lane == 0) {
        s_cnt[warp][0] = __popc(m0);
        s_cnt[warp][1] = __popc(m1);
    }
    const unsigned any = m0 | m1;
    const int wb = __shfl_sync(kFull, b, any ? (__ffs(any) - 1) : 0);
    if (lane == 0) s_b[warp] = any ? wb : -1;
    __syncthreads();
    if (!f) return;

    const TierDev &tier = p.tier[tr];
    unsigned slot;
    if (f & kFlagMiss) {
        long long r = __ldg(a.idx + static_cast<size_t>(lane) * B + s);
        if (r < 0 || r >= __ldg(p.rows + lane)) r = 0;
        bool claimed;
        slot = claim_slot(tier, make_key(p.table_base + lane, r), claimed);
        p.pos_slot[pos] = slot | (claimed ? kClaimedBit : 0u);
        atomicMax(&tier.ctl->prot, (static_cast<unsigned long long>(f & 0x3Fu) << 32) | static_cast<unsigned>(pos));
    } else {
        slot = p.pos_slot[pos];
    }

    unsigned pre = 0;
    for (int w = 0; w < warp; ++w)
        if (s_b[w] == b) pre += s_cnt[w][tr];
    const unsigned rank = __popc((tr ? m1 : m0) & ((1u << lane) - 1u));
    TierCtl *c = tier.ctl;
    const unsigned long long q =
        c->tail_prev[b] + p.hist[static_cast<size_t>(tr * kMaxBuckets + b) * p.n_chunks_max + blockIdx.x] + pre + rank;
    tier.ring[static_cast<size_t>(b) * tier.ring_cap + (q & (tier.ring_cap - 1))] = slot;
    const unsigned long long mine = pack_meta(b, q);
    const unsigned long long old = atomicMax(&tier.slots[slot].meta, mine);
    if (mine > old) {
        const int ob = meta_bucket(old);
        if (ob != b) {
            atomicAdd(&c->count[b], 1u);
            if (ob >= 0) atomicSub(&c->count[ob], 1u);
            else atomicAdd(&c->stat_inserts, 1ull);
        }
    }
}

// ---- k_fetch -----------------------------------------------------------------------------
// Copy one backing-store row (row_bytes, arbitrary alignment) into a 16-byte aligned staging
// row of `stride` bytes, zero padded.  Zero-copy reads when the store lives in pinned host memory.
__device__ __forceinline__ void fetch_row(const unsigned char *__restrict__ src, unsigned row_bytes,
                                          unsigned char *stage, unsigned stride, int lane) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
    if (((a | row_bytes) & 3u) == 0) {
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

template <int PREC>
__device__ __forceinline__ void fetch_one(const TierDev &tier, const Params &p, const BatchArgs &a, int pos, int lane,
                                          int glane, int gsize, unsigned char *stage, bool vec, const CodecLut *lut) {
    // executed by a group of `gsize` consecutive lanes (glane = lane within the group); the group
    // is the whole warp on the staged (unaligned) path
    const int T = p.T, D = p.D;
    const int s = pos / T, t = pos - s * T;
    long long r = __ldg(a.idx + static_cast<size_t>(t) * a.B + s);
    if (r < 0 || r >= __ldg(p.rows + t)) r = 0;
    const unsigned ps = p.pos_slot[pos];
    const unsigned slot = ps & ~kClaimedBit;
    const bool claimed = (ps & kClaimedBit) != 0u;
    const unsigned char *src = tier.store[t] + static_cast<size_t>(r) * tier.row_bytes;
    float *orow = a.out + static_cast<size_t>(s) * a.out_stride + t * D;
    const int cpr = static_cast<int>(tier.row_stride >> 4);
    unsigned char *dst = tier.slab + static_cast<size_t>(slot) * tier.row_stride;
    if (stage == nullptr) {                       // 16-byte aligned rows: straight through registers
        for (int c = glane; c < cpr; c += gsize) {
            const uint4 v = ldg16(src + (c << 4));
            if (claimed) *reinterpret_cast<uint4 *>(dst + (c << 4)) = v;
            decode_store<PREC>(v, orow, c, D, vec, lut);
        }
    } else {
        fetch_row(src, tier.row_bytes, stage, tier.row_stride, lane);
        __syncwarp();
        for (int c = lane; c < cpr; c += 32) {
            const uint4 v = *reinterpret_cast<const uint4 *>(stage + (c << 4));
            if (claimed) *reinterpret_cast<uint4 *>(dst + (c << 4)) = v;
            decode_store<PREC>(v, orow, c, D, vec, lut);
        }
        __syncwarp();
    }
}

// A warp scans 32 positions' flags at a time and serves the misses among them.
// Dynamic shared memory: warps * max(row_stride) bytes of staging.
template <int P0, int P1>
__global__ void __launch_bounds__(256) k_fetch(const __grid_constant__ Params p) {
    extern __shared__ __align__(16) unsigned char s_stage[];
    __shared__ CodecLut s_lut;
    codec_lut_init<P0, P1>(&s_lut);
    __syncthreads();
    const BatchArgs a = *p.args;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpc = blockDim.x >> 5;
    const int N = a.B * p.T;
    const TierDev &t0 = p.tier[0];
    const TierDev &t1 = p.tier[1];
    const unsigned max_stride = (P1 != 0 && t1.row_stride > t0.row_stride) ? t1.row_stride : t0.row_stride;
    unsigned char *stage = s_stage + static_cast<size_t>(warp) * max_stride;
    const bool vec = ((reinterpret_cast<uintptr_t>(a.out) & 15u) == 0) && ((a.out_stride & 3) == 0) && ((p.D & 3) == 0);
    // rows of a tier can go straight through registers when every row start is 16-byte aligned
    bool al0 = (t0.row_bytes & 15u) == 0, al1 = (P1 != 0) && (t1.row_bytes & 15u) == 0;
    for (int t = 0; t < p.T; ++t) {
        al0 = al0 && ((reinterpret_cast<uintptr_t>(t0.store[t]) & 15u) == 0);
        if (P1 != 0) al1 = al1 && ((reinterpret_cast<uintptr_t>(t1.store[t]) & 15u) == 0);
    }
    auto group_of = [](unsigned stride) {
        int g = 1;
        while (g < static_cast<int>(stride >> 4) && g < 32) g <<= 1;
        return g;
    };
    const int g0 = group_of(t0.row_stride), g1 = (P1 != 0) ? group_of(t1.row_stride) : 32;

    for (int base = (blockIdx.x * wpc + warp) * 32; base < N; base += gridDim.x * wpc * 32) {
        const int pos = base + lane;
        const unsigned f = (pos < N) ? p.flags[pos] : 0u;
        unsigned m0 = __ballot_sync(kFull, (f & kFlagMiss) && !(f & kFlagTier));
        unsigned m1 = __ballot_sync(kFull, (f & kFlagMiss) && (f & kFlagTier));
        // tier 0 misses
        if (al0) {
            const int ngrp = 32 / g0, grp = lane / g0, gl = lane - grp * g0;
            while (m0) {
                unsigned mm = m0;
                int mine = -1;
                for (int k = 0; k < ngrp && mm; ++k) {
                    const int bit = __ffs(mm) - 1;
                    mm &= mm - 1;
                    if (k == grp) mine = bit;
                }
                m0 = mm;
                if (mine >= 0) fetch_one<P0>(t0, p, a, base + mine, lane, gl, g0, nullptr, vec, &s_lut);
            }
        } else {
            while (m0) {
                const int bit = __ffs(m0) - 1;
                m0 &= m0 - 1;
                fetch_one<P0>(t0, p, a, base + bit, lane, lane, 32, stage, vec, &s_lut);
            }
        }
        if (P1 != 0) {
            if (al1) {
                const int ngrp = 32 / g1, grp = lane / g1, gl = lane - grp * g1;
                while (m1) {
                    unsigned mm = m1;
                    int mine = -1;
                    for (int k = 0; k < ngrp && mm; ++k) {
                        const int bit = __ffs(mm) - 1;
                        mm &= mm - 1;
                        if (k == grp) mine = bit;
                    }
                    m1 = mm;
                    if (mine >= 0) fetch_one<(P1 != 0 ? P1 : 32)>(t1, p, a, base + mine, lane, gl, g1, nullptr, vec, &s_lut);
                }
            } else {
                while (m1) {
                    const int bit = __ffs(m1) - 1;
                    m1 &= m1 - 1;
                    fetch_one<(P1 != 0 ? P1 : 32)>(t1, p, a, base + bit, lane, lane, 32, stage, vec, &s_lut);
                }
            }
        }
    }
}

// ---- k_evict ---------------------------------------------------------------------------
__device__ __forceinline__ void evict_slot(const TierDev &tier, unsigned slot, unsigned long long key) {
    const unsigned mask = tier.hash_mask;
    for (unsigned j = hash_key(key, mask); j != slot; j = (j + 1) & mask) atomicAdd(&tier.slots[j].kw, ~kPassOne + 1ull);
    tier.slots[slot].meta = 0ull;
    atomicOr(&tier.slots[slot].kw, kEmptyKey);        // keep the pass bits: other keys may cross this slot
}

// Pop up to `want` live records from the head of bucket b (whole CTA).  The protected slot is
// skipped but keeps its place.  Returns the number popped (uniform across the CTA).
__device__ unsigned pop_bucket(const TierDev &tier, int b, unsigned want, unsigned prot_slot,
                               unsigned long long *out_keys, unsigned out_base) {
    __shared__ unsigned s_wsum[32];
    __shared__ unsigned s_total;
    __shared__ unsigned long long s_first_kept;
    constexpr int R = kEvictPerThread;
    TierCtl *c = tier.ctl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const unsigned long long h = c->head[b], tl = c->tail[b];
    __syncthreads();                       // everyone has read head/tail before thread 0 rewrites them
    if (threadIdx.x == 0) s_first_kept = ~0ull;
    __syncthreads();
    unsigned got = 0;
    unsigned long long base = h;
    const unsigned *ring = tier.ring + static_cast<size_t>(b) * tier.ring_cap;
    for (; base < tl && got < want; base += static_cast<unsigned long long>(blockDim.x) * R) {
        const unsigned long long q0 = base + static_cast<unsigned long long>(threadIdx.x) * R;
        unsigned slot[R];
        uint4 sv[R];
        bool live[R], cand[R];
#pragma unroll
        for (int r = 0; r < R; ++r) slot[r] = (q0 + r < tl) ? ring[(q0 + r) & (tier.ring_cap - 1)] : kNoSlot;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (slot[r] <= tier.hash_mask) sv[r] = *reinterpret_cast<const uint4 *>(tier.slots + slot[r]);
        unsigned mycnt = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            live[r] = (slot[r] <= tier.hash_mask) && (u64_of(sv[r].z, sv[r].w) == pack_meta(b, q0 + r));
            cand[r] = live[r] && (slot[r] != prot_slot);
            mycnt += cand[r] ? 1u : 0u;
        }
        // block-wide exclusive scan of mycnt
        unsigned incl = mycnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned n = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += n;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned v = (lane < nwarp) ? s_wsum[lane] : 0u;
            unsigned wi = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned n = __shfl_up_sync(kFull, wi, d);
                if (lane >= d) wi += n;
            }
            s_wsum[lane] = wi - v;
            if (lane == 31) s_total = wi;
        }
        __syncthreads();
        unsigned idx = s_wsum[warp] + incl - mycnt;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool take = cand[r] && (got + idx < want);
            if (take) {
                const unsigned long long key = u64_of(sv[r].x, sv[r].y) & kKeyMask;
                evict_slot(tier, slot[r], key);
                if (out_keys != nullptr) out_keys[out_base + got + idx] = key;
            }
            if (live[r] && !take) atomicMin(&s_first_kept, q0 + r);
            idx += cand[r] ? 1u : 0u;
        }
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

// One CTA per tier.
__global__ void __launch_bounds__(kEvictThreads) k_evict(const __grid_constant__ Params p) {
    __shared__ unsigned s_size;
    const TierDev &tier = p.tier[blockIdx.x];
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
        flushed = pop_bucket(tier, top, want, kNoSlot, tier.flushed, 0);
        size -= flushed;
        if (threadIdx.x == 0) c->n_perfect = c->count[top];
        __syncthreads();
    }

    // evict down to capacity, lowest bucket first, FIFO inside a bucket; the highest ranked
    // new key is never the victim (the reference evicts before it inserts)
    unsigned prot_slot = kNoSlot;
    if (n_new > 0) prot_slot = p.pos_slot[static_cast<unsigned>(prot & 0xFFFFFFFFull)] & ~kClaimedBit;
    unsigned need = size > tier.cap ? size - tier.cap : 0u;
    unsigned ev = 0;
    for (int b = 0; b <= top && need > 0; ++b) {
        if (c->count[b] == 0) continue;
        const unsigned got = pop_bucket(tier, b, need, prot_slot, tier.evicted, ev);
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
        c->n_new = 0;
        c->prot = 0ull;
        c->any_perfect = 0;
        if (need > 0) c->error = 4u;
        if (blockIdx.x == 0 && p.g != nullptr) atomicAdd(&p.g->batches, 1ull);
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
    unsigned *ring = tier.ring + static_cast<size_t>(b) * tier.ring_cap;
    for (unsigned long long base = h; base < tl; base += blockDim.x) {
        const unsigned long long q = base + threadIdx.x;
        bool valid = false;
        unsigned slot = kNoSlot;
        if (q < tl) {
            slot = ring[q & (tier.ring_cap - 1)];
            if (slot <= tier.hash_mask) valid = (tier.slots[slot].meta == pack_meta(b, q));
        }
        const unsigned bal = __ballot_sync(kFull, valid);
        if (lane == 0) s_wsum[warp] = __popc(bal);
        __syncthreads();                    // also: every record of this window has been read
        if (warp == 0) {
            const unsigned v = (lane < nwarp) ? s_wsum[lane] : 0u;
            unsigned incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned n = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += n;
            }
            s_wsum[lane] = incl - v;
            if (lane == 31) s_total = incl;
        }
        __syncthreads();
        if (valid) {
            const unsigned long long nq = dst + s_wsum[warp] + __popc(bal & ((1u << lane) - 1u));
            ring[nq & (tier.ring_cap - 1)] = slot;
            tier.slots[slot].meta = pack_meta(b, nq);
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
        tier.slots[j].kw = kEmptyKey;
        tier.slots[j].meta = 0ull;
    }
}

}  // namespace evs
