// k_update: everything of a batch that changes the cache, in one launch.
//
//   all CTAs   same warp <-> sample, lane <-> table mapping as k_serve.  A flagged position
//              (a) claims an index slot if it missed (CAS; same-batch duplicates converge on one slot),
//              (b) appends a record to the FIFO ring of bucket agg_hit(sample) at the position its
//                  rank among ALL flagged positions of the batch dictates (sample-major, table-minor:
//                  the order EvLFU_C1.py processes them), computed from k_serve's per-CTA counts,
//              (c) atomicMax on the slot's meta picks the winning occurrence of a key (highest
//                  (agg_hit, position)), bucket counters follow,
//              (d) if it missed, fetches its row from the host-pinned backing store (zero-copy) into
//                  the slab row it claimed and, dequantised, into the output.
// k_evict (one 1024-thread CTA per tier) then advances the ring tails, applies the flush rule,
// evicts down to capacity and inserts the victims into C3.
#pragma once
#include "evs_c3.cuh"
#include "evs_kernels.cuh"

namespace evs {

template <int P0, int P1>
__global__ void __launch_bounds__(kLookupThreads) k_update(const __grid_constant__ Params p) {
    extern __shared__ __align__(16) unsigned char s_stage[];    // warps * max(row_stride), unaligned rows only
    __shared__ unsigned s_cnt[kSamplesPerCta][kMaxTiers];
    __shared__ int s_b[kSamplesPerCta];
    __shared__ int s_delta[kSeqs];
    __shared__ unsigned s_new[kMaxTiers], s_ins[kMaxTiers];
    __shared__ unsigned long long s_prot[kMaxTiers];
    __shared__ CodecLut s_lut;
    __shared__ int s_first_any;

    const unsigned long long t_start = gtime();
    unsigned long long t_pre = t_start, t_app = t_start;
    if (blockIdx.x == 0 && threadIdx.x == 0) p.dbg[2] = t_start;
    const BatchArgs a = *p.args;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = p.T, B = a.B, D = p.D;
    const int n_chunks = (B + kSamplesPerCta - 1) / kSamplesPerCta;
    if (static_cast<int>(blockIdx.x) >= n_chunks) return;
    const int s = blockIdx.x * kSamplesPerCta + warp;
    const bool act = (s < B) && (lane < T);
    const int pos = s * T + lane;
    const unsigned f = act ? p.flags[pos] : 0u;
    long long r = 0;
    if (f & kFlagMiss) {
        r = __ldg(a.idx + static_cast<size_t>(lane) * B + s);
        if (r < 0 || r >= __ldg(p.rows + lane)) r = 0;
    }
    if (threadIdx.x < kSeqs) s_delta[threadIdx.x] = 0;
    if (threadIdx.x < kMaxTiers) {
        s_new[threadIdx.x] = 0;
        s_ins[threadIdx.x] = 0;
        s_prot[threadIdx.x] = 0ull;
    }
    codec_lut_init<P0, P1>(&s_lut);

    const int tr = (f & kFlagTier) ? 1 : 0;
    const int b = static_cast<int>(f & 0x3Fu) - 1;
    const unsigned m0 = __ballot_sync(kFull, f != 0u && tr == 0);
    const unsigned m1 = __ballot_sync(kFull, f != 0u && tr == 1);
    const unsigned any = m0 | m1;
    const int wb = __shfl_sync(kFull, b, any ? (__ffs(any) - 1) : 0);    // all flagged lanes share it
    if (threadIdx.x == 0) s_first_any = 99;
    __syncwarp();
    if (lane == 0) {
        s_cnt[warp][0] = __popc(m0);
        s_cnt[warp][1] = __popc(m1);
        s_b[warp] = any ? wb : -1;
    }
    __syncthreads();
    if (lane == 0 && any) atomicMin(&s_first_any, warp);
    __syncthreads();

    if (any) {
        // records of earlier chunks in my bucket's sequences (k_scan already made them prefixes for
        // very large batches), then of earlier samples of this chunk
        unsigned base0 = 0, base1 = 0;
        const unsigned *h0 = p.hist + static_cast<size_t>(wb) * p.n_chunks_max;
        const unsigned *h1 = p.hist + static_cast<size_t>(kMaxBuckets + wb) * p.n_chunks_max;
        if (n_chunks > kQuadMaxChunks) {
            if (m0) base0 = __ldcg(h0 + blockIdx.x);
            if (m1) base1 = __ldcg(h1 + blockIdx.x);
        } else {
            unsigned p0 = 0, p1 = 0;
            for (int c = lane; c < static_cast<int>(blockIdx.x); c += 32) {
                if (m0) p0 += __ldcg(h0 + c);
                if (m1) p1 += __ldcg(h1 + c);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                p0 += __shfl_xor_sync(kFull, p0, d);
                p1 += __shfl_xor_sync(kFull, p1, d);
            }
            base0 = p0;
            base1 = p1;
        }
        t_pre = gtime();
        for (int w = 0; w < warp; ++w)
            if (s_b[w] == wb) {
                base0 += s_cnt[w][0];
                base1 += s_cnt[w][1];
            }

        unsigned slotword = 0;
        if (f) {
            const TierDev &tier = p.tier[tr];
            unsigned slot;
            if (f & kFlagMiss) {
                bool claimed;
                slot = claim_slot(tier, make_key(p.table_base + lane, r), claimed);
                slotword = slot | (claimed ? kClaimedBit : 0u);
                p.pos_slot[pos] = slotword;
                if (claimed) atomicAdd(&s_new[tr], 1u);
                atomicMax(&s_prot[tr], (static_cast<unsigned long long>(f & 0x3Fu) << 32) | static_cast<unsigned>(pos));
            } else {
                slot = p.pos_slot[pos];
            }
            const unsigned rank = __popc((tr ? m1 : m0) & ((1u << lane) - 1u));
            const volatile TierCtl *c = tier.ctl;
            const unsigned long long tbase = (n_chunks > kQuadMaxChunks) ? c->tail_prev[b] : c->tail[b];
            const unsigned long long q = tbase + (tr ? base1 : base0) + rank;
            tier.ring[static_cast<size_t>(b) * tier.ring_cap + (q & (tier.ring_cap - 1))] = slot;
            const unsigned long long mine = pack_meta(b, q);
            const unsigned long long old = atomicMax(&tier.slots[slot].meta, mine);
            if (mine > old) {
                const int ob = meta_bucket(old);
                if (ob != b) {
                    atomicAdd(&s_delta[tr * kMaxBuckets + b], 1);
                    if (ob >= 0) atomicSub(&s_delta[tr * kMaxBuckets + ob], 1);
                    else atomicAdd(&s_ins[tr], 1u);
                }
            }
        }

        t_app = gtime();
        // miss fetch: host-pinned rows -> slab (claimer) + output
        const unsigned mm0 = __ballot_sync(kFull, (f & kFlagMiss) && tr == 0);
        const unsigned mm1 = __ballot_sync(kFull, (f & kFlagMiss) && tr == 1);
        if (mm0 | mm1) {
            const bool vec = ((reinterpret_cast<uintptr_t>(a.out) & 15u) == 0) && ((a.out_stride & 3) == 0) && ((D & 3) == 0);
            const TierDev &t0 = p.tier[0];
            const TierDev &t1 = p.tier[1];
            const unsigned max_stride = (P1 != 0 && t1.row_stride > t0.row_stride) ? t1.row_stride : t0.row_stride;
            unsigned char *stage = s_stage + static_cast<size_t>(warp) * max_stride;
            auto group_of = [](unsigned stride) {
                int g = 1;
                while (g < static_cast<int>(stride >> 4) && g < 32) g <<= 1;
                return g;
            };
            if (mm0) fetch_misses<P0>(t0, a, D, s, mm0, r, slotword, lane, (p.store_aligned & 1) != 0, group_of(t0.row_stride), stage, vec, &s_lut);
            if (P1 != 0 && mm1)
                fetch_misses<(P1 != 0 ? P1 : 32)>(t1, a, D, s, mm1, r, slotword, lane, (p.store_aligned & 2) != 0, group_of(t1.row_stride), stage, vec, &s_lut);
        }
    }

    // ---- publish this CTA's counter deltas ----------------------------------------------------
    const unsigned long long t_fetch = gtime();
    if (any && lane == 0 && warp == s_first_any) {
        atomicAdd(&p.dbg[9], t_pre - t_start);
        atomicAdd(&p.dbg[10], t_app - t_start);
        atomicAdd(&p.dbg[11], t_fetch - t_start);
        atomicAdd(&p.dbg[8], 1ull);
    }
    if (any && lane == 0) {
        atomicMax(&p.dbg[12], t_pre - t_start);
        atomicMax(&p.dbg[13], t_app - t_pre);
        atomicMax(&p.dbg[14], t_fetch - t_app);
    }
    __syncthreads();
    if (threadIdx.x < kSeqs) {
        const int d = s_delta[threadIdx.x];
        if (d != 0) {
            const int t = threadIdx.x / kMaxBuckets, bb = threadIdx.x - t * kMaxBuckets;
            atomicAdd(&p.tier[t].ctl->count[bb], static_cast<unsigned>(d));
        }
    } else if (threadIdx.x < kSeqs + kMaxTiers) {
        const int t = threadIdx.x - kSeqs;
        if (t < p.n_tiers) {
            if (s_new[t]) atomicAdd(&p.tier[t].ctl->n_new, s_new[t]);
            if (s_ins[t]) atomicAdd(&p.tier[t].ctl->stat_inserts, static_cast<unsigned long long>(s_ins[t]));
            if (s_prot[t]) atomicMax(&p.tier[t].ctl->prot, s_prot[t]);
        }
    }
    if (threadIdx.x == 0) atomicMax(&p.dbg[3], gtime());
}

// ---- k_evict -------------------------------------------------------------------------------
// One CTA per tier: advance the ring tails by the batch totals, flush rule, evict down to
// capacity; the CTA of tier 0 then inserts the victims of both tiers into C3.
__global__ void __launch_bounds__(kEvictThreads) k_evict(const __grid_constant__ Params p) {
    const int t = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int B = p.args->B;
    const int n_chunks = (B + kSamplesPerCta - 1) / kSamplesPerCta;
    if (t == 0 && threadIdx.x == 0) p.dbg[4] = gtime();
    if (n_chunks <= kQuadMaxChunks) {
        for (int bb = warp; bb < p.tier[t].n_buckets; bb += kEvictThreads / 32) {
            const unsigned *h = p.hist + static_cast<size_t>(t * kMaxBuckets + bb) * p.n_chunks_max;
            unsigned tot = 0;
            for (int c = lane; c < n_chunks; c += 32) tot += __ldcg(h + c);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(kFull, tot, d);
            if (lane == 0 && tot) {
                atomicAdd(&p_dbg_appends, static_cast<unsigned long long>(tot));
                volatile TierCtl *c = p.tier[t].ctl;
                const unsigned long long tl = c->tail[bb] + tot;
                if (tl - c->head[bb] > p.tier[t].ring_cap) c->error = 4u;
                c->tail[bb] = tl;
            }
        }
    }
    __syncthreads();
    evict_tier(p.tier[t], p);
    if (t == 0 && threadIdx.x == 0) p.dbg[5] = gtime();
    if (p.c3.active) {
        // C3 needs the victims of both tiers: the tier-1 CTA publishes completion, tier 0 waits for it
        if (t == 1) {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicExch(p.done, 1u);
            return;
        }
        if (threadIdx.x == 0) {
            while (atomicAdd(p.done, 0u) == 0u) {}
            *p.done = 0u;
            __threadfence();
        }
        __syncthreads();
        c3_update(p);
    }
    if (t == 0 && threadIdx.x == 0) {
        p.dbg[6] = gtime();
        p.dbg[7] = p.dbg[1];
        p.dbg[1] = 0ull;
        atomicAdd(&p.g->batches, 1ull);
    }
}

}  // namespace evs
