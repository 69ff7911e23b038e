// k_update: everything of a batch that changes the cache, in one launch.
//
//   all CTAs   same warp <-> sample, lane <-> table mapping as k_serve.  A flagged position
//              (a) claims an index slot if it missed (CAS; same-batch duplicates converge on one slot),
//              (b) appends a record to the FIFO ring of bucket agg_hit(sample) at the position its
//                  rank among ALL flagged positions of the batch dictates (sample-major, table-minor:
//                  the order EvLFU_C1.py processes them), computed from k_serve's per-CTA counts,
//              (c) atomicMax on the slot's meta picks the winning occurrence of a key (highest
//                  (agg_hit, position)), bucket counters follow,
// The rows of the missing keys are fetched meanwhile by k_fetch on a side stream (output + miss
// staging buffer); k_fill then moves the claimers' rows into their slab rows, next to k_evict.
// k_evict (one 1024-thread CTA per tier) then advances the ring tails, applies the flush rule,
// evicts down to capacity and inserts the victims into C3.
#pragma once
#include "evs_c3.cuh"
#include "evs_kernels.cuh"

namespace evs {

__global__ void __launch_bounds__(kLookupThreads) k_update(const __grid_constant__ Params p) {
    __shared__ unsigned s_cnt[kSamplesPerCta][kMaxTiers];
    __shared__ int s_b[kSamplesPerCta];
    __shared__ int s_delta[kSeqs];
    __shared__ unsigned s_new[kMaxTiers], s_ins[kMaxTiers];
    __shared__ unsigned long long s_prot[kMaxTiers];

    if (blockIdx.x == 0 && threadIdx.x == 0) p.dbg[2] = gtime();
    const BatchArgs a = *p.args;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = p.T, B = a.B;
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

    const int tr = (f & kFlagTier) ? 1 : 0;
    const int b = static_cast<int>(f & 0x3Fu) - 1;
    const unsigned m0 = __ballot_sync(kFull, f != 0u && tr == 0);
    const unsigned m1 = __ballot_sync(kFull, f != 0u && tr == 1);
    const unsigned any = m0 | m1;
    const int wb = __shfl_sync(kFull, b, any ? (__ffs(any) - 1) : 0);    // all flagged lanes share it
    if (lane == 0) {
        s_cnt[warp][0] = __popc(m0);
        s_cnt[warp][1] = __popc(m1);
        s_b[warp] = any ? wb : -1;
    }
    __syncthreads();

    if (any) {
        // the claim (a chain of dependent accesses) goes first so that it overlaps the prefix sums
        unsigned slot = 0;
        bool claimed = false;
        if (f & kFlagMiss) {
            slot = claim_slot(p.tier[tr], make_key(p.table_base + lane, r), claimed);
            p.pos_slot[pos] = slot | (claimed ? kClaimedBit : 0u);
            if (claimed) atomicAdd(&s_new[tr], 1u);
            atomicMax(&s_prot[tr], (static_cast<unsigned long long>(f & 0x3Fu) << 32) | static_cast<unsigned>(pos));
        } else if (f) {
            slot = p.pos_slot[pos];
        }

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
            const int nc = static_cast<int>(blockIdx.x);
            for (int c0 = 0; c0 < nc; c0 += 128) {             // 4 independent loads per lane and round
                unsigned x0[4] = {0, 0, 0, 0}, x1[4] = {0, 0, 0, 0};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = c0 + u * 32 + lane;
                    if (c < nc) {
                        if (m0) x0[u] = __ldcg(h0 + c);
                        if (m1) x1[u] = __ldcg(h1 + c);
                    }
                }
                p0 += x0[0] + x0[1] + x0[2] + x0[3];
                p1 += x1[0] + x1[1] + x1[2] + x1[3];
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                p0 += __shfl_xor_sync(kFull, p0, d);
                p1 += __shfl_xor_sync(kFull, p1, d);
            }
            base0 = p0;
            base1 = p1;
        }
        for (int w = 0; w < warp; ++w)
            if (s_b[w] == wb) {
                base0 += s_cnt[w][0];
                base1 += s_cnt[w][1];
            }

        if (f) {
            const TierDev &tier = p.tier[tr];
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
    }

    // ---- publish this CTA's counter deltas ----------------------------------------------------
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
    if (t == 0 && threadIdx.x == 0) {
        p.dbg[4] = gtime();
        if (p.dbg[0] > p.dbg[2]) p.dbg[14] += 1000ull;      // the next batch's k_serve already started (must not happen)
    }
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
        // running sums over batches (ns): serve, gap, update, gap, evict(+c3), count
        p.dbg[8] += p.dbg[7] - p.dbg[0];
        p.dbg[9] += p.dbg[2] - p.dbg[7];
        p.dbg[10] += p.dbg[3] - p.dbg[2];
        p.dbg[11] += p.dbg[4] - p.dbg[3];
        p.dbg[12] += p.dbg[6] - p.dbg[4];
        p.dbg[13] += 1ull;
        atomicAdd(&p.g->batches, 1ull);
    }
}

}  // namespace evs
