// C3, the approximate-embedding map (mixed_precs_caching/aprx_embedding.cpp): key -> alternative
// key, a FIFO queue of keys (lists_C3, may hold stale duplicates) and second-chance eviction
// (recency_aware_eviction, :360-388).  The reference fills it asynchronously from worker threads
// in groups of IO_JOB_Q_SIZE = 50 evicted keys (:125-173, 304-329); here the keys a batch evicted
// from C2 and then C1 (normal evictions only, evlfu_8.cpp:284-287) form ONE group that is inserted
// synchronously at the end of the batch:
//     n_erase = size + |group| - cap;  evict_one_key() x n_erase;  push + vals[key] = {alt, false}.
#pragma once
#include "evs_kernels.cuh"

namespace evs {


__device__ __forceinline__ void c3_erase(const C3Dev &c, unsigned slot, unsigned long long key) {
    const unsigned mask = c.hash_mask;
    for (unsigned j = hash_key(key, mask); j != slot; j = (j + 1) & mask) atomicAdd(&c.slots[j].kw, ~kPassOne + 1ull);
    atomicOr(&c.slots[slot].kw, kEmptyKey);
}

__device__ __forceinline__ unsigned c3_claim(const C3Dev &c, unsigned long long key, bool &claimed) {
    const unsigned mask = c.hash_mask;
    const unsigned home = hash_key(key, mask);
    unsigned i = home;
    claimed = false;
    while (true) {
        const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&c.slots[i].kw);
        const unsigned long long k = cur & kKeyMask;
        if (k == key) break;
        if (k == kEmptyKey) {
            const unsigned long long old = atomicCAS(&c.slots[i].kw, cur, (cur & ~kKeyMask) | key);
            if (old == cur) {
                claimed = true;
                break;
            }
            continue;
        }
        i = (i + 1) & mask;
    }
    if (claimed)
        for (unsigned j = home; j != i; j = (j + 1) & mask) atomicAdd(&c.slots[j].kw, kPassOne);
    return i;
}

// Block-wide exclusive scan of one value per thread (kC3Threads threads); returns the exclusive
// prefix and writes the total to *total (uniform).
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned *s_w, unsigned *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned n = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += n;
    }
    __syncthreads();                            // s_w may still be read from a previous call
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const unsigned x = (lane < nwarp) ? s_w[lane] : 0u;
        unsigned wi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned n = __shfl_up_sync(kFull, wi, d);
            if (lane >= d) wi += n;
        }
        s_w[lane] = wi - x;
        if (lane == 31) s_w[32] = wi;
    }
    __syncthreads();
    *total = s_w[32];
    return s_w[warp] + incl - v;
}

// Whole CTA (any multiple of 32 threads up to 1024); called by the last CTA of k_update after the
// evictions, so the victims' key lists are complete.
__device__ void c3_update(const Params &p) {
    __shared__ unsigned s_w[33];
    __shared__ unsigned s_dup;
    __shared__ unsigned long long s_head, s_tail;
    __shared__ unsigned s_evicted, s_size;
    const C3Dev &c = p.c3;
    volatile C3Ctl *ctl = c.ctl;
    __syncthreads();
    const unsigned n1 = reinterpret_cast<volatile TierCtl *>(p.tier[1].ctl)->n_evicted_last;
    const unsigned n0 = reinterpret_cast<volatile TierCtl *>(p.tier[0].ctl)->n_evicted_last;
    const unsigned n = n1 + n0;
    if (n == 0) return;
    if (threadIdx.x == 0) {
        s_head = ctl->head;
        s_tail = ctl->tail;
        s_size = ctl->size;
        s_evicted = 0;
    }
    __syncthreads();
    const unsigned size0 = s_size;
    unsigned n_erase = (size0 + n > c.cap) ? size0 + n - c.cap : 0u;
    if (n_erase > size0) n_erase = size0;

    // ---- second-chance eviction over windows of the FIFO -----------------------------------
    unsigned done = 0;
    while (done < n_erase) {
        const unsigned long long head = s_head, tail = s_tail;
        if (head >= tail) {
            if (threadIdx.x == 0) ctl->error = 6u;
            break;
        }
        const unsigned long long q = head + threadIdx.x;
        const bool inw = q < tail;
        unsigned long long key = 0;
        unsigned slot = 0, alt = 0;
        bool present = false;
        if (inw) {
            key = __ldcg(c.ring + (q & (c.ring_cap - 1)));
            present = c3_find(c, key, slot, alt);
        }
        if (threadIdx.x == 0) s_dup = 0;
        if (present) atomicMin(&c.scratch[slot], threadIdx.x);
        __syncthreads();
        if (present && __ldcg(c.scratch + slot) != threadIdx.x) s_dup = 1u;
        __syncthreads();
        if (present) c.scratch[slot] = 0xFFFFFFFFu;
        const bool dup = s_dup != 0u;
        const unsigned budget = n_erase - done;
        if (!dup) {
            const bool flag = present && (__ldcg(&c.slots[slot].flag) & 1u);
            const bool cand = present && !flag;
            unsigned tot_c;
            const unsigned ci = block_excl_scan(cand ? 1u : 0u, s_w, &tot_c);
            // processed = every record up to and including the one that makes the budget
            const bool processed = inw && (ci < budget);            // records after the budget-th victim stay
            const bool evict = cand && processed;
            const bool again = present && flag && processed;
            unsigned tot_a;
            const unsigned ai = block_excl_scan(again ? 1u : 0u, s_w, &tot_a);
            if (evict) c3_erase(c, slot, key);
            if (again) {
                c.slots[slot].flag = 0u;
                c.ring[(tail + ai) & (c.ring_cap - 1)] = key;
            }
            unsigned tot_p;
            block_excl_scan(processed ? 1u : 0u, s_w, &tot_p);
            const unsigned ev = min(tot_c, budget);
            if (threadIdx.x == 0) {
                s_head = head + tot_p;
                s_tail = tail + tot_a;
                s_evicted += ev;
                s_size -= ev;
            }
            done += ev;
            __syncthreads();
        } else {
            // same key twice in one window: the second record's fate depends on the first -> one
            // thread replays the window in order (rare)
            if (threadIdx.x == 0) {
                unsigned long long h = head, tl = tail;
                unsigned ev = 0;
                const unsigned long long wend = (head + blockDim.x < tail) ? head + blockDim.x : tail;
                while (h < wend && ev < budget) {
                    const unsigned long long k = __ldcg(c.ring + (h & (c.ring_cap - 1)));
                    unsigned sl, al;
                    if (c3_find(c, k, sl, al)) {
                        if (__ldcg(&c.slots[sl].flag) & 1u) {
                            c.slots[sl].flag = 0u;
                            c.ring[tl & (c.ring_cap - 1)] = k;
                            ++tl;
                        } else {
                            c3_erase(c, sl, k);
                            ++ev;
                        }
                    }
                    ++h;
                }
                s_head = h;
                s_tail = tl;
                s_evicted += ev;
                s_size -= ev;
                s_w[32] = ev;
            }
            __syncthreads();
            done += s_w[32];
            __syncthreads();
        }
    }
    __syncthreads();

    // ---- insert the group: C2's victims first, then C1's (evlfu_8.cpp:617-620, 654-658) ------
    const unsigned long long tail = s_tail;
    if (tail + n - s_head > c.ring_cap) {
        if (threadIdx.x == 0) ctl->error = 6u;
        return;
    }
    unsigned my_new = 0;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long key = (i < n1) ? __ldcg(p.tier[1].evicted + i) : __ldcg(p.tier[0].evicted + (i - n1));
        c.ring[(tail + i) & (c.ring_cap - 1)] = key;
        const int tbl = p.loc[static_cast<int>(key >> kKeyShift) & (kMaxTables - 1)];
        const unsigned long long row = key & ((1ull << kKeyShift) - 1ull);
        const unsigned alt = __ldg(c.alt[tbl] + row);
        // upsert: a key that is already mapped keeps its slot (claiming the first free slot of its probe
        // path would map it twice when an erased slot lies before its own)
        bool claimed = false;
        unsigned slot, old_alt;
        if (!c3_find(c, key, slot, old_alt)) slot = c3_claim(c, key, claimed);
        c.slots[slot].alt = alt;
        c.slots[slot].flag = 0u;
        my_new += claimed ? 1u : 0u;
    }
    unsigned tot_new;
    block_excl_scan(my_new, s_w, &tot_new);
    if (threadIdx.x == 0) {
        ctl->head = s_head;
        ctl->tail = tail + n;
        ctl->size = s_size + tot_new;
        ctl->stat_inserts += n;
        ctl->stat_evictions += s_evicted;
    }
}

__global__ void k_init_c3(C3Dev c) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t j = i; j <= c.hash_mask; j += stride) {
        c.slots[j].kw = kEmptyKey;
        c.slots[j].alt = 0;
        c.slots[j].flag = 0;
        c.scratch[j] = 0xFFFFFFFFu;
    }
}

}  // namespace evs
