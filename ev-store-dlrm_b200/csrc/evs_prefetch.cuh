// k_prefetch: look-ahead for the NEXT batch (evs_prefetch), on its own stream while the kernels of the current
// batch run.
//
// The batches of a handle are strictly ordered -- batch n+1 must see the evictions of batch n -- so their policy
// kernels cannot overlap.  What CAN run ahead is everything that only moves data:
//   * the next batch's keys are probed against the index as it is NOW (racing with the current batch's inserts and
//     evictions: the answer is a guess, never a decision);
//   * a key that looks absent has its backing-store row read over PCIe into a staging row in HBM, addressed by the
//     key's position in the next batch and tagged with the announcement's generation once the row is complete.  The fetch
//     role of the next batch's k_evict copies a staged row instead of reading the link; a key that turned out to
//     be resident after all leaves a stale staging row nobody reads, and a key evicted in between is simply not
//     staged and is fetched the usual way.  Staged bytes are the backing store's bytes, so results do not change;
//   * a key that looks resident has its slab row prefetched into L2, and the index batch itself and the probed
//     slot sectors are in L2 afterwards, so the three dependent accesses of the next k_serve (index -> slot -> row)
//     hit L2 instead of HBM.
// The miss path is the longest part of a step (the link gives 55-110 row reads / us); with the next batch announced
// it runs entirely under the current batch's kernels.
#pragma once
#include "evs_kernels.cuh"

namespace evs {

constexpr int kPfThreads = 256;
constexpr int kPfListCap = 2560;                 // positions of one tile (<= 2048 + a sample group's worth)

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// samples per tile: ~2048 positions, a multiple of 64 samples
__host__ __device__ inline int pf_tile_samples(int T) {
    const int m = 2048 / (64 * T);
    return 64 * (m < 1 ? 1 : m);
}

template <int P0, int P1>
__global__ void __launch_bounds__(kPfThreads) k_prefetch(const __grid_constant__ Params p, const __grid_constant__ PrefetchArgs a) {
    __shared__ unsigned s_list[kPfListCap];
    __shared__ unsigned s_n;
    const int T = p.T, B = a.B;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TierDev &t0 = p.tier[0];
    const TierDev &t1 = p.tier[1];
    const unsigned par = a.seq & 1u;
    unsigned *tags = p.pf_tag + static_cast<size_t>(par) * p.n_max;
    unsigned char *rows = p.pf_rows + static_cast<size_t>(par) * p.n_max * p.stage_stride;
    const int full = (P1 != 0) ? static_cast<int>(*reinterpret_cast<volatile unsigned *>(&t0.ctl->full_at_start)) : 0;
    // The staging rows of this parity belong to batch seq - 2 until its miss-fetch role has finished.  The wait is a
    // device-side spin on a word that role sets, not a stream dependency: a stream wait would release this kernel at the
    // very moment the next batch's k_serve becomes runnable, and the two launches then queue behind each other.
    if (a.seq > 2u) {
        __shared__ int s_go;
        if (threadIdx.x == 0) {
            const volatile unsigned *f = &p.g->fetch_done_seq;
            int go = 1;
            if (static_cast<int>(*f - (a.seq - 2u)) < 0) {
                const unsigned long long t0 = gtime();
                while (static_cast<int>(*f - (a.seq - 2u)) < 0) {
                    if (gtime() - t0 > 20000000ull) {           // 20 ms: give the look-ahead up, the batch fetches for itself
                        go = 0;
                        break;
                    }
                    __nanosleep(200);
                }
            }
            s_go = go;
        }
        __syncthreads();
        if (!s_go) return;
        __threadfence();
    }
    const int S = pf_tile_samples(T);
    const int n_tiles = (B + S - 1) / S;
    const int gsize = fetch_gsize(p, P1 == 0 ? 1 : 2);
    const int rpw = 32 / gsize, grp = lane / gsize, gl = lane - grp * gsize;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const int s_base = tile * S;
        const int ns = min(S, B - s_base);
        const int n_pos = ns * T;
        // ---- probe: element e = (table t, sample sl) with the sample fastest, so the index reads are coalesced ----
        constexpr int U = 4;
        for (int e0 = threadIdx.x; e0 < n_pos; e0 += kPfThreads * U) {
            unsigned long long key[U];
            unsigned i0[U], i1[U];
            uint4 v0[U][2], v1[U][2];
            int tt[U], ss[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * kPfThreads;
                ok[u] = e < n_pos;
                tt[u] = ok[u] ? e / ns : 0;
                ss[u] = s_base + (e - tt[u] * ns);
                long long r = ok[u] ? __ldg(a.idx + static_cast<size_t>(tt[u]) * B + ss[u]) : 0;
                if (r < 0) ok[u] = false;                   // no key at this position (a slice of ragged bags), or an index error
                if (r < 0 || r >= __ldg(p.rows + tt[u])) r = 0;
                key[u] = make_key(p.tid[tt[u]], r);
                i0[u] = hash_key(key[u], t0.hash_mask);
                if (ok[u]) {                                 // the first two slots of the probe path (one 32-byte sector)
                    v0[u][0] = __ldcg(reinterpret_cast<const uint4 *>(t0.slots + i0[u]));
                    v0[u][1] = __ldcg(reinterpret_cast<const uint4 *>(t0.slots + ((i0[u] + 1) & t0.hash_mask)));
                }
                if (P1 != 0) {
                    i1[u] = hash_key(key[u], t1.hash_mask);
                    if (ok[u]) {
                        v1[u][0] = __ldcg(reinterpret_cast<const uint4 *>(t1.slots + i1[u]));
                        v1[u][1] = __ldcg(reinterpret_cast<const uint4 *>(t1.slots + ((i1[u] + 1) & t1.hash_mask)));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!ok[u]) continue;
                unsigned slot = 0;
                unsigned long long meta = 0;
                int r0 = probe_check<2>(v0[u], t0.hash_mask, key[u], i0[u], slot, meta);
                bool h0 = (r0 >= 0) ? (r0 == 1) : probe_rest(t0.slots, t0.hash_mask, key[u], i0[u] + 2, slot, meta);
                if (h0) {
                    const unsigned char *row = t0.slab + static_cast<size_t>(slot) * t0.row_stride;
                    for (unsigned o = 0; o < t0.row_stride; o += 128u) prefetch_l2(row + o);
                    continue;
                }
                if (P1 != 0) {
                    unsigned slot1 = 0;
                    int r1 = probe_check<2>(v1[u], t1.hash_mask, key[u], i1[u], slot1, meta);
                    bool h1 = (r1 >= 0) ? (r1 == 1) : probe_rest(t1.slots, t1.hash_mask, key[u], i1[u] + 2, slot1, meta);
                    if (h1) {
                        const unsigned char *row = t1.slab + static_cast<size_t>(slot1) * t1.row_stride;
                        for (unsigned o = 0; o < t1.row_stride; o += 128u) prefetch_l2(row + o);
                        continue;
                    }
                }
                // probable miss.  Which tier will fetch it is decided by the batch itself (evlfu_8.cpp:573-602); the
                // guess follows the common case: C1 while it is not full, then odd tables C1, even tables C2.
                const unsigned tr = (P1 != 0 && full && !(p.tid[tt[u]] & 1)) ? 1u : 0u;
                const unsigned k = atomicAdd(&s_n, 1u);
                if (k < kPfListCap) s_list[k] = static_cast<unsigned>(ss[u] * T + tt[u]) | (tr << 31);
            }
        }
        __syncthreads();
        // ---- fetch: a group of lanes per row, 16 bytes per lane --------------------------------------------------
        const unsigned n = (a.mode == 1) ? 0u : min(s_n, static_cast<unsigned>(kPfListCap));
        for (unsigned k0 = warp * rpw; k0 < n; k0 += (kPfThreads / 32) * rpw) {
            const unsigned k = k0 + grp;
            const bool on = k < n;
            unsigned e = on ? s_list[k] : 0u;
            const unsigned tr = (P1 == 0) ? 0u : (e >> 31);
            const int pos = static_cast<int>(e & 0x7FFFFFFFu);
            if (on) {
                const TierDev &tier = tr ? t1 : t0;
                const int s = pos / T, t = pos - s * T;
                long long r = __ldg(a.idx + static_cast<size_t>(t) * B + s);
                if (r < 0 || r >= __ldg(p.rows + t)) r = 0;
                const unsigned char *src = tier.store[t] + static_cast<size_t>(r) * tier.row_bytes;
                unsigned char *dst = rows + static_cast<size_t>(pos) * p.stage_stride;
                const int cpr = static_cast<int>(tier.row_stride >> 4);
                for (int c = gl; c < cpr; c += gsize) *reinterpret_cast<uint4 *>(dst + (c << 4)) = ldg16(src + (c << 4));
            }
            __syncwarp();                                   // the group's stores are ordered before the tag
            if (on && gl == 0) {
                __threadfence();
                *reinterpret_cast<volatile unsigned *>(tags + pos) = (a.gen << 1) | tr;
            }
        }
        __syncthreads();
    }
}

}  // namespace evs
