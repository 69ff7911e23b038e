// Device-side data layout of one EvLFU cache tier (C1 or C2) in HBM.
//
// What the reference keeps in std::unordered_map<string, Cache_data> vals_C1 plus
// vector<unordered_set<string>> lists_C1 (evlfu_32.hpp:47-49) is here:
//
//   slots[]     open-addressing index, 16 B per slot {key, rowword, pass}
//               rowword = (bucket+1) << 27 | slab row;  pass = number of resident keys whose
//               probe path crosses this slot (lets a probe stop without tombstones)
//   slab[]      the rows themselves, row_stride bytes each (precision dependent)
//   row_meta[]  per slab row: (bucket+1) << 56 | position of its live record in the bucket ring
//   ring[b][]   per agg_hit bucket b a FIFO log of slab rows (lists_C1[b]); a record is live
//               iff row_meta[row] still points at it, so promotion/eviction never edits a log
//   free_rows[] stack of unused slab rows
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace evs {

constexpr int kMaxTables = 32;
constexpr int kMaxBuckets = 32;              // agg_hit 0..n_tables_total (<= 31)
constexpr int kSamplesPerCta = 8;            // one warp per sample
constexpr int kLookupThreads = kSamplesPerCta * 32;
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned kRowBits = 27;
constexpr unsigned kRowMask = (1u << kRowBits) - 1u;
constexpr int kKeyShift = 40;
constexpr unsigned kNoRow = 0xFFFFFFFFu;

// flags[p]: 0 = nothing to do; low 6 bits = bucket+1 of the append this position asks for;
// bit 7 = the position missed and needs a row
constexpr uint8_t kFlagMiss = 0x80;

struct __align__(16) Slot {
    unsigned long long key;
    unsigned int rowword;
    unsigned int pass;
};

struct TierCtl {
    unsigned long long head[kMaxBuckets];
    unsigned long long tail[kMaxBuckets];
    unsigned long long tail_prev[kMaxBuckets];   // tail before this batch's appends (k_hist_scan)
    unsigned int count[kMaxBuckets];       // live entries per bucket
    unsigned int free_top;
    unsigned int n_perfect;                // n_perfect_item_C1 (evlfu_32.hpp:51)
    // per batch (reset by the evict kernel)
    unsigned int miss_count;
    unsigned int n_new;
    unsigned long long prot;               // max over new keys of (bucket+1) << 32 | position
    unsigned int any_perfect;
    unsigned int n_evicted_last;
    unsigned int n_flushed_last;
    unsigned int error;
    unsigned int full_at_start;            // size >= cap when the batch began (two-tier routing)
    // cumulative
    unsigned long long stat_inserts, stat_evictions, stat_flushed;
};

struct TierDev {
    Slot *slots;
    unsigned int hash_mask;
    unsigned char *slab;
    unsigned int row_stride;               // bytes, multiple of 16
    unsigned int row_bytes;                // dim*prec/8 (bytes of one backing-store row)
    int prec;                              // 32/16/8/4
    unsigned long long *row_meta;
    unsigned long long *row_key;
    unsigned int *row_slot;
    unsigned int *free_rows;
    unsigned int *ring;                    // [n_buckets][ring_cap]
    unsigned int ring_cap;                 // power of two
    unsigned int cap;                      // policy capacity (entries)
    unsigned int rows_total;               // cap + spare rows for one batch of inserts
    unsigned int max_perfect;              // int(cap * 0.95)
    unsigned int flush_n;                  // int(0.3 * cap) + 1
    int n_buckets;                         // n_tables_total + 1
    TierCtl *ctl;
    const unsigned char *const *store;     // [n_tables] device-visible backing rows at this precision
    // per-batch scratch
    uint8_t *flags;                        // [N]
    unsigned int *pos_slot;                // [N]
    unsigned int *miss_list;               // [N]
    unsigned int *hist;                    // [n_chunks][kMaxBuckets]
    unsigned long long *evicted;           // [N] keys evicted by the last batch (rank order)
    unsigned long long *flushed;           // [flush_n] keys flushed by the last batch (or null)
};

struct GlobalCtl {
    unsigned long long lookups, samples, hits[2], c3_hits, approx_subst, misses, perfect_hits, batches;
    unsigned int error;
    unsigned int pad;
};

struct LookupArgs {
    const long long *idx;                  // [T][B]
    const long long *rows;                 // [T] cardinalities (device)
    float *out;
    long long out_stride;                  // floats between samples
    uint8_t *hit;                          // [B][T] or null
    const uint8_t *agg_in;                 // [B] or null
    uint8_t *agg_out;                      // [B] (scratch; always written)
    int B, T, D;
    int table_base;
    int n_perfect_agg;                     // agg value that counts as a perfect hit (n_tables_total)
    int approx_thres;
    GlobalCtl *g;
};

__host__ __device__ inline unsigned long long pack_meta(int bucket, unsigned long long q) {
    return (static_cast<unsigned long long>(bucket + 1) << 56) | (q & 0x00FFFFFFFFFFFFFFull);
}
__host__ __device__ inline int meta_bucket(unsigned long long m) { return static_cast<int>(m >> 56) - 1; }
__host__ __device__ inline unsigned long long meta_pos(unsigned long long m) { return m & 0x00FFFFFFFFFFFFFFull; }

__host__ __device__ inline unsigned long long make_key(int table, long long row) {
    return (static_cast<unsigned long long>(table) << kKeyShift) | static_cast<unsigned long long>(row);
}

// 64-bit finaliser (splitmix64); the slot is the low bits
__host__ __device__ inline unsigned int hash_key(unsigned long long k, unsigned int mask) {
    k ^= k >> 30; k *= 0xbf58476d1ce4e5b9ull;
    k ^= k >> 27; k *= 0x94d049bb133111ebull;
    k ^= k >> 31;
    return static_cast<unsigned int>(k) & mask;
}

}  // namespace evs
