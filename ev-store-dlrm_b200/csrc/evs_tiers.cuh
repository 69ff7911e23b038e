// Two- and three-tier lookups (C1 + C2 [+ C3]).  Filled in after the single-tier path.
#pragma once
#include "evs_host.h"
#include "evs_kernels.cuh"

namespace evs {

struct C3Dev {
    int dummy;
};

inline int launch_multi_tier(evs_handle h, const LookupArgs &a, cudaStream_t st) {
    (void)h; (void)a; (void)st;
    set_error("multi-tier lookup not built yet");
    return EVS_ERR_INVALID;
}
inline int c3_build(evs_handle h) { (void)h; return EVS_ERR_INVALID; }
inline void c3_stats(evs_handle h, uint64_t *size, uint64_t *cap) { (void)h; *size = 0; *cap = 0; }
inline int c3_dump(evs_handle h, int64_t *keys, uint32_t *alt, uint8_t *recency, int64_t *n) {
    (void)h; (void)keys; (void)alt; (void)recency; *n = 0; return EVS_OK;
}

}  // namespace evs
