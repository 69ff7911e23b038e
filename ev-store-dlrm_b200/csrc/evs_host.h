// Host-side objects behind the C-ABI (include/evstore_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/evstore_b200.h"
#include "evs_types.cuh"

namespace evs {

void set_error(const std::string &msg);

#define EVS_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            ::evs::set_error(std::string(#call) + " -> " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                             std::to_string(__LINE__) + ")");                                       \
            return EVS_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

// Capacities in entries, following the reference's constructors (see evs_api.cu:compute_caps).
struct Caps {
    long long c1 = 0, c2 = 0, c3 = 0;
};
int compute_caps(const evs_config &cfg, Caps &out);

struct Tier {
    TierDev dev{};
    int prec = 0;
    unsigned long long ub_used = 0;          // host upper bound of max_b(tail-head)
    std::vector<void *> allocs;              // device allocations to free
    std::vector<const unsigned char *> store_dev;   // per table, device-visible backing rows
};

struct C3Dev;                                // evs_tiers.cuh

}  // namespace evs

struct evs_handle_s {
    evs_config cfg{};
    std::vector<int64_t> rows;
    int n_tiers = 0;
    bool c3_active = false;
    evs::Caps caps;
    evs::Tier tier[EVS_MAX_TIERS];
    evs::C3Dev *c3 = nullptr;                // host copy of the device view
    std::vector<void *> c3_allocs;
    cudaStream_t stream = nullptr;
    evs::GlobalCtl *g = nullptr;             // device
    long long *d_rows = nullptr;
    uint8_t *d_agg = nullptr;                // [max_batch]
    // staging for the host-buffer path
    long long *d_idx = nullptr;
    float *d_out = nullptr;
    uint8_t *d_hit = nullptr;
    std::vector<void *> registered;          // host ranges we page-locked
    std::vector<void *> dev_allocs;
    const uint32_t **d_alt = nullptr;        // [n_tables] device-visible alt-key tables
    uint64_t batches = 0;
};
