// Microbenchmark: random 64-byte row reads from host-pinned (zero-copy) memory.
// Each group of 4 lanes reads one 64 B row with 16 B loads.  Reports rows/s vs the number of rows.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__global__ void rows64(const uint4 *host, const unsigned *rowid, int n, uint4 *out) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, l = threadIdx.x & 3;
    if (g >= n) return;
    const uint4 v = __ldg(host + (size_t)rowid[g] * 4 + l);
    out[(size_t)g * 4 + l] = v;
}
__global__ void rows64_w(const uint4 *host, const unsigned *rowid, int n, uint4 *out, int per_warp) {
    // one warp handles per_warp rows one after another (8 rows at a time), like a serial consumer
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    for (int k = 0; k < per_warp; k += 8) {
        const int g = w * per_warp + k + (lane >> 2);
        if (g < n) out[(size_t)g * 4 + (lane & 3)] = __ldg(host + (size_t)rowid[g] * 4 + (lane & 3));
    }
}
int main() {
    const size_t rows = 32u << 20;                   // 2 GB table of 64 B rows
    uint4 *h; cudaHostAlloc(&h, rows * 64, cudaHostAllocMapped);
    for (size_t i = 0; i < rows * 4; i += 1024) h[i].x = (unsigned)i;
    uint4 *hd; cudaHostGetDevicePointer(&hd, h, 0);
    const int maxn = 1 << 20;
    std::vector<unsigned> ids(maxn);
    uint64_t s = 88172645463325252ull;
    for (auto &x : ids) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (unsigned)(s % rows); }
    unsigned *d_ids; cudaMalloc(&d_ids, maxn * 4); cudaMemcpy(d_ids, ids.data(), maxn * 4, cudaMemcpyHostToDevice);
    uint4 *d_out; cudaMalloc(&d_out, (size_t)maxn * 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int n : {64, 256, 1024, 1300, 4096, 16384, 65536, 262144, 1048576}) {
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            rows64<<<(n * 4 + 255) / 256, 256>>>(hd, d_ids, n, d_out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("rows %8d: %8.2f us  -> %7.1f M rows/s, %6.2f GB/s\n", n, best * 1e3, n / best / 1e3, n * 64.0 / best / 1e6);
    }
    for (int pw : {8, 32}) {
        const int n = 16384;
        cudaEventRecord(e0);
        rows64_w<<<(n / pw * 32 + 255) / 256, 256>>>(hd, d_ids, n, d_out, pw);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("serial-warp per_warp=%d rows %d: %.2f us\n", pw, n, ms * 1e3);
    }
    // compare: cudaMemcpyAsync H2D of a contiguous block
    for (size_t bytes : {(size_t)83200, (size_t)1 << 20, (size_t)64 << 20}) {
        cudaEventRecord(e0);
        cudaMemcpyAsync(d_out, h, bytes, cudaMemcpyHostToDevice);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("memcpy H2D %zu B: %.2f us (%.2f GB/s)\n", bytes, ms * 1e3, bytes / ms / 1e6);
    }
    for (size_t bytes : {(size_t)3461120, (size_t)64 << 20}) {
        cudaEventRecord(e0);
        cudaMemcpyAsync(h, d_out, bytes, cudaMemcpyDeviceToHost);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("memcpy D2H %zu B: %.2f us (%.2f GB/s)\n", bytes, ms * 1e3, bytes / ms / 1e6);
    }
    return 0;
}
