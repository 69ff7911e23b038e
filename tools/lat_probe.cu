// Microbenchmark: dependent-load latency vs footprint, returning-atomic latency, zero-copy host
// read latency on the device at hand.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

__global__ void chase(const unsigned *next, unsigned start, int hops, unsigned long long *cycles, unsigned *sink) {
    unsigned i = start;
    for (int k = 0; k < 64; ++k) i = next[i];            // warm
    const long long t0 = clock64();
    for (int k = 0; k < hops; ++k) i = next[i];
    const long long t1 = clock64();
    *cycles = t1 - t0;
    *sink = i;
}
__global__ void chase_atomic(unsigned *buf, unsigned n_mask, int hops, unsigned long long *cycles, unsigned *sink) {
    unsigned i = 12345u & n_mask;
    const long long t0 = clock64();
    for (int k = 0; k < hops; ++k) { unsigned o = atomicAdd(&buf[i], 1u); i = (i * 1664525u + 1013904223u + o) & n_mask; }
    const long long t1 = clock64();
    *cycles = t1 - t0;
    *sink = i;
}
int main() {
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("sm clock attr %d kHz\n", clk);
    unsigned long long *d_cyc; unsigned *d_sink; cudaMalloc(&d_cyc, 8); cudaMalloc(&d_sink, 4);
    const int hops = 2000;
    for (size_t mb : {1, 32, 96, 256, 1024, 4096, 16384}) {
        const size_t n = mb * 1024 * 1024 / 128;          // one entry per 128 B line
        std::vector<unsigned> perm(n);
        for (size_t i = 0; i < n; ++i) perm[i] = (unsigned)i;
        std::mt19937_64 rng(1);
        std::shuffle(perm.begin(), perm.end(), rng);
        // next[] stored strided: entry i lives at word i*32
        unsigned *d; cudaMalloc(&d, n * 128);
        std::vector<unsigned> host(n * 32, 0);
        for (size_t i = 0; i < n; ++i) host[perm[i] * 32ull] = perm[(i + 1) % n] * 32u;
        cudaMemcpy(d, host.data(), n * 128, cudaMemcpyHostToDevice);
        unsigned long long c;
        chase<<<1, 1>>>(d, perm[0] * 32u, hops, d_cyc, d_sink);
        cudaDeviceSynchronize();
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("footprint %6zu MB: first pass %.0f cycles/hop", mb, (double)c / hops);
        chase<<<1, 1>>>(d, perm[0] * 32u, hops, d_cyc, d_sink);
        cudaDeviceSynchronize();
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf(", second pass (L2 resident chain) %.0f cycles/hop\n", (double)c / hops);
        cudaFree(d);
    }
    {   // returning atomics, random addresses in 64 MB and same address
        unsigned *d; cudaMalloc(&d, 64u << 20); cudaMemset(d, 0, 64u << 20);
        chase_atomic<<<1, 1>>>(d, (16u << 20) - 1, hops, d_cyc, d_sink); cudaDeviceSynchronize();
        unsigned long long c; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("atomicAdd(return) random 64MB: %.0f cycles/op\n", (double)c / hops);
        chase_atomic<<<1, 1>>>(d, 0, hops, d_cyc, d_sink); cudaDeviceSynchronize();
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("atomicAdd(return) same address: %.0f cycles/op\n", (double)c / hops);
        cudaFree(d);
    }
    {   // zero-copy host reads
        const size_t n = (256u << 20) / 128;
        unsigned *hbuf; cudaHostAlloc(&hbuf, n * 128, cudaHostAllocMapped);
        std::vector<unsigned> perm(n);
        for (size_t i = 0; i < n; ++i) perm[i] = (unsigned)i;
        std::mt19937_64 rng(2); std::shuffle(perm.begin(), perm.end(), rng);
        for (size_t i = 0; i < n; ++i) hbuf[perm[i] * 32ull] = perm[(i + 1) % n] * 32u;
        unsigned *dptr; cudaHostGetDevicePointer(&dptr, hbuf, 0);
        chase<<<1, 1>>>(dptr, perm[0] * 32u, 500, d_cyc, d_sink); cudaDeviceSynchronize();
        unsigned long long c; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("zero-copy host read: %.0f cycles/hop\n", (double)c / 500);
        cudaFreeHost(hbuf);
    }
    // kernel launch + empty kernel durations
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 1000; ++i) chase<<<1, 1>>>(d_sink, 0, 0, d_cyc, d_sink);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("back-to-back tiny kernels: %.2f us each\n", ms);
    return 0;
}
