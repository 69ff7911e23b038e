// Microbenchmark for the miss fetch (k_fetch_list): zero-copy row reads from host-pinned memory as a function of
// the ROWS KEPT IN FLIGHT and the row size.  A fixed number of lane groups (16 B per lane, one row per group and
// round) walk a list of random rows, exactly like k_fetch_list does; the grid is `ctas` x 256 threads, so
// rows in flight = ctas * 256 / lanes_per_row.  Prints rows/us and GB/s per (row bytes, rows in flight).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/zc_inflight_probe.cu -o tools/zc_inflight_probe
//   ./tools/zc_inflight_probe [n_rows_per_launch = 1300]
//
// Round 1 measured the optimum only for 64-byte rows (~512 rows = 32 KB in flight, profiles/r1_fetch_list_ab.md);
// this probe is what sizes the grid for 128- and 256-byte rows (Terabyte shape, d = 64).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__global__ void fetch_rows(const uint4 *__restrict__ host, const unsigned *__restrict__ rowid, int n, int chunks_per_row,
                           int gsize, uint4 *__restrict__ out) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int rpw = 32 / gsize, grp = lane / gsize, gl = lane - grp * gsize;
    for (int i0 = warp * rpw; i0 < n; i0 += n_warps * rpw) {
        const int i = i0 + grp;
        if (i < n && gl < chunks_per_row) {
            const unsigned r = __ldcg(rowid + i);
            const uint4 v = __ldg(host + static_cast<size_t>(r) * chunks_per_row + gl);
            out[static_cast<size_t>(i) * chunks_per_row + gl] = v;
        }
    }
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1300;
    const size_t bytes = 2ull << 30;                          // 2 GB of pinned rows
    uint4 *h = nullptr;
    if (cudaHostAlloc(&h, bytes, cudaHostAllocMapped) != cudaSuccess) { printf("cudaHostAlloc failed\n"); return 1; }
    for (size_t i = 0; i < bytes / 16; i += 256) h[i].x = static_cast<unsigned>(i);
    uint4 *hd = nullptr;
    cudaHostGetDevicePointer(&hd, h, 0);
    unsigned *d_ids = nullptr;
    uint4 *d_out = nullptr;
    cudaMalloc(&d_ids, static_cast<size_t>(n) * 4);
    cudaMalloc(&d_out, static_cast<size_t>(n) * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("%d rows per launch (one batch's misses); best of 7 launches, a fresh random row list per launch\n", n);
    printf("%9s %9s %6s %10s %9s %8s\n", "row_bytes", "in_flight", "ctas", "us/launch", "rows/us", "GB/s");
    uint64_t s = 88172645463325252ull;
    for (int row_bytes : {64, 128, 256}) {
        const int cpr = row_bytes / 16;
        int gsize = 1;
        while (gsize < cpr) gsize <<= 1;
        const size_t rows = bytes / row_bytes;
        for (int inflight : {64, 128, 256, 384, 512, 768, 1024, 2048, 4096}) {
            const int rows_per_cta = 256 / gsize;
            const int ctas = (inflight + rows_per_cta - 1) / rows_per_cta;
            float best = 1e9f;
            for (int rep = 0; rep < 7; ++rep) {
                std::vector<unsigned> ids(n);
                for (auto &x : ids) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = static_cast<unsigned>(s % rows); }
                cudaMemcpy(d_ids, ids.data(), static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice);
                cudaEventRecord(e0);
                fetch_rows<<<ctas, 256>>>(hd, d_ids, n, cpr, gsize, d_out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            printf("%9d %9d %6d %10.2f %9.1f %8.2f\n", row_bytes, ctas * rows_per_cta, ctas, best * 1e3f, n / (best * 1e3f),
                   static_cast<double>(n) * row_bytes / (best * 1e-3) / 1e9);
        }
    }
    if (cudaGetLastError() != cudaSuccess) { printf("CUDA error\n"); return 1; }
    return 0;
}
