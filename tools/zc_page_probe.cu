// Is the zero-copy miss fetch bound by address translation?  zc_inflight_probe measured ~65 row reads / us whatever the
// row size (64..256 B) -- a per-row cost, not a byte or sector cost.  Every miss of a Zipf trace lands on a different 4 KB
// page of a multi-GB pinned table, so the suspect is the GPU MMU's walk for system-memory pages.  This probe reads the
// same number of random 64 B rows
//   (a) from regions of growing size inside one cudaHostAlloc block (2 MB .. 4 GB): if the rate falls with the region,
//       translation reach is the limit;
//   (b) from a block backed by transparent huge pages (aligned_alloc + madvise(MADV_HUGEPAGE) + cudaHostRegister) and,
//       if the box has them, explicit 2 MB pages (mmap MAP_HUGETLB).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/zc_page_probe.cu -o tools/zc_page_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <vector>
#include <cuda_runtime.h>

__global__ void fetch_rows(const uint4 *__restrict__ host, const unsigned *__restrict__ rowid, int n, uint4 *__restrict__ out) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int grp = lane >> 2, gl = lane & 3;
    for (int i0 = warp * 8; i0 < n; i0 += n_warps * 8) {
        const int i = i0 + grp;
        if (i < n) {
            const unsigned r = __ldcg(rowid + i);
            out[static_cast<size_t>(i) * 4 + gl] = __ldg(host + static_cast<size_t>(r) * 4 + gl);
        }
    }
}

static uint64_t s = 88172645463325252ull;
static unsigned rnd(uint64_t m) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return static_cast<unsigned>(s % m); }

static void sweep(const char *what, uint4 *hd, size_t bytes, int n, unsigned *d_ids, uint4 *d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (size_t region = 2ull << 20; region <= bytes; region *= 4) {
        for (int ctas : {8, 16, 64}) {
            float best = 1e9f;
            for (int rep = 0; rep < 6; ++rep) {
                std::vector<unsigned> ids(n);
                for (auto &x : ids) x = rnd(region / 64);
                cudaMemcpy(d_ids, ids.data(), static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice);
                cudaEventRecord(e0);
                fetch_rows<<<ctas, 256>>>(hd, d_ids, n, d_out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            printf("%-28s region %6zu MB  in_flight %5d  %8.2f us  %7.1f rows/us\n", what, region >> 20, ctas * 64, best * 1e3f,
                   n / (best * 1e3f));
        }
    }
    fflush(stdout);
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1300;
    const size_t bytes = 2ull << 30;
    unsigned *d_ids = nullptr;
    uint4 *d_out = nullptr;
    cudaMalloc(&d_ids, static_cast<size_t>(n) * 4);
    cudaMalloc(&d_out, static_cast<size_t>(n) * 64);
    {
        uint4 *h = nullptr, *hd = nullptr;
        if (cudaHostAlloc(&h, bytes, cudaHostAllocMapped) == cudaSuccess) {
            memset(h, 1, bytes);
            cudaHostGetDevicePointer(&hd, h, 0);
            sweep("cudaHostAlloc", hd, bytes, n, d_ids, d_out);
            cudaFreeHost(h);
        }
    }
    {
        void *p = aligned_alloc(2ull << 20, bytes);
        if (p) {
            int rc = madvise(p, bytes, MADV_HUGEPAGE);
            memset(p, 1, bytes);
            cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
            printf("THP: madvise rc %d, cudaHostRegister %s\n", rc, cudaGetErrorString(e));
            FILE *f = fopen("/proc/meminfo", "r");
            char ln[256];
            while (f && fgets(ln, sizeof ln, f))
                if (strstr(ln, "AnonHugePages") || strstr(ln, "HugePages_Total") || strstr(ln, "Hugepagesize")) printf("  %s", ln);
            if (f) fclose(f);
            if (e == cudaSuccess) {
                uint4 *hd = nullptr;
                cudaHostGetDevicePointer(&hd, p, 0);
                sweep("aligned_alloc+THP+register", hd, bytes, n, d_ids, d_out);
                cudaHostUnregister(p);
            }
            free(p);
        }
    }
    {
        void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_HUGETLB, -1, 0);
        if (p == MAP_FAILED) {
            printf("MAP_HUGETLB: not available on this box\n");
        } else {
            memset(p, 1, bytes);
            cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
            printf("MAP_HUGETLB: cudaHostRegister %s\n", cudaGetErrorString(e));
            if (e == cudaSuccess) {
                uint4 *hd = nullptr;
                cudaHostGetDevicePointer(&hd, p, 0);
                sweep("mmap MAP_HUGETLB+register", hd, bytes, n, d_ids, d_out);
                cudaHostUnregister(p);
            }
            munmap(p, bytes);
        }
    }
    // the same rows from HBM, for scale
    {
        uint4 *d = nullptr;
        if (cudaMalloc(&d, bytes) == cudaSuccess) {
            cudaMemset(d, 1, bytes);
            sweep("HBM (cudaMalloc)", d, bytes, n, d_ids, d_out);
            cudaFree(d);
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
