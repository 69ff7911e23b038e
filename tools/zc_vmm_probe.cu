// Can the GPU map pinned host memory with pages larger than 4 KB?  zc_page_probe showed the zero-copy row rate falling
// from ~115 to ~55 rows / us once the random rows spread over more than ~512 MB of a cudaHostAlloc block, whatever the
// host page size behind it (THP made no difference): the GPU-side mapping of cudaHostAlloc / cudaHostRegister memory is
// made of 4 KB entries.  This probe repeats the sweep over host memory obtained in two other ways:
//   (a) the virtual-memory-management API: cuMemCreate(location = HOST_NUMA node 0, pinned) + cuMemMap + cuMemSetAccess
//       for the device and for the host; the allocation granularity it reports is the candidate GPU page size;
//   (b) cudaMallocManaged with preferred location = CPU and accessed-by = the device (a direct mapping, no migration).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/zc_vmm_probe.cu -o tools/zc_vmm_probe -lcuda
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
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

static void sweep(const char *what, const uint4 *hd, size_t bytes, int n, unsigned *d_ids, uint4 *d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (size_t region = 32ull << 20; region <= bytes; region *= 4) {
        for (int ctas : {16, 64}) {
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
            printf("%-34s n %6d region %6zu MB  in_flight %5d  %8.2f us  %7.1f rows/us\n", what, n, region >> 20, ctas * 64, best * 1e3f,
                   n / (best * 1e3f));
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  (%s: %s)\n", what, cudaGetErrorString(e));
    fflush(stdout);
}

#define CU(x)                                                                  \
    do {                                                                       \
        CUresult r_ = (x);                                                     \
        if (r_ != CUDA_SUCCESS) {                                              \
            const char *m_ = nullptr;                                          \
            cuGetErrorString(r_, &m_);                                         \
            printf("%s -> %s\n", #x, m_ ? m_ : "?");                           \
            ok = false;                                                        \
        }                                                                      \
    } while (0)

int main(int argc, char **argv) {
    const size_t bytes = 2ull << 30;
    cudaFree(0);
    unsigned *d_ids = nullptr;
    uint4 *d_out = nullptr;
    const int n_max = 16384;
    cudaMalloc(&d_ids, static_cast<size_t>(n_max) * 4);
    cudaMalloc(&d_out, static_cast<size_t>(n_max) * 64);
    const int ns[2] = {1300, 16384};
    // reference point
    {
        uint4 *h = nullptr, *hd = nullptr;
        if (cudaHostAlloc(&h, bytes, cudaHostAllocMapped) == cudaSuccess) {
            memset(h, 1, bytes);
            cudaHostGetDevicePointer(&hd, h, 0);
            for (int n : ns) sweep("cudaHostAlloc", hd, bytes, n, d_ids, d_out);
            cudaFreeHost(h);
        }
    }
    // (a) VMM host allocation
    {
        bool ok = true;
        CUmemAllocationProp prop{};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_HOST_NUMA;
        prop.location.id = 0;
        size_t gmin = 0, grec = 0;
        CU(cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
        CU(cuMemGetAllocationGranularity(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
        printf("VMM HOST_NUMA granularity: minimum %zu, recommended %zu\n", gmin, grec);
        CUmemGenericAllocationHandle hnd{};
        if (ok) CU(cuMemCreate(&hnd, bytes, &prop, 0));
        CUdeviceptr va = 0;
        if (ok) CU(cuMemAddressReserve(&va, bytes, grec ? grec : (2ull << 20), 0, 0));
        if (ok) CU(cuMemMap(va, bytes, 0, hnd, 0));
        if (ok) {
            CUmemAccessDesc acc[2]{};
            acc[0].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            acc[0].location.id = 0;
            acc[0].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            acc[1].location.type = CU_MEM_LOCATION_TYPE_HOST_NUMA;
            acc[1].location.id = 0;
            acc[1].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            CU(cuMemSetAccess(va, bytes, acc, 2));
        }
        if (ok) {
            memset(reinterpret_cast<void *>(va), 1, bytes);          // the host writes through the same address
            for (int n : ns) sweep("cuMemCreate HOST_NUMA", reinterpret_cast<const uint4 *>(va), bytes, n, d_ids, d_out);
            cuMemUnmap(va, bytes);
            cuMemAddressFree(va, bytes);
            cuMemRelease(hnd);
        } else {
            printf("VMM host allocation: not available\n");
        }
    }
    // (b) managed memory pinned to the CPU, mapped into the GPU
    {
        uint4 *m = nullptr;
        if (cudaMallocManaged(&m, bytes) == cudaSuccess) {
            cudaMemLocation cpu{};
            cpu.type = cudaMemLocationTypeHost;
            cudaMemLocation gpu{};
            gpu.type = cudaMemLocationTypeDevice;
            gpu.id = 0;
            cudaError_t e1 = cudaMemAdvise_v2(m, bytes, cudaMemAdviseSetPreferredLocation, cpu);
            cudaError_t e2 = cudaMemAdvise_v2(m, bytes, cudaMemAdviseSetAccessedBy, gpu);
            printf("managed: advise preferred CPU %s, accessed-by GPU %s\n", cudaGetErrorString(e1), cudaGetErrorString(e2));
            memset(m, 1, bytes);
            cudaDeviceSynchronize();
            for (int n : ns) sweep("managed, preferred CPU", m, bytes, n, d_ids, d_out);
            cudaFree(m);
        }
    }
    // the same rows from HBM, for scale
    {
        uint4 *d = nullptr;
        if (cudaMalloc(&d, bytes) == cudaSuccess) {
            cudaMemset(d, 1, bytes);
            for (int n : ns) sweep("HBM (cudaMalloc)", d, bytes, n, d_ids, d_out);
            cudaFree(d);
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
