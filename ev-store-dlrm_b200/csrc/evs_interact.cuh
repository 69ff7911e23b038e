// Pairwise-dot feature interaction (dlrm_s_pytorch_C1_C2_C3.py:625-658, op "dot",
// arch_interaction_itself = 0):  T = [x ; ly] in [B, n_f+1, d],  Z = T T^t,
// R = [x, Z[i][j] for i > j in row-major order]  ->  [B, d + (n_f+1) n_f / 2].
//
// 2*(n_f+1)^2*d FLOP against (n_f+1)*d*4 + (d + pairs)*4 bytes per sample is 7-11 FLOP/B,
// far below the tensor-core ridge, so this is an fp32 FMA kernel bound by HBM: one warp per
// sample, T staged in shared memory (padded rows), each lane owns every 32nd (i, j) pair.
#pragma once
#include "evs_host.h"

namespace evs {

constexpr int kInteractWarps = 4;

__global__ void __launch_bounds__(kInteractWarps * 32) k_interact(const float *__restrict__ x, const float *__restrict__ ly,
                                                                  float *__restrict__ r, int B, int n_f, int D) {
    extern __shared__ float s_t[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nt = n_f + 1, ld = D + 1;
    float *t = s_t + static_cast<size_t>(warp) * nt * ld;
    const int n_pairs = nt * (nt - 1) / 2;
    const int out_w = D + n_pairs;
    for (int s = blockIdx.x * kInteractWarps + warp; s < B; s += gridDim.x * kInteractWarps) {
        const float *xs = x + static_cast<size_t>(s) * D;
        const float *ls = ly + static_cast<size_t>(s) * n_f * D;
        float *rs = r + static_cast<size_t>(s) * out_w;
        for (int e = lane; e < D; e += 32) {
            const float v = __ldg(xs + e);
            t[e] = v;
            rs[e] = v;
        }
        for (int e = lane; e < n_f * D; e += 32) {
            const int f = e / D, k = e - f * D;
            t[(f + 1) * ld + k] = __ldg(ls + e);
        }
        __syncwarp();
        for (int pr = lane; pr < n_pairs; pr += 32) {
            // invert pr = i(i-1)/2 + j, 0 <= j < i
            int i = static_cast<int>((1.0f + sqrtf(1.0f + 8.0f * static_cast<float>(pr))) * 0.5f);
            while (i * (i - 1) / 2 > pr) --i;
            while ((i + 1) * i / 2 <= pr) ++i;
            const int j = pr - i * (i - 1) / 2;
            const float *a = t + i * ld, *b = t + j * ld;
            float acc = 0.0f;
            for (int k = 0; k < D; ++k) acc = fmaf(a[k], b[k], acc);
            rs[D + pr] = acc;
        }
        __syncwarp();
    }
}

inline int launch_interact(const float *x, const float *ly, float *r, int B, int n_f, int D, cudaStream_t st) {
    if (B < 0 || n_f < 1 || D < 1 || x == nullptr || ly == nullptr || r == nullptr) return EVS_ERR_INVALID;
    if (B == 0) return EVS_OK;
    const size_t smem = static_cast<size_t>(kInteractWarps) * (n_f + 1) * (D + 1) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_interact, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) {
            set_error(std::string("k_interact smem: ") + cudaGetErrorString(e));
            return EVS_ERR_CUDA;
        }
    }
    const int ctas = std::min((B + kInteractWarps - 1) / kInteractWarps, 148 * 8);
    k_interact<<<ctas, kInteractWarps * 32, smem, st>>>(x, ly, r, B, n_f, D);
    EVS_CUDA(cudaGetLastError());
    return EVS_OK;
}

}  // namespace evs
