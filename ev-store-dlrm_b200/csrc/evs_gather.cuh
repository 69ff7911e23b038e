// Sum-pooling gather: the op the cache output stands in for, nn.EmbeddingBag(mode="sum") of
// apply_emb_ori_dlrm (dlrm_s_pytorch_C1_C2_C3.py:191-223):  out[b] = sum_{j in bag b} w_j * W[idx_j],
// j ascending (the order the fp32 sum is defined in), W stored at 32 / 16 / 8 / 4 bits in the
// reference's formats (evs_codec.cuh).  The table may live in HBM or in mapped pinned host memory
// (the no-cache path of apply_emb_evstore, storage_manager.request_to_emb_storage).
//
// A group of lanes owns one bag; lane c of the group owns 16-byte chunk c of the row, so a warp
// reads whole rows with 128-bit loads and consecutive lanes write consecutive floats.
#pragma once
#include "evs_codec.cuh"
#include "evs_host.h"

namespace evs {

__device__ __forceinline__ uint4 gather_ldg16(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

template <int PREC>
__device__ __forceinline__ void decode_regs(uint4 v, float (&o)[ElemsPerChunk<PREC>::value], const CodecLut *l) {
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    if (PREC == 32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __uint_as_float(w[i]);
    } else if (PREC == 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[2 * i] = dec16(w[i] & 0xFFFFu);
            o[2 * i + 1] = dec16(w[i] >> 16);
        }
    } else if (PREC == 8) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[4 * i + k] = l->lut8[(w[i] >> (8 * k)) & 0xFFu];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {       // byte k: element 2k in the high nibble, 2k+1 in the low one
                const unsigned byte = (w[i] >> (8 * k)) & 0xFFu;
                o[8 * i + 2 * k] = l->lut4[byte >> 4];
                o[8 * i + 2 * k + 1] = l->lut4[byte & 15u];
            }
        }
    }
}

// one element of a raw row (rows whose byte length is not a multiple of 16)
template <int PREC>
__device__ __forceinline__ float decode_elem(const unsigned char *row, int e, const CodecLut *l) {
    if (PREC == 32) {
        unsigned u = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) u |= static_cast<unsigned>(__ldg(row + 4 * e + k)) << (8 * k);
        return __uint_as_float(u);
    } else if (PREC == 16) {
        return dec16(static_cast<unsigned>(__ldg(row + 2 * e)) | (static_cast<unsigned>(__ldg(row + 2 * e + 1)) << 8));
    } else if (PREC == 8) {
        return l->lut8[__ldg(row + e)];
    } else {
        const unsigned byte = __ldg(row + (e >> 1));
        return l->lut4[(e & 1) ? (byte & 15u) : (byte >> 4)];
    }
}

template <int PREC>
__global__ void __launch_bounds__(256) k_gather(const unsigned char *__restrict__ table, long long rows, int D,
                                                const long long *__restrict__ idx, const long long *__restrict__ off,
                                                long long nnz, int B, const float *__restrict__ psw, float *__restrict__ out,
                                                long long out_stride, unsigned *err) {
    __shared__ CodecLut s_lut;
    codec_lut_init<PREC, 0>(&s_lut);
    __syncthreads();
    constexpr int EPC = ElemsPerChunk<PREC>::value;
    const unsigned row_bytes = static_cast<unsigned>(D) * PREC / 8;
    const int lane = threadIdx.x & 31;
    const long long gwarp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const bool aligned = ((row_bytes & 15u) == 0) && ((reinterpret_cast<uintptr_t>(table) & 15u) == 0);
    if (aligned) {
        const int cpr = static_cast<int>(row_bytes >> 4);
        int gsize = 1;
        while (gsize < cpr && gsize < 32) gsize <<= 1;
        const int gpw = 32 / gsize, grp = lane / gsize, gl = lane - grp * gsize;
        const bool vec = ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) && ((out_stride & 3) == 0);
        for (long long bag = gwarp * gpw + grp; bag < B; bag += n_warps * gpw) {
            const long long j0 = __ldg(off + bag);
            const long long j1 = (bag + 1 < B) ? __ldg(off + bag + 1) : nnz;
            float *orow = out + bag * out_stride;
            for (int c = gl; c < cpr; c += gsize) {
                float acc[EPC];
#pragma unroll
                for (int i = 0; i < EPC; ++i) acc[i] = 0.0f;
                // U rows of the bag per round: the index loads, then the row loads, are all issued before the first
                // add, so a bag costs ceil(P / U) dependent round trips instead of P; the sum still runs in ascending j.
                // fp32 rows: the reference's 10 indices per lookup (dlrm_data_pytorch.py:961-1007) in ONE round -- the
                // kernel is a single wave of short-lived warps, so its time is the length of this chain (ncu,
                // profiles/r2_gather_d64: DRAM 34 % busy, issue slots 31 %)
                constexpr int U = 5;   // (10 in one round was slower: 80 registers, 18.5 vs 14.5 us at d = 64)
                for (long long jb = j0; jb < j1; jb += U) {
                    long long rr[U];
                    uint4 v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        rr[u] = (jb + u < j1) ? __ldg(idx + jb + u) : 0;
                        if (rr[u] < 0 || rr[u] >= rows) {
                            *err = 1u;
                            rr[u] = 0;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (jb + u < j1) v[u] = gather_ldg16(table + static_cast<size_t>(rr[u]) * row_bytes + (c << 4));
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (jb + u >= j1) break;
                        float x[EPC];
                        decode_regs<PREC>(v[u], x, &s_lut);
                        if (psw != nullptr) {
                            const float w = __ldg(psw + jb + u);
#pragma unroll
                            for (int i = 0; i < EPC; ++i) acc[i] = __fadd_rn(acc[i], __fmul_rn(w, x[i]));
                        } else {
#pragma unroll
                            for (int i = 0; i < EPC; ++i) acc[i] = __fadd_rn(acc[i], x[i]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < EPC; i += 4) store4(orow, vec, c * EPC + i, D, acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
            }
        }
    } else {
        for (long long bag = gwarp; bag < B; bag += n_warps) {
            const long long j0 = __ldg(off + bag);
            const long long j1 = (bag + 1 < B) ? __ldg(off + bag + 1) : nnz;
            for (int e = lane; e < D; e += 32) {
                float acc = 0.0f;
                for (long long j = j0; j < j1; ++j) {
                    long long r = __ldg(idx + j);
                    if (r < 0 || r >= rows) {
                        *err = 1u;
                        r = 0;
                    }
                    const float x = decode_elem<PREC>(table + static_cast<size_t>(r) * row_bytes, e, &s_lut);
                    acc = (psw != nullptr) ? __fadd_rn(acc, __fmul_rn(__ldg(psw + j), x)) : __fadd_rn(acc, x);
                }
                out[bag * out_stride + e] = acc;
            }
        }
    }
}

inline int launch_gather(const void *table, long long rows, int D, int prec, const long long *idx, const long long *off,
                         long long nnz, int B, const float *psw, float *out, long long out_stride, unsigned *err,
                         cudaStream_t st) {
    if (table == nullptr || rows < 1 || D < 1 || B < 0 || nnz < 0 || (B > 0 && (off == nullptr || out == nullptr)) ||
        (nnz > 0 && idx == nullptr) || (static_cast<long long>(D) * prec) % 8 != 0)
        return EVS_ERR_INVALID;
    if (B == 0) return EVS_OK;
    if (out_stride <= 0) out_stride = D;
    const unsigned char *tb = static_cast<const unsigned char *>(table);
    const int ctas = static_cast<int>(std::min<long long>((static_cast<long long>(B) + 7) / 8, 148 * 8));
    switch (prec) {
        case 32: k_gather<32><<<ctas, 256, 0, st>>>(tb, rows, D, idx, off, nnz, B, psw, out, out_stride, err); break;
        case 16: k_gather<16><<<ctas, 256, 0, st>>>(tb, rows, D, idx, off, nnz, B, psw, out, out_stride, err); break;
        case 8: k_gather<8><<<ctas, 256, 0, st>>>(tb, rows, D, idx, off, nnz, B, psw, out, out_stride, err); break;
        case 4: k_gather<4><<<ctas, 256, 0, st>>>(tb, rows, D, idx, off, nnz, B, psw, out, out_stride, err); break;
        default: return EVS_ERR_INVALID;
    }
    EVS_CUDA(cudaGetLastError());
    return EVS_OK;
}

}  // namespace evs
