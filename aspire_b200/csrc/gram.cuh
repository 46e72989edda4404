// Warp-level fp32 Gram-tile helpers shared by pair_cost.cu and ot_fused.cu: the 32 lanes split the embedding
// dimension, each lane keeps partial dot products for a TI x TJ tile of sentence pairs plus the partial squared
// norms, and a transpose-reduce over the warp leaves each finished sum in exactly one lane.
#pragma once
#include "common.cuh"

namespace asp {

template <int TI, int TJ>
struct GramTile {
    static constexpr int kEntries = TI * TJ;
    static constexpr int kVals = kEntries + TI + TJ;           // dots, |q_i|^2, |c_j|^2
    static constexpr int NV = ((kVals + 31) / 32) * 32;        // padded for the transpose-reduce
};

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
    return acc;
}

// Accumulate this lane's share (k = 4*lane + 128*m) of a TI x TJ tile.
// STREAM_C: the candidate rows are read exactly once by this kernel (single-tile pairs) -> do not allocate them in L1;
// false when several tiles / warps re-read the same rows (long documents) and the caches should keep them.
template <int TI, int TJ, bool STREAM_C = true>
__device__ __forceinline__ void gram_accumulate(const float* __restrict__ q, int nq, const float* __restrict__ c,
                                                int nc, int D, int lane, float (&v)[GramTile<TI, TJ>::NV]) {
    using T = GramTile<TI, TJ>;
    const int d4 = D >> 2;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k4 = lane; k4 < d4; k4 += 32) {
        float4 cv[TJ];
#pragma unroll
        for (int j = 0; j < TJ; ++j)
            cv[j] = (j < nc) ? (STREAM_C ? ldg_stream(reinterpret_cast<const float4*>(c + (size_t)j * D) + k4)
                                         : __ldg(reinterpret_cast<const float4*>(c + (size_t)j * D) + k4))
                             : zero4;
#pragma unroll
        for (int i = 0; i < TI; ++i) {
            const float4 qv = (i < nq) ? __ldg(reinterpret_cast<const float4*>(q + (size_t)i * D) + k4) : zero4;
#pragma unroll
            for (int j = 0; j < TJ; ++j) v[i * TJ + j] = dot4(qv, cv[j], v[i * TJ + j]);
            v[T::kEntries + i] = dot4(qv, qv, v[T::kEntries + i]);
        }
#pragma unroll
        for (int j = 0; j < TJ; ++j) v[T::kEntries + TI + j] = dot4(cv[j], cv[j], v[T::kEntries + TI + j]);
    }
}

// Sum v[] over the 32 lanes.  After the call lane l holds, in v[m], the total of original slot
// 32*m + bitrev5(l).  Cost: NV*(1/2+1/4+...) ~ NV shuffles.
template <int NV>
__device__ __forceinline__ void transpose_reduce(float (&v)[NV], int lane) {
    int n = NV;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const bool upper = (lane & s) != 0;
        n >>= 1;
#pragma unroll
        for (int m = 0; m < NV / 2; ++m) {
            if (m < n) {
                const float a = v[2 * m], b = v[2 * m + 1];
                const float send = upper ? a : b;
                const float keep = upper ? b : a;
                v[m] = keep + __shfl_xor_sync(0xffffffffu, send, s);
            }
        }
    }
}

}  // namespace asp
