// Shared by the two Q x C all-pairs kernels (l2max_allpairs.cu: tsAspire, ot_allpairs.cu: otAspire): the tcgen05 main
// loop's stage geometry, the 64B-swizzle shared-memory descriptor, and the operand preparation (fp32 rows -> bf16 hi / lo
// halves + exact fp32 squared norms).  Kernels defined here are `static`: each translation unit launches its own copy
// (the library is built without relocatable device code).
#pragma once
#include <algorithm>
#include "common.cuh"
#include "bert/tc05.cuh"

namespace asp {

using namespace tc;

int make_tmap_bf16_k32(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows);

// One pipeline stage = one 32-wide K block of ALL FOUR operand tiles (A_hi, A_lo, B_hi, B_lo), each loaded exactly once
// and used by the four split-precision MMAs of that block (an earlier version reloaded A and B for every term: twice
// the L2 -> shared-memory traffic, which is what bounded it).  64-byte rows, 64B swizzle.
constexpr int kApBlockM = 128, kApBlockN = 160, kApBlockK = 32;
constexpr int kApABytes = kApBlockM * kApBlockK * 2, kApBBytes = kApBlockN * kApBlockK * 2;  // 8 KB, 10 KB
constexpr int kApStage = 2 * kApABytes + 2 * kApBBytes;                                     // 36 KB

// Shared-memory matrix descriptor of a K-major tile stored as 64-byte rows with the 64B swizzle (what TMA writes with
// CU_TENSOR_MAP_SWIZZLE_64B and a 64-byte inner box): 8-row groups are 512 B apart, layout type 4 = SWIZZLE_64B.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

// fp32 rows -> bf16 hi / lo halves + exact fp32 squared norm.  One warp per row; D % 4 == 0.
static __global__ void __launch_bounds__(128)
split_rows_kernel(const float* __restrict__ x, long long rows, int D, __nv_bfloat16* __restrict__ hi,
                  __nv_bfloat16* __restrict__ lo, float* __restrict__ norms) {
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float4* src = reinterpret_cast<const float4*>(x + r * D);
    float s = 0.f;
    for (int k4 = lane; k4 < (D >> 2); k4 += 32) {
        const float4 v = ldg_stream(src + k4);
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __low2float(h0), v.y - __high2float(h0));
        const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __low2float(h1), v.w - __high2float(h1));
        uint2 ph, pl;
        ph.x = *reinterpret_cast<const uint32_t*>(&h0); ph.y = *reinterpret_cast<const uint32_t*>(&h1);
        pl.x = *reinterpret_cast<const uint32_t*>(&l0); pl.y = *reinterpret_cast<const uint32_t*>(&l1);
        reinterpret_cast<uint2*>(hi + r * D)[k4] = ph;
        reinterpret_cast<uint2*>(lo + r * D)[k4] = pl;
    }
    s = warp_sum(s);
    if (lane == 0) norms[r] = s;
}

// Operand copies of one side in a caller-owned workspace: [rows, D] bf16 hi, [rows, D] bf16 lo, [rows] fp32 |x|^2.
struct SplitRows {
    __nv_bfloat16 *hi, *lo;
    float* norms;
};
inline size_t split_rows_bytes(size_t rows, int D) { return rows * D * 2 * 2 + rows * sizeof(float) + 512; }
// carve `rows` rows out of *w (256-byte aligned) and launch the split; advances *w
inline int split_rows_launch(const float* x, size_t rows, int D, char** w, SplitRows* out, cudaStream_t stream) {
    char* p = *w;
    p += (256 - (reinterpret_cast<uintptr_t>(p) & 255)) & 255;
    out->hi = reinterpret_cast<__nv_bfloat16*>(p); p += rows * D * 2;
    out->lo = reinterpret_cast<__nv_bfloat16*>(p); p += rows * D * 2;
    out->norms = reinterpret_cast<float*>(p); p += rows * sizeof(float);
    *w = p;
    if (rows == 0) return ASP_OK;
    split_rows_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, stream>>>(x, (long long)rows, D, out->hi, out->lo, out->norms);
    ASP_LAUNCH_CHECK("split_rows_kernel");
    return ASP_OK;
}

}  // namespace asp
