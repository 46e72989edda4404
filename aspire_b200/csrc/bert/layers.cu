// K0 row-wise layers of the encoder: embedding sum + LayerNorm, and residual LayerNorm.
//
// Replaces BertEmbeddings (word + position + token-type lookup, LayerNorm eps 1e-12) and the two post-LN
// LayerNorms of every BertLayer as reached from examples/ex_aspire_consent.py:72.  HBM-bound: one warp per token
// row, 128-bit loads, the whole row lives in registers (two-pass mean / variance, no re-read), and the row is
// written once as fp32 (the residual stream) plus its bf16 hi (and, in bf16x3 mode, lo) copy -- the operand format
// of the tcgen05 GEMMs -- so no separate cast pass ever touches HBM.
#include "../common.cuh"
#include <cuda_bf16.h>

namespace asp {

constexpr int kMaxHiddenPerLane = 32;  // hidden <= 32 lanes * 32 floats = 1024

template <int VPL>  // float4 vectors per lane: hidden = 128 * VPL
__device__ __forceinline__ void ln_row_store(float4 (&x)[VPL], int lane, const float* __restrict__ gamma,
                                             const float* __restrict__ beta, float eps, float* __restrict__ out_f32,
                                             __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                             float2* __restrict__ stats = nullptr) {
    // out_f32 == nullptr: the fp32 row is not written -- its consumer (the residual epilogue of a GEMM) rebuilds it from
    // the pre-LayerNorm row and *stats = (mean, rstd) with the expression below (gemm.cu: ln_on_read)
    constexpr float inv_n = 1.0f / (128 * VPL);
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) s += (x[v].x + x[v].y) + (x[v].z + x[v].w);
    const float mean = warp_sum(s) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const float a = x[v].x - mean, b = x[v].y - mean, c = x[v].z - mean, d = x[v].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_n + eps);
    if (stats && lane == 0) *stats = make_float2(mean, rstd);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int k4 = v * 32 + lane;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + k4);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + k4);
        float4 y;
        y.x = (x[v].x - mean) * rstd * g.x + b.x;
        y.y = (x[v].y - mean) * rstd * g.y + b.y;
        y.z = (x[v].z - mean) * rstd * g.z + b.z;
        y.w = (x[v].w - mean) * rstd * g.w + b.w;
        if (out_f32) reinterpret_cast<float4*>(out_f32)[k4] = y;
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(y.x, y.y), h1 = __floats2bfloat162_rn(y.z, y.w);
        uint2 ph;
        ph.x = *reinterpret_cast<const uint32_t*>(&h0);
        ph.y = *reinterpret_cast<const uint32_t*>(&h1);
        reinterpret_cast<uint2*>(out_hi)[k4] = ph;
        if (out_lo) {
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(y.x - __low2float(h0), y.y - __high2float(h0));
            const __nv_bfloat162 l1 = __floats2bfloat162_rn(y.z - __low2float(h1), y.w - __high2float(h1));
            uint2 pl;
            pl.x = *reinterpret_cast<const uint32_t*>(&l0);
            pl.y = *reinterpret_cast<const uint32_t*>(&l1);
            reinterpret_cast<uint2*>(out_lo)[k4] = pl;
        }
    }
}

template <int VPL>
__global__ void __launch_bounds__(128)
embed_ln_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ type_ids, int T, int L, int vocab, int max_pos,
                const float* __restrict__ word_emb, const float* __restrict__ pos_emb, const float* __restrict__ type_emb,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* __restrict__ out_f32,
                __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
    constexpr int H = 128 * VPL;
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= T) return;
    const int id = min(max(ids[t], 0), vocab - 1);
    const int pos = min(t % L, max_pos - 1);
    const int tt = type_ids ? min(max(type_ids[t], 0), 1) : 0;
    const float4* w = reinterpret_cast<const float4*>(word_emb + (size_t)id * H);
    const float4* p = reinterpret_cast<const float4*>(pos_emb + (size_t)pos * H);
    const float4* y = reinterpret_cast<const float4*>(type_emb + (size_t)tt * H);
    float4 x[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int k4 = v * 32 + lane;
        const float4 a = __ldg(w + k4), b = __ldg(p + k4), c = __ldg(y + k4);
        x[v] = make_float4((a.x + c.x) + b.x, (a.y + c.y) + b.y, (a.z + c.z) + b.z, (a.w + c.w) + b.w);
    }
    ln_row_store<VPL>(x, lane, gamma, beta, eps, out_f32 + (size_t)t * H, out_hi + (size_t)t * H,
                      out_lo ? out_lo + (size_t)t * H : nullptr);
}

// in: fp32 [T,H] = GEMM output + bias + residual (written by the GEMM epilogue).  May alias out_f32.
template <int VPL>
__global__ void __launch_bounds__(128)
ln_kernel(const float* in, int T, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* out_f32,
          __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, float2* __restrict__ stats) {
    constexpr int H = 128 * VPL;
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= T) return;
    const float4* r = reinterpret_cast<const float4*>(in + (size_t)t * H);
    float4 x[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) x[v] = r[v * 32 + lane];
    ln_row_store<VPL>(x, lane, gamma, beta, eps, out_f32 ? out_f32 + (size_t)t * H : nullptr, out_hi + (size_t)t * H,
                      out_lo ? out_lo + (size_t)t * H : nullptr, stats ? stats + t : nullptr);
}

int embed_ln_launch(const int32_t* ids, const int32_t* type_ids, int T, int L, int H, int vocab, int max_pos,
                    const float* word_emb, const float* pos_emb, const float* type_emb, const float* gamma, const float* beta,
                    float eps, float* out_f32, void* out_hi, void* out_lo, cudaStream_t stream) {
    ASP_REQUIRE(H == 768, "encoder: hidden size %d not built (768 only)", H);
    const int wpb = 4, blocks = (T + wpb - 1) / wpb;
    ASP_CUDA(launch_pdl(embed_ln_kernel<6>, dim3(blocks), dim3(wpb * 32), 0, stream, ids, type_ids, T, L, vocab, max_pos, word_emb,
                        pos_emb, type_emb, gamma, beta, eps, out_f32, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo));
    ASP_LAUNCH_CHECK("embed_ln_kernel");
    return ASP_OK;
}

int ln_launch(const float* in, int T, int H, const float* gamma, const float* beta, float eps, float* out_f32, void* out_hi,
              void* out_lo, cudaStream_t stream, void* stats) {
    ASP_REQUIRE(out_f32 || stats, "encoder: a LayerNorm without fp32 output must leave its row statistics");
    ASP_REQUIRE(H == 768, "encoder: hidden size %d not built (768 only)", H);
    const int wpb = 4, blocks = (T + wpb - 1) / wpb;
    ASP_CUDA(launch_pdl(ln_kernel<6>, dim3(blocks), dim3(wpb * 32), 0, stream, in, T, gamma, beta, eps, out_f32,
                        (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, (float2*)stats));
    ASP_LAUNCH_CHECK("ln_kernel");
    return ASP_OK;
}

}  // namespace asp
