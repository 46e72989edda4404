// K0 self-attention of the encoder: softmax(Q K^T / sqrt(64) + key-padding mask) V per (document, head).
//
// Replaces BertSelfAttention as reached from examples/ex_aspire_consent.py:72 (12 heads x 64, L <= 512, additive
// mask that removes padded keys).  Flash-style: one CTA = 16 * QW query rows (QW = 4 warps) of one
// (document, head); K/V tiles of 64 keys stream through a two-stage cp.async ring in shared memory (tile j+1 loads
// while tile j is multiplied), scores / probabilities never leave registers, online softmax in fp32.
// The two small contractions (64x64x64 per tile) run on mma.sync.m16n8k16 bf16 tensor-core fragments: at L <= 512
// attention is ~3-6 % of the encoder FLOPs, the tcgen05 budget goes to the GEMMs (gemm.cu).
// PRECISE ("bf16x3"): Q, K, V and P are carried as (hi, lo) bf16 pairs and every product is hi.hi + hi.lo + lo.hi,
// which keeps the layer fp32-equivalent like the GEMMs around it.
#include "../common.cuh"
#include <cuda_bf16.h>

namespace asp {

constexpr int kHeadDim = 64;
constexpr int kAttnTile = 64;   // keys per staged tile
constexpr int kAttnLd = 72;     // bf16 row stride of the staged tiles (144 B: conflict-free fragment loads)

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void split_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = pack_bf16(a - __low2float(h), b - __high2float(h));
}

// 16-byte global -> shared copy without a register round trip; src_bytes = 0 writes zeros (keys past the sequence).
__device__ __forceinline__ void attn_cp_async16(void* smem_dst, const void* gmem_src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(src_bytes)
                 : "memory");
}

template <int NP>
struct AttnSmem {
    __nv_bfloat16 k[2][NP][kAttnTile][kAttnLd];
    __nv_bfloat16 v[2][NP][kAttnTile][kAttnLd];
};

template <bool PRECISE, int QW>
__global__ void __launch_bounds__(32 * QW)
attention_kernel(const __nv_bfloat16* __restrict__ qkv_hi, const __nv_bfloat16* __restrict__ qkv_lo,
                 const int32_t* __restrict__ seq_lens, int L, int H, __nv_bfloat16* __restrict__ ctx_hi,
                 __nv_bfloat16* __restrict__ ctx_lo) {
    constexpr int NP = PRECISE ? 2 : 1;
    extern __shared__ __align__(16) uint8_t attn_smem_raw[];
    AttnSmem<NP>& sm = *reinterpret_cast<AttnSmem<NP>*>(attn_smem_raw);
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * (16 * QW);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int kv_len = min(max(seq_lens[b], 1), L);
    const size_t ld = (size_t)3 * H;
    const __nv_bfloat16* base[2] = {qkv_hi + (size_t)b * L * ld + head * kHeadDim,
                                    PRECISE ? qkv_lo + (size_t)b * L * ld + head * kHeadDim : nullptr};

    // Q fragments of this warp's 16 rows (rows beyond L read as zero)
    uint32_t qa[NP][4][4];
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int c = kk * 16 + 2 * t;
            qa[p][kk][0] = r0 < L ? *reinterpret_cast<const uint32_t*>(base[p] + (size_t)r0 * ld + c) : 0u;
            qa[p][kk][1] = r1 < L ? *reinterpret_cast<const uint32_t*>(base[p] + (size_t)r1 * ld + c) : 0u;
            qa[p][kk][2] = r0 < L ? *reinterpret_cast<const uint32_t*>(base[p] + (size_t)r0 * ld + c + 8) : 0u;
            qa[p][kk][3] = r1 < L ? *reinterpret_cast<const uint32_t*>(base[p] + (size_t)r1 * ld + c + 8) : 0u;
        }

    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const float sc = 0.125f * kLog2e;  // 1/sqrt(64), base-2 exponent

    // K and V rows j0..j0+63 of both halves into ring stage st (keys >= L are zero-filled, never read from memory)
    auto stage_tile = [&](int j0, int st) {
        for (int e = threadIdx.x; e < NP * 2 * kAttnTile * 8; e += 32 * QW) {
            const int c8 = e & 7, row = (e >> 3) & 63, which = (e >> 9) & 1, p = e >> 10;
            const int key = j0 + row;
            const __nv_bfloat16* src = (p ? base[NP - 1] : base[0]) + (size_t)min(key, L - 1) * ld + (which + 1) * H + c8 * 8;
            attn_cp_async16(which ? &sm.v[st][p][row][c8 * 8] : &sm.k[st][p][row][c8 * 8], src, key < L ? 16u : 0u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage_tile(0, 0);

    for (int j0 = 0, st = 0; j0 < kv_len; j0 += kAttnTile, st ^= 1) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // tile j0 is visible to every warp, and every warp is done with the stage refilled next
        if (j0 + kAttnTile < kv_len) stage_tile(j0 + kAttnTile, st ^ 1);
        const auto& Ks = sm.k[st];
        const auto& Vs = sm.v[st];

        // ---- S = Q K^T ----
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Ks[0][n * 8 + g][kk * 16 + 2 * t]);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Ks[0][n * 8 + g][kk * 16 + 8 + 2 * t]);
                mma_bf16(s[n], qa[0][kk], b0, b1);
                if (PRECISE) {
                    const uint32_t c0 = *reinterpret_cast<const uint32_t*>(&Ks[NP - 1][n * 8 + g][kk * 16 + 2 * t]);
                    const uint32_t c1 = *reinterpret_cast<const uint32_t*>(&Ks[NP - 1][n * 8 + g][kk * 16 + 8 + 2 * t]);
                    mma_bf16(s[n], qa[0][kk], c0, c1);        // hi . lo
                    mma_bf16(s[n], qa[NP - 1][kk], b0, b1);   // lo . hi
                }
            }
        // ---- mask padded keys (last tile only), online softmax (rows g and g+8 of this warp's 16) ----
        // The 1/sqrt(64) scale and the base-2 conversion ride in the exponent's FFMA: the running maximum is kept on
        // the raw scores (the scale is positive), p = exp2(s * sc - m * sc).
        if (j0 + kAttnTile > kv_len) {  // block-uniform
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int key = j0 + n * 8 + 2 * t;
                if (key >= kv_len) s[n][0] = s[n][2] = -INFINITY;
                if (key + 1 >= kv_len) s[n][1] = s[n][3] = -INFINITY;
            }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
            mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: every tile holds >= 1 valid key
        const float a0 = exp2f((m0 - mn0) * sc), a1 = exp2f((m1 - mn1) * sc);
        m0 = mn0;
        m1 = mn1;
        const float ms0 = -mn0 * sc, ms1 = -mn1 * sc;
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            s[n][0] = exp2f(fmaf(s[n][0], sc, ms0));
            s[n][1] = exp2f(fmaf(s[n][1], sc, ms0));
            s[n][2] = exp2f(fmaf(s[n][2], sc, ms1));
            s[n][3] = exp2f(fmaf(s[n][3], sc, ms1));
            sum0 += s[n][0] + s[n][1];
            sum1 += s[n][2] + s[n][3];
            o[n][0] *= a0; o[n][1] *= a0; o[n][2] *= a1; o[n][3] *= a1;
        }
        l0 = l0 * a0 + sum0;
        l1 = l1 * a1 + sum1;
        // ---- O += P V ----
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t pa[NP][4];
            if (PRECISE) {
                split_bf16(s[2 * kk][0], s[2 * kk][1], pa[0][0], pa[NP - 1][0]);
                split_bf16(s[2 * kk][2], s[2 * kk][3], pa[0][1], pa[NP - 1][1]);
                split_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1], pa[0][2], pa[NP - 1][2]);
                split_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3], pa[0][3], pa[NP - 1][3]);
            } else {
                pa[0][0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
                pa[0][1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
                pa[0][2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
                pa[0][3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
            }
#pragma unroll
            for (int dn = 0; dn < 8; dn += 2) {
                // lanes 0-15: rows kk*16+lane of d-tile dn; lanes 16-31: the same rows of d-tile dn+1
                uint32_t vb[4];
                ldmatrix_x4_trans(vb, &Vs[0][kk * 16 + (lane & 15)][(dn + (lane >> 4)) * 8]);
                mma_bf16(o[dn], pa[0], vb[0], vb[1]);
                mma_bf16(o[dn + 1], pa[0], vb[2], vb[3]);
                if (PRECISE) {
                    uint32_t vl[4];
                    ldmatrix_x4_trans(vl, &Vs[NP - 1][kk * 16 + (lane & 15)][(dn + (lane >> 4)) * 8]);
                    mma_bf16(o[dn], pa[0], vl[0], vl[1]);            // hi . lo
                    mma_bf16(o[dn + 1], pa[0], vl[2], vl[3]);
                    mma_bf16(o[dn], pa[NP - 1], vb[0], vb[1]);       // lo . hi
                    mma_bf16(o[dn + 1], pa[NP - 1], vb[2], vb[3]);
                }
            }
        }
    }
    // ---- finalize: divide by the row sums (reduced over the quad), write the context rows ----
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const int col = head * kHeadDim + n * 8 + 2 * t;
        uint32_t hi, lo;
        if (r0 < L) {
            split_bf16(o[n][0] * i0, o[n][1] * i0, hi, lo);
            *reinterpret_cast<uint32_t*>(ctx_hi + ((size_t)b * L + r0) * H + col) = hi;
            if (PRECISE) *reinterpret_cast<uint32_t*>(ctx_lo + ((size_t)b * L + r0) * H + col) = lo;
        }
        if (r1 < L) {
            split_bf16(o[n][2] * i1, o[n][3] * i1, hi, lo);
            *reinterpret_cast<uint32_t*>(ctx_hi + ((size_t)b * L + r1) * H + col) = hi;
            if (PRECISE) *reinterpret_cast<uint32_t*>(ctx_lo + ((size_t)b * L + r1) * H + col) = lo;
        }
    }
}

template <bool PRECISE, int QW>
static int attention_launch_as(const void* qkv_hi, const void* qkv_lo, const int32_t* seq_lens, int B, int L, int H, int heads,
                               void* ctx_hi, void* ctx_lo, cudaStream_t stream) {
    constexpr int kSmem = (int)sizeof(AttnSmem<PRECISE ? 2 : 1>);
    static thread_local int attr_dev = -1;
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        ASP_CUDA(cudaFuncSetAttribute(attention_kernel<PRECISE, QW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_dev = dev;
    }
    dim3 grid((L + 16 * QW - 1) / (16 * QW), heads, B);
    ASP_CUDA(launch_pdl(attention_kernel<PRECISE, QW>, grid, dim3(32 * QW), kSmem, stream, (const __nv_bfloat16*)qkv_hi,
                        (const __nv_bfloat16*)qkv_lo, seq_lens, L, H, (__nv_bfloat16*)ctx_hi, (__nv_bfloat16*)ctx_lo));
    ASP_LAUNCH_CHECK("attention_kernel");
    return ASP_OK;
}

bool attention_tc_supported(const void* qkv_lo, int L, int H, int heads);
int attention_tc_launch(const void* qkv_hi, const int32_t* seq_lens, int B, int L, int H, int heads, void* ctx_hi,
                        cudaStream_t stream);

int attention_launch(const void* qkv_hi, const void* qkv_lo, const int32_t* seq_lens, int B, int L, int H, int heads,
                     void* ctx_hi, void* ctx_lo, cudaStream_t stream) {
    ASP_REQUIRE(H == heads * kHeadDim, "attention: head size must be 64 (hidden %d, heads %d)", H, heads);
    // plain bf16, L <= 256: both contractions on tcgen05 (attention_tc.cu); longer sequences and the bf16x3 mode stay here
    if (attention_tc_supported(qkv_lo, L, H, heads)) return attention_tc_launch(qkv_hi, seq_lens, B, L, H, heads, ctx_hi, stream);
    // (128 query rows per CTA -- QW = 8 -- halve the K/V staging per query but measured 1 % slower at B=32, L=256: two
    // 8-warp CTAs per SM hide the softmax latency worse than four 4-warp ones.)
    if (qkv_lo) return attention_launch_as<true, 4>(qkv_hi, qkv_lo, seq_lens, B, L, H, heads, ctx_hi, ctx_lo, stream);
    return attention_launch_as<false, 4>(qkv_hi, nullptr, seq_lens, B, L, H, heads, ctx_hi, nullptr, stream);
}

}  // namespace asp
