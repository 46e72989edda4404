// K0 GEMMs: out[M,N] = epilogue(A[M,K] . W[N,K]^T + bias) on the 5th-generation tensor cores.
//
// Replaces the nn.Linear calls inside HF BertModel as reached from AspireConSent.consent_reps_bert
// (examples/ex_aspire_consent.py:72): Q/K/V, attention output, and the two feed-forward projections of each of the
// 12 layers.  W keeps the PyTorch Linear layout [out_features, in_features], so both operands are K-major.
//
// Three kernels share the operand format, the epilogues and the results (bit-identical, tests/test_gemm_gpu.py):
//   * gemm_tn_persistent_kernel -- the product path: one CTA per SM walks 128 x {128,192,256} tiles with a TMA ring that
//     runs ahead across tiles, a double-buffered TMEM accumulator and 8 epilogue warps (optionally 2/4-CTA clusters that
//     share W tiles by TMA multicast);
//   * gemm_tn_pair_kernel       -- the same pipeline on CTA pairs (tcgen05 cta_group::2, M = 256 over two SMs);
//   * gemm_tn_kernel            -- one tile per CTA, the round's first version, kept as the measured baseline:
//
// One CTA (128 threads) computes a 128 x BLOCK_N output tile:
//   * warp 0 / lane 0  -- TMA producer: cp.async.bulk.tensor 128x64 (A) and BLOCK_Nx64 (W) bf16 boxes, 128B swizzle,
//                         into a STAGES-deep shared-memory ring, completion on mbarriers (expect_tx);
//   * warp 1 / lane 0  -- MMA issuer: 4 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BLOCK_N, K=16) per stage,
//                         fp32 accumulator in TMEM; tcgen05.commit releases the stage and finally signals the epilogue;
//   * warp 2           -- allocates / frees the TMEM columns;
//   * all 4 warps      -- epilogue: tcgen05.ld (32 lanes x 32 columns per warp), bias / GELU / residual, stores.
// Two CTAs are resident per SM (3 stages x 32 KB each), so one tile's epilogue overlaps the other's main loop.
//
// fp32-equivalent mode ("bf16x3"): every fp32 operand x is carried as hi = bf16(x), lo = bf16(x - hi); the main loop
// then runs three passes per K block into the SAME accumulator: hi.hi + hi.lo + lo.hi (the lo.lo term is below fp32
// resolution).  The producer simply picks the tensor map of the pass; nothing else changes.
#include <algorithm>
#include "../common.cuh"
#include <unordered_map>
#include "tc05.cuh"

namespace asp {

using namespace tc;

enum { EPI_BF16 = 0, EPI_GELU_BF16 = 1, EPI_RESID_F32 = 2, EPI_F32 = 3 };

struct GemmArgs {
    int M, N, K, nterms;
    const float* bias;
    const float* residual;
    __nv_bfloat16* out_hi;
    __nv_bfloat16* out_lo;
    float* out_f32;
    // LayerNorm on read (EPI_RESID_F32 only): `residual` then holds the PRE-LayerNorm rows and the epilogue adds
    // (r - mean) * rstd * gamma + beta -- the expression ln_row_store evaluates -- from the per-row (mean, rstd) the
    // LayerNorm kernel left behind, so the normalised fp32 tensor is never written to or read from memory.
    const float2* ln_stats;
    const float* ln_gamma;
    const float* ln_beta;
};

__device__ __forceinline__ float4 ln_on_read(const float4& r, const float2& st, const float4& g, const float4& b) {
    return make_float4((r.x - st.x) * st.y * g.x + b.x, (r.y - st.x) * st.y * g.y + b.y, (r.z - st.x) * st.y * g.z + b.z,
                       (r.w - st.x) * st.y * g.w + b.w);
}

constexpr int kGemmStages = 3;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 bytes = one swizzle row

template <int BLOCK_N>
struct GemmSmem {
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStage = kABytes + kBBytes;
    static constexpr int kBarOff = kGemmStages * kStage;
    static constexpr int kTotal = kBarOff + 128 + 1024;  // barriers + slack for the 1024-byte alignment
};

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) -- the exact (erf) form HF BERT uses.  erf by Abramowitz & Stegun 7.1.26
// (|error| <= 1.5e-7, i.e. at fp32 rounding level; libdevice's branch-free erff costs about twice as much, and the
// epilogue of the FFN1 GEMM evaluates it 16k-32k times per tile).  Two values at a time on the packed fp32 pipe
// (FFMA2 / FMUL2 / FADD2): 13 packed instructions and 4 MUFU per pair; with the scalar form the FFN1 epilogue was
// issue-bound (issue slots 49 % busy, tile time 1.25x the QKV GEMM's at equal main loops).
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
    const float2 z = __fmul2_rn(make_float2(fabsf(x.x), fabsf(x.y)), make_float2(0.70710678118654752f, 0.70710678118654752f));
    const float2 d = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
    float2 t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(d.y));
    float2 p = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
    p = __ffma2_rn(p, t, make_float2(1.421413741f, 1.421413741f));
    p = __ffma2_rn(p, t, make_float2(-0.284496736f, -0.284496736f));
    p = __ffma2_rn(p, t, make_float2(0.254829592f, 0.254829592f));
    const float2 w = __fmul2_rn(__fmul2_rn(z, z), make_float2(-1.4426950408889634f, -1.4426950408889634f));
    float2 e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(w.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(w.y));
    const float2 r = __fmul2_rn(__fmul2_rn(p, t), e);
    const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
    const float2 one_minus_r = __fadd2_rn(make_float2(1.0f, 1.0f), make_float2(-r.x, -r.y));
    return __ffma2_rn(make_float2(fabsf(h.x), fabsf(h.y)), one_minus_r, h);
}

// Epilogue of 32 consecutive columns of one output row: bias, GELU / residual, stores (hi and optional lo halves).
template <int EPI>
__device__ __forceinline__ void epilogue_store32(const GemmArgs& g, float (&v)[32], bool row_ok, size_t off, int col) {
    if (g.bias) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col + i));
            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
    }
    if (!row_ok) return;
    if (EPI == EPI_GELU_BF16) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            const float2 y = gelu_erf2(make_float2(v[i], v[i + 1]));
            v[i] = y.x;
            v[i + 1] = y.y;
        }
    }
    if (EPI == EPI_RESID_F32 || EPI == EPI_F32) {
        float4* o = reinterpret_cast<float4*>(g.out_f32 + off);
        const float4* r = (EPI == EPI_RESID_F32) ? reinterpret_cast<const float4*>(g.residual + off) : nullptr;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            float4 x = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            if (EPI == EPI_RESID_F32) {
                float4 rr = __ldg(r + i / 4);
                if (g.ln_stats)
                    rr = ln_on_read(rr, __ldg(g.ln_stats + off / g.N), __ldg(reinterpret_cast<const float4*>(g.ln_gamma + col + i)),
                                    __ldg(reinterpret_cast<const float4*>(g.ln_beta + col + i)));
                x.x += rr.x; x.y += rr.y; x.z += rr.z; x.w += rr.w;
            }
            o[i / 4] = x;
        }
    } else {
        uint4* oh = reinterpret_cast<uint4*>(g.out_hi + off);
        uint4* ol = g.out_lo ? reinterpret_cast<uint4*>(g.out_lo + off) : nullptr;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            __nv_bfloat162 h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = v[i + 2 * j], b = v[i + 2 * j + 1];
                h[j] = __floats2bfloat162_rn(a, b);
                l[j] = __floats2bfloat162_rn(a - __low2float(h[j]), b - __high2float(h[j]));
            }
            oh[i / 8] = *reinterpret_cast<uint4*>(h);
            if (ol) ol[i / 8] = *reinterpret_cast<uint4*>(l);
        }
    }
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(128)
gemm_tn_kernel(const __grid_constant__ CUtensorMap ta_hi, const __grid_constant__ CUtensorMap ta_lo,
               const __grid_constant__ CUtensorMap tb_hi, const __grid_constant__ CUtensorMap tb_lo, const GemmArgs g) {
    using S = GemmSmem<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
    uint64_t* empty = full + kGemmStages;
    uint64_t* accum = empty + kGemmStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * kBlockM, n0 = blockIdx.x * BLOCK_N;
    const int total = (g.K / kBlockK) * g.nterms;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&ta_hi);
        tma_prefetch_desc(&tb_hi);
        for (int s = 0; s < kGemmStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, BLOCK_N);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer ----------------
        for (int it = 0; it < total; ++it) {
            const int s = it % kGemmStages, ph = (it / kGemmStages) & 1;
            const int kb = it / g.nterms, term = it - kb * g.nterms;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], S::kStage);
            uint8_t* sa = smem + s * S::kStage;
            tma_load_2d(sa, term == 2 ? &ta_lo : &ta_hi, &full[s], kb * kBlockK, m0);
            tma_load_2d(sa + S::kABytes, term == 1 ? &tb_lo : &tb_hi, &full[s], kb * kBlockK, n0);
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N);
        for (int it = 0; it < total; ++it) {
            const int s = it % kGemmStages, ph = (it / kGemmStages) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after_sync();
            const uint32_t sa = smem_u32(smem + s * S::kStage);
            const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + S::kABytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)  // 16 bf16 = 32 bytes = +2 in the (>>4) start-address field
                umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
            umma_commit(&empty[s]);
        }
        umma_commit(accum);
    }
    __syncwarp();

    // ---------------- epilogue (all warps): TMEM lane = tile row, TMEM column = tile column ----------------
    mbar_wait(accum, 0);
    tc_fence_after_sync();
    const int row = m0 + warp * 32 + lane;
    const bool row_ok = row < g.M;
    const size_t row_off = (size_t)row * g.N + n0;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        epilogue_store32<EPI>(g, v, row_ok, row_off + c0, n0 + c0);  // no early exit before this: tcgen05.ld is warp-wide
        __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, BLOCK_N);
}

// Coalesced variant for the persistent kernel.  TMEM hands every lane one ROW of the tile, so storing straight from
// registers makes each warp-wide 16-byte store touch 32 different lines with a quarter sector each; at 64 KB per tile
// that alone took twice as long as the tile's main loop.  Here the 32x32 chunk is turned through a per-warp staging
// buffer (80-byte row pitch: 64 data bytes + 16 pad, conflict-free 16-byte writes) so that afterwards four adjacent
// lanes own 64 consecutive bytes of one row: every store instruction writes whole sectors of 8 rows, and the residual
// is read the same way.  bf16 outputs pass the chunk through once per half (hi, lo); fp32 outputs in two 16-column
// halves.
constexpr int kEpiPitch = 80;
constexpr int kEpiBytesPerWarp = 32 * kEpiPitch;

// The residual values a lane adds to one 32-column chunk: rows row0 + 8 i + lane / 4, columns col + 16 h + 4 (lane % 4) ..+3.
__device__ __forceinline__ void epilogue_load_residual(const GemmArgs& g, int lane, int row0, int col, float4 (&rr)[8]) {
    const int r_sub = lane >> 2, c_sub = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = min(row0 + 8 * i + r_sub, g.M - 1);
            rr[4 * h + i] = __ldg(reinterpret_cast<const float4*>(g.residual + (size_t)row * g.N + col + 16 * h + 4 * c_sub));
        }
}

// ln_st: the (mean, rstd) of this lane's four rows (row0 + 8 i + lane / 4) when the caller has already fetched them -- the
// persistent kernel does, once per tile and before it waits for the accumulator -- else nullptr.
template <int EPI>
__device__ __forceinline__ void epilogue_store32_staged(const GemmArgs& g, float (&v)[32], uint8_t* stage, int lane, int row0,
                                                        int col, const float2* ln_st = nullptr, const float4* rr_pre = nullptr) {
    if (g.bias) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col + i));
            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
    }
    if (EPI == EPI_GELU_BF16) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            const float2 y = gelu_erf2(make_float2(v[i], v[i + 1]));
            v[i] = y.x;
            v[i + 1] = y.y;
        }
    }
    const int r_sub = lane >> 2, c_sub = lane & 3;
    uint8_t* wr = stage + lane * kEpiPitch;
    const uint8_t* rd = stage + r_sub * kEpiPitch + c_sub * 16;
    if (EPI == EPI_RESID_F32 || EPI == EPI_F32) {
        // LayerNorm on read: the statistics of this lane's four rows once per chunk, gamma / beta of its columns once per half
        const bool lnr = EPI == EPI_RESID_F32 && g.ln_stats != nullptr;  // uniform
        float2 st[4];
        if (lnr) {
#pragma unroll
            for (int i = 0; i < 4; ++i) st[i] = ln_st ? ln_st[i] : __ldg(g.ln_stats + min(row0 + 8 * i + r_sub, g.M - 1));
        }
        // the residual loads of BOTH halves go out before the chunk is turned through the staging buffer: one L2 round trip
        // per chunk instead of one per half (the epilogue of the K = 768 residual GEMM is a chain of such round trips)
        float4 rr[8], gm[2], bt[2];
        if (EPI == EPI_RESID_F32) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = min(row0 + 8 * i + r_sub, g.M - 1);
                    rr[4 * h + i] = rr_pre ? rr_pre[4 * h + i]
                                           : __ldg(reinterpret_cast<const float4*>(g.residual + (size_t)row * g.N + col + 16 * h + 4 * c_sub));
                }
                if (lnr) {
                    gm[h] = __ldg(reinterpret_cast<const float4*>(g.ln_gamma + col + 16 * h + 4 * c_sub));
                    bt[h] = __ldg(reinterpret_cast<const float4*>(g.ln_beta + col + 16 * h + 4 * c_sub));
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(wr + 16 * j) =
                    make_float4(v[16 * h + 4 * j], v[16 * h + 4 * j + 1], v[16 * h + 4 * j + 2], v[16 * h + 4 * j + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = row0 + 8 * i + r_sub;
                float4 x = *reinterpret_cast<const float4*>(rd + 8 * i * kEpiPitch);
                if (row < g.M) {
                    const size_t off = (size_t)row * g.N + col + 16 * h + 4 * c_sub;
                    if (EPI == EPI_RESID_F32) {
                        const float4 r4 = lnr ? ln_on_read(rr[4 * h + i], st[i], gm[h], bt[h]) : rr[4 * h + i];
                        x.x += r4.x; x.y += r4.y; x.z += r4.z; x.w += r4.w;
                    }
                    *reinterpret_cast<float4*>(g.out_f32 + off) = x;
                }
            }
            __syncwarp();
        }
    } else {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float a = v[2 * j], b = v[2 * j + 1];
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __low2float(h2), b - __high2float(h2));
            hi[j] = *reinterpret_cast<const uint32_t*>(&h2);
            lo[j] = *reinterpret_cast<const uint32_t*>(&l2);
        }
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            __nv_bfloat16* out = pass == 0 ? g.out_hi : g.out_lo;
            if (out == nullptr) break;  // warp-uniform
            const uint32_t(&w)[16] = pass == 0 ? hi : lo;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(wr + 16 * j) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = row0 + 8 * i + r_sub;
                const uint4 x = *reinterpret_cast<const uint4*>(rd + 8 * i * kEpiPitch);
                if (row < g.M) *reinterpret_cast<uint4*>(out + (size_t)row * g.N + col + 8 * c_sub) = x;
            }
            __syncwarp();
        }
    }
}

// ---- persistent variant -------------------------------------------------------------------------------------------
// One CTA per SM walks the output tiles (tile = blockIdx.x + i * gridDim.x, N fastest so the CTAs of a wave share
// A row blocks and all of W in L2).  Three roles, three pipelines:
//   * warp 0 / lane 0 -- TMA producer, a kStages-deep ring that runs ahead ACROSS tiles (the next tile's first K blocks
//                        load while the current tile's last MMAs and epilogue run);
//   * warp 1 / lane 0 -- MMA issuer; the fp32 accumulator is double buffered in TMEM (2 x BLOCK_N columns), so tile
//                        t+1's main loop overlaps tile t's epilogue (tmem_full / tmem_empty mbarriers);
//   * warps 2..9      -- epilogue: warp w reads TMEM lane quadrant w % 4 (the hardware's per-warp lane window) and
//                        column half (w - 2) / 4 of the tile.
// The one-tile-per-CTA kernel above pays TMEM allocation, barrier set-up, a cold TMA pipeline and a serial epilogue per
// tile; here they are paid once per SM.  192 KB of operand ring per SM (6 x 32 KB or 4 x 48 KB).
#ifndef ASP_GEMM_PAIR_FFN2
#define ASP_GEMM_PAIR_FFN2 1
#endif
#ifndef ASP_GEMM_RESID_AHEAD
#define ASP_GEMM_RESID_AHEAD 1
#endif
constexpr int kPersistEpiWarps = 8;
constexpr int kPersistThreads = 64 + 32 * kPersistEpiWarps;

template <int BLOCK_N>
struct PersistSmem {
    static constexpr int kStages = BLOCK_N == 256 ? 4 : BLOCK_N == 192 ? 5 : 6;  // 4 x 48 KB, 5 x 40 KB, 6 x 32 KB
    static constexpr int kTmemCols = BLOCK_N == 128 ? 256 : 512;                 // two accumulators, power of two
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStage = kABytes + kBBytes;
    static constexpr int kBarOff = kStages * kStage;
    static constexpr int kEpiOff = kBarOff + 256;  // per-warp staging buffers of the epilogue
    static constexpr int kTotal = kEpiOff + kPersistEpiWarps * kEpiBytesPerWarp + 1024;
};

template <int BLOCK_N, int EPI, int CL>
__global__ void __launch_bounds__(kPersistThreads, 1)
gemm_tn_persistent_kernel(const __grid_constant__ CUtensorMap ta_hi, const __grid_constant__ CUtensorMap ta_lo,
                          const __grid_constant__ CUtensorMap tb_hi, const __grid_constant__ CUtensorMap tb_lo,
                          const GemmArgs g) {
    using S = PersistSmem<BLOCK_N>;
    constexpr int kStages = S::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
    uint64_t* empty = full + kStages;
    uint64_t* tmem_full = empty + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CL > 1: a cluster of CL CTAs works on CL vertically adjacent tiles (same W tile, CL different A row blocks); each
    // CTA fetches 1/CL of the W tile and multicasts it to the whole cluster, so W costs one L2 read per cluster.  A
    // stage may only be refilled once EVERY CTA's MMAs have consumed it: the release commit is multicast too.
    const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
    const int group = blockIdx.x / CL, groups = gridDim.x / CL;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
    constexpr int kSliceRows = BLOCK_N / CL;
    const int n_tiles = g.N / BLOCK_N, m_tiles = (g.M + kBlockM - 1) / kBlockM;
    const int tiles = n_tiles * ((m_tiles + CL - 1) / CL);  // per cluster
    const int per_tile = (g.K / kBlockK) * g.nterms;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&ta_hi);
        tma_prefetch_desc(&tb_hi);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CL);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], kPersistEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, S::kTmemCols);
    tc_fence_before_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // every CTA's barriers exist before a peer multicasts into them
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();  // everything above overlapped the previous kernel's tail; from here on its outputs are read

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            int it = 0;
            for (int tile = group; tile < tiles; tile += groups) {
                const int m0 = ((tile / n_tiles) * CL + rank) * kBlockM, n0 = (tile % n_tiles) * BLOCK_N;
                for (int k = 0; k < per_tile; ++k, ++it) {
                    const int s = it % kStages, ph = (it / kStages) & 1;
                    const int kb = k / g.nterms, term = k - kb * g.nterms;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], S::kStage);
                    uint8_t* sa = smem + s * S::kStage;
                    tma_load_2d(sa, term == 2 ? &ta_lo : &ta_hi, &full[s], kb * kBlockK, m0);  // rows past M read as zero
                    if (CL == 1)
                        tma_load_2d(sa + S::kABytes, term == 1 ? &tb_lo : &tb_hi, &full[s], kb * kBlockK, n0);
                    else
                        tma_load_2d_multicast(sa + S::kABytes + rank * (kSliceRows * kBlockK * 2), term == 1 ? &tb_lo : &tb_hi,
                                              &full[s], kb * kBlockK, n0 + rank * kSliceRows, kMask);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer ----------------
            constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N);
            int it = 0, t = 0;
            for (int tile = group; tile < tiles; tile += groups, ++t) {
                const int buf = t & 1;
                mbar_wait(&tmem_empty[buf], ((t >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
                tc_fence_after_sync();
                const uint32_t acc = tmem_base + (uint32_t)(buf * BLOCK_N);
                for (int k = 0; k < per_tile; ++k, ++it) {
                    const int s = it % kStages, ph = (it / kStages) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(smem + s * S::kStage);
                    const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + S::kABytes);
#pragma unroll
                    for (int kk = 0; kk < kBlockK / 16; ++kk)
                        umma_bf16(acc, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0);
                    if (CL == 1)
                        umma_commit(&empty[s]);
                    else
                        umma_commit_multicast(&empty[s], kMask);
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else {
        // ---------------- epilogue warps ----------------
        const int quad = warp & 3, half = (warp - 2) >> 2;
        constexpr int kCols = BLOCK_N / 2;
        uint8_t* stage = smem + S::kEpiOff + (warp - 2) * kEpiBytesPerWarp;
        int t = 0;
        for (int tile = group; tile < tiles; tile += groups, ++t) {
            const int buf = t & 1;
            const int row0 = ((tile / n_tiles) * CL + rank) * kBlockM + quad * 32;
            const int n0 = (tile % n_tiles) * BLOCK_N + half * kCols;
            float2 ln_st[4];  // LayerNorm on read: row statistics of this lane's rows, fetched under the wait for the accumulator
            if (EPI == EPI_RESID_F32 && g.ln_stats) {
#pragma unroll
                for (int i = 0; i < 4; ++i) ln_st[i] = __ldg(g.ln_stats + min(row0 + 8 * i + (lane >> 2), g.M - 1));
            }
            mbar_wait(&tmem_full[buf], (t >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BLOCK_N + half * kCols);
            // two chunks per round, the second one's TMEM load in flight under the first one's stores; the accumulator is
            // handed back to the MMA warp right after its last read (kCols = 64, 96 or 128)
            auto release = [&]() {
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            };
            float va[32];
            if (EPI == EPI_RESID_F32 && ASP_GEMM_RESID_AHEAD) {
                // Residual epilogue: what it waits for is the residual rows (a global load, ~1 us under load), not tensor
                // memory.  The register budget (168) holds ONE accumulator chunk and TWO chunks of residual values: the next
                // chunk's residual loads are in flight while this chunk is added, staged and stored.
                float4 ra[8], rb[8];
                epilogue_load_residual(g, lane, row0, n0, ra);
#pragma unroll
                for (int c = 0; c < kCols / 32; ++c) {
                    tmem_ld32(acc + (uint32_t)(32 * c), va);
                    if (c + 1 == kCols / 32) release();
                    if (c + 1 < kCols / 32) epilogue_load_residual(g, lane, row0, n0 + 32 * (c + 1), (c & 1) ? ra : rb);
                    epilogue_store32_staged<EPI>(g, va, stage, lane, row0, n0 + 32 * c, ln_st, (c & 1) ? rb : ra);
                }
                continue;
            }
            float vb[32];
            tmem_ld32_issue(acc, va);
#pragma unroll 1
            for (int c0 = 0; c0 < kCols; c0 += 64) {
                const bool second = c0 + 32 < kCols;  // warp-uniform
                tmem_ld32_wait(va);
                if (second)
                    tmem_ld32_issue(acc + (uint32_t)(c0 + 32), vb);
                else
                    release();
                epilogue_store32_staged<EPI>(g, va, stage, lane, row0, n0 + c0, ln_st);
                if (second) {
                    tmem_ld32_wait(vb);
                    if (c0 + 64 < kCols)
                        tmem_ld32_issue(acc + (uint32_t)(c0 + 64), va);
                    else
                        release();
                    epilogue_store32_staged<EPI>(g, vb, stage, lane, row0, n0 + c0 + 32, ln_st);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer can still write its shared memory or barriers
    if (warp == 1) tmem_dealloc(tmem_base, S::kTmemCols);
}

// ---- CTA-pair variant (cta_group::2) ---------------------------------------------------------------------------------
// Same roles and pipelines as the persistent kernel, but two CTAs on the two SMs of a TPC share one 256 x BLOCK_N tile:
// each loads its own 128 rows of A and only HALF of the W tile (BLOCK_N / 2 rows), the leader's MMA thread issues
// M = 256 tcgen05.mma.cta_group::2 instructions that read both halves, and each CTA's TMEM receives its own 128
// accumulator rows.  What this buys is bytes INTO each SM per MMA: 32 KB instead of 48 KB per K block at BLOCK_N = 256,
// 24 KB instead of 32 KB at BLOCK_N = 128 -- the quantity that bounds the single-CTA kernel.
template <int BLOCK_N>
struct PairSmem {
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;
    static constexpr int kStage = kABytes + kBBytes;
    static constexpr int kStages = (192 * 1024) / kStage;
    static constexpr int kBarOff = kStages * kStage;
    static constexpr int kEpiOff = kBarOff + 256;
    static constexpr int kTotal = kEpiOff + kPersistEpiWarps * kEpiBytesPerWarp + 1024;
};

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kPersistThreads, 1)
gemm_tn_pair_kernel(const __grid_constant__ CUtensorMap ta_hi, const __grid_constant__ CUtensorMap ta_lo,
                    const __grid_constant__ CUtensorMap tb_hi, const __grid_constant__ CUtensorMap tb_lo, const GemmArgs g) {
    using S = PairSmem<BLOCK_N>;
    constexpr int kStages = S::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOff);  // used in the leader only
    uint64_t* empty = full + kStages;
    uint64_t* tmem_full = empty + kStages;
    uint64_t* tmem_empty = tmem_full + 2;  // used in the leader only
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const bool leader = rank == 0;
    const int group = blockIdx.x / 2, groups = gridDim.x / 2;
    const int n_tiles = g.N / BLOCK_N, m_tiles = (g.M + kBlockM - 1) / kBlockM;
    const int tiles = n_tiles * ((m_tiles + 1) / 2);  // per pair
    const int per_tile = (g.K / kBlockK) * g.nterms;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&ta_hi);
        tma_prefetch_desc(&tb_hi);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 2 * kPersistEpiWarps);
        }
        fence_barrier_init();
    }
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers exist before the pair-scoped allocation and any remote arrival
    if (warp == 1) tmem_alloc_pair(tmem_slot, 2 * BLOCK_N);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer (both CTAs; bytes are counted on the leader's barrier) ----------------
            int it = 0;
            for (int tile = group; tile < tiles; tile += groups) {
                const int m0 = ((tile / n_tiles) * 2 + rank) * kBlockM;
                const int n0 = (tile % n_tiles) * BLOCK_N + rank * (BLOCK_N / 2);
                for (int k = 0; k < per_tile; ++k, ++it) {
                    const int s = it % kStages, ph = (it / kStages) & 1;
                    const int kb = k / g.nterms, term = k - kb * g.nterms;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (leader) mbar_arrive_expect_tx(&full[s], 2 * S::kStage);
                    uint8_t* sa = smem + s * S::kStage;
                    const uint32_t leader_full = cluster_map_addr(&full[s], 0);
                    tma_load_2d_pair(sa, term == 2 ? &ta_lo : &ta_hi, leader_full, kb * kBlockK, m0);
                    tma_load_2d_pair(sa + S::kABytes, term == 1 ? &tb_lo : &tb_hi, leader_full, kb * kBlockK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // ---------------- MMA issuer (leader only, for both CTAs) ----------------
            constexpr uint32_t idesc = umma_idesc_bf16(2 * kBlockM, BLOCK_N);
            int it = 0, t = 0;
            for (int tile = group; tile < tiles; tile += groups, ++t) {
                const int buf = t & 1;
                mbar_wait(&tmem_empty[buf], ((t >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this accumulator
                tc_fence_after_sync();
                const uint32_t acc = tmem_base + (uint32_t)(buf * BLOCK_N);
                for (int k = 0; k < per_tile; ++k, ++it) {
                    const int s = it % kStages, ph = (it / kStages) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(smem + s * S::kStage);
                    const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + S::kABytes);
#pragma unroll
                    for (int kk = 0; kk < kBlockK / 16; ++kk)
                        umma_bf16_pair(acc, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0);
                    umma_commit_pair(&empty[s]);
                }
                umma_commit_pair(&tmem_full[buf]);
            }
        }
    } else {
        // ---------------- epilogue warps (both CTAs, own 128 rows) ----------------
        const int quad = warp & 3, half = (warp - 2) >> 2;
        constexpr int kCols = BLOCK_N / 2;
        uint8_t* stage = smem + S::kEpiOff + (warp - 2) * kEpiBytesPerWarp;
        int t = 0;
        for (int tile = group; tile < tiles; tile += groups, ++t) {
            const int buf = t & 1;
            const int row0 = ((tile / n_tiles) * 2 + rank) * kBlockM + quad * 32;
            const int n0 = (tile % n_tiles) * BLOCK_N + half * kCols;
            mbar_wait(&tmem_full[buf], (t >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BLOCK_N + half * kCols);
            float va[32];
            if (EPI == EPI_RESID_F32 && ASP_GEMM_RESID_AHEAD) {  // as in the one-CTA kernel: the double buffer is on the residual loads
                float4 ra[8], rb[8];
                epilogue_load_residual(g, lane, row0, n0, ra);
#pragma unroll
                for (int c = 0; c < kCols / 32; ++c) {
                    tmem_ld32(acc + (uint32_t)(32 * c), va);
                    if (c + 1 == kCols / 32) {
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(&tmem_empty[buf], 0);
                    }
                    if (c + 1 < kCols / 32) epilogue_load_residual(g, lane, row0, n0 + 32 * (c + 1), (c & 1) ? ra : rb);
                    epilogue_store32_staged<EPI>(g, va, stage, lane, row0, n0 + 32 * c, nullptr, (c & 1) ? rb : ra);
                }
                continue;
            }
            float vb[32];
            tmem_ld32_issue(acc, va);
#pragma unroll 1
            for (int c0 = 0; c0 < kCols; c0 += 64) {
                tmem_ld32_wait(va);
                tmem_ld32_issue(acc + (uint32_t)(c0 + 32), vb);
                epilogue_store32_staged<EPI>(g, va, stage, lane, row0, n0 + c0);
                tmem_ld32_wait(vb);
                if (c0 + 64 < kCols) {
                    tmem_ld32_issue(acc + (uint32_t)(c0 + 64), va);
                } else {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(&tmem_empty[buf], 0);  // the leader's barrier, from either CTA
                }
                epilogue_store32_staged<EPI>(g, vb, stage, lane, row0, n0 + c0 + 32);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other can still touch its memory or barriers
    if (warp == 1) tmem_dealloc_pair(tmem_base, 2 * BLOCK_N);
}

// ---- host side ----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Row-major bf16 matrix [rows, cols] (cols contiguous) -> tensor map with a [box_rows x 64] box and 128B swizzle.
// Encoded maps are kept per thread, keyed by (pointer, shape, box): a forward pass asks for the same ~100 maps (every
// weight, the activation workspaces) on every call, and cuTensorMapEncodeTiled costs about a microsecond each -- 288
// encodes per forward were a fifth of the host time of a small batch.
struct TmapKey {
    const void* ptr;
    uint64_t rows, cols;
    uint32_t box_rows;
    bool operator==(const TmapKey& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && box_rows == o.box_rows; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h ^= k.rows * 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        h ^= k.cols * 0xc2b2ae3d27d4eb4full + (h << 6) + (h >> 2);
        return h ^ (k.box_rows * 0x165667b19e3779f9ull);
    }
};

int make_tmap_bf16(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    const TmapKey key{ptr, rows, cols, box_rows};
    auto it = cache.find(key);
    if (it != cache.end()) {
        *m = it->second;
        return ASP_OK;
    }
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return ASP_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with code %d (ptr %p rows %llu cols %llu)", (int)r, ptr,
                  (unsigned long long)rows, (unsigned long long)cols);
        return ASP_ERR_CUDA;
    }
    if (cache.size() >= 4096) cache.clear();  // bounded: a long-lived process that keeps reallocating its buffers
    cache.emplace(key, *m);
    return ASP_OK;
}

// Same matrix, [box_rows x 32] box (64-byte rows) with the 64B swizzle -- the all-pairs kernel's K block.
int make_tmap_bf16_k32(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return ASP_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {32u, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (k32) failed with code %d (ptr %p rows %llu cols %llu)", (int)r, ptr,
                  (unsigned long long)rows, (unsigned long long)cols);
        return ASP_ERR_CUDA;
    }
    return ASP_OK;
}

template <int BLOCK_N, int EPI>
static int launch_gemm(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi, const CUtensorMap& tb_lo,
                       const GemmArgs& g, cudaStream_t stream) {
    using S = GemmSmem<BLOCK_N>;
    static thread_local int attr_dev = -1;
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        ASP_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
        attr_dev = dev;
    }
    dim3 grid(g.N / BLOCK_N, (g.M + kBlockM - 1) / kBlockM);
    gemm_tn_kernel<BLOCK_N, EPI><<<grid, 128, S::kTotal, stream>>>(ta_hi, ta_lo, tb_hi, tb_lo, g);
    ASP_LAUNCH_CHECK("gemm_tn_kernel");
    return ASP_OK;
}

template <int BLOCK_N, int EPI, int CL>
static int launch_gemm_persistent(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi,
                                  const CUtensorMap& tb_lo, const GemmArgs& g, cudaStream_t stream) {
    using S = PersistSmem<BLOCK_N>;
    static thread_local int attr_dev = -1, groups_max = 0;
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    auto kernel = gemm_tn_persistent_kernel<BLOCK_N, EPI, CL>;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(kPersistThreads);
    cfg.dynamicSmemBytes = S::kTotal;
    cfg.stream = stream;
    int n_attr = 0;
    if (CL > 1) ++n_attr;
    if (g_pdl) {
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    if (attr_dev != dev) {
        int sms = 0;
        ASP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
        ASP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        groups_max = sms / CL;
        if (CL > 1) {  // clusters cannot straddle a GPC: ask how many fit at once
            cfg.gridDim = dim3(sms / CL * CL);
            int n = 0;
            ASP_CUDA(cudaOccupancyMaxActiveClusters(&n, kernel, &cfg));
            ASP_REQUIRE(n >= 1, "gemm: a cluster of %d CTAs does not fit on this device", CL);
            groups_max = std::min(groups_max, n);
        }
        attr_dev = dev;
    }
    const int m_tiles = (g.M + kBlockM - 1) / kBlockM;
    const int tiles = (g.N / BLOCK_N) * ((m_tiles + CL - 1) / CL);
    cfg.gridDim = dim3(CL * std::min(tiles, groups_max));
    ASP_CUDA(cudaLaunchKernelEx(&cfg, kernel, ta_hi, ta_lo, tb_hi, tb_lo, g));
    ASP_LAUNCH_CHECK("gemm_tn_persistent_kernel");
    return ASP_OK;
}

template <int BLOCK_N, int CL>
static int dispatch_persistent(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi,
                               const CUtensorMap& tb_lo, const GemmArgs& g, int epilogue, cudaStream_t stream) {
    switch (epilogue) {
        case EPI_BF16: return launch_gemm_persistent<BLOCK_N, EPI_BF16, CL>(ta_hi, ta_lo, tb_hi, tb_lo, g, stream);
        case EPI_GELU_BF16: return launch_gemm_persistent<BLOCK_N, EPI_GELU_BF16, CL>(ta_hi, ta_lo, tb_hi, tb_lo, g, stream);
        case EPI_RESID_F32: return launch_gemm_persistent<BLOCK_N, EPI_RESID_F32, CL>(ta_hi, ta_lo, tb_hi, tb_lo, g, stream);
        case EPI_F32: return launch_gemm_persistent<BLOCK_N, EPI_F32, CL>(ta_hi, ta_lo, tb_hi, tb_lo, g, stream);
    }
    set_error("gemm: unknown epilogue %d", epilogue);
    return ASP_ERR_INVALID;
}

template <int BLOCK_N, int EPI>
static int launch_gemm_pair(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi,
                            const CUtensorMap& tb_lo, const GemmArgs& g, cudaStream_t stream) {
    using S = PairSmem<BLOCK_N>;
    static thread_local int attr_dev = -1, pairs_max = 0;
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    auto kernel = gemm_tn_pair_kernel<BLOCK_N, EPI>;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    int n_attr = 1;
    if (g_pdl) {
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        ++n_attr;
    }
    cfg.blockDim = dim3(kPersistThreads);
    cfg.dynamicSmemBytes = S::kTotal;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    if (attr_dev != dev) {
        int sms = 0, n = 0;
        ASP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
        ASP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        cfg.gridDim = dim3(sms / 2 * 2);
        ASP_CUDA(cudaOccupancyMaxActiveClusters(&n, kernel, &cfg));
        ASP_REQUIRE(n >= 1, "gemm: a CTA pair does not fit on this device");
        pairs_max = std::min(sms / 2, n);
        attr_dev = dev;
    }
    const int m_tiles = (g.M + kBlockM - 1) / kBlockM;
    const int tiles = (g.N / BLOCK_N) * ((m_tiles + 1) / 2);
    cfg.gridDim = dim3(2 * std::min(tiles, pairs_max));
    ASP_CUDA(cudaLaunchKernelEx(&cfg, kernel, ta_hi, ta_lo, tb_hi, tb_lo, g));
    ASP_LAUNCH_CHECK("gemm_tn_pair_kernel");
    return ASP_OK;
}

template <int BLOCK_N>
static int dispatch_pair(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi, const CUtensorMap& tb_lo,
                         const GemmArgs& g, int epilogue, cudaStream_t stream) {
    switch (epilogue) {
        case EPI_BF16: return launch_gemm_pair<BLOCK_N, EPI_BF16>(ta_hi, ta_lo, tb_hi, tb_lo, g, stream);
        case EPI_GELU_BF16: return launch_gemm_pair<BLOCK_N, EPI_GELU_BF16>(ta_hi, ta_lo, tb_hi, tb_lo, g, stream);
        case EPI_RESID_F32: return launch_gemm_pair<BLOCK_N, EPI_RESID_F32>(ta_hi, ta_lo, tb_hi, tb_lo, g, stream);
        case EPI_F32: return launch_gemm_pair<BLOCK_N, EPI_F32>(ta_hi, ta_lo, tb_hi, tb_lo, g, stream);
    }
    set_error("gemm: unknown epilogue %d", epilogue);
    return ASP_ERR_INVALID;
}

// asp_set_option("gemm_kernel"): 0 one tile per CTA, 1 persistent with 128-wide tiles, 2 / 4 persistent with 256- / 192-wide
// tiles where N allows, 3 (default) persistent, tile width picked per shape.
// asp_set_option("gemm_pair"): 0 off, 1 CTA pairs (cta_group::2) with 128-wide pair tiles, 2 with 256-wide ones where N
// allows.
// asp_set_option("gemm_cluster"): 1, 2 or 4 CTAs per cluster sharing each W tile by TMA multicast (persistent kernels).
int g_gemm_kernel = 3;
constexpr int kCost192 = 520;  // (FFN2 - attention-output time at 192 columns) / (36 K blocks x 2 rounds)
int g_gemm_cluster = 1;
// -1 (default): CTA pairs with 256-wide pair tiles for the GEMMs of >= 16384 rows whose main loop is what is left to gain -- the
// bf16-output ones (QKV, FFN1) and the K >= 2048 residual one (FFN2); attn-out (K = 768, epilogue-bound) stays on one CTA.  The 1-CTA MMA reads
// A (4 KB) and B (8 KB) from shared memory for every M 128 x N 256 x K 16 instruction -- 96 B/clk of a 128 B/clk pipe, which is
// the ~75 % "practical" tensor rate; a pair splits B between two SMs (64 B/clk each).  At 8192 rows the pair kernel gains nothing
// (round 1), at 32768 rows the forward is 6 % faster (profiles/r02_4o_gemm_pair_32k.txt).
// 0 = never, 1 / 2 = always (128- / 256-wide pair tiles).
int g_gemm_pair = -1;
int g_pdl = 1;

int gemm_bf16_tn(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                 const float* residual, int M, int N, int K, int epilogue, void* out_hi, void* out_lo, float* out_f32,
                 cudaStream_t stream, const LnOnRead* ln) {
    constexpr int BN = 128;
    const float2* ln_stats = ln ? reinterpret_cast<const float2*>(ln->stats) : nullptr;
    const float *ln_g = ln ? ln->gamma : nullptr, *ln_b = ln ? ln->beta : nullptr;
    if (ln) ASP_REQUIRE(epilogue == EPI_RESID_F32 && ln->stats && ln->gamma && ln->beta, "gemm: LayerNorm-on-read needs the residual epilogue");
    ASP_REQUIRE(a_hi && w_hi, "gemm: NULL operand");
    ASP_REQUIRE(M >= 1 && N >= BN && (N % BN) == 0 && K >= kBlockK && (K % kBlockK) == 0,
                "gemm: need N %% %d == 0 and K %% %d == 0 (got M=%d N=%d K=%d)", BN, kBlockK, M, N, K);
    ASP_REQUIRE((a_lo == nullptr) == (w_lo == nullptr), "gemm: give both lo operands (bf16x3) or neither (bf16)");
    ASP_REQUIRE(aligned16(a_hi) && aligned16(w_hi) && aligned16(a_lo) && aligned16(w_lo), "gemm: operands must be 16-byte aligned");
    if (epilogue == EPI_RESID_F32) ASP_REQUIRE(residual && out_f32, "gemm: residual epilogue needs residual and out_f32");
    if (epilogue == EPI_F32) ASP_REQUIRE(out_f32, "gemm: fp32 epilogue needs out_f32");
    if (epilogue == EPI_BF16 || epilogue == EPI_GELU_BF16) ASP_REQUIRE(out_hi, "gemm: bf16 epilogue needs out_hi");
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    int rc;
    // Tile width of the persistent kernel.  Measured main-loop cost per 64-wide K block (tools/gemm_bench.py, slope of
    // time over K): 547 clk at 128 columns, 520 at 192, 690 at 256 -- the 256-wide loop runs at the tensor pipe's
    // practical rate (1.7 PFLOP/s, what cuBLAS reaches on this part; CTA pairs add nothing to it), the 128-wide one at
    // 62 % of that.  Pick the width that minimises rounds x cost + the last tile's exposed epilogue.
    int bn = BN;
    if (g_gemm_kernel == 2 && (N % 256) == 0) bn = 256;
    if (g_gemm_kernel == 4 && (N % 192) == 0) bn = 192;
    if (g_gemm_kernel == 3) {
        int dev = 0, sms = 148;
        ASP_CUDA(cudaGetDevice(&dev));
        ASP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const long m_tiles = (M + kBlockM - 1) / kBlockM, k_iters = (long)(K / kBlockK) * (a_lo ? 3 : 1);
        const bool f32_out = epilogue == EPI_RESID_F32 || epilogue == EPI_F32;
        long best = -1;
        const int widths[3] = {128, 192, 256}, cost[3] = {547, kCost192, 690};
        for (int i = 0; i < 3; ++i) {
            if (N % widths[i]) continue;
            const long tiles = m_tiles * (N / widths[i]), rounds = (tiles + sms - 1) / sms;
            const long c = rounds * k_iters * cost[i] + (long)widths[i] * (f32_out ? 24 : 12);
            if (best < 0 || c < best) {
                best = c;
                bn = widths[i];
            }
        }
    }
    const bool pair_auto = g_gemm_pair < 0 && g_gemm_kernel == 3 && g_gemm_cluster == 1 && (N % 256) == 0 && M >= 16384 &&
                           (epilogue == EPI_BF16 || epilogue == EPI_GELU_BF16 || (ASP_GEMM_PAIR_FFN2 && epilogue == EPI_RESID_F32 && K >= 2048));
    if ((g_gemm_pair > 0 || pair_auto) && g_gemm_kernel >= 1) {
        const int pbn = ((g_gemm_pair == 2 || pair_auto) && (N % 256) == 0) ? 256 : 128;
        if ((rc = make_tmap_bf16(&ta_hi, a_hi, M, K, kBlockM))) return rc;
        if ((rc = make_tmap_bf16(&tb_hi, w_hi, N, K, pbn / 2))) return rc;
        if ((rc = make_tmap_bf16(&ta_lo, a_lo ? a_lo : a_hi, M, K, kBlockM))) return rc;
        if ((rc = make_tmap_bf16(&tb_lo, w_lo ? w_lo : w_hi, N, K, pbn / 2))) return rc;
        GemmArgs gp{M, N, K, a_lo ? 3 : 1, bias, residual, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, out_f32, ln_stats, ln_g, ln_b};
        return pbn == 256 ? dispatch_pair<256>(ta_hi, ta_lo, tb_hi, tb_lo, gp, epilogue, stream)
                          : dispatch_pair<128>(ta_hi, ta_lo, tb_hi, tb_lo, gp, epilogue, stream);
    }
    if ((rc = make_tmap_bf16(&ta_hi, a_hi, M, K, kBlockM))) return rc;
    const int w_box = (g_gemm_kernel >= 1 && bn != 192) ? bn / g_gemm_cluster : bn;  // each CTA of a cluster fetches its slice of the W tile
    if ((rc = make_tmap_bf16(&tb_hi, w_hi, N, K, w_box))) return rc;
    if ((rc = make_tmap_bf16(&ta_lo, a_lo ? a_lo : a_hi, M, K, kBlockM))) return rc;
    if ((rc = make_tmap_bf16(&tb_lo, w_lo ? w_lo : w_hi, N, K, w_box))) return rc;
    GemmArgs g{M, N, K, a_lo ? 3 : 1, bias, residual, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, out_f32, ln_stats, ln_g, ln_b};
    if (g_gemm_kernel >= 1) {
        const int cl = bn == 192 ? 1 : g_gemm_cluster;
        if (bn == 192) return dispatch_persistent<192, 1>(ta_hi, ta_lo, tb_hi, tb_lo, g, epilogue, stream);
        if (bn == 256) {
            if (cl == 4) return dispatch_persistent<256, 4>(ta_hi, ta_lo, tb_hi, tb_lo, g, epilogue, stream);
            if (cl == 2) return dispatch_persistent<256, 2>(ta_hi, ta_lo, tb_hi, tb_lo, g, epilogue, stream);
            return dispatch_persistent<256, 1>(ta_hi, ta_lo, tb_hi, tb_lo, g, epilogue, stream);
        }
        if (cl == 4) return dispatch_persistent<128, 4>(ta_hi, ta_lo, tb_hi, tb_lo, g, epilogue, stream);
        if (cl == 2) return dispatch_persistent<128, 2>(ta_hi, ta_lo, tb_hi, tb_lo, g, epilogue, stream);
        return dispatch_persistent<128, 1>(ta_hi, ta_lo, tb_hi, tb_lo, g, epilogue, stream);
    }
    switch (epilogue) {
        case EPI_BF16: return launch_gemm<BN, EPI_BF16>(ta_hi, ta_lo, tb_hi, tb_lo, g, stream);
        case EPI_GELU_BF16: return launch_gemm<BN, EPI_GELU_BF16>(ta_hi, ta_lo, tb_hi, tb_lo, g, stream);
        case EPI_RESID_F32: return launch_gemm<BN, EPI_RESID_F32>(ta_hi, ta_lo, tb_hi, tb_lo, g, stream);
        case EPI_F32: return launch_gemm<BN, EPI_F32>(ta_hi, ta_lo, tb_hi, tb_lo, g, stream);
    }
    set_error("gemm: unknown epilogue %d", epilogue);
    return ASP_ERR_INVALID;
}

}  // namespace asp

extern "C" int asp_gemm_bf16_tn(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                                const float* residual, int M, int N, int K, int epilogue, void* out_hi, void* out_lo,
                                float* out_f32, asp_stream_t stream) {
    return asp::gemm_bf16_tn(a_hi, a_lo, w_hi, w_lo, bias, residual, M, N, K, epilogue, out_hi, out_lo, out_f32,
                             (cudaStream_t)stream, nullptr);
}
