// K0 self-attention on tcgen05 (plain bf16 mode, L <= 256): softmax(Q K^T / 8 + key-padding mask) V per (document, head).
//
// Replaces BertSelfAttention as reached from examples/ex_aspire_consent.py:72 -- the same contraction attention.cu runs
// on mma.sync fragments (38.5 us per layer at 8192 tokens, 17 % of the encoder's time for 5 % of its FLOPs).  One CTA =
// 128 query rows of one (document, head), all <= 256 keys at once, no online softmax:
//   TMA      Q [128 x 64], K [256 x 64], V [256 x 64] boxes straight out of the fused QKV activation [tokens, 3 * hidden]
//            (128-byte rows, 128B swizzle); rows past the batch are zero-filled, keys past the document are masked
//   MMA 1    S[128 x 256] = Q K^T: 4 tcgen05.mma (M 128, N 256, K 16), fp32 in 256 TMEM columns
//   V        is the B operand of P V as it arrived: [key][d] rows of 128 bytes with the 128B swizzle are the canonical
//            MN-major SW128 layout (8-key groups 1 KB apart), selected by bit 16 of the instruction descriptor.  (The first
//            version transposed V into a K-major V^T with ldmatrix.trans / stmatrix while MMA 1 ran -- kept as attn_tc=1.)
//   softmax  one thread per query row: row max, then p = 2^((s - max) / 8 * log2 e) straight from TMEM (tcgen05.ld, 32
//            columns at a time), row sum in fp32, P as bf16 into shared memory in the A-operand layout (it reuses the
//            Q / K / V staging, which MMA 1 and the transposition are done with)
//   MMA 2    O[128 x 64] = P V: 16 tcgen05.mma (M 128, N 64, K 16) into TMEM columns 0-63 (S is consumed by then)
//   epilogue O / row sum -> bf16 context rows
// 112 KB of shared memory and 256 TMEM columns per CTA: two CTAs per SM, one's softmax under the other's loads and MMAs.
#include <algorithm>
#include <cuda_bf16.h>
#include "../common.cuh"
#include "tc05.cuh"

namespace asp {

using namespace tc;

int make_tmap_bf16(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows);

constexpr int kAtQ = 128, kAtKeys = 256, kAtD = 64;
constexpr int kAtQBytes = kAtQ * 128, kAtKBytes = kAtKeys * 128;              // 16 KB, 32 KB
constexpr int kAtOffK = kAtQBytes, kAtOffV = kAtOffK + kAtKBytes, kAtOffVt = kAtOffV + kAtKBytes;  // 16, 48, 80 KB
constexpr int kAtVtBlock = kAtD * 128;                                        // one 64-key K block of V^T: 8 KB
constexpr int kAtPBlock = kAtQ * 128;                                         // one 64-key K block of P: 16 KB (P at offset 0)
constexpr int kAtOffBar = kAtOffVt + 4 * kAtVtBlock;                          // 3 mbarriers + the TMEM address
constexpr int kAtSmem = kAtOffBar + 64;  // 112 KB + 64 B: with the 1 KB the system keeps per CTA, two CTAs fit in 227 KB.
                                         // No static shared memory and no alignment slack: the dynamic window of a kernel
                                         // without static shared memory starts 1024-byte aligned (checked at run time).

__device__ __forceinline__ void at_ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void at_stmatrix_x4(uint32_t addr, const uint32_t (&r)[4]) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3])
                 : "memory");
}

constexpr int kAtThreads = 256;
constexpr float kAtSafeExp = 100.f;  // largest exponent of 2 a probability may carry before the exact path takes over

// one 32-key piece of a row: p = 2^((s - m) c) into the P tile as bf16; returns the sum of the probabilities and tracks
// the largest exponent.  nvalid = keys of the piece inside the document: 32 (no per-element test), 1..31, or 0 (zeros).
// TM: the piece goes to 16 TMEM columns at p_tmem (A operand of P V read from tensor memory) instead of shared memory.
template <bool TM = false>
__device__ __forceinline__ float at_softmax_piece(const float (&v)[32], int c0, int nvalid, float sc, float off, uint8_t* smem, int r,
                                                  float& xmax, uint32_t p_tmem = 0u) {
    uint32_t pk[16];
    float sum = 0.f;
    if (nvalid >= 32) {  // warp-uniform
        float s1 = 0.f;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
            const float x0 = fmaf(v[e], sc, -off), x1 = fmaf(v[e + 1], sc, -off);
            xmax = fmaxf(xmax, fmaxf(x0, x1));
            const float p0 = ex2(x0), p1 = ex2(x1);
            sum += p0;
            s1 += p1;
            const __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
            pk[e >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        }
        sum += s1;
    } else if (nvalid > 0) {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
            const float x0 = fmaf(v[e], sc, -off), x1 = fmaf(v[e + 1], sc, -off);
            const bool in0 = e < nvalid, in1 = e + 1 < nvalid;
            xmax = fmaxf(xmax, fmaxf(in0 ? x0 : -INFINITY, in1 ? x1 : -INFINITY));
            const float p0 = in0 ? ex2(x0) : 0.f, p1 = in1 ? ex2(x1) : 0.f;
            sum += p0 + p1;
            const __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
            pk[e >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        }
    } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = 0u;
    }
    if (TM) {
        tmem_st16(p_tmem, pk);
        return sum;
    }
    uint8_t* prow = smem + (c0 >> 6) * kAtPBlock + (r >> 3) * 1024 + (r & 7) * 128;
    const int ch0 = (c0 & 63) >> 3;  // first 16-byte chunk (8 keys) of this 32-key piece inside its 64-key block
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(prow + ((((ch0 + q) ^ r) & 7) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    return sum;
}

// VMN: V is consumed as an MN-major B operand straight from its TMA tile ([key][d], 128-byte rows, 128B swizzle: the canonical
// MN-major SW128 layout with 8-key groups 1024 bytes apart), so the V -> V^T pass does not exist.  The tile then lands where
// V^T would have been (P overwrites Q, K and the unused gap).
// PT (with VMN): P never touches shared memory.  Each thread overwrites the part of S it has already consumed with its
// probabilities as packed bf16 pairs (half h of row r: S columns 128h..128h+127 -> P columns 128h..128h+63, piece by
// piece behind the read position), MMA 2 takes A from tensor memory and writes O to columns 64-127 (consumed S of half 0).
// The rare exact path needs S again: MMA 1 is simply re-issued, Q and K are still in shared memory.
template <bool VMN, bool PT = false>
__global__ void __launch_bounds__(kAtThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tkv,
                    const int32_t* __restrict__ seq_lens, int L, int H, __nv_bfloat16* __restrict__ ctx) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (smem_u32(smem) & 1023u) __trap();  // the swizzled tiles need 1024-byte alignment
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAtOffBar);
    uint64_t &bar_load = bars[0], &bar_s = bars[1], &bar_o = bars[2];
    uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 3);
    // [2 halves][128 rows] row maxima, then row sums: in the part of the V staging that P does not cover (free once V^T exists)
    float* xch = reinterpret_cast<float*>(smem + 4 * kAtPBlock);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kAtQ, head = blockIdx.y, b = blockIdx.z;
    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tq);
        tma_prefetch_desc(&tkv);
        mbar_init(&bar_load, 1);
        mbar_init(&bar_s, 1);
        mbar_init(&bar_o, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t sbase = smem_u32(smem);
    const int nkeys = min(max(seq_lens[b], 1), min(L, kAtKeys));

    if (threadIdx.x == 0) {
        pdl_wait();  // the QKV projection has written its output
        mbar_arrive_expect_tx(&bar_load, kAtQBytes + 2 * kAtKBytes);
        tma_load_2d(smem, &tq, &bar_load, head * kAtD, b * L + q0);
        tma_load_2d(smem + kAtOffK, &tkv, &bar_load, H + head * kAtD, b * L);
        tma_load_2d(smem + (VMN ? kAtOffVt : kAtOffV), &tkv, &bar_load, 2 * H + head * kAtD, b * L);
    }
    mbar_wait(&bar_load, 0);
    auto issue_s = [&]() {
        tc_fence_after_sync();
        constexpr uint32_t idesc = umma_idesc_bf16(kAtQ, kAtKeys);
        const uint64_t qd = umma_desc_sw128(sbase), kd = umma_desc_sw128(sbase + kAtOffK);
#pragma unroll
        for (int kk = 0; kk < kAtD / 16; ++kk) umma_bf16(tmem_base, qd + 2 * kk, kd + 2 * kk, idesc, kk != 0);
        umma_commit(&bar_s);
    };
    if (threadIdx.x == 0) issue_s();
    __syncwarp();
    // ---- V -> V^T while MMA 1 runs: 8 x 8 blocks, four at a time (same 8 keys, four 8-wide slices of d) ----
    if (!VMN) {
        const int mtx = lane >> 3, i = lane & 7;  // ldmatrix / stmatrix: lanes 8m..8m+7 address the rows of matrix m
#pragma unroll 4
        for (int t = warp; t < 2 * (kAtKeys / 8); t += kAtThreads / 32) {
            const int k0 = (t >> 1) * 8, d0 = (t & 1) * 32 + mtx * 8;
            const int key = k0 + i;
            const uint32_t src = sbase + kAtOffV + key * 128 + ((((d0 >> 3) ^ key) & 7) << 4);
            uint32_t r[4];
            at_ldmatrix_x4_trans(r, src);
            const int drow = d0 + i;
            const uint32_t dst = sbase + kAtOffVt + (k0 >> 6) * kAtVtBlock + drow * 128 + (((((k0 & 63) >> 3)) ^ drow) & 7) * 16;
            at_stmatrix_x4(dst, r);
        }
    }
    if (!PT) __syncthreads();  // V^T complete, nobody reads the V staging any more: P may overwrite Q / K / V
    // ---- softmax: two threads per query row (warps w and w + 4 share TMEM lane quadrant w), 128 keys each.  S is read
    //      from TMEM ONCE (64 B per clk per SM: a second pass over the 128 KB tile would cost as much as everything else):
    //      the shift m is the maximum over the FIRST 32 keys of each half, not the row maximum.  Any m within 2^100 of the
    //      row maximum gives the same normalised probabilities (bf16 and fp32 share the exponent range, the row sum is
    //      fp32); a row whose exponents exceed 100 anyway sends the tile through the exact two-pass form. ----
    mbar_wait(&bar_s, 0);
    tc_fence_after_sync();
    const int half = warp >> 2, r = (warp & 3) * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int cbeg = half * 128;
    const float sc = 0.125f * kLog2e;
    float v[32];
    float mloc = -INFINITY;
    if (cbeg < nkeys) {  // warp-uniform
        tmem_ld32(trow + cbeg, v);
#pragma unroll
        for (int e = 0; e < 32; ++e)
            if (cbeg + e < nkeys) mloc = fmaxf(mloc, v[e]);
    }
    xch[half * 128 + r] = mloc;
    __syncthreads();
    float m = fmaxf(xch[r], xch[128 + r]);
    float off = m * sc, sum = 0.f, xmax = -INFINITY;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + 128; c0 += 32) {
        if (c0 != cbeg && c0 < nkeys) tmem_ld32(trow + c0, v);   // (the first piece is still in registers)
        sum += at_softmax_piece<PT>(v, c0, min(max(nkeys - c0, 0), 32), sc, off, smem, r, xmax,  // beyond the document: zeros
                                    trow + (uint32_t)(cbeg + ((c0 - cbeg) >> 1)));
    }
    if (PT) {
        tmem_wait_st();
        tc_fence_before_sync();
    }
    if (__syncthreads_or(xmax > kAtSafeExp)) {
        // exact form (rare): row maximum over all keys first, then the probabilities again
        if (PT) {  // S was overwritten by P: compute it again
            if (threadIdx.x == 0) issue_s();
            __syncwarp();
            mbar_wait(&bar_s, 1);
            tc_fence_after_sync();
        }
        float me = -INFINITY;
#pragma unroll 1
        for (int c0 = cbeg; c0 < cbeg + 128 && c0 < nkeys; c0 += 32) {
            tmem_ld32(trow + c0, v);
#pragma unroll
            for (int e = 0; e < 32; ++e)
                if (c0 + e < nkeys) me = fmaxf(me, v[e]);
        }
        __syncthreads();
        xch[half * 128 + r] = me;
        __syncthreads();
        m = fmaxf(xch[r], xch[128 + r]);
        off = m * sc;
        sum = 0.f;
#pragma unroll 1
        for (int c0 = cbeg; c0 < cbeg + 128; c0 += 32) {
            if (c0 < nkeys) tmem_ld32(trow + c0, v);
            sum += at_softmax_piece<PT>(v, c0, min(max(nkeys - c0, 0), 32), sc, off, smem, r, xmax,
                                        trow + (uint32_t)(cbeg + ((c0 - cbeg) >> 1)));
        }
        if (PT) tmem_wait_st();
    }
    __syncthreads();
    xch[half * 128 + r] = sum;
    if (!PT) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // P and V^T were written through the generic proxy
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after_sync();
        constexpr uint32_t idesc = umma_idesc_bf16(kAtQ, kAtD) | (VMN ? (1u << 16) : 0u);  // bit 16: B is MN-major
#pragma unroll
        for (int kb = 0; kb < kAtKeys / 64; ++kb) {
            const uint64_t pd = umma_desc_sw128(sbase + kb * kAtPBlock);
            // K-major V^T: 64-key blocks of [d][key] rows, 32 bytes per K = 16 step; MN-major V: 16 keys = 16 rows = 2 KB per step
            const uint64_t vd = umma_desc_sw128(sbase + kAtOffVt + (VMN ? kb * 64 * 128 : kb * kAtVtBlock));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (PT)  // keys 64 kb + 16 kk .. + 15: packed P columns 128 * (kb / 2) + 32 * (kb % 2) + 8 kk
                    umma_bf16_ts(tmem_base + 64u, tmem_base + (uint32_t)(128 * (kb >> 1) + 32 * (kb & 1) + 8 * kk),
                                 vd + (2048 >> 4) * kk, idesc, (kb | kk) != 0);
                else
                    umma_bf16(tmem_base, pd + 2 * kk, vd + (VMN ? (2048 >> 4) * kk : 2 * kk), idesc, (kb | kk) != 0);
            }
        }
        umma_commit(&bar_o);
    }
    __syncwarp();
    const float inv = 1.0f / (xch[r] + xch[128 + r]);
    mbar_wait(&bar_o, 0);
    tc_fence_after_sync();
    // ---- epilogue: the half-th 32 columns of O for row r ----
    {
        tmem_ld32(trow + (PT ? 64 : 0) + half * 32, v);
        if (q0 + r < L) {
            __nv_bfloat16* dst = ctx + ((size_t)b * L + q0 + r) * H + head * kAtD + half * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * q + 2 * e] * inv, v[8 * q + 2 * e + 1] * inv);
                    w[e] = *reinterpret_cast<const uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(dst + 8 * q) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ---- persistent, warp-specialised version -----------------------------------------------------------------------------
// One CTA per SM walks (document, head, query block) tiles.  Warp 0 is the producer: TMA loads of tile k + 1, MMA 1 of tile
// k + 1 and MMA 2 of tile k are issued while the eight worker warps (256 threads, two per query row) run the transposition,
// softmax and epilogue of tile k.  Two stages of shared memory (112 KB each: Q | K | V | V^T, P aliasing Q / K / V) and two
// 256-column TMEM buffers (S, then O in its first 64 columns), one mbarrier per hand-over and stage.
constexpr int kAtpThreads = 288;
constexpr int kAtStage = kAtOffVt + 4 * kAtVtBlock;               // 112 KB
constexpr int kAtpOffBar = 2 * kAtStage;
constexpr int kAtpSmem = kAtpOffBar + 128;

__device__ __forceinline__ void atp_workers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(kAtpThreads, 1)
attention_tc_persistent_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tkv,
                               const int32_t* __restrict__ seq_lens, int B, int L, int H, int heads,
                               __nv_bfloat16* __restrict__ ctx) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (smem_u32(smem) & 1023u) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAtpOffBar);
    uint64_t *loaded = bars, *s_ready = bars + 2, *p_ready = bars + 4, *o_ready = bars + 6, *stage_free = bars + 8;
    uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 10);
    int* flag = reinterpret_cast<int*>(bars + 11);  // [2]: a row of the stage's tile needs the exact softmax
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nqb = (L + kAtQ - 1) / kAtQ;
    const int ntiles = nqb * heads * B;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tq);
        tma_prefetch_desc(&tkv);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&loaded[i], 1);
            mbar_init(&s_ready[i], 1);
            mbar_init(&p_ready[i], 256);
            mbar_init(&o_ready[i], 1);
            mbar_init(&stage_free[i], 256);
            flag[i] = 0;
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t sbase = smem_u32(smem);

    if (warp == 0) {
        // ------------------------------ producer: TMA + MMA issue -------------------------------------------------
        if (lane == 0) {
            auto load_tile = [&](int k) {
                const int t = blockIdx.x + k * gridDim.x, s = k & 1;
                const int qb = t % nqb, head = (t / nqb) % heads, b = t / (nqb * heads);
                uint8_t* st = smem + s * kAtStage;
                mbar_arrive_expect_tx(&loaded[s], kAtQBytes + 2 * kAtKBytes);
                tma_load_2d(st, &tq, &loaded[s], head * kAtD, b * L + qb * kAtQ);
                tma_load_2d(st + kAtOffK, &tkv, &loaded[s], H + head * kAtD, b * L);
                tma_load_2d(st + kAtOffV, &tkv, &loaded[s], 2 * H + head * kAtD, b * L);
            };
            auto mma2 = [&](int k) {
                const int s = k & 1;
                mbar_wait(&p_ready[s], (k >> 1) & 1);
                tc_fence_after_sync();
                constexpr uint32_t idesc = umma_idesc_bf16(kAtQ, kAtD);
                const uint32_t st = sbase + s * kAtStage;
#pragma unroll
                for (int kb = 0; kb < kAtKeys / 64; ++kb) {
                    const uint64_t pd = umma_desc_sw128(st + kb * kAtPBlock), vd = umma_desc_sw128(st + kAtOffVt + kb * kAtVtBlock);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base + s * 256, pd + 2 * kk, vd + 2 * kk, idesc, (kb | kk) != 0);
                }
                umma_commit(&o_ready[s]);
            };
            pdl_wait();  // the QKV projection has written its output
            if (my_tiles > 0) load_tile(0);
            for (int k = 0; k < my_tiles; ++k) {
                const int s = k & 1;
                mbar_wait(&loaded[s], (k >> 1) & 1);
                tc_fence_after_sync();
                {
                    constexpr uint32_t idesc = umma_idesc_bf16(kAtQ, kAtKeys);
                    const uint32_t st = sbase + s * kAtStage;
                    const uint64_t qd = umma_desc_sw128(st), kd = umma_desc_sw128(st + kAtOffK);
#pragma unroll
                    for (int kk = 0; kk < kAtD / 16; ++kk) umma_bf16(tmem_base + s * 256, qd + 2 * kk, kd + 2 * kk, idesc, kk != 0);
                    umma_commit(&s_ready[s]);
                }
                if (k >= 1) mma2(k - 1);
                if (k + 1 < my_tiles) {
                    // the other stage (and TMEM buffer) is free once the workers have stored tile k - 1
                    if (k >= 1) mbar_wait(&stage_free[s ^ 1], ((k - 1) >> 1) & 1);
                    load_tile(k + 1);
                }
            }
            if (my_tiles > 0) mma2(my_tiles - 1);
        }
    } else {
        // ------------------------------ workers: transposition, softmax, epilogue ---------------------------------
        const int wid = warp - 1;                       // 0..7
        const int half = wid >> 2, quad = warp & 3;     // TMEM lane quadrant is fixed by the hardware warp id
        const int r = quad * 32 + lane;
        const float sc = 0.125f * kLog2e;
        for (int k = 0; k < my_tiles; ++k) {
            const int t = blockIdx.x + k * gridDim.x, s = k & 1;
            const uint32_t par = (k >> 1) & 1;
            const int qb = t % nqb, head = (t / nqb) % heads, b = t / (nqb * heads);
            const int q0 = qb * kAtQ;
            const int nkeys = min(max(seq_lens[b], 1), min(L, kAtKeys));
            uint8_t* st = smem + s * kAtStage;
            const uint32_t stu = sbase + s * kAtStage;
            float* xch = reinterpret_cast<float*>(st + 4 * kAtPBlock);
            mbar_wait(&loaded[s], par);
            {   // V -> V^T
                const int mtx = lane >> 3, i = lane & 7;
#pragma unroll 4
                for (int tt = wid; tt < 2 * (kAtKeys / 8); tt += 8) {
                    const int k0 = (tt >> 1) * 8, d0 = (tt & 1) * 32 + mtx * 8;
                    const int key = k0 + i;
                    uint32_t rr[4];
                    at_ldmatrix_x4_trans(rr, stu + kAtOffV + key * 128 + ((((d0 >> 3) ^ key) & 7) << 4));
                    const int drow = d0 + i;
                    at_stmatrix_x4(stu + kAtOffVt + (k0 >> 6) * kAtVtBlock + drow * 128 + (((((k0 & 63) >> 3)) ^ drow) & 7) * 16, rr);
                }
            }
            atp_workers_sync();          // V^T complete; the V staging is free
            mbar_wait(&s_ready[s], par); // MMA 1 is done with Q and K: P may overwrite them
            tc_fence_after_sync();
            const uint32_t trow = tmem_base + s * 256 + ((uint32_t)(quad * 32) << 16);
            const int cbeg = half * 128;
            float v[32];
            float mloc = -INFINITY;
            if (cbeg < nkeys) {
                tmem_ld32(trow + cbeg, v);
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    if (cbeg + e < nkeys) mloc = fmaxf(mloc, v[e]);
            }
            xch[half * 128 + r] = mloc;
            atp_workers_sync();
            float m = fmaxf(xch[r], xch[128 + r]);
            float off = m * sc, sum = 0.f, xmax = -INFINITY;
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + 128; c0 += 32) {
                if (c0 != cbeg && c0 < nkeys) tmem_ld32(trow + c0, v);
                sum += at_softmax_piece(v, c0, min(max(nkeys - c0, 0), 32), sc, off, st, r, xmax);
            }
            if (xmax > kAtSafeExp) flag[s] = 1;
            atp_workers_sync();
            if (flag[s]) {  // exact two-pass form (rare), block-uniform
                float me = -INFINITY;
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + 128 && c0 < nkeys; c0 += 32) {
                    tmem_ld32(trow + c0, v);
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (c0 + e < nkeys) me = fmaxf(me, v[e]);
                }
                atp_workers_sync();
                xch[half * 128 + r] = me;
                atp_workers_sync();
                m = fmaxf(xch[r], xch[128 + r]);
                off = m * sc;
                sum = 0.f;
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + 128; c0 += 32) {
                    if (c0 < nkeys) tmem_ld32(trow + c0, v);
                    sum += at_softmax_piece(v, c0, min(max(nkeys - c0, 0), 32), sc, off, st, r, xmax);
                }
                atp_workers_sync();
                if (threadIdx.x == 32) flag[s] = 0;
            }
            atp_workers_sync();
            xch[half * 128 + r] = sum;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // P, V^T: generic-proxy writes the MMA will read
            tc_fence_before_sync();
            mbar_arrive(&p_ready[s]);
            mbar_wait(&o_ready[s], par);
            tc_fence_after_sync();
            const float inv = 1.0f / (xch[r] + xch[128 + r]);  // (both halves stored their sums before arriving on p_ready)
            tmem_ld32(trow + half * 32, v);
            if (q0 + r < L) {
                __nv_bfloat16* dst = ctx + ((size_t)b * L + q0 + r) * H + head * kAtD + half * 32;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * q + 2 * e] * inv, v[8 * q + 2 * e + 1] * inv);
                        w[e] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    *reinterpret_cast<uint4*>(dst + 8 * q) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&stage_free[s]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- pipelined version: one CTA per SM, a producer warp and TWO worker groups -------------------------------------------
// The one-tile-per-CTA kernel is a serial chain per tile -- prologue, load (HBM / L2 latency), MMA 1, softmax, MMA 2, store --
// with only the second resident CTA to hide it: a fifth of the warps' time is the wait for the TMA loads, a CTA prologue is
// paid per tile.  Here the two tiles in flight per SM (two 256-column TMEM buffers) are worked on by two groups of eight
// warps, and a producer thread keeps the operands one tile AHEAD of both:
//   shared memory  three Q|K slots (48 KB) + two V slots (32 KB).  Q and K of a tile are dead once MMA 1 has run and V is only
//                  needed by MMA 2, so tile k+2's Q|K land while tiles k and k+1 are being worked on, and its V takes the slot
//                  tile k's MMA 2 has just released.  P never touches shared memory (TMEM A operand, V MN-major: as in PT above).
//   warp 0, lane 0 TMA loads and every tcgen05.mma: S(k+2) as soon as group k % 2 has stored O(k), P V(k) as soon as the
//                  group has written P(k).
//   groups 0 / 1   warps 1-8 / 9-16, tile k goes to group k % 2: softmax from TMEM (two threads per row), P back into the
//                  consumed part of S, O / row sum -> bf16 context rows.
// A tile whose exponents leave the safe range asks the producer to run MMA 1 again (S was overwritten by P; Q and K are still
// in their slot) and redoes the softmax in the exact two-pass form.
constexpr int kApThreads = 32 + 2 * 256;
constexpr int kApQK = kAtQBytes + kAtKBytes;                     // 48 KB
constexpr int kApOffV = 3 * kApQK;                               // 144 KB
constexpr int kApOffX = kApOffV + 2 * kAtKBytes;                 // 208 KB: [2 groups][2 halves][128] floats
constexpr int kApOffBar = kApOffX + 2 * 2 * 128 * 4;
constexpr int kApSmem = kApOffBar + 256;

__device__ __forceinline__ void ap_group_sync(int g) { asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory"); }

__global__ void __launch_bounds__(kApThreads, 1)
attention_tc_pipe_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tkv,
                         const int32_t* __restrict__ seq_lens, int B, int L, int H, int heads, __nv_bfloat16* __restrict__ ctx) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (smem_u32(smem) & 1023u) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kApOffBar);
    uint64_t *qk_loaded = bars, *v_loaded = bars + 3, *s_ready = bars + 5, *p_ready = bars + 7, *o_ready = bars + 9,
             *tmem_free = bars + 11;
    uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 13);
    volatile int* flag = reinterpret_cast<volatile int*>(bars + 14);  // [0..1] a row of the group's tile overflowed; [2..3] redo request
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nqb = (L + kAtQ - 1) / kAtQ;
    const int ntiles = nqb * heads * B;  // tile t = (document, head, query block), query block fastest: neighbouring CTAs
                                         // work on the two query blocks of one (document, head) and share its K / V in L2
    const int n = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto tile_of = [&](int k, int& qb, int& head, int& b) {
        const int t = blockIdx.x + k * gridDim.x;
        const int u = t / nqb;
        qb = t - u * nqb;
        b = u / heads;
        head = u - b * heads;
    };
    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tq);
        tma_prefetch_desc(&tkv);
        for (int i = 0; i < 3; ++i) mbar_init(&qk_loaded[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&v_loaded[i], 1);
            mbar_init(&s_ready[i], 1);
            mbar_init(&p_ready[i], 256);
            mbar_init(&o_ready[i], 1);
            mbar_init(&tmem_free[i], 256);
        }
        for (int i = 0; i < 4; ++i) flag[i] = 0;
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t sbase = smem_u32(smem);

    if (warp == 0) {
        if (lane == 0 && n > 0) {
            // ------------------------------ producer: TMA + MMA issue -------------------------------------------------
            uint32_t ph_qk = 0, ph_v = 0, ph_p = 0, ph_o = 0, ph_f = 0;  // phase bits, one per barrier
            auto wait = [&](uint64_t* bar, uint32_t& ph, int i) {
                mbar_wait(&bar[i], (ph >> i) & 1u);
                ph ^= 1u << i;
            };
            auto load_qk = [&](int k) {
                int qb, head, b;
                tile_of(k, qb, head, b);
                const int q = k % 3;
                uint8_t* st = smem + q * kApQK;
                mbar_arrive_expect_tx(&qk_loaded[q], kApQK);
                tma_load_2d(st, &tq, &qk_loaded[q], head * kAtD, b * L + qb * kAtQ);
                tma_load_2d(st + kAtQBytes, &tkv, &qk_loaded[q], H + head * kAtD, b * L);
            };
            auto load_v = [&](int k) {
                int qb, head, b;
                tile_of(k, qb, head, b);
                const int s = k & 1;
                mbar_arrive_expect_tx(&v_loaded[s], kAtKBytes);
                tma_load_2d(smem + kApOffV + s * kAtKBytes, &tkv, &v_loaded[s], 2 * H + head * kAtD, b * L);
            };
            auto mma_s = [&](int k) {  // S(k) = Q K^T into TMEM buffer k & 1
                const int s = k & 1;
                tc_fence_after_sync();
                constexpr uint32_t idesc = umma_idesc_bf16(kAtQ, kAtKeys);
                const uint32_t st = sbase + (k % 3) * kApQK;
                const uint64_t qd = umma_desc_sw128(st), kd = umma_desc_sw128(st + kAtQBytes);
#pragma unroll
                for (int kk = 0; kk < kAtD / 16; ++kk) umma_bf16(tmem_base + s * 256, qd + 2 * kk, kd + 2 * kk, idesc, kk != 0);
                umma_commit(&s_ready[s]);
            };
            pdl_wait();  // the QKV projection has written its output
            load_qk(0);
            load_v(0);
            if (n > 1) {
                load_qk(1);
                load_v(1);
            }
            if (n > 2) load_qk(2);
            wait(qk_loaded, ph_qk, 0);
            mma_s(0);
            if (n > 1) {
                wait(qk_loaded, ph_qk, 1);
                mma_s(1);
            }
            for (int k = 0; k < n; ++k) {
                const int s = k & 1;
                wait(p_ready, ph_p, s);
                while (flag[2 + s]) {  // the group asks for S(k) again (exact softmax): Q and K are still in their slot
                    flag[2 + s] = 0;
                    mma_s(k);
                    wait(p_ready, ph_p, s);
                }
                wait(v_loaded, ph_v, s);
                tc_fence_after_sync();
                {
                    constexpr uint32_t idesc = umma_idesc_bf16(kAtQ, kAtD) | (1u << 16);  // B = V, MN-major
                    const uint32_t tb = tmem_base + s * 256;
#pragma unroll
                    for (int kb = 0; kb < kAtKeys / 64; ++kb) {
                        const uint64_t vd = umma_desc_sw128(sbase + kApOffV + s * kAtKBytes + kb * 64 * 128);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16_ts(tb + 64u, tb + (uint32_t)(128 * (kb >> 1) + 32 * (kb & 1) + 8 * kk), vd + (2048 >> 4) * kk, idesc,
                                         (kb | kk) != 0);
                    }
                    umma_commit(&o_ready[s]);
                }
                if (k + 3 < n) load_qk(k + 3);  // slot k % 3: S(k) is final now
                if (k + 2 < n) {
                    wait(o_ready, ph_o, s);     // P V(k) has read V(k): its slot takes V(k+2)
                    load_v(k + 2);
                    wait(tmem_free, ph_f, s);   // the group has read O(k): the TMEM buffer takes S(k+2)
                    wait(qk_loaded, ph_qk, (k + 2) % 3);
                    mma_s(k + 2);
                }
            }
        }
    } else {
        // ------------------------------ worker groups: softmax, epilogue ------------------------------------------
        const int g = (warp - 1) >> 3, wid = (warp - 1) & 7;
        const int half = wid >> 2, quad = warp & 3;     // TMEM lane quadrant is fixed by the hardware warp id
        const int r = quad * 32 + lane;
        const float sc = 0.125f * kLog2e;
        float* xch = reinterpret_cast<float*>(smem + kApOffX) + g * 256;
        const uint32_t trow = tmem_base + g * 256 + ((uint32_t)(quad * 32) << 16);
        const int cbeg = half * 128;
        uint32_t ph_s = 0, ph_o = 0;
        // the next tile's coordinates and key count are fetched one tile ahead: a global load in front of the first use of
        // nkeys cost every tile ~10 % of its time (long-scoreboard stall right after the wait for S)
        int qb = 0, head = 0, b = 0, len_next = 1;
        if (g < n) {
            tile_of(g, qb, head, b);
            len_next = __ldg(seq_lens + b);
        }
        for (int k = g; k < n; k += 2) {
            const int q0 = qb * kAtQ, head_k = head, b_k = b;
            const int nkeys = min(max(len_next, 1), min(L, kAtKeys));
            if (k + 2 < n) {
                tile_of(k + 2, qb, head, b);
                len_next = __ldg(seq_lens + b);
            }
            mbar_wait(&s_ready[g], ph_s);
            ph_s ^= 1u;
            tc_fence_after_sync();
            float v[32];
            float mloc = -INFINITY;
            if (cbeg < nkeys) {
                tmem_ld32(trow + cbeg, v);
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    if (cbeg + e < nkeys) mloc = fmaxf(mloc, v[e]);
            }
            xch[half * 128 + r] = mloc;
            ap_group_sync(g);
            float m = fmaxf(xch[r], xch[128 + r]);
            float off = m * sc, sum = 0.f, xmax = -INFINITY;
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + 128; c0 += 32) {
                if (c0 != cbeg && c0 < nkeys) tmem_ld32(trow + c0, v);
                sum += at_softmax_piece<true>(v, c0, min(max(nkeys - c0, 0), 32), sc, off, nullptr, r, xmax,
                                              trow + (uint32_t)(cbeg + ((c0 - cbeg) >> 1)));
            }
            if (xmax > kAtSafeExp) flag[g] = 1;
            tmem_wait_st();
            ap_group_sync(g);  // (also: every thread has read the maxima, xch may take the sums)
            const bool redo = flag[g] != 0;  // group-uniform
            if (redo) {
                if (wid == 0 && lane == 0) flag[2 + g] = 1;  // published by the arrive below
                tc_fence_before_sync();
                mbar_arrive(&p_ready[g]);
                mbar_wait(&s_ready[g], ph_s);  // S(k) again
                ph_s ^= 1u;
                tc_fence_after_sync();
                float me = -INFINITY;
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + 128 && c0 < nkeys; c0 += 32) {
                    tmem_ld32(trow + c0, v);
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (c0 + e < nkeys) me = fmaxf(me, v[e]);
                }
                xch[half * 128 + r] = me;
                ap_group_sync(g);
                m = fmaxf(xch[r], xch[128 + r]);
                off = m * sc;
                sum = 0.f;
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + 128; c0 += 32) {
                    if (c0 < nkeys) tmem_ld32(trow + c0, v);
                    sum += at_softmax_piece<true>(v, c0, min(max(nkeys - c0, 0), 32), sc, off, nullptr, r, xmax,
                                                  trow + (uint32_t)(cbeg + ((c0 - cbeg) >> 1)));
                }
                tmem_wait_st();
                ap_group_sync(g);
                if (wid == 0 && lane == 0) flag[g] = 0;
            }
            xch[half * 128 + r] = sum;
            tc_fence_before_sync();
            mbar_arrive(&p_ready[g]);
            mbar_wait(&o_ready[g], ph_o);
            ph_o ^= 1u;
            tc_fence_after_sync();
            const float inv = 1.0f / (xch[r] + xch[128 + r]);  // (both halves stored their sums before arriving on p_ready)
            tmem_ld32(trow + 64 + half * 32, v);
            if (q0 + r < L) {
                __nv_bfloat16* dst = ctx + ((size_t)b_k * L + q0 + r) * H + head_k * kAtD + half * 32;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * q + 2 * e] * inv, v[8 * q + 2 * e + 1] * inv);
                        w[e] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    *reinterpret_cast<uint4*>(dst + 8 * q) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&tmem_free[g]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// asp_set_option("attn_tc"): plain-bf16 attention with L <= 256 on tcgen05 -- all variants agree bit for bit:
//   5 (default)  the pipelined kernel above: one CTA per SM, producer thread + two worker groups, P in tensor memory
//                (64 us against 90 us for variant 3 per layer at 32768 tokens when timed alone, profiles/r02_3y_*; 1.5-3 % of
//                a whole forward, r02_3z_*)
//   3            one tile per CTA, two CTAs per SM, V consumed as an MN-major operand straight from its TMA tile
//   4            as 3 with P in tensor memory (same speed as 3)
//   1            as 3 with V transposed in shared memory first (~1 % of a forward slower: r02_3t_*)
//   2            the first persistent kernel: eight worker warps run a tile's transposition, softmax and epilogue back to
//                back (slower than 1: profiles/r02_3h_attention_persistent_ab.txt)
//   0            always mma.sync (attention.cu)
int g_attn_tc = 5;

bool attention_tc_supported(const void* qkv_lo, int L, int H, int heads) {
    return g_attn_tc && qkv_lo == nullptr && L >= 1 && L <= kAtKeys && H == heads * kAtD;
}

int attention_tc_launch(const void* qkv_hi, const int32_t* seq_lens, int B, int L, int H, int heads, void* ctx_hi,
                        cudaStream_t stream) {
    // tensor maps of the QKV activation, cached: the encoder calls this once per layer with the same workspace
    struct Cached {
        const void* ptr = nullptr;
        uint64_t rows = 0, cols = 0;
        CUtensorMap q, kv;
    };
    static thread_local Cached cache;
    const uint64_t rows = (uint64_t)B * L, cols = 3ull * H;
    if (cache.ptr != qkv_hi || cache.rows != rows || cache.cols != cols) {
        int rc;
        if ((rc = make_tmap_bf16(&cache.q, qkv_hi, rows, cols, kAtQ))) return rc;
        if ((rc = make_tmap_bf16(&cache.kv, qkv_hi, rows, cols, kAtKeys))) return rc;
        cache.ptr = qkv_hi;
        cache.rows = rows;
        cache.cols = cols;
    }
    static thread_local int attr_dev = -1;
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        ASP_CUDA(cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
        ASP_CUDA(cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
        ASP_CUDA(cudaFuncSetAttribute(attention_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
        attr_dev = dev;
    }
    if (g_attn_tc == 5) {
        static thread_local int attr_dev5 = -1;
        if (attr_dev5 != dev) {
            ASP_CUDA(cudaFuncSetAttribute(attention_tc_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kApSmem));
            attr_dev5 = dev;
        }
        const int ntiles = ((L + kAtQ - 1) / kAtQ) * heads * B;
        ASP_CUDA(launch_pdl(attention_tc_pipe_kernel, dim3(std::min(ntiles, sm_count())), dim3(kApThreads), (size_t)kApSmem, stream,
                            cache.q, cache.kv, seq_lens, B, L, H, heads, (__nv_bfloat16*)ctx_hi));
        ASP_LAUNCH_CHECK("attention_tc_pipe_kernel");
        return ASP_OK;
    }
    if (g_attn_tc == 2) {
        static thread_local int attr_dev2 = -1;
        if (attr_dev2 != dev) {
            ASP_CUDA(cudaFuncSetAttribute(attention_tc_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtpSmem));
            attr_dev2 = dev;
        }
        const int ntiles = ((L + kAtQ - 1) / kAtQ) * heads * B;
        ASP_CUDA(launch_pdl(attention_tc_persistent_kernel, dim3(std::min(ntiles, sm_count())), dim3(kAtpThreads), (size_t)kAtpSmem,
                            stream, cache.q, cache.kv, seq_lens, B, L, H, heads, (__nv_bfloat16*)ctx_hi));
        ASP_LAUNCH_CHECK("attention_tc_persistent_kernel");
        return ASP_OK;
    }
    dim3 grid((L + kAtQ - 1) / kAtQ, heads, B);
    if (g_attn_tc == 4)
        ASP_CUDA(launch_pdl(attention_tc_kernel<true, true>, grid, dim3(kAtThreads), (size_t)kAtSmem, stream, cache.q, cache.kv,
                            seq_lens, L, H, (__nv_bfloat16*)ctx_hi));
    else if (g_attn_tc == 3)
        ASP_CUDA(launch_pdl(attention_tc_kernel<true>, grid, dim3(kAtThreads), (size_t)kAtSmem, stream, cache.q, cache.kv, seq_lens,
                            L, H, (__nv_bfloat16*)ctx_hi));
    else
        ASP_CUDA(launch_pdl(attention_tc_kernel<false>, grid, dim3(kAtThreads), (size_t)kAtSmem, stream, cache.q, cache.kv, seq_lens,
                            L, H, (__nv_bfloat16*)ctx_hi));
    ASP_LAUNCH_CHECK("attention_tc_kernel");
    return ASP_OK;
}

}  // namespace asp
