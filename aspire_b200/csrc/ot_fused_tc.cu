// MEASURED PROTOTYPE, off by default (asp_set_option("ot_fused_tc", 1) turns it on; parity-tested either way): the 1 x N
// Gram tile on tcgen05 instead of FFMA2.  It is correct and slower -- 1.04e8 vs 1.34e8 pairs/s -- for a reason worth
// keeping on file (see g_ot_fused_tc below and DESIGN.md section 4).
//
// K2+K4 fused, tensor-core variant for POOLS: one query document against many candidates (q_group >= 128), the shape of
// caching_score (src/learning/facetid_models/disent_models.py:274-297: one query replicated over a 1k pool) and of the
// headline bench.  Same results as ot_fused.cu (pair_distances.py:21-92 + geomloss), different machine mapping:
//
//   ot_fused.cu (v7) computes the 10 x 10 Gram tile of a pair with 1320 FFMA2 and is bound by fp32 issue (and, under a
//   sustained run, by the 1 kW power cap) at 0.65-0.75 of the HBM roofline.  Here the Gram tile comes from tcgen05:
//   128 candidate sentence rows (12.8 documents) x the 10 query rows (N = 16) per MMA, bf16 hi/lo split operands
//   (hi.hi + hi.lo + lo.hi, fp32 accumulation in TMEM).  The candidate rows cannot come in by TMA -- they are fp32 in HBM
//   and the tensor core wants bf16 halves, and a conversion pass through shared memory would cost more shared-memory
//   bandwidth than the HBM roofline leaves -- so PRODUCER WARPS load them with coalesced 128-bit
//   LDGs (one K block per thread in flight in registers, 64 KB per SM), split them in registers and store the bf16 halves
//   straight into the 128B-swizzled operand tiles the MMA reads.  Shared-memory traffic per pair: 30 KB written + 45 KB read by
//   the MMA + 29 KB read by the Sinkhorn solve, ~0.6 of what the HBM roofline allows.
//
//   One persistent CTA per SM over a CONTIGUOUS range of pairs (the query operand changes once or twice per CTA):
//     warps 0-3    drain: accumulator row (candidate sentence) per thread -> sqrt(max(|q|^2+|c|^2-2q.c, 1e-8)) -> the pair's
//                  10 x 10 cost tile in shared memory
//     warp  4      MMA issuer (warps 5-7 idle: warpgroup granularity of setmaxnreg)
//     warps 8-15   producers (also stage the query's bf16 halves + norms when the query changes)
//     warps 16-23  Sinkhorn: one pair per thread (solve_pair_thread_stream, the solver of the other fused kernels);
//                  groups of 32 consecutive pairs go to the warps round robin
#include <algorithm>
#include <type_traits>
#include "bert/tc05.cuh"
#include "ot_pair.cuh"

namespace asp {

using namespace tc;

// Where the generic-proxy -> async-proxy fence for the operand tiles sits.
//   0  every producer thread fences before it arrives (the documented pattern).  The fence is a MEMBAR that waits for the
//      thread's loads in flight, so the next K block's loads can only be issued after it.
//   1  the MMA-issuing thread fences after its acquire of the stage barrier; producers reload each register piece as soon
//      as it has been converted (loads stay in flight across the hand-over).
#ifndef ASP_TC_FENCE
#define ASP_TC_FENCE 0
#endif
constexpr int kTcFT = 10;                 // max sentences per document
constexpr int kTcLd = kTcFT * kTcFT;      // cost floats per pair
constexpr int kTcRows = 128;              // candidate sentence rows per tile (MMA M)
constexpr int kTcN = 16;                  // MMA N: query rows padded to 16
constexpr int kTcKB = 64;                 // K block: 64 elements = 128-byte bf16 rows (one 128B swizzle atom)
constexpr int kTcStages = 2;
constexpr int kTcATile = kTcRows * kTcKB * 2;       // 16 KB (hi or lo)
constexpr int kTcStage = 2 * kTcATile;              // hi + lo
constexpr int kTcBTile = kTcN * kTcKB * 2;          // 2 KB (hi or lo) per K block
constexpr int kTcMaxD = 768;
constexpr int kTcSink = 8, kTcProd = 8;
constexpr int kTcWarps = 8 + kTcProd + kTcSink;     // 24
constexpr int kTcAccCols = 32;                      // TMEM columns per accumulator buffer (16 used)
constexpr int kTcSmemB = (kTcMaxD / kTcKB) * 2 * kTcBTile;  // query operand, all K blocks: 48 KB
constexpr int kTcSmem = kTcStages * kTcStage + kTcSmemB + kTcSink * 32 * kTcLd * 4 + 1024;
// Register budgets (setmaxnreg).  The pool is what the CTA was launched with -- 24 warps x 80 registers = 1920 per lane
// slot -- not the SM's register file: drain 48, MMA warpgroup 24, producers 96, Sinkhorn 104: 4*48 + 4*24 + 8*96 + 8*104 = 1888

struct TcArgs {
    const float* q;
    const int32_t* q_lens;
    const float* c;
    const int32_t* c_lens;
    const int32_t* c_index;
    int q_group, B, Sq, Sc, D;
    float inv_temp;
};

// byte offset of (row r, byte o of the 128-byte row) in a K-major tile stored with the 128B swizzle
__device__ __forceinline__ uint32_t tc_sw128(int r, int o) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((o >> 4) ^ r) & 7) << 4) + (o & 15));
}
__device__ __forceinline__ void tc_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_mbar_arrive_n(uint64_t* bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void tc_tmem_ld16(uint32_t taddr, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                   "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :
                 : "memory");
}

// fp32 x4 -> bf16 hi x4 (8 bytes) and bf16 lo x4 = bf16(x - hi), hi = x TRUNCATED to its upper 16 bits: the hi halves are
// a byte permute (no conversion instruction), x - hi is exact in fp32, and hi + lo carries x to 2^-17 relative.  Packed
// fp32 arithmetic (FADD2 / FFMA2): ~14 instructions per four floats.  sq accumulates the squares (for |c|^2).
__device__ __forceinline__ void tc_split4(const float4& v, uint2& hi, uint2& lo, float& sq) {
    const uint32_t x = __float_as_uint(v.x), y = __float_as_uint(v.y), z = __float_as_uint(v.z), w = __float_as_uint(v.w);
    hi.x = __byte_perm(x, y, 0x7632);  // upper halves of (x, y): low 16 bits = bf16(x), high 16 bits = bf16(y)
    hi.y = __byte_perm(z, w, 0x7632);
    const float2 r01 = __fadd2_rn(make_float2(v.x, v.y), make_float2(-__uint_as_float(x & 0xffff0000u), -__uint_as_float(y & 0xffff0000u)));
    const float2 r23 = __fadd2_rn(make_float2(v.z, v.w), make_float2(-__uint_as_float(z & 0xffff0000u), -__uint_as_float(w & 0xffff0000u)));
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(r01.x, r01.y), l1 = __floats2bfloat162_rn(r23.x, r23.y);
    lo.x = *reinterpret_cast<const uint32_t*>(&l0);
    lo.y = *reinterpret_cast<const uint32_t*>(&l1);
    const float2 s2 = __ffma2_rn(make_float2(v.x, v.y), make_float2(v.x, v.y), __fmul2_rn(make_float2(v.z, v.w), make_float2(v.z, v.w)));
    sq += s2.x + s2.y;
}

// The tile sequence of a CTA: candidate sentence rows [rel, rel + nrows) RELATIVE to the first row of its pair range
// (row b0 * Sc), cut at query boundaries.  32-bit state: a CTA's range is at most a few million rows.
struct TcTileIter {
    int rel, end, bound, grows, q;  // current row, end row, next query boundary, rows per query group, query index
    __device__ __forceinline__ TcTileIter(long long b0, long long b1, int q_group, int Sc) {
        const long long R0 = b0 * Sc, g = (long long)q_group * Sc;
        rel = 0;
        end = (int)((b1 - b0) * Sc);
        q = (int)(R0 / g);
        const long long to_bound = (long long)(q + 1) * g - R0;
        bound = (int)min(to_bound, (long long)end);
        grows = (int)min(g, (long long)0x3fffffff);
    }
    __device__ __forceinline__ bool done() const { return rel >= end; }
    __device__ __forceinline__ int nrows() const { return min(kTcRows, min(bound, end) - rel); }
    __device__ __forceinline__ int qidx() const { return q; }
    __device__ __forceinline__ void next() {
        rel += nrows();
        if (rel >= bound && rel < end) {
            bound = (int)min((long long)bound + grows, (long long)end);
            ++q;
        }
    }
};

__device__ __forceinline__ void tc_phase2(float* Cs, int ql, int cl, int b, int Sq, int Sc, const float* eps_s, int n_eps,
                                          float inv_temp, const OtOut* out) {
    solve_pair_thread_stream<kTcFT, kTcFT, false>(Cs, ql, cl, b, Sq, Sc, eps_s, n_eps, inv_temp, *out);
}
__device__ __forceinline__ void tc_phase2_full(float* Cs, int b, const float* eps_s, int n_eps, float inv_temp,
                                               const OtOut* out) {
    solve_pair_thread_stream<kTcFT, kTcFT, true>(Cs, kTcFT, kTcFT, b, kTcFT, kTcFT, eps_s, n_eps, inv_temp, *out);
}

__global__ void __launch_bounds__(kTcWarps * 32, 1)
ot_fused_tc_kernel(const TcArgs a, const EpsSched sched, const OtOut out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ float eps_s[ASP_MAX_EPS];
    __shared__ OtOut out_s;
    __shared__ float cn_s[2][2][kTcRows];  // |c|^2 of the tile's rows: [accumulator buffer][producer set: even / odd K blocks]
    __shared__ float qn_s[2][kTcN];      // |q|^2 of the query rows, by query parity
    __shared__ uint64_t full[kTcStages], empty[kTcStages], acc_full[2], acc_empty[2], cfull[kTcSink], cempty[kTcSink];
    __shared__ uint32_t tmem_slot;
    uint8_t* stage0 = smem;
    uint8_t* bop = smem + kTcStages * kTcStage;                                // [kb][hi 2 KB | lo 2 KB]
    float* cost = reinterpret_cast<float*>(bop + kTcSmemB);                    // [kTcSink][32][100]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Sq = a.Sq, Sc = a.Sc, D = a.D;
    const int kblocks = D / kTcKB;
    const long long b0 = (long long)a.B * blockIdx.x / gridDim.x, b1 = (long long)a.B * (blockIdx.x + 1) / gridDim.x;
    const int npairs = (int)(b1 - b0);

    for (int k = threadIdx.x; k < sched.n; k += blockDim.x) eps_s[k] = sched.eps[k];
    if (threadIdx.x == 0) {
        out_s = out;
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(&full[s], (kTcProd / 2) * 32);  // stage s is filled by producer set s
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 128);
        }
        for (int w = 0; w < kTcSink; ++w) {
            mbar_init(&cfull[w], 32 * Sc);  // one arrival per candidate sentence row of the group's 32 pairs
            mbar_init(&cempty[w], 1);
        }
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(&tmem_slot, 2 * kTcAccCols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;

    if (warp < 4) {
        // ============================== drain ====================================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        const int r = warp * 32 + lane;
        int k = 0;
        for (TcTileIter ti(b0, b1, a.q_group, Sc); !ti.done(); ti.next(), ++k) {
            const int nrows = ti.nrows(), buf = k & 1, qpar = ti.qidx() & 1;
            const bool rv = r < nrows;
            const int rr = ti.rel + r;
            const int p = rv ? rr / Sc : 0;                       // pair index within the CTA's range
            const int j = rv ? rr - p * Sc : 0;
            const int G = p >> 5, w = G % kTcSink, use = G / kTcSink;
            if (rv) mbar_wait_parked(&cempty[w], (use & 1) ^ 1);  // the Sinkhorn warp is done with the slot's previous group
            mbar_wait_parked(&acc_full[buf], (k >> 1) & 1);
            __syncwarp();
            tc_fence_after_sync();
            float v[16];
            tc_tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * kTcAccCols), v);
            const float cn = cn_s[buf][0][r] + cn_s[buf][1][r];  // read before the accumulator (and with it this buffer of norms) is released
            tc_fence_before_sync();
            mbar_arrive(&acc_empty[buf]);
            if (rv) {
                float* dst = cost + (size_t)(w * 32 + (p & 31)) * kTcLd + j;
#pragma unroll
                for (int i = 0; i < kTcFT; ++i)
                    if (i < Sq) dst[i * kTcFT] = sqrtf(fmaxf(qn_s[qpar][i] + cn - 2.f * v[i], 1e-8f));
                mbar_arrive(&cfull[w]);
            }
        }
        // the last group of the range may hold fewer than 32 pairs: supply the arrivals of its missing rows
        if (threadIdx.x == 0 && (npairs & 31)) {
            const int G = (npairs - 1) >> 5;
            tc_mbar_arrive_n(&cfull[G % kTcSink], (uint32_t)((32 - (npairs & 31)) * Sc));
        }
    } else if (warp < 8) {
        // ============================== MMA issuer ===============================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (warp == 4 && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kTcRows, kTcN);
            const uint32_t bop_u = smem_u32(bop);
            int it = 0, k = 0;
            for (TcTileIter ti(b0, b1, a.q_group, Sc); !ti.done(); ti.next(), ++k) {
                const int buf = k & 1;
                mbar_wait_parked(&acc_empty[buf], ((k >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t acc = tmem_base + (uint32_t)(buf * kTcAccCols);
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % kTcStages, ph = (it / kTcStages) & 1;
                    mbar_wait_parked(&full[s], ph);
                    if (ASP_TC_FENCE == 1) tc_fence_async_smem();
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(stage0 + s * kTcStage);
                    const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + kTcATile);
                    const uint64_t b_hi = umma_desc_sw128(bop_u + kb * 2 * kTcBTile), b_lo = umma_desc_sw128(bop_u + kb * 2 * kTcBTile + kTcBTile);
#pragma unroll
                    for (int kk = 0; kk < kTcKB / 16; ++kk) {  // 16 bf16 = 32 bytes along K inside the swizzle atom
                        umma_bf16(acc, a_lo + 2 * kk, b_hi + 2 * kk, idesc, (kb | kk) != 0);  // small terms first
                        umma_bf16(acc, a_hi + 2 * kk, b_lo + 2 * kk, idesc, true);
                        umma_bf16(acc, a_hi + 2 * kk, b_hi + 2 * kk, idesc, true);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else if (warp < 8 + kTcProd) {
        // ============================== producers ================================================================
        // Two sets of four warps; set h converts the K blocks kb = h, h + 2, ... of every tile into stage h.  Per K block
        // a thread loads its 16 pieces (8 rows x 2 x 16 bytes), splits them, stores the halves, fences and arrives, and
        // only THEN issues the loads of its next K block: the proxy fence (MEMBAR) waits for every load the thread has
        // in flight, so a load issued before it would put a full HBM latency inside the fence (measured: 2950 clk per K
        // block that way).  The latency is hidden across threads instead: 256 threads x 256 bytes = 64 KB in flight per SM.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        const int pw = warp - 8, pt = threadIdx.x - 256;       // producer warp / thread (0..255)
        const int set = pw >> 2, pws = pw & 3;
        const int c8 = lane & 7;
        // tile row of pass p: 16 p + pws + 4 (lane >> 3)  (rows 4 apart inside a warp: conflict-free swizzled stores)
        const int rbase = pws + 4 * (lane >> 3);
        float4 bufA[8], bufB[8];   // floats [4 c8, +4) and [32 + 4 c8, +4) of the K block, for each of the thread's 8 rows
        float nrm[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) nrm[p] = 0.f;
        // swizzled byte offsets of this thread's two 8-byte pieces in tile row rbase; row 16 p + rbase is 2048 p bytes further
        const uint32_t o0 = tc_sw128(rbase, 8 * c8), o1 = tc_sw128(rbase, 64 + 8 * c8);
        // ---- load stream: this set's K blocks, one ahead of the convert stream ----
        TcTileIter lt(b0, b1, a.q_group, Sc);
        int lkb = set;
        int lrow[8];   // corpus row (document * Sc + sentence) behind each of this thread's 8 tile rows; -1 = none
        auto tile_ptrs = [&]() {
            const int nrows = lt.done() ? 0 : lt.nrows();
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int r = p * 16 + rbase;
                lrow[p] = -1;
                if (r < nrows) {
                    const int rr = lt.rel + r;
                    const int pl = rr / Sc, j = rr - pl * Sc;
                    const long long b = b0 + pl;
                    const int ci = a.c_index ? a.c_index[b] : (int)b;
                    if (j < min(a.c_lens[ci], Sc)) lrow[p] = ci * Sc + j;
                }
            }
        };
        const float* cbase = a.c + 4 * c8;
        auto issue_loads = [&]() {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                if (lrow[p] >= 0) {
                    const float* src = cbase + (size_t)lrow[p] * D + lkb * kTcKB;
                    bufA[p] = ldg_stream(reinterpret_cast<const float4*>(src));
                    bufB[p] = ldg_stream(reinterpret_cast<const float4*>(src + 32));
                } else {
                    bufA[p] = make_float4(0.f, 0.f, 0.f, 0.f);
                    bufB[p] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            lkb += 2;
            if (lkb >= kblocks) {
                lkb = set;
                lt.next();
                tile_ptrs();
            }
        };
        tile_ptrs();
        if (!lt.done()) issue_loads();
        // ---- convert stream ----
        int k = 0, cur_q = -1;
        for (TcTileIter ti(b0, b1, a.q_group, Sc); !ti.done(); ti.next(), ++k) {
            if (ti.qidx() != cur_q) {
                // new query: every MMA that reads the old operand must be done -> the "empty" phases of the last two K blocks
                cur_q = ti.qidx();
                const int it0 = k * kblocks;
                if (it0 >= 1) mbar_wait_parked(&empty[(it0 - 1) % kTcStages], ((it0 - 1) / kTcStages) & 1);
                if (it0 >= 2) mbar_wait_parked(&empty[(it0 - 2) % kTcStages], ((it0 - 2) / kTcStages) & 1);
                const float* qsrc = a.q + (size_t)cur_q * Sq * D;
                const int ql = min(max(a.q_lens[cur_q], 0), Sq);
                // 16 rows x D floats, 4 floats per thread and step; rows >= ql are zero
                for (int e = pt; e < kTcN * (D / 4); e += kTcProd * 32) {
                    const int row = e / (D / 4), f4 = e - row * (D / 4);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row < ql) v = __ldg(reinterpret_cast<const float4*>(qsrc + (size_t)row * D) + f4);
                    uint2 hi, lo;
                    float unused = 0.f;
                    tc_split4(v, hi, lo, unused);
                    const int kb = (4 * f4) / kTcKB, o = ((4 * f4) % kTcKB) * 2;
                    uint8_t* tb = bop + kb * 2 * kTcBTile;
                    *reinterpret_cast<uint2*>(tb + tc_sw128(row, o)) = hi;
                    *reinterpret_cast<uint2*>(tb + kTcBTile + tc_sw128(row, o)) = lo;
                }
                // query norms: one warp per two rows
                for (int row = pw; row < kTcN; row += kTcProd) {
                    float sum = 0.f;
                    if (row < ql)
                        for (int d = lane; d < D; d += 32) {
                            const float x = __ldg(qsrc + (size_t)row * D + d);
                            sum = fmaf(x, x, sum);
                        }
                    sum = warp_sum(sum);
                    if (lane == 0) qn_s[cur_q & 1][row] = sum;
                }
                tc_fence_async_smem();
                asm volatile("bar.sync 1, 256;" ::: "memory");  // operand complete before any producer releases a stage
            }
            const int buf = k & 1;
            uint8_t* th = stage0 + set * kTcStage;
#pragma unroll 1
            for (int kb = set; kb < kblocks; kb += 2) {
                const int it = k * kblocks + kb;
                mbar_wait_parked(&empty[set], ((it / kTcStages) & 1) ^ 1);
                const bool reload = ASP_TC_FENCE != 0 && kb + 2 < kblocks;  // next K block of this set is in the same tile
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    uint2 hi, lo;
                    tc_split4(bufA[p], hi, lo, nrm[p]);
                    *reinterpret_cast<uint2*>(th + p * 2048 + o0) = hi;
                    *reinterpret_cast<uint2*>(th + kTcATile + p * 2048 + o0) = lo;
                    tc_split4(bufB[p], hi, lo, nrm[p]);
                    *reinterpret_cast<uint2*>(th + p * 2048 + o1) = hi;
                    *reinterpret_cast<uint2*>(th + kTcATile + p * 2048 + o1) = lo;
                    if (reload && lrow[p] >= 0) {
                        const float* src = cbase + (size_t)lrow[p] * D + (kb + 2) * kTcKB;
                        bufA[p] = ldg_stream(reinterpret_cast<const float4*>(src));
                        bufB[p] = ldg_stream(reinterpret_cast<const float4*>(src + 32));
                    }
                }
                if (kb + 2 >= kblocks) {
                    // this set's share of the squared norms of the tile's rows: sum the 8 lanes of each row
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        float sum = nrm[p];
                        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
                        if (c8 == 0) cn_s[buf][set][p * 16 + rbase] = sum;
                        nrm[p] = 0.f;
                    }
                }
                if (ASP_TC_FENCE == 0) tc_fence_async_smem();
                mbar_arrive(&full[set]);
                if (reload) {            // the loads are already on their way: only advance the load stream's position
                    lkb += 2;
                    if (lkb >= kblocks) {
                        lkb = set;
                        lt.next();
                        tile_ptrs();
                    }
                } else if (!lt.done()) {
                    issue_loads();
                }
            }
        }
    } else {
        // ============================== Sinkhorn: one pair per thread ============================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        const int w = warp - (8 + kTcProd);
        float* Cs = cost + (size_t)(w * 32 + lane) * kTcLd;
        const int ngroups = (npairs + 31) >> 5;
        for (int G = w, use = 0; G < ngroups; G += kTcSink, ++use) {
            const long long b = b0 + 32LL * G + lane;
            const bool valid = b < b1;
            int ql = 0, cl = 0;
            if (valid) {
                ql = min(max(a.q_lens[b / a.q_group], 0), Sq);
                cl = min(max(a.c_lens[a.c_index ? a.c_index[b] : b], 0), Sc);
            }
            const bool all_full = __all_sync(0xffffffffu, !valid || (ql == kTcFT && cl == kTcFT));
            mbar_wait_parked(&cfull[w], use & 1);
            if (valid) {
                if (all_full) tc_phase2_full(Cs, (int)b, eps_s, sched.n, a.inv_temp, &out_s);
                else tc_phase2(Cs, ql, cl, (int)b, Sq, Sc, eps_s, sched.n, a.inv_temp, &out_s);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&cempty[w]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 2 * kTcAccCols);
}

// asp_set_option("ot_fused_tc"): 1 = pools (q_group >= 128) take this kernel, 0 (default) = always ot_fused.cu.
// OFF by default: measured 1.04e8 pairs/s against 1.34e8 for the FFMA2 kernel on the bench workload (256 x 1k pairs,
// sustained; profiles/r02_v_fused_tc_ab.txt, r02_u_fused_tc_ncu_summary.txt).  What bounds it is bytes in flight: the fp32
// rows must pass through registers to be split, 256 producer threads x 256 B = 64 KB per SM is all the register file
// leaves, the proxy fence keeps a thread from holding loads across the hand-over, and HBM under this load answers in
// ~2.5 us -- 3.2 TB/s.  The FFMA2 kernel stages the same rows by cp.async with 80 KB permanently in flight.
int g_ot_fused_tc = 0;

bool ot_fused_tc_supported(int q_group, int B, int Sq, int Sc, int D) {
    return g_ot_fused_tc && q_group >= 128 && Sq <= kTcFT && Sc <= kTcFT && D >= 2 * kTcKB && (D % (2 * kTcKB)) == 0 && D <= kTcMaxD &&
           B >= 64 * sm_count();
}

int ot_fused_tc_launch(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens,
                       const int32_t* c_index, int B, int Sq, int Sc, int D, const EpsSched& sched, float temp, const OtOut& out,
                       cudaStream_t stream) {
    static thread_local int attr_dev = -1;
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        ASP_CUDA(cudaFuncSetAttribute(ot_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
        attr_dev = dev;
    }
    TcArgs a{q, q_lens, c, c_lens, c_index, q_group, B, Sq, Sc, D, 1.0f / temp};
    ot_fused_tc_kernel<<<sm_count(), kTcWarps * 32, kTcSmem, stream>>>(a, sched, out);
    ASP_LAUNCH_CHECK("ot_fused_tc_kernel");
    return ASP_OK;
}

}  // namespace asp
