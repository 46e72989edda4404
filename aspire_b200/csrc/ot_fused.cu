// K2+K4 fused: otAspire score straight from the sentence representations -- the headline kernel.
//
// Replaces, in ONE launch and without ever writing the cost tensor to HBM,
//   pad mask + -cdist            src/learning/facetid_models/pair_distances.py:39-50
//   softmax marginals            pair_distances.py:56-60
//   geomloss SamplesLoss(...)    pair_distances.py:68-72 / 88-91   (eps-scaling Sinkhorn, restated in ot_pair.cuh)
//   plan + primal value          pair_distances.py:76-85
// for documents of at most kFT sentences (the reference's abstracts: 10-sentence synthetic config, CSFCube ~7).
//
// Every candidate row is read from HBM exactly once (30 KB per pair), only 4-8 bytes per pair are written, and the
// query never comes from memory in the inner loop.
//
//   ot_fused_v7_kernel -- 12 warps per SM, specialised by pipe and sized with setmaxnreg:
//     * 8 Gram warps (two per scheduler, 200 registers), phase 1 on the FMA pipe: half-tiles of <= 16 pairs from a
//       launch-owned atomic counter; per pair the two HALF-WARPS take 5 query rows each and their 16 lanes split the
//       embedding dimension.  Candidate rows are staged by cp.async through a per-warp shared-memory ring (4 slices =
//       10 KB in flight per warp, 80 KB per SM; the stream runs across pair boundaries) and read back with one
//       128-bit LDS per row.  The QUERY rows live in TENSOR MEMORY: each lane parks its 240 floats of the current
//       query in its own TMEM lane once per (warp, query) with tcgen05.st and pulls one 20-float slice per step back
//       with tcgen05.ld, one slice ahead.  The 5x10 Gram tile + squared norms accumulate in packed fp32 (FFMA2), are
//       transpose-reduced over the 16 lanes and leave sqrt(max(|q|^2+|c|^2-2q.c, 1e-8)) in a shared half-tile.
//     * 4 Sinkhorn warps (one per scheduler, 104 registers), phase 2 on the MUFU pipe: one pair per THREAD, the cost
//       tile streamed from shared memory every step (ot_pair.cuh: solve_pair_thread_stream), one ex2 per (i,j) and
//       step; full 10x10 tiles run a mask-free specialisation.
//     * hand-over: per scheduler a ring of four half-tiles, ticketed slots, full/empty mbarriers, parked waits.
// (The round-1 predecessor -- 8 warps per SM, every warp running both phases, phase-staggered -- is gone: the A/B that
// kept it alive is recorded in profiles/r01_al_sustained_ab.txt.)
#include <algorithm>
#include "bert/tc05.cuh"
#include "gram.cuh"
#include "ot_pair.cuh"

namespace asp {

constexpr int kFT = 10;        // max sentences per document on the fused path
constexpr int kHR = kFT / 2;   // query rows per half-warp
constexpr int kCostLd = 101;   // floats per pair in a padded cost tile (default of phase1's LD)
constexpr int kRedVals = 64;   // 50 dot products + 5 candidate norms, padded for the 16-lane transpose-reduce
#ifndef ASP_V7_TTABLE
#define ASP_V7_TTABLE 1  // 1: the solvers read log2e / eps from a per-CTA table instead of dividing every step
#endif
#if ASP_V7_TTABLE
#define ASP_V7_TSCHED (eps_s + ASP_MAX_EPS)
#else
#define ASP_V7_TSCHED nullptr
#endif
#ifndef ASP_V7_RING
#define ASP_V7_RING 5
#endif
#ifndef ASP_V7_SLOTS
#define ASP_V7_SLOTS 4
#endif
constexpr int kRing = ASP_V7_RING;         // candidate slices in the per-warp cp.async ring (kRing-1 in flight)
constexpr int kSliceFloats = kFT * 64;     // one slice: 64 floats of each of the kFT rows
constexpr int kSliceBytes = kSliceFloats * 4;
constexpr int kCounterSlots = 1024;
constexpr int kQCols = 20;        // TMEM columns per query slice and lane: kHR rows x 4 floats
constexpr int kTmemCols = 512;    // the whole tensor memory of the SM: 256 columns per warp of a lane quadrant;
                                  // D/64 * kQCols <= 256  =>  D <= 768
constexpr int kMaxFusedD = 768;

__device__ unsigned int g_tile_counter[kCounterSlots];  // launch-owned: a launch's slot is zeroed on its own stream

struct FusedArgs {
    const float* q;
    const int32_t* q_lens;
    const float* c;
    const int32_t* c_lens;
    const int32_t* c_index;  // optional: pair b scores candidate document c_index[b] of a resident corpus (c, c_lens are
                             // then indexed by document id); NULL = candidate b
    int q_group, B, Sq, Sc, D;
    unsigned int* counter;  // tile counter of this launch
    int tile_pairs;  // pairs per warp tile (<= 32; chosen by the launcher so that the tiles fill whole waves of warps)
    float inv_temp;
};

// Sum v[] over the W lanes of each aligned lane group (W = 16: half-warps).  Afterwards lane l (index within its
// group) holds, in v[m], the total of original slot W*m + bitrev(l).
template <int NV, int W>
__device__ __forceinline__ void transpose_reduce_w(float (&v)[NV], int lane) {
    int n = NV;
#pragma unroll
    for (int s = W / 2; s > 0; s >>= 1) {
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

// 16-byte cp.async (L1 bypass) with compile-time byte offsets on both sides: one LDGSTS, no address arithmetic.
template <int SOFF, int GOFF>
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const float* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0+%2], [%1+%3], 16;" ::"r"(smem_dst), "l"(gmem_src), "n"(SOFF), "n"(GOFF)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const float* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
// 16-byte cp.async that reads `bytes` (0 or 16) from global memory and zero-fills the rest of the destination
__device__ __forceinline__ void cp_async16_zfill(uint32_t smem_dst, const float* gmem_src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// The kHR rows of query `qidx` that this half-warp multiplies go into the lane's private TMEM columns: slice s
// (floats [64s, 64s+64) of every row; the lane owns 4 of them per row) occupies columns [20s, 20s+20) in row-major
// (row, component) order, so one x16 + one x4 tcgen05.ld brings a whole slice back.  Rows >= nq are stored as zeros.
// Squared row norms are accumulated on the way and left in qn_s[0..kFT).  Nothing else of the query is ever re-read:
// per (warp, query) the 30 KB come from L2 once instead of once per pair.
template <int DT, bool FULL>
__device__ __forceinline__ void query_to_tmem(const FusedArgs& a, int qidx, int nq, int lane, uint32_t tq, float* qn_s) {
    const int h = lane >> 4, l16 = lane & 15;
    const int D = DT ? DT : a.D, d4 = D >> 2, nit = d4 >> 4;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* qb = reinterpret_cast<const float4*>(a.q + (size_t)qidx * a.Sq * D) + (size_t)(kHR * h) * d4 + l16;
    float2 s[kHR];
#pragma unroll
    for (int i = 0; i < kHR; ++i) s[i] = make_float2(0.f, 0.f);
#pragma unroll 3
    for (int it = 0; it < nit; ++it) {
        float4 v[kHR];
#pragma unroll
        for (int i = 0; i < kHR; ++i) v[i] = (FULL || kHR * h + i < nq) ? __ldg(qb + (size_t)i * d4 + (it << 4)) : zero4;
#pragma unroll
        for (int i = 0; i < kHR; ++i) {
            s[i] = __ffma2_rn(make_float2(v[i].x, v[i].y), make_float2(v[i].x, v[i].y), s[i]);
            s[i] = __ffma2_rn(make_float2(v[i].z, v[i].w), make_float2(v[i].z, v[i].w), s[i]);
            tc::tmem_st4(tq + it * kQCols + i * 4, v[i]);
        }
    }
    tc::tmem_wait_st();
#pragma unroll
    for (int i = 0; i < kHR; ++i) {
        float t = s[i].x + s[i].y;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (l16 == 0) qn_s[kHR * h + i] = t;
    }
    __syncwarp();
}

// Phase 1 of one tile: distances of `npairs` pairs -> Cs[p][kCostLd].  FULL: all documents have kFT sentences.
//
// Candidate rows stream through a per-warp shared-memory ring of kRing slices (one slice = the same 64 floats of
// all kFT rows = 2.5 KB) filled with cp.async (16 B per lane, L1 bypassed): kRing-1 slices are always in flight per
// warp (~80 KB per SM), which is what Little's law asks for at HBM latency.  The stream runs straight across pair
// boundaries of the tile; producer and consumer positions are plain running pointers (the five copies of a slice
// differ by compile-time offsets).
// The query comes from tensor memory (query_to_tmem): its slices are tcgen05.ld'ed one slice ahead into a pair of
// register buffers, so the inner loop issues no global or shared load for the query at all.
// Register plan per lane: 5x10 packed accumulators (100) + 10 packed candidate norms (20) + the current candidate
// slice (40, one 128-bit LDS per row; lanes l and l+16 read the same address) + two 5-row query slices (40).
template <int DT, bool FULL, int LD = kCostLd>
__device__ __forceinline__ void phase1(const FusedArgs& a, int base, int npairs, int my_ql, int my_cl, int my_ci, int lane, float* Cs,
                                       float* red, float* qn_s, float* ring, uint32_t tq, const int* lut_s) {
    const int h = lane >> 4, l16 = lane & 15;
    const int D = DT ? DT : a.D, d4 = D >> 2;
    const int nit = d4 >> 4;  // 64-float slices per row (16 lanes x float4)
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float2 zero2 = make_float2(0.f, 0.f);
    const int rev4 = (int)(__brev((unsigned)l16) >> 28);
    const size_t doc = (size_t)a.Sc * D;  // floats per candidate document

    // Half-warp h keeps candidate row (j + kHR*h) % kFT in position j: the labels of its accumulators rotate (the
    // epilogue's table undoes that), and each half squares only its positions 0..4, i.e. rows kHR*h .. kHR*h+4 -- every
    // candidate norm is accumulated exactly once per warp instead of once per half.
    float2 acc[kHR][kFT], cn[kHR];
    float4 qa[kHR], qb4[kHR];

    // ---- producer side of the ring: slices are issued in stream order (pair ip, slice iit) ----
    // lane l copies the 16-byte piece (row (l>>4) + 2m, float4 column l&15) for m = 0..4
    int ip = 0, iit = 0, islot = 0;
    const size_t lane_off = (size_t)h * D + l16 * 4;
    const float* gsrc = a.c + (size_t)__shfl_sync(0xffffffffu, my_ci, 0) * doc + lane_off;  // lane p holds pair p's document
    const uint32_t sring = tc::smem_u32(ring) + (uint32_t)lane * 16u;
    uint32_t sdst = sring;
    auto issue_next = [&]() {
        if (ip < npairs) {
            if (FULL) {
                if (DT) {
                    cp_async16<0 * 512, 0 * 8 * DT>(sdst, gsrc);
                    cp_async16<1 * 512, 1 * 8 * DT>(sdst, gsrc);
                    cp_async16<2 * 512, 2 * 8 * DT>(sdst, gsrc);
                    cp_async16<3 * 512, 3 * 8 * DT>(sdst, gsrc);
                    cp_async16<4 * 512, 4 * 8 * DT>(sdst, gsrc);
                } else {
#pragma unroll
                    for (int m = 0; m < kHR; ++m) cp_async16(sdst + m * 512, gsrc + (size_t)(2 * m) * D);
                }
            } else {
                // ragged tile: rows the document does not have are ZERO-FILLED in the ring (src-size 0: nothing is read),
                // so the consumer below runs the same unpredicated instruction stream as for full tiles
                const int ncp = __shfl_sync(0xffffffffu, my_cl, ip);
#pragma unroll
                for (int m = 0; m < kHR; ++m) {
                    const bool ok = h + 2 * m < ncp;
                    cp_async16_zfill(sdst + m * 512, gsrc + (ok ? (size_t)(2 * m) * D : (size_t)0), ok ? 16 : 0);
                }
            }
            gsrc += 64;
            if (++iit == nit) {  // next document of the tile (consecutive without an index list, anywhere with one)
                iit = 0;
                ++ip;
                gsrc = a.c + (size_t)__shfl_sync(0xffffffffu, my_ci, min(ip, npairs - 1)) * doc + lane_off;
            }
        }
        cp_async_commit();  // always commit: the wait below counts groups
        sdst += kSliceBytes;
        if (++islot == kRing) {
            islot = 0;
            sdst = sring;
        }
    };
    // consumer side
    int cslot = 0;
    const float4* const cring = reinterpret_cast<const float4*>(ring) + l16;
    const float4* cptr = cring;
    const int rotA = kHR * h * 16, rotB = kHR * (1 - h) * 16;  // float4 offsets of rows kHR*h.. and of the other five

    // one 64-float slice of the current pair (nc_cur rows valid): wait for it, refill the slot freed by the previous
    // slice, pull the rows out of the ring and multiply with the query slice q
    auto slice = [&](const float4 (&q)[kHR], int nc_cur) {
        cp_async_wait<kRing - 2>();
        __syncwarp();
        issue_next();
        float4 cv[kFT];
#pragma unroll
        for (int j = 0; j < kHR; ++j) {
            cv[j] = cptr[rotA + j * 16];
            cv[kHR + j] = cptr[rotB + j * 16];
        }
        cptr += kSliceFloats / 4;
        if (++cslot == kRing) {
            cslot = 0;
            cptr = cring;
        }
#pragma unroll
        for (int j = 0; j < kFT; ++j) {
            const float2 c0 = make_float2(cv[j].x, cv[j].y), c1 = make_float2(cv[j].z, cv[j].w);
#pragma unroll
            for (int i = 0; i < kHR; ++i) acc[i][j] = __ffma2_rn(make_float2(q[i].x, q[i].y), c0, acc[i][j]);
            if (j < kHR) cn[j] = __ffma2_rn(c0, c0, cn[j]);
#pragma unroll
            for (int i = 0; i < kHR; ++i) acc[i][j] = __ffma2_rn(make_float2(q[i].z, q[i].w), c1, acc[i][j]);
            if (j < kHR) cn[j] = __ffma2_rn(c1, c1, cn[j]);
        }
    };

#pragma unroll 1
    for (int k = 0; k < kRing - 1; ++k) issue_next();

    // per-pair state (uniform across the warp): current pair / next pair
    int cur_q = base / a.q_group;
    int nq_p = kFT, nc_p = kFT;
    if (!FULL) {
        nq_p = __shfl_sync(0xffffffffu, my_ql, 0);
        nc_p = __shfl_sync(0xffffffffu, my_cl, 0);
    }
    query_to_tmem<DT, FULL>(a, cur_q, nq_p, lane, tq, qn_s);
    tc::tmem_ld20_issue(tq, qa);

    for (int p = 0; p < npairs; ++p) {
#pragma unroll
        for (int i = 0; i < kHR; ++i)
#pragma unroll
            for (int j = 0; j < kFT; ++j) acc[i][j] = zero2;
#pragma unroll
        for (int j = 0; j < kHR; ++j) cn[j] = zero2;
        const bool more = p + 1 < npairs;
        const int next_q = more ? (base + p + 1) / a.q_group : cur_q;
        const bool same_q = more && next_q == cur_q;
        // slices 0 .. nit-1 of this pair, two per iteration (query slices ping-pong between qa and qb4; each is
        // requested from TMEM one slice ahead, right after the wait for the one about to be used)
        for (int it = 0; it < nit; it += 2) {
            tc::tmem_ld20_wait(qa);
            tc::tmem_ld20_issue(tq + (it + 1) * kQCols, qb4);  // it+1 < nit because nit is even
            slice(qa, nc_p);
            tc::tmem_ld20_wait(qb4);
            if (it + 2 < nit)
                tc::tmem_ld20_issue(tq + (it + 2) * kQCols, qa);
            else if (same_q)
                tc::tmem_ld20_issue(tq, qa);  // last slice of the pair: the next pair's first query slice
            slice(qb4, nc_p);
        }
        // pair finished: reduce over each half-warp, turn Gram values into distances
        float v[kRedVals];
#pragma unroll
        for (int i = 0; i < kHR; ++i)
#pragma unroll
            for (int j = 0; j < kFT; ++j) v[i * kFT + j] = acc[i][j].x + acc[i][j].y;
#pragma unroll
        for (int j = 0; j < kHR; ++j) v[kHR * kFT + j] = cn[j].x + cn[j].y;
#pragma unroll
        for (int e = kHR * kFT + kHR; e < kRedVals; ++e) v[e] = 0.f;
        transpose_reduce_w<kRedVals, 16>(v, lane);
        __syncwarp();  // the previous pair's readers of red[] are done
#pragma unroll
        for (int m = 0; m < kRedVals / 16; ++m) red[h * kRedVals + 16 * m + rev4] = v[m];
        __syncwarp();
        // entry e = lane + 32k of the 10x10 tile; lut_s packs (index of its dot product in red[], index of |c_j|^2 in
        // red[], j, i)
        float* row = Cs + p * LD + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int pk = lut_s[k * 32 + lane];
            if (pk >= 0) {
                const int i = pk >> 24, j = (pk >> 16) & 0xff;
                const float d2 = qn_s[i] + red[(pk >> 8) & 0xff] - 2.f * red[pk & 0xff];
                const float d = sqrt_approx(fmaxf(d2, 1e-8f));
                row[32 * k] = (FULL || (i < nq_p && j < nc_p)) ? d : 1.0e30f;
            }
        }
        if (more) {
            if (!FULL) {
                nq_p = __shfl_sync(0xffffffffu, my_ql, p + 1);
                nc_p = __shfl_sync(0xffffffffu, my_cl, p + 1);
            }
            if (!same_q) {  // the tile crosses into the next query's pool: swap the query held in TMEM
                __syncwarp();  // every lane has read qn_s for the finished pair
                cur_q = next_q;
                query_to_tmem<DT, FULL>(a, cur_q, nq_p, lane, tq, qn_s);
                tc::tmem_ld20_issue(tq, qa);
            }
        }
    }
    cp_async_wait<0>();  // only empty groups can still be pending; leave the ring quiescent for the next tile
    __syncwarp();
}

// ---- v7: Gram warps + Sinkhorn warps ---------------------------------------------------------------------------------------
// 12 warps per SM, all within 168 registers.  Warps 0-7 ("Gram" warps, two per scheduler) run phase 1 back to back on
// half-tiles of 16 pairs and between them keep the scheduler's FMA pipe busy (one warp alone cannot: it issues at
// most one FFMA2 per ~2.8 clk).  Warps 8-11 ("Sinkhorn" warps, one per scheduler) run phase 2 on the MUFU pipe, one
// pair per thread, with the cost tile STREAMED from shared memory in every step (solve_pair_thread_stream: ~100
// registers instead of ~200 -- that is what makes three warps per scheduler fit).  The hand-over is a ring of four
// 16-pair half-tiles per scheduler: Gram warps claim ring slots with a ticket, the Sinkhorn warp consumes tickets in
// order, two at a time (lanes 0-15 / 16-31); full/empty mbarriers per slot.
constexpr int kV7Gram = 8, kV7Warps = 12;
constexpr int kV7Half = 16;                 // pairs per Gram tile
constexpr int kV7Ld = kFT * kFT;            // cost floats per pair (16-byte aligned rows; LDS.128 conflict-free)
constexpr int kV7Slots = ASP_V7_SLOTS;      // half-tiles per scheduler
constexpr int kV7SlotFloats = kV7Half * kV7Ld;
// dynamic shared memory, floats: [4 schedulers][kV7Slots] half-tiles | per Gram warp: reduced values, query norms, ring
constexpr int kV7GramSmem = 2 * kRedVals + 16 + kRing * kSliceFloats;
constexpr int kV7RowSlots = 3;               // pairs a Sinkhorn warp solves ten-lanes-per-pair when a pass is that small
constexpr int kV7Scratch = kV7RowSlots * kV7Ld;  // floats of exponential scratch per Sinkhorn warp
constexpr int kV7Smem = 4 * kV7Slots * kV7SlotFloats + kV7Gram * kV7GramSmem + 4 * kV7Scratch;
static_assert((4 * kV7Slots * kV7SlotFloats) % 4 == 0 && (2 * kRedVals + 16) % 4 == 0 && kV7GramSmem % 4 == 0, "16-byte alignment");

constexpr int kV7GramRegs = 200, kV7SinkRegs = 104;  // 8 x 200 + 4 x 104 = 2016 <= 2048 registers per lane slot
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// (inlined into the Sinkhorn branch: ptxas must see them under that branch's setmaxnreg budget)
__device__ __forceinline__ void v7_phase2(float* Cs, int ql, int cl, int b, int Sq, int Sc, const float* eps_s, int n_eps,
                                       float inv_temp, const OtOut* out) {
    solve_pair_thread_stream<kFT, kFT, false>(Cs, ql, cl, b, Sq, Sc, eps_s, n_eps, inv_temp, *out, ASP_V7_TSCHED);
}
__device__ __forceinline__ void v7_phase2_full(float* Cs, int b, const float* eps_s, int n_eps, float inv_temp,
                                            const OtOut* out) {
    solve_pair_thread_stream<kFT, kFT, true>(Cs, kFT, kFT, b, kFT, kFT, eps_s, n_eps, inv_temp, *out, ASP_V7_TSCHED);
}

// ROWS: instantiation for launches of at most one pair per Gram warp (a single query against a <= 1k pool): the
// Sinkhorn warps then solve with ten lanes per pair (solve_pairs_rows, ~1/3 shorter launch).  Kept out of the
// large-batch instantiation, whose 104-register Sinkhorn loop it would crowd.
template <int DT, bool ROWS>
__global__ void __launch_bounds__(kV7Warps * 32, 1)
ot_fused_v7_kernel(const FusedArgs a, const EpsSched sched, const OtOut out) {
    extern __shared__ float smem[];
    __shared__ float eps_s[2 * ASP_MAX_EPS];  // eps[k], then log2e / eps[k] (the per-step division, done once here)
    __shared__ OtOut out_s;
    __shared__ int lut_s[128];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t full_bar[4][kV7Slots], empty_bar[4][kV7Slots];
    __shared__ int meta_s[4][kV7Slots][4];   // base, npairs (0 = end of a Gram warp's stream), full_tile
    __shared__ unsigned int ticket_s[4];
    for (int k = threadIdx.x; k < sched.n; k += blockDim.x) {
        eps_s[k] = sched.eps[k];
        eps_s[ASP_MAX_EPS + k] = kLog2e / sched.eps[k];
    }
    if (threadIdx.x == 0) out_s = out;
    if (threadIdx.x < 4) ticket_s[threadIdx.x] = 0u;
    {
        const int e = threadIdx.x;
        int pk = -1;
        if (e < kFT * kFT) {
            const int i = e / kFT, j = e - i * kFT, hh = i / kHR, ii = i - hh * kHR;
            const int jpos = (j + kFT - kHR * hh) % kFT;
            const int dot = hh * kRedVals + ii * kFT + jpos, nrm = (j / kHR) * kRedVals + kHR * kFT + (j % kHR);
            pk = dot | (nrm << 8) | (j << 16) | (i << 24);
        }
        if (e < 128) lut_s[e] = pk;
    }
    if (threadIdx.x < 4 * kV7Slots) {
        tc::mbar_init(&full_bar[0][0] + threadIdx.x, 1);
        tc::mbar_init(&empty_bar[0][0] + threadIdx.x, 1);
        tc::fence_barrier_init();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, kTmemCols);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const int w = warp & 3;  // scheduler (SM sub-partition) of this warp
    float* slots = smem + (size_t)w * kV7Slots * kV7SlotFloats;
    const int ntiles = (a.B + a.tile_pairs - 1) / a.tile_pairs;

    if (warp < kV7Gram) {
        // ------------------------------ Gram warp --------------------------------------------------------------------
        setmaxnreg_inc<kV7GramRegs>();
        float* red = smem + 4 * kV7Slots * kV7SlotFloats + (size_t)warp * kV7GramSmem;
        float* qn_s = red + 2 * kRedVals;
        float* ring = qn_s + 16;
        const uint32_t tq = tmem_slot + ((uint32_t)(w * 32) << 16) + (uint32_t)((warp >> 2) * 256);
        for (;;) {
            int tile = 0;
            unsigned int ticket = 0;
            if (lane == 0) {
                tile = (int)atomicAdd(a.counter, 1u);
                ticket = atomicAdd(&ticket_s[w], 1u);
            }
            tile = __shfl_sync(0xffffffffu, tile, 0);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            const int sl = ticket % kV7Slots;
            tc::mbar_wait_parked(&empty_bar[w][sl], ((ticket / kV7Slots) & 1u) ^ 1u);  // Sinkhorn warp done with the slot
            if (tile >= ntiles) {
                if (lane == 0) {
                    meta_s[w][sl][1] = 0;
                    tc::mbar_arrive(&full_bar[w][sl]);
                }
                break;
            }
            const int base = tile * a.tile_pairs;
            const int npairs = min(a.tile_pairs, a.B - base);
            int my_ql = 0, my_cl = 0, my_ci = 0;
            if (lane < npairs) {
                my_ci = a.c_index ? a.c_index[base + lane] : base + lane;
                my_ql = min(max(a.q_lens[(base + lane) / a.q_group], 0), a.Sq);
                my_cl = min(max(a.c_lens[my_ci], 0), a.Sc);
            }
            const bool full_tile = __all_sync(0xffffffffu, lane >= npairs || (my_ql == kFT && my_cl == kFT)) &&
                                   a.Sq == kFT && a.Sc == kFT;
            float* Cs = slots + sl * kV7SlotFloats;
            if (full_tile)
                phase1<DT, true, kV7Ld>(a, base, npairs, my_ql, my_cl, my_ci, lane, Cs, red, qn_s, ring, tq, lut_s);
            else
                phase1<DT, false, kV7Ld>(a, base, npairs, my_ql, my_cl, my_ci, lane, Cs, red, qn_s, ring, tq, lut_s);
            __syncwarp();
            if (lane == 0) {
                meta_s[w][sl][0] = base;
                meta_s[w][sl][1] = npairs;
                meta_s[w][sl][2] = full_tile ? 1 : 0;
                tc::mbar_arrive(&full_bar[w][sl]);  // release: tile + descriptor visible to the waiter
            }
        }
    } else {
        // ------------------------------ Sinkhorn warp: tickets in order, two half-tiles per pass ---------------------
        setmaxnreg_dec<kV7SinkRegs>();
        int ended = 0;  // Gram warps of this scheduler that have posted their end marker
        const int hsel = lane >> 4, l16 = lane & 15;
        for (unsigned int t0 = 0; ended < 2; t0 += 2) {
            int base[2] = {0, 0}, np[2] = {0, 0}, ft[2] = {1, 1};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (ended >= 2) break;
                const unsigned int t = t0 + k;
                const int sl = t % kV7Slots;
                tc::mbar_wait_parked(&full_bar[w][sl], (t / kV7Slots) & 1u);
                base[k] = meta_s[w][sl][0];
                np[k] = meta_s[w][sl][1];
                ft[k] = meta_s[w][sl][2];
                if (np[k] == 0) ++ended;
            }
            const int my_np = hsel ? np[1] : np[0], my_base = hsel ? base[1] : base[0];
            const int sl_mine = (t0 + hsel) % kV7Slots;
            bool solved = false;
            if (ROWS && np[0] + np[1] <= kV7RowSlots && np[0] + np[1] > 0) {
                // low-latency path: ten lanes per pair (lanes 10p .. 10p+9 <-> the p-th pair of this pass)
                const int p = lane / kFT, r = lane - p * kFT;
                const bool active = p < np[0] + np[1];
                const int from_b = (active && p >= np[0]) ? 1 : 0;            // which half-tile the pair sits in
                const int idx = active ? (from_b ? p - np[0] : p) : 0;        // its index there
                const int sl_p = (t0 + from_b) % kV7Slots;
                const float* Cp = slots + sl_p * kV7SlotFloats + idx * kV7Ld;
                const int b = (from_b ? base[1] : base[0]) + idx;
                int ql = 0, cl = 0;
                if (active) {
                    ql = min(max(a.q_lens[b / a.q_group], 0), a.Sq);
                    cl = min(max(a.c_lens[a.c_index ? a.c_index[b] : b], 0), a.Sc);
                }
                float* scratch = smem + 4 * kV7Slots * kV7SlotFloats + kV7Gram * kV7GramSmem + w * kV7Scratch +
                                 (active ? p : 0) * kV7Ld;
                solved = solve_pairs_rows<kFT>(Cp, ql, cl, b, a.Sq, a.Sc, active, min(p, kV7RowSlots - 1) * kFT, r, scratch,
                                               eps_s, sched.n, a.inv_temp, out_s, ASP_V7_TSCHED);
            }
            if (!solved && l16 < my_np) {
                float* Cs = slots + sl_mine * kV7SlotFloats + l16 * kV7Ld;
                const int b = my_base + l16;
                // warp-uniform choice keeps the two specialisations from serialising inside a warp
                if ((np[0] == 0 || ft[0]) && (np[1] == 0 || ft[1])) {
                    v7_phase2_full(Cs, b, eps_s, sched.n, a.inv_temp, &out_s);
                } else {
                    const int ql = min(max(a.q_lens[b / a.q_group], 0), a.Sq);
                    const int cl = min(max(a.c_lens[a.c_index ? a.c_index[b] : b], 0), a.Sc);
                    v7_phase2(Cs, ql, cl, b, a.Sq, a.Sc, eps_s, sched.n, a.inv_temp, &out_s);
                }
            }
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    if (np[k] > 0) tc::mbar_arrive(&empty_bar[w][(t0 + k) % kV7Slots]);
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_slot, kTmemCols);
}

bool ot_fused_supported(int Sq, int Sc, int D) {
    return Sq <= kFT && Sc <= kFT && D >= 128 && (D % 128) == 0 && D <= kMaxFusedD;
}

int ot_fused_launch(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens,
                    const int32_t* c_index, int B, int Sq, int Sc, int D, const EpsSched& sched, float temp, const OtOut& out,
                    cudaStream_t stream) {
    if (ot_fused_tc_supported(q_group, B, Sq, Sc, D))
        return ot_fused_tc_launch(q, q_lens, q_group, c, c_lens, c_index, B, Sq, Sc, D, sched, temp, out, stream);
    static std::atomic<unsigned int> next_slot{0};
    const int smem = kV7Smem * (int)sizeof(float);
    static thread_local int attr_dev = -1;
    static thread_local unsigned int* counters = nullptr;
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        ASP_CUDA(cudaFuncSetAttribute(ot_fused_v7_kernel<768, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ASP_CUDA(cudaFuncSetAttribute(ot_fused_v7_kernel<768, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ASP_CUDA(cudaFuncSetAttribute(ot_fused_v7_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ASP_CUDA(cudaFuncSetAttribute(ot_fused_v7_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ASP_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&counters), g_tile_counter));
        attr_dev = dev;
    }
    // Tile size: the batch is cut into the smallest number of whole waves of Gram warps (waves = ceil(B / (16 * warps)))
    // and every tile gets ceil(B / (waves * warps)) <= 16 pairs, so no warp runs one tile more than the others; small
    // batches spread down to one pair per warp so that a single 1 x 1k call still uses the whole GPU.
    const int max_ctas = sm_count();
    const int nwarps = max_ctas * kV7Gram;
    const int waves = (B + kV7Half * nwarps - 1) / (kV7Half * nwarps);
    const int tile_pairs = std::min(kV7Half, std::max(1, (B + waves * nwarps - 1) / (waves * nwarps)));
    // launch-owned tile counter, zeroed on this launch's stream: an aborted earlier launch cannot leave it armed, and
    // launches in flight on different streams never share one (1024 slots, round robin)
    unsigned int* counter = counters + (next_slot.fetch_add(1) % kCounterSlots);
    ASP_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
    FusedArgs a{q, q_lens, c, c_lens, c_index, q_group, B, Sq, Sc, D, counter, tile_pairs, 1.0f / temp};
    const int ntiles = (B + tile_pairs - 1) / tile_pairs;
    const int ctas = std::min(max_ctas, (ntiles + kV7Gram - 1) / kV7Gram);
    const bool rows = tile_pairs == 1;  // at most one pair per Gram warp: the low-latency instantiation
    if (D == 768 && rows)
        ot_fused_v7_kernel<768, true><<<ctas, kV7Warps * 32, smem, stream>>>(a, sched, out);
    else if (D == 768)
        ot_fused_v7_kernel<768, false><<<ctas, kV7Warps * 32, smem, stream>>>(a, sched, out);
    else if (rows)
        ot_fused_v7_kernel<0, true><<<ctas, kV7Warps * 32, smem, stream>>>(a, sched, out);
    else
        ot_fused_v7_kernel<0, false><<<ctas, kV7Warps * 32, smem, stream>>>(a, sched, out);
    ASP_LAUNCH_CHECK("ot_fused_v7_kernel");
    return ASP_OK;
}

}  // namespace asp

extern "C" size_t asp_ot_score_workspace_bytes(int B, int Sq, int Sc, int D) {
    if (asp::ot_fused_supported(Sq, Sc, D) && asp::g_ot_kernel != 1) return 0;
    if (asp::ot_varlen_supported(Sq, Sc, D) && asp::g_ot_kernel == 0) return asp::ot_varlen_workspace_bytes(B);
    return (size_t)B * Sq * Sc * sizeof(float);
}

extern "C" int asp_ot_score(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens,
                            int B, int Sq, int Sc, int D, const float* eps_host, int n_eps, float temp,
                            const asp_ot_outputs* out, void* workspace, size_t workspace_bytes, asp_stream_t stream) {
    if (B == 0) return ASP_OK;
    int rc = asp::check_pair_args(q, q_lens, c, c_lens, B, Sq, Sc, D);
    if (rc) return rc;
    ASP_REQUIRE(out, "asp_ot_score: out is NULL");
    ASP_REQUIRE(q_group >= 1, "asp_ot_score: q_group must be >= 1 (got %d)", q_group);
    ASP_REQUIRE(temp > 0.f, "asp_ot_score: temp must be > 0");
    if (Sq > ASP_MAX_SENTS || Sc > ASP_MAX_SENTS) {
        asp::set_error("ot: %dx%d sentences exceeds ASP_MAX_SENTS=%d", Sq, Sc, ASP_MAX_SENTS);
        return ASP_ERR_UNSUPPORTED;
    }
    asp::EpsSched sched;
    rc = asp::make_sched(eps_host, n_eps, &sched);
    if (rc) return rc;
    if (B == 0) return ASP_OK;
    const asp::OtOut o = asp::to_out(out);
    if (asp::ot_fused_supported(Sq, Sc, D) && asp::g_ot_kernel != 1)
        return asp::ot_fused_launch(q, q_lens, q_group, c, c_lens, nullptr, B, Sq, Sc, D, sched, temp, o,
                                    (cudaStream_t)stream);
    if (asp::ot_varlen_supported(Sq, Sc, D) && asp::g_ot_kernel == 0)
        return asp::ot_varlen_launch(q, q_lens, q_group, c, c_lens, nullptr, B, Sq, Sc, D, sched, temp, o, workspace,
                                     workspace_bytes, (cudaStream_t)stream);
    const size_t need = (size_t)B * Sq * Sc * sizeof(float);
    ASP_REQUIRE(workspace && workspace_bytes >= need, "asp_ot_score: workspace of %zu bytes needed for %dx%d sentences",
                need, Sq, Sc);
    float* cost = static_cast<float*>(workspace);
    rc = asp::pair_cost_launch(q, q_lens, q_group, c, c_lens, B, Sq, Sc, D, cost, (cudaStream_t)stream);
    if (rc) return rc;
    return asp::launch_sinkhorn(cost, q_lens, q_group, c_lens, B, Sq, Sc, sched, temp, o, (cudaStream_t)stream);
}

// Q x C all-pairs mode: every query document against every candidate document, dual values only,
// scores[i * NC + j] = OT_eps(query i, candidate j).  Documents of <= 10 sentences with the split-operand workspace go to
// the tcgen05 kernel (ot_allpairs.cu: a candidate tile in shared memory serves 12 query documents; Gram matrices on the
// tensor cores, Sinkhorn on the MUFU pipe).  Otherwise -- longer documents, a single query, or a caller that only brought
// the asp_ot_score workspace -- one 1 x N launch per query on the caller's stream.
extern "C" int asp_ot_score_allpairs(const float* q, const int32_t* q_lens, int NQ, const float* c, const int32_t* c_lens,
                                     int NC, int Sq, int Sc, int D, const float* eps_host, int n_eps, float temp,
                                     float* scores, void* workspace, size_t workspace_bytes, asp_stream_t stream) {
    ASP_REQUIRE(scores, "asp_ot_score_allpairs: scores is NULL");
    ASP_REQUIRE(NQ >= 0 && NC >= 0, "asp_ot_score_allpairs: bad shape NQ=%d NC=%d", NQ, NC);
    if (NQ == 0 || NC == 0) return ASP_OK;
    if (NQ >= 2 && asp::g_ot_kernel == 0 && asp::ot_allpairs_supported(Sq, Sc, D) && workspace &&
        workspace_bytes >= asp::ot_allpairs_workspace_bytes(NQ, NC, Sq, Sc, D) && (long long)NQ * NC <= 0x7fffffffLL &&
        (long long)NC * Sc <= 0x7fffffffLL && (long long)NQ * Sq <= 0x7fffffffLL) {
        int rc = asp::check_pair_args(q, q_lens, c, c_lens, NC, Sq, Sc, D);
        if (rc) return rc;
        ASP_REQUIRE(temp > 0.f, "asp_ot_score_allpairs: temp must be > 0");
        asp::EpsSched sched;
        rc = asp::make_sched(eps_host, n_eps, &sched);
        if (rc) return rc;
        return asp::ot_allpairs_launch(q, q_lens, NQ, c, c_lens, NC, Sq, Sc, D, sched, temp, scores, workspace,
                                       (cudaStream_t)stream);
    }
    for (int i = 0; i < NQ; ++i) {
        asp_ot_outputs out = {};
        out.dual = scores + (size_t)i * NC;
        const int rc = asp_ot_score(q + (size_t)i * Sq * D, q_lens + i, NC > 0 ? NC : 1, c, c_lens, NC, Sq, Sc, D, eps_host,
                                    n_eps, temp, &out, workspace, workspace_bytes, stream);
        if (rc) return rc;
    }
    return ASP_OK;
}

// Pools as index lists into a corpus that stays in HBM: pair b = (query b / q_group, corpus document c_index[b]).  c is
// the WHOLE corpus [N,Sc,D] and c_lens its [N] lengths; nothing is gathered or copied -- the kernel's producer walks the
// index list.  Fused shapes only (Sq, Sc <= 10, D % 128 == 0, D <= 768); other shapes: gather on the caller's side.
extern "C" int asp_ot_score_indexed(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens,
                                    const int32_t* c_index, int B, int Sq, int Sc, int D, const float* eps_host, int n_eps,
                                    float temp, const asp_ot_outputs* out, asp_stream_t stream) {
    if (B == 0) return ASP_OK;
    int rc = asp::check_pair_args(q, q_lens, c, c_lens, B, Sq, Sc, D);
    if (rc) return rc;
    ASP_REQUIRE(out && c_index, "asp_ot_score_indexed: out / c_index is NULL");
    ASP_REQUIRE(q_group >= 1 && temp > 0.f, "asp_ot_score_indexed: q_group >= 1 and temp > 0 required");
    const bool fused = asp::ot_fused_supported(Sq, Sc, D);
    if (!fused && !asp::ot_varlen_supported(Sq, Sc, D)) {
        asp::set_error("asp_ot_score_indexed: %dx%d sentences, D=%d is not a fused shape", Sq, Sc, D);
        return ASP_ERR_UNSUPPORTED;
    }
    asp::EpsSched sched;
    rc = asp::make_sched(eps_host, n_eps, &sched);
    if (rc) return rc;
    if (!fused)
        return asp::ot_varlen_launch(q, q_lens, q_group, c, c_lens, c_index, B, Sq, Sc, D, sched, temp, asp::to_out(out),
                                     nullptr, 0, (cudaStream_t)stream);
    return asp::ot_fused_launch(q, q_lens, q_group, c, c_lens, c_index, B, Sq, Sc, D, sched, temp, asp::to_out(out),
                                (cudaStream_t)stream);
}
