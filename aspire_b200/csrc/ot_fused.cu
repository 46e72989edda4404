// K2+K4 fused: otAspire score straight from the sentence representations -- the headline kernel.
//
// Replaces, in ONE launch and without ever writing the cost tensor to HBM,
//   pad mask + -cdist            src/learning/facetid_models/pair_distances.py:39-50
//   softmax marginals            pair_distances.py:56-60
//   geomloss SamplesLoss(...)    pair_distances.py:68-72 / 88-91   (eps-scaling Sinkhorn, restated in ot_pair.cuh)
//   plan + primal value          pair_distances.py:76-85
// for documents of at most kFT sentences (the reference's abstracts: 10-sentence synthetic config, CSFCube ~7).
//
// Work decomposition (HBM-bound design: every candidate row is read exactly once, 30 KB per pair, and only
// 4-8 bytes per pair are written):
//   * persistent grid of independent WARPS (2 CTAs x 4 warps per SM); a warp takes tiles of 32 pairs from a global
//     atomic counter, so memory-phase warps and math-phase warps of the same SM overlap (FMA pipe vs MUFU pipe);
//   * phase 1 (streaming): for each of its 32 pairs the two HALF-WARPS take 5 query rows each and their 16 lanes
//     split the embedding dimension; candidate rows are staged by cp.async through a per-warp shared-memory ring
//     (4 slices = 10 KB in flight per warp, ~80 KB per SM: enough bytes in flight for HBM latency; the stream runs
//     across pair boundaries) and read back with one 128-bit LDS per row.  The 5x10 Gram tile + squared norms accumulate in
//     packed fp32 (FFMA2: two k-partials per register pair), are transpose-reduced over the 16 lanes and leave
//     sqrt(max(|q|^2+|c|^2-2q.c, 1e-8)) in a shared cost tile [32][101] (odd stride: conflict-free in phase 2);
//   * phase 2 (math): each THREAD solves one pair entirely in registers (ot_pair.cuh): one ex2 per (i,j)
//     and step, no shuffles, no shared memory in the loop.
// The query rows come through L1 (30 KB per query, re-read by every pair of its pool); their squared norms are
// computed once per (warp, query) and kept in shared memory.
#include <algorithm>
#include "gram.cuh"
#include "ot_pair.cuh"

namespace asp {

constexpr int kFT = 10;        // max sentences per document on the fused path
constexpr int kHR = kFT / 2;   // query rows per half-warp
constexpr int kFusedWarps = 4; // warps per CTA
constexpr int kCostLd = 101;   // floats per pair in the shared cost tile
constexpr int kRedVals = 64;   // 50 dot products + 10 candidate norms, padded for the 16-lane transpose-reduce
constexpr int kRing = 5;                   // candidate slices in the per-warp cp.async ring (kRing-1 in flight)
constexpr int kSliceFloats = kFT * 64;     // one slice: 64 floats of each of the kFT rows
constexpr int kWarpSmem = 32 * kCostLd + 2 * kRedVals + 32 + kRing * kSliceFloats;  // cost tile, reduced values per
                                                                                   // half, 2 x query norms, the ring
constexpr int kCounterSlots = 256;

__device__ unsigned int g_tile_counter[kCounterSlots];
__device__ unsigned int g_done_counter[kCounterSlots];

struct FusedArgs {
    const float* q;
    const int32_t* q_lens;
    const float* c;
    const int32_t* c_lens;
    int q_group, B, Sq, Sc, D, slot;
    int tile_pairs;  // pairs per warp tile (32 when the batch fills the machine, fewer for small batches)
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

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Squared norms of the kFT rows of query `qidx` -> qn_s[0..kFT) (each half-warp takes kHR rows).
__device__ __forceinline__ void query_norms(const FusedArgs& a, int qidx, int nq, int D, int lane, float* qn_s) {
    const int h = lane >> 4, l16 = lane & 15, d4 = D >> 2;
    const float4* qb = reinterpret_cast<const float4*>(a.q + (size_t)qidx * a.Sq * D) + (size_t)(kHR * h) * d4 + l16;
    float2 s[kHR];
#pragma unroll
    for (int i = 0; i < kHR; ++i) s[i] = make_float2(0.f, 0.f);
    for (int k = 0; k < d4; k += 16) {
#pragma unroll
        for (int i = 0; i < kHR; ++i) {
            if (kHR * h + i < nq) {
                const float4 v = __ldg(qb + (size_t)i * d4 + k);
                s[i] = __ffma2_rn(make_float2(v.x, v.y), make_float2(v.x, v.y), s[i]);
                s[i] = __ffma2_rn(make_float2(v.z, v.w), make_float2(v.z, v.w), s[i]);
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kHR; ++i) {
        float t = s[i].x + s[i].y;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (l16 == 0) qn_s[kHR * h + i] = t;
    }
    __syncwarp();
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Phase 1 of one tile: distances of `npairs` pairs -> Cs[p][kCostLd].  FULL: all documents have kFT sentences.
//
// The candidate rows stream through a per-warp shared-memory ring of kRing slices (one slice = the same 64 floats of
// all kFT rows = 2.5 KB) filled with cp.async (16 B per lane, L1 bypassed): kRing-1 slices are always in flight per
// warp (~80 KB per SM), which is what Little's law asks for at HBM latency -- registers could only hold one slice
// ahead.  The stream runs straight across pair boundaries of the tile.
// Register plan per lane: 5x10 packed accumulators (100) + 10 packed candidate norms (20) + the current candidate
// slice (40, read from the ring with one 128-bit LDS per row; lanes l and l+16 read the same address) + two 5-row
// query slices (40, double buffered loads through L1).
template <int DT, bool FULL>
__device__ __forceinline__ void phase1(const FusedArgs& a, int base, int npairs, int my_ql, int my_cl, int lane, float* Cs,
                                       float* red, float* qn_s, float* ring) {
    const int h = lane >> 4, l16 = lane & 15;
    const int D = DT ? DT : a.D, d4 = D >> 2;
    const int nit = d4 >> 4;  // 64-float slices per row (16 lanes x float4)
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float2 zero2 = make_float2(0.f, 0.f);
    const int rev4 = (int)(__brev((unsigned)l16) >> 28);
    const size_t doc = (size_t)a.Sc * D;  // floats per candidate document

    float2 acc[kHR][kFT], cn[kFT];
    float4 qa[kHR], qb4[kHR];
    int cur_q = -1, qslot = 0;  // qn_s holds two sets of query norms: pair_setup runs one pair ahead of the epilogue

    // ---- producer side of the ring: slices are issued in stream order (pair ip, slice iit) ----
    int ip = 0, iit = 0, islot = 0;
    auto issue_next = [&]() {
        if (ip < npairs) {
            const int ncp = FULL ? kFT : __shfl_sync(0xffffffffu, my_cl, ip);
            const float* src = a.c + (size_t)(base + ip) * doc + (iit << 6);
            float* dst = ring + islot * kSliceFloats;
#pragma unroll
            for (int m = 0; m < kFT * 16 / 32; ++m) {  // 160 16-byte pieces per slice, 5 per lane
                const int id = lane + 32 * m, row = id >> 4, c16 = id & 15;
                if (FULL || row < ncp) cp_async16(dst + id * 4, src + (size_t)row * D + c16 * 4);
            }
            if (++iit == nit) {
                iit = 0;
                ++ip;
            }
        }
        cp_async_commit();  // always commit: the wait below counts groups
        islot = (islot + 1 == kRing) ? 0 : islot + 1;
    };
    int cslot = 0;

    // per-pair state (uniform across the warp)
    int nq = kFT, nc = kFT, qidx = 0;
    const float4* qptr = nullptr;  // query slice 0 of the pair being multiplied (+ i*d4 per row, + 16 per slice)

    auto pair_setup = [&](int p) {
        qidx = (base + p) / a.q_group;
        if (!FULL) {
            nq = __shfl_sync(0xffffffffu, my_ql, p);
            nc = __shfl_sync(0xffffffffu, my_cl, p);
        }
        qptr = reinterpret_cast<const float4*>(a.q + (size_t)qidx * a.Sq * D) + (size_t)(kHR * h) * d4 + l16;
        if (qidx != cur_q) {
            qslot ^= 1;
            query_norms(a, qidx, FULL ? kFT : nq, D, lane, qn_s + 16 * qslot);
            cur_q = qidx;
        }
    };
    auto load_q = [&](float4 (&dst)[kHR], const float4* src) {
#pragma unroll
        for (int i = 0; i < kHR; ++i) dst[i] = (FULL || kHR * h + i < nq) ? __ldg(src + (size_t)i * d4) : zero4;
    };
    // one 64-float slice of the current pair (nc_cur rows valid): wait for it, refill the slot freed by the previous
    // slice, pull the rows out of the ring and multiply with the query slice q
    auto slice = [&](const float4 (&q)[kHR], int nc_cur) {
        cp_async_wait<kRing - 2>();
        __syncwarp();
        issue_next();
        const float4* sl = reinterpret_cast<const float4*>(ring + cslot * kSliceFloats) + l16;
        cslot = (cslot + 1 == kRing) ? 0 : cslot + 1;
        float4 cv[kFT];
#pragma unroll
        for (int j = 0; j < kFT; ++j) cv[j] = (FULL || j < nc_cur) ? sl[j * 16] : zero4;
#pragma unroll
        for (int j = 0; j < kFT; ++j) {
            const float2 c0 = make_float2(cv[j].x, cv[j].y), c1 = make_float2(cv[j].z, cv[j].w);
#pragma unroll
            for (int i = 0; i < kHR; ++i) acc[i][j] = __ffma2_rn(make_float2(q[i].x, q[i].y), c0, acc[i][j]);
            cn[j] = __ffma2_rn(c0, c0, cn[j]);
#pragma unroll
            for (int i = 0; i < kHR; ++i) acc[i][j] = __ffma2_rn(make_float2(q[i].z, q[i].w), c1, acc[i][j]);
            cn[j] = __ffma2_rn(c1, c1, cn[j]);
        }
    };

#pragma unroll 1
    for (int k = 0; k < kRing - 1; ++k) issue_next();
    pair_setup(0);
    load_q(qa, qptr);

    for (int p = 0; p < npairs; ++p) {
#pragma unroll
        for (int i = 0; i < kHR; ++i)
#pragma unroll
            for (int j = 0; j < kFT; ++j) acc[i][j] = zero2;
#pragma unroll
        for (int j = 0; j < kFT; ++j) cn[j] = zero2;
        const int nq_p = nq, nc_p = nc;         // lengths / query-norm slot of the pair being accumulated
        const float* qn_p = qn_s + 16 * qslot;  // (pair_setup below moves on to the next pair)
        // slices 0 .. nit-1 of this pair, two per iteration (query slices ping-pong between qa and qb4)
        for (int it = 0; it < nit; it += 2) {
            load_q(qb4, qptr + ((it + 1) << 4));  // it+1 < nit because nit is even
            slice(qa, nc_p);
            if (it + 2 < nit) {
                load_q(qa, qptr + ((it + 2) << 4));
            } else if (p + 1 < npairs) {  // last slice of the pair: fetch the next pair's first query slice
                pair_setup(p + 1);
                load_q(qa, qptr);
            }
            slice(qb4, nc_p);
        }
        // pair finished: reduce over each half-warp, turn Gram values into distances
        float v[kRedVals];
#pragma unroll
        for (int i = 0; i < kHR; ++i)
#pragma unroll
            for (int j = 0; j < kFT; ++j) v[i * kFT + j] = acc[i][j].x + acc[i][j].y;
#pragma unroll
        for (int j = 0; j < kFT; ++j) v[kHR * kFT + j] = cn[j].x + cn[j].y;
#pragma unroll
        for (int e = kHR * kFT + kFT; e < kRedVals; ++e) v[e] = 0.f;
        transpose_reduce_w<kRedVals, 16>(v, lane);
        __syncwarp();  // the previous pair's readers of red[] are done
#pragma unroll
        for (int m = 0; m < kRedVals / 16; ++m) red[h * kRedVals + 16 * m + rev4] = v[m];
        __syncwarp();
        float* row = Cs + p * kCostLd;
        for (int e = lane; e < kFT * kFT; e += 32) {
            const int i = e / kFT, j = e - i * kFT;
            const int hh = i / kHR, ii = i - hh * kHR;
            const float d2 = qn_p[i] + red[kHR * kFT + j] - 2.f * red[hh * kRedVals + ii * kFT + j];
            row[e] = (i < nq_p && j < nc_p) ? sqrtf(fmaxf(d2, 1e-8f)) : 1.0e30f;
        }
    }
    cp_async_wait<0>();  // only empty groups can still be pending; leave the ring quiescent for the next tile
    __syncwarp();
}

// Phase 2 lives in its own (non-inlined) function so that it gets a register allocation of its own: the solver wants
// ~200 registers for the 10x10 tile and the potentials, and must not share them with phase 1's live state.
__device__ __noinline__ void fused_phase2(const float* row, int ql, int cl, int b, int Sq, int Sc, const float* eps_s,
                                          int n_eps, float inv_temp, const OtOut* out) {
    solve_pair_thread<kFT, kFT>([&](int i, int j) { return row[i * kFT + j]; }, ql, cl, b, Sq, Sc, eps_s, n_eps, inv_temp,
                                *out);
}

template <int DT>  // embedding size known at compile time (0 = runtime a.D); D % 128 == 0
__global__ void __launch_bounds__(kFusedWarps * 32, 2)
ot_fused_kernel(const FusedArgs a, const EpsSched sched, const OtOut out) {
    extern __shared__ float smem[];
    __shared__ float eps_s[ASP_MAX_EPS];
    __shared__ OtOut out_s;
    for (int k = threadIdx.x; k < sched.n; k += blockDim.x) eps_s[k] = sched.eps[k];
    if (threadIdx.x == 0) out_s = out;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* Cs = smem + (size_t)warp * kWarpSmem;  // cost tile of this warp's 32 pairs
    float* red = Cs + 32 * kCostLd;               // [2][kRedVals] reduced Gram values of the pair being finished
    float* qn_s = red + 2 * kRedVals;             // [2][16] squared norms of the current / next query's rows
    float* ring = qn_s + 32;                      // [kRing][kSliceFloats] candidate slices (16-byte aligned)
    const int ntiles = (a.B + a.tile_pairs - 1) / a.tile_pairs;

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = (int)atomicAdd(&g_tile_counter[a.slot], 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= ntiles) break;
        const int base = tile * a.tile_pairs;
        const int npairs = min(a.tile_pairs, a.B - base);
        // lane p keeps the lengths of pair base+p
        int my_ql = 0, my_cl = 0;
        if (lane < npairs) {
            my_ql = min(max(a.q_lens[(base + lane) / a.q_group], 0), a.Sq);
            my_cl = min(max(a.c_lens[base + lane], 0), a.Sc);
        }
        // uniform fast path: every pair of the tile has all kFT x kFT sentences (no predicates, no zero fill)
        const bool full_tile = __all_sync(0xffffffffu, lane >= npairs || (my_ql == kFT && my_cl == kFT)) &&
                               a.Sq == kFT && a.Sc == kFT;
        if (full_tile)
            phase1<DT, true>(a, base, npairs, my_ql, my_cl, lane, Cs, red, qn_s, ring);
        else
            phase1<DT, false>(a, base, npairs, my_ql, my_cl, lane, Cs, red, qn_s, ring);
        __syncwarp();

        // ---------------- phase 2: one pair per thread ---------------------------------------------------------
        if (lane < npairs) {
            fused_phase2(Cs + lane * kCostLd, my_ql, my_cl, base + lane, a.Sq, a.Sc, eps_s, sched.n, a.inv_temp, &out_s);
        }
        __syncwarp();
    }
    // the last warp to leave re-arms the counters for the next launch that uses this slot
    if (lane == 0) {
        const unsigned int total_warps = gridDim.x * kFusedWarps;
        if (atomicAdd(&g_done_counter[a.slot], 1u) == total_warps - 1) {
            g_tile_counter[a.slot] = 0u;
            g_done_counter[a.slot] = 0u;
            __threadfence();
        }
    }
}

bool ot_fused_supported(int Sq, int Sc, int D) { return Sq <= kFT && Sc <= kFT && D >= 128 && (D % 128) == 0; }

int ot_fused_launch(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens, int B,
                    int Sq, int Sc, int D, const EpsSched& sched, float temp, const OtOut& out, cudaStream_t stream) {
    static std::atomic<unsigned int> next_slot{0};
    const int smem = kFusedWarps * kWarpSmem * (int)sizeof(float);
    static thread_local int attr_dev = -1;
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        ASP_CUDA(cudaFuncSetAttribute(ot_fused_kernel<768>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ASP_CUDA(cudaFuncSetAttribute(ot_fused_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_dev = dev;
    }
    // Tile size: 32 pairs per warp once the batch can feed every resident warp; smaller batches are spread over more
    // warps (down to one pair per warp) so that a single-query call (1 x 1k candidates) still uses the whole GPU.
    const int max_ctas = 2 * sm_count();
    const int tile_pairs = std::min(32, std::max(1, (B + max_ctas * kFusedWarps - 1) / (max_ctas * kFusedWarps)));
    FusedArgs a{q, q_lens, c, c_lens, q_group, B, Sq, Sc, D, (int)(next_slot.fetch_add(1) % kCounterSlots), tile_pairs,
                1.0f / temp};
    const int ntiles = (B + tile_pairs - 1) / tile_pairs;
    const int ctas = std::min(max_ctas, (ntiles + kFusedWarps - 1) / kFusedWarps);
    if (D == 768)
        ot_fused_kernel<768><<<ctas, kFusedWarps * 32, smem, stream>>>(a, sched, out);
    else
        ot_fused_kernel<0><<<ctas, kFusedWarps * 32, smem, stream>>>(a, sched, out);
    ASP_LAUNCH_CHECK("ot_fused_kernel");
    return ASP_OK;
}

}  // namespace asp

extern "C" size_t asp_ot_score_workspace_bytes(int B, int Sq, int Sc, int D) {
    if (asp::ot_fused_supported(Sq, Sc, D) && asp::g_ot_kernel != 1) return 0;
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
        return asp::ot_fused_launch(q, q_lens, q_group, c, c_lens, B, Sq, Sc, D, sched, temp, o, (cudaStream_t)stream);
    const size_t need = (size_t)B * Sq * Sc * sizeof(float);
    ASP_REQUIRE(workspace && workspace_bytes >= need, "asp_ot_score: workspace of %zu bytes needed for %dx%d sentences",
                need, Sq, Sc);
    float* cost = static_cast<float*>(workspace);
    rc = asp::pair_cost_launch(q, q_lens, q_group, c, c_lens, B, Sq, Sc, D, cost, (cudaStream_t)stream);
    if (rc) return rc;
    return asp::launch_sinkhorn(cost, q_lens, q_group, c_lens, B, Sq, Sc, sched, temp, o, (cudaStream_t)stream);
}
