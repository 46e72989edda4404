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
//     split the embedding dimension; candidate rows arrive through 128-bit non-allocating loads (both halves read the
//     same addresses, so a row still crosses L2->SM once), the next 10-row slice is always in flight while the current
//     one is multiplied (slot j of cv[] is refilled as soon as row j has been consumed; the stream runs across pair
//     boundaries) and the pair after that is prefetched into L2.  The 5x10 Gram tile + squared norms accumulate in
//     packed fp32 (FFMA2: two k-partials per register pair), are transpose-reduced over the 16 lanes and leave
//     sqrt(max(|q|^2+|c|^2-2q.c, 1e-8)) in a shared cost tile [32][101] (odd stride: conflict-free in phase 2);
//   * phase 2 (math): each THREAD solves one pair entirely in registers (ot_pair.cuh): one ex2 per (i,j)
//     and step, no shuffles, no shared memory in the loop.
// The query rows come through L1 (30 KB per query, re-read by every pair of its pool); their squared norms are
// computed once per (warp, query) and kept in shared memory.
#include "gram.cuh"
#include "ot_pair.cuh"

namespace asp {

constexpr int kFT = 10;        // max sentences per document on the fused path
constexpr int kHR = kFT / 2;   // query rows per half-warp
constexpr int kFusedWarps = 4; // warps per CTA
constexpr int kCostLd = 101;   // floats per pair in the shared cost tile
constexpr int kRedVals = 64;   // 50 dot products + 10 candidate norms, padded for the 16-lane transpose-reduce
constexpr int kWarpSmem = 32 * kCostLd + 2 * kRedVals + 16;  // cost tile + reduced values per half + query norms
constexpr int kCounterSlots = 256;

__device__ unsigned int g_tile_counter[kCounterSlots];
__device__ unsigned int g_done_counter[kCounterSlots];

struct FusedArgs {
    const float* q;
    const int32_t* q_lens;
    const float* c;
    const int32_t* c_lens;
    int q_group, B, Sq, Sc, D, slot;
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

// Phase 2 lives in its own (non-inlined) function so that it gets a register allocation of its own: the solver wants
// ~200 registers for the 10x10 tile and the potentials, and must not share them with phase 1's live state.
__device__ __noinline__ void fused_phase2(const float* row, int ql, int cl, int b, int Sq, int Sc, const float* eps_s,
                                          int n_eps, float inv_temp, const OtOut* out) {
    solve_pair_thread<kFT, kFT>([&](int i, int j) { return row[i * kFT + j]; }, ql, cl, b, Sq, Sc, eps_s, n_eps, inv_temp,
                                *out);
}

template <int DT>  // embedding size known at compile time (0 = runtime a.D); D % 64 == 0
__global__ void __launch_bounds__(kFusedWarps * 32, 2)
ot_fused_kernel(const FusedArgs a, const EpsSched sched, const OtOut out) {
    extern __shared__ float smem[];
    __shared__ float eps_s[ASP_MAX_EPS];
    __shared__ OtOut out_s;
    for (int k = threadIdx.x; k < sched.n; k += blockDim.x) eps_s[k] = sched.eps[k];
    if (threadIdx.x == 0) out_s = out;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane >> 4, l16 = lane & 15;
    float* Cs = smem + (size_t)warp * kWarpSmem;  // cost tile of this warp's 32 pairs
    float* red = Cs + 32 * kCostLd;               // [2][kRedVals] reduced Gram values of the pair being finished
    float* qn_s = red + 2 * kRedVals;             // squared norms of the current query's rows
    const int D = DT ? DT : a.D, d4 = D >> 2;
    const int nit = d4 >> 4;  // 64-float slices per row (16 lanes x float4)
    const int ntiles = (a.B + 31) >> 5;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float2 zero2 = make_float2(0.f, 0.f);
    const int rev4 = (int)(__brev((unsigned)l16) >> 28);

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = (int)atomicAdd(&g_tile_counter[a.slot], 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= ntiles) break;
        const int base = tile << 5;
        const int npairs = min(32, a.B - base);
        // lane p keeps the lengths of pair base+p
        int my_ql = 0, my_cl = 0;
        if (lane < npairs) {
            my_ql = min(max(a.q_lens[(base + lane) / a.q_group], 0), a.Sq);
            my_cl = min(max(a.c_lens[base + lane], 0), a.Sc);
        }
        int cur_q = -1;  // query whose row norms are in qn_s

        // ---------------- phase 1: cost tiles of the 32 pairs ------------------------------------------------
        float2 acc[kHR][kFT], cn[kFT], qn[kHR];
        float4 cv[kFT];
        const int total = npairs * nit;
        auto chunk_ptr = [&](int n, int& nc) -> const float4* {
            const int p = n / nit, it = n - p * nit;
            const int cl_p = __shfl_sync(0xffffffffu, my_cl, p & 31);  // every lane takes part in the shuffle
            nc = (n < total) ? cl_p : 0;
            return reinterpret_cast<const float4*>(a.c + (size_t)(base + p) * a.Sc * D) + (it << 4) + l16;
        };
        {
            int nc0;
            const float4* cb0 = chunk_ptr(0, nc0);
#pragma unroll
            for (int j = 0; j < kFT; ++j) cv[j] = (j < nc0) ? ldg_stream(cb0 + (size_t)j * d4) : zero4;
        }
        bool need_qn = false;
        for (int n = 0; n < total; ++n) {
            const int p = n / nit, it = n - p * nit;
            const int nq = __shfl_sync(0xffffffffu, my_ql, p), nc = __shfl_sync(0xffffffffu, my_cl, p);
            const int qidx = (base + p) / a.q_group;
            int nc_next;
            const float4* cb_next = chunk_ptr(n + 1, nc_next);
            if (it == 0) {
#pragma unroll
                for (int i = 0; i < kHR; ++i) {
                    qn[i] = zero2;
#pragma unroll
                    for (int j = 0; j < kFT; ++j) acc[i][j] = zero2;
                }
#pragma unroll
                for (int j = 0; j < kFT; ++j) cn[j] = zero2;
                need_qn = (qidx != cur_q);
                if (p + 1 < npairs) {  // pull the next pair's candidate rows into L2 (128-byte lines)
                    const char* nxt = reinterpret_cast<const char*>(a.c + (size_t)(base + p + 1) * a.Sc * D);
                    const int nbytes = __shfl_sync(0xffffffffu, my_cl, (p + 1) & 31) * D * 4;
                    for (int o = lane * 128; o < nbytes; o += 32 * 128) prefetch_l2(nxt + o);
                }
            }
            const float4* qb = reinterpret_cast<const float4*>(a.q + (size_t)qidx * a.Sq * D) + (size_t)(kHR * h) * d4 +
                               (it << 4) + l16;
            float4 qv[kHR];
#pragma unroll
            for (int i = 0; i < kHR; ++i) qv[i] = (kHR * h + i < nq) ? __ldg(qb + (size_t)i * d4) : zero4;
#pragma unroll
            for (int j = 0; j < kFT; ++j) {
                const float4 cj = cv[j];
                cv[j] = (j < nc_next) ? ldg_stream(cb_next + (size_t)j * d4) : zero4;
                const float2 c0 = make_float2(cj.x, cj.y), c1 = make_float2(cj.z, cj.w);
#pragma unroll
                for (int i = 0; i < kHR; ++i) {
                    acc[i][j] = __ffma2_rn(make_float2(qv[i].x, qv[i].y), c0, acc[i][j]);
                    acc[i][j] = __ffma2_rn(make_float2(qv[i].z, qv[i].w), c1, acc[i][j]);
                }
                cn[j] = __ffma2_rn(c0, c0, cn[j]);
                cn[j] = __ffma2_rn(c1, c1, cn[j]);
            }
            if (need_qn) {
#pragma unroll
                for (int i = 0; i < kHR; ++i) {
                    qn[i] = __ffma2_rn(make_float2(qv[i].x, qv[i].y), make_float2(qv[i].x, qv[i].y), qn[i]);
                    qn[i] = __ffma2_rn(make_float2(qv[i].z, qv[i].w), make_float2(qv[i].z, qv[i].w), qn[i]);
                }
            }
            if (it == nit - 1) {  // pair finished: reduce over each half-warp, turn Gram values into distances
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
                if (need_qn) {  // once per (tile, query): row norms of the query, reduced over the 16 lanes
#pragma unroll
                    for (int i = 0; i < kHR; ++i) {
                        float s = qn[i].x + qn[i].y;
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                        if (l16 == 0) qn_s[kHR * h + i] = s;
                    }
                    cur_q = qidx;
                }
                __syncwarp();
                float* row = Cs + p * kCostLd;
                for (int e = lane; e < kFT * kFT; e += 32) {
                    const int i = e / kFT, j = e - i * kFT;
                    const int hh = i / kHR, ii = i - hh * kHR;
                    const float d2 = qn_s[i] + red[kHR * kFT + j] - 2.f * red[hh * kRedVals + ii * kFT + j];
                    row[e] = (i < nq && j < nc) ? sqrtf(fmaxf(d2, 1e-8f)) : 1.0e30f;
                }
            }
        }
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

bool ot_fused_supported(int Sq, int Sc, int D) { return Sq <= kFT && Sc <= kFT && D >= 64 && (D % 64) == 0; }

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
    FusedArgs a{q, q_lens, c, c_lens, q_group, B, Sq, Sc, D, (int)(next_slot.fetch_add(1) % kCounterSlots), 1.0f / temp};
    const int ntiles = (B + 31) / 32;
    const int max_ctas = 2 * sm_count();
    const int ctas = min(max_ctas, (ntiles + kFusedWarps - 1) / kFusedWarps);
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
