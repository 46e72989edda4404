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
//     atomic counter, so memory-phase warps and math-phase warps of the same SM overlap;
//   * phase 1 (streaming): for each of its 32 pairs the warp's lanes split the embedding dimension, read the
//     candidate rows with 128-bit non-allocating loads (the next 10-row chunk is always in flight while the
//     current one is multiplied: register double buffer that runs across pair boundaries), accumulate the
//     10x10 Gram tile + squared norms in exact fp32 FMA, transpose-reduce over the warp and leave
//     sqrt(max(|q|^2+|c|^2-2q.c, 1e-8)) in a shared cost tile [32][101] (odd stride: conflict-free in phase 2);
//   * phase 2 (math): each THREAD solves one pair entirely in registers (ot_pair.cuh): one ex2 per (i,j)
//     and step, no shuffles, no shared memory in the loop.
// The query rows come through L1 (30 KB per query, re-read by every pair of its pool).
#include "gram.cuh"
#include "ot_pair.cuh"

namespace asp {

constexpr int kFT = 10;        // max sentences per document on the fused path
constexpr int kFusedWarps = 4; // warps per CTA
constexpr int kCostLd = 101;   // floats per pair in the shared cost tile
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

__global__ void __launch_bounds__(kFusedWarps * 32, 2)
ot_fused_kernel(const FusedArgs a, const EpsSched sched, const OtOut out) {
    using T = GramTile<kFT, kFT>;
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* Cs = smem + (size_t)warp * (32 * kCostLd + T::NV);  // cost tile of this warp's 32 pairs
    float* red = Cs + 32 * kCostLd;                            // reduced Gram values of the pair being finished
    const int D = a.D, d4 = D >> 2;
    const int nit = (d4 + 31) >> 5;  // 128-float chunks per row
    const int ntiles = (a.B + 31) >> 5;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

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

        // ---------------- phase 1: cost tiles of the 32 pairs ------------------------------------------------
        // Software pipeline over "chunks" (pair p, 128-float slice it): while chunk n is multiplied, the candidate
        // rows of chunk n+1 are already in flight -- slot j of cv[] is refilled as soon as row j has been consumed,
        // and the stream runs straight across pair boundaries.
        float v[T::NV];
        float4 cv[kFT];
        const int total = npairs * nit;
        auto chunk_ptr = [&](int n, int& nc) -> const float4* {
            const int p = n / nit, it = n - p * nit;
            const int k4 = (it << 5) + lane;
            nc = (n < total && k4 < d4) ? __shfl_sync(0xffffffffu, my_cl, p & 31) : 0;
            return reinterpret_cast<const float4*>(a.c + (size_t)(base + p) * a.Sc * D) + k4;
        };
        {
            int nc0;
            const float4* cb0 = chunk_ptr(0, nc0);
#pragma unroll
            for (int j = 0; j < kFT; ++j) cv[j] = (j < nc0) ? ldg_stream(cb0 + (size_t)j * d4) : zero4;
        }
        for (int n = 0; n < total; ++n) {
            const int p = n / nit, it = n - p * nit;
            const int k4 = (it << 5) + lane;
            const int nq = __shfl_sync(0xffffffffu, my_ql, p), nc = __shfl_sync(0xffffffffu, my_cl, p);
            int nc_next;
            const float4* cb_next = chunk_ptr(n + 1, nc_next);
            if (it == 0) {
#pragma unroll
                for (int e = 0; e < T::NV; ++e) v[e] = 0.f;
            }
            const float4* qb = reinterpret_cast<const float4*>(a.q + (size_t)((base + p) / a.q_group) * a.Sq * D) + k4;
            float4 qv[kFT];
#pragma unroll
            for (int i = 0; i < kFT; ++i) qv[i] = (i < nq && k4 < d4) ? __ldg(qb + (size_t)i * d4) : zero4;
#pragma unroll
            for (int j = 0; j < kFT; ++j) {
                const float4 cj = cv[j];
                cv[j] = (j < nc_next) ? ldg_stream(cb_next + (size_t)j * d4) : zero4;
#pragma unroll
                for (int i = 0; i < kFT; ++i) v[i * kFT + j] = dot4(qv[i], cj, v[i * kFT + j]);
                v[T::kEntries + kFT + j] = dot4(cj, cj, v[T::kEntries + kFT + j]);
            }
#pragma unroll
            for (int i = 0; i < kFT; ++i) v[T::kEntries + i] = dot4(qv[i], qv[i], v[T::kEntries + i]);
            if (it == nit - 1) {  // pair finished: reduce over the warp, turn Gram values into distances
                transpose_reduce<T::NV>(v, lane);
                const int rev = __brev((unsigned)lane) >> 27;
                __syncwarp();
#pragma unroll
                for (int m = 0; m < T::NV / 32; ++m) red[32 * m + rev] = v[m];
                __syncwarp();
                float* row = Cs + p * kCostLd;
                for (int e = lane; e < T::kEntries; e += 32) {
                    const int i = e / kFT, j = e - i * kFT;
                    const float d2 = red[T::kEntries + i] + red[T::kEntries + kFT + j] - 2.f * red[e];
                    row[e] = (i < nq && j < nc) ? sqrtf(fmaxf(d2, 1e-8f)) : 1.0e30f;
                }
            }
        }
        __syncwarp();

        // ---------------- phase 2: one pair per thread ---------------------------------------------------------
        if (lane < npairs) {
            const float* row = Cs + lane * kCostLd;
            solve_pair_thread<kFT, kFT>([&](int i, int j) { return row[i * kFT + j]; }, my_ql, my_cl, base + lane, a.Sq,
                                        a.Sc, sched, a.inv_temp, out);
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

bool ot_fused_supported(int Sq, int Sc, int D) { return Sq <= kFT && Sc <= kFT && (D % 4) == 0; }

int ot_fused_launch(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens, int B,
                    int Sq, int Sc, int D, const EpsSched& sched, float temp, const OtOut& out, cudaStream_t stream) {
    static std::atomic<unsigned int> next_slot{0};
    using T = GramTile<kFT, kFT>;
    const int smem = kFusedWarps * (32 * kCostLd + T::NV) * (int)sizeof(float);
    static thread_local int attr_dev = -1;
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        ASP_CUDA(cudaFuncSetAttribute(ot_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_dev = dev;
    }
    FusedArgs a{q, q_lens, c, c_lens, q_group, B, Sq, Sc, D, (int)(next_slot.fetch_add(1) % kCounterSlots), 1.0f / temp};
    const int ntiles = (B + 31) / 32;
    const int max_ctas = 2 * sm_count();
    const int ctas = min(max_ctas, (ntiles + kFusedWarps - 1) / kFusedWarps);
    ot_fused_kernel<<<ctas, kFusedWarps * 32, smem, stream>>>(a, sched, out);
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
