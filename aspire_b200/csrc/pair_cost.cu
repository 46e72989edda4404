// K2 / K3: pairwise sentence-sentence L2 cost (exact fp32 FMA) and the tsAspire masked max/argmax.
//
// One warp owns one (query, candidate) pair.  The 32 lanes split the embedding dimension (128-bit
// coalesced loads straight from global memory, no shared-memory staging: every candidate row is read
// exactly once), each lane keeps partial dot products for a TI x TJ tile of sentence pairs plus the
// partial squared norms, and a transpose-reduce over the warp (NV-ish shuffles instead of 5*NV) leaves
// each finished sum in exactly one lane.  Reference arithmetic replaced: torch.cdist
// (pair_distances.py:49-50,167) / geomloss distances() = sqrt(clamp_min(|x|^2-2x.y+|y|^2, 1e-8)).
#include <algorithm>
#include "gram.cuh"

namespace asp {

enum { MODE_COST = 0, MODE_L2MAX = 1 };

// red: per-warp shared scratch of NV floats.  Leaves dots / norms of the tile in red[].
template <int TI, int TJ, bool STREAM_C = true>
__device__ __forceinline__ void gram_tile_to_smem(const float* q, int nq, const float* c, int nc, int D, int lane,
                                                  float* red) {
    using T = GramTile<TI, TJ>;
    float v[T::NV];
#pragma unroll
    for (int e = 0; e < T::NV; ++e) v[e] = 0.f;
    gram_accumulate<TI, TJ, STREAM_C>(q, nq, c, nc, D, lane, v);
    transpose_reduce<T::NV>(v, lane);
    const int rev = __brev((unsigned)lane) >> 27;
    __syncwarp();
#pragma unroll
    for (int m = 0; m < T::NV / 32; ++m) red[32 * m + rev] = v[m];
    __syncwarp();
}

template <int TI, int TJ, int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
pair_cost_kernel(const float* __restrict__ q, const int32_t* __restrict__ q_lens, int q_group,
                 const float* __restrict__ c, const int32_t* __restrict__ c_lens, int B, int Sq, int Sc, int D,
                 float* __restrict__ cost, float* __restrict__ best, int32_t* __restrict__ flat_idx) {
    using T = GramTile<TI, TJ>;
    __shared__ float red_all[WARPS][T::NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* red = red_all[warp];
    for (int b = blockIdx.x * WARPS + warp; b < B; b += gridDim.x * WARPS) {
        const int qb = b / q_group;
        const int ql = min(max(q_lens[qb], 0), Sq), cl = min(max(c_lens[b], 0), Sc);
        const float* qbase = q + (size_t)qb * Sq * D;
        const float* cbase = c + (size_t)b * Sc * D;
        float* out = cost ? cost + (size_t)b * Sq * Sc : nullptr;
        float run_best = kPadNeg, run_ns = 0.f;
        int run_idx = 0x7fffffff;
        if (MODE == MODE_L2MAX && (ql == 0 || cl == 0)) run_idx = 0;
        for (int ti = 0; ti < ql; ti += TI) {
            for (int tj = 0; tj < cl; tj += TJ) {
                const int nq = min(TI, ql - ti), nc = min(TJ, cl - tj);
                gram_tile_to_smem<TI, TJ>(qbase + (size_t)ti * D, nq, cbase + (size_t)tj * D, nc, D, lane, red);
                for (int e = lane; e < T::kEntries; e += 32) {
                    const int i = e / TJ, j = e - i * TJ;
                    if (i < nq && j < nc) {
                        const float d2 = red[T::kEntries + i] + red[T::kEntries + TI + j] - 2.f * red[e];
                        const float dist = sqrtf(fmaxf(d2, 1e-8f));
                        const int gi = ti + i, gj = tj + j;
                        if (MODE == MODE_COST) {
                            out[gi * Sc + gj] = dist;
                        } else {
                            const float sim = -dist;
                            const int idx = gi * Sc + gj;
                            if (out) out[idx] = sim;
                            if (sim > run_best || (sim == run_best && idx < run_idx)) {
                                run_best = sim;
                                run_idx = idx;
                                run_ns = red[T::kEntries + i] + red[T::kEntries + TI + j];
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }
        // padding of the [Sq,Sc] output
        if (out) {
            const float padv = (MODE == MODE_COST) ? 0.f : kPadNeg;
            for (int e = lane; e < Sq * Sc; e += 32) {
                const int i = e / Sc, j = e - i * Sc;
                if (i >= ql || j >= cl) out[e] = padv;
            }
        }
        if (MODE == MODE_L2MAX) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, run_best, s);
                const int oi = __shfl_xor_sync(0xffffffffu, run_idx, s);
                const float on = __shfl_xor_sync(0xffffffffu, run_ns, s);
                if (ob > run_best || (ob == run_best && oi < run_idx)) {
                    run_best = ob;
                    run_idx = oi;
                    run_ns = on;
                }
            }
            if (l2max_needs_refine(run_best, run_ns) && run_idx != 0x7fffffff)  // warp-uniform after the butterfly
                run_best = l2max_refine(qbase + (size_t)(run_idx / Sc) * D, cbase + (size_t)(run_idx % Sc) * D, D, lane);
            if (lane == 0) {
                best[b] = run_best;
                if (flat_idx) flat_idx[b] = (run_idx == 0x7fffffff) ? 0 : run_idx;
            }
        }
    }
}

// Long documents (more than one TI x TJ tile per pair): ONE CTA per pair, its warps take the pair's tiles round-robin.
// With a warp per pair, 1184 resident warps each walk their own pair's ~100-180 KB several times and together
// overflow the L2 (measured: 8.7 % L2 hit rate, 1.7x the unique bytes read from HBM); a CTA per pair keeps four times
// fewer pairs in flight and the tiles of one pair re-use each other's rows while they are still in L1/L2.
template <int TI, int TJ, int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
pair_cost_cta_kernel(const float* __restrict__ q, const int32_t* __restrict__ q_lens, int q_group,
                     const float* __restrict__ c, const int32_t* __restrict__ c_lens, int B, int Sq, int Sc, int D,
                     float* __restrict__ cost, float* __restrict__ best, int32_t* __restrict__ flat_idx) {
    using T = GramTile<TI, TJ>;
    __shared__ float red_all[WARPS][T::NV];
    __shared__ float best_s[WARPS], ns_s[WARPS];
    __shared__ int idx_s[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* red = red_all[warp];
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const int qb = b / q_group;
        const int ql = min(max(q_lens[qb], 0), Sq), cl = min(max(c_lens[b], 0), Sc);
        const float* qbase = q + (size_t)qb * Sq * D;
        const float* cbase = c + (size_t)b * Sc * D;
        float* out = cost ? cost + (size_t)b * Sq * Sc : nullptr;
        float run_best = kPadNeg, run_ns = 0.f;
        int run_idx = 0x7fffffff;
        const int nti = (ql + TI - 1) / TI, ntj = (cl + TJ - 1) / TJ;
        for (int t = warp; t < nti * ntj; t += WARPS) {
            const int ti = (t / ntj) * TI, tj = (t % ntj) * TJ;
            const int nq = min(TI, ql - ti), nc = min(TJ, cl - tj);
            gram_tile_to_smem<TI, TJ, false>(qbase + (size_t)ti * D, nq, cbase + (size_t)tj * D, nc, D, lane, red);
            for (int e = lane; e < T::kEntries; e += 32) {
                const int i = e / TJ, j = e - i * TJ;
                if (i < nq && j < nc) {
                    const float d2 = red[T::kEntries + i] + red[T::kEntries + TI + j] - 2.f * red[e];
                    const float dist = sqrtf(fmaxf(d2, 1e-8f));
                    const int idx = (ti + i) * Sc + tj + j;
                    if (MODE == MODE_COST) {
                        out[idx] = dist;
                    } else {
                        const float sim = -dist;
                        if (out) out[idx] = sim;
                        if (sim > run_best || (sim == run_best && idx < run_idx)) {
                            run_best = sim;
                            run_idx = idx;
                            run_ns = red[T::kEntries + i] + red[T::kEntries + TI + j];
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (out) {  // padding of the [Sq,Sc] output
            const float padv = (MODE == MODE_COST) ? 0.f : kPadNeg;
            for (int e = threadIdx.x; e < Sq * Sc; e += WARPS * 32) {
                const int i = e / Sc, j = e - i * Sc;
                if (i >= ql || j >= cl) out[e] = padv;
            }
        }
        if (MODE == MODE_L2MAX) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, run_best, s);
                const int oi = __shfl_xor_sync(0xffffffffu, run_idx, s);
                const float on = __shfl_xor_sync(0xffffffffu, run_ns, s);
                if (ob > run_best || (ob == run_best && oi < run_idx)) {
                    run_best = ob;
                    run_idx = oi;
                    run_ns = on;
                }
            }
            if (lane == 0) {
                best_s[warp] = run_best;
                idx_s[warp] = run_idx;
                ns_s[warp] = run_ns;
            }
            __syncthreads();
            if (warp == 0) {  // every lane of warp 0 merges the WARPS candidates identically, then refines together
                run_best = best_s[0];
                run_idx = idx_s[0];
                run_ns = ns_s[0];
                for (int w = 1; w < WARPS; ++w)
                    if (best_s[w] > run_best || (best_s[w] == run_best && idx_s[w] < run_idx)) {
                        run_best = best_s[w];
                        run_idx = idx_s[w];
                        run_ns = ns_s[w];
                    }
                if (l2max_needs_refine(run_best, run_ns) && run_idx != 0x7fffffff)
                    run_best = l2max_refine(qbase + (size_t)(run_idx / Sc) * D, cbase + (size_t)(run_idx % Sc) * D, D, lane);
                if (lane == 0) {
                    best[b] = run_best;
                    if (flat_idx) flat_idx[b] = (run_idx == 0x7fffffff) ? 0 : run_idx;
                }
            }
            __syncthreads();
        }
    }
}

template <int MODE>
int launch_pair_cost(const float* q, const int32_t* q_lens, int q_group, const float* c,
                            const int32_t* c_lens, int B, int Sq, int Sc, int D, float* cost, float* best,
                            int32_t* flat_idx, cudaStream_t stream) {
    constexpr int WARPS = 4;
    const int blocks = (B + WARPS - 1) / WARPS;
    if (Sq <= 10 && Sc <= 10) {
        pair_cost_kernel<10, 10, MODE, WARPS><<<blocks, WARPS * 32, 0, stream>>>(q, q_lens, q_group, c, c_lens,
                                                                              B, Sq, Sc, D, cost, best, flat_idx);
    } else {
        const int ctas = std::min(B, 8 * sm_count());
        pair_cost_cta_kernel<8, 8, MODE, WARPS><<<ctas, WARPS * 32, 0, stream>>>(q, q_lens, q_group, c, c_lens, B, Sq, Sc,
                                                                                D, cost, best, flat_idx);
    }
    ASP_LAUNCH_CHECK("pair_cost_kernel");
    return ASP_OK;
}

int check_pair_args(const float* q, const int32_t* q_lens, const float* c, const int32_t* c_lens, int B,
                           int Sq, int Sc, int D) {
    ASP_REQUIRE(q && c && q_lens && c_lens, "pair kernels: NULL input pointer");
    ASP_REQUIRE(B >= 0 && Sq >= 1 && Sc >= 1, "pair kernels: bad shape B=%d Sq=%d Sc=%d", B, Sq, Sc);
    ASP_REQUIRE(D >= 4 && (D % 4) == 0, "pair kernels: D=%d must be a positive multiple of 4", D);
    ASP_REQUIRE(aligned16(q) && aligned16(c), "pair kernels: q/c must be 16-byte aligned");
    return ASP_OK;
}

int pair_cost_launch(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens, int B,
                     int Sq, int Sc, int D, float* cost, cudaStream_t stream) {
    return launch_pair_cost<MODE_COST>(q, q_lens, q_group, c, c_lens, B, Sq, Sc, D, cost, nullptr, nullptr, stream);
}

}  // namespace asp

extern "C" int asp_pair_cost(const float* q, const int32_t* q_lens, int q_broadcast, const float* c,
                             const int32_t* c_lens, int B, int Sq, int Sc, int D, float* cost,
                             asp_stream_t stream) {
    int rc = asp::check_pair_args(q, q_lens, c, c_lens, B, Sq, Sc, D);
    if (rc) return rc;
    ASP_REQUIRE(cost, "asp_pair_cost: cost is NULL");
    if (B == 0) return ASP_OK;
    return asp::launch_pair_cost<asp::MODE_COST>(q, q_lens, q_broadcast ? B : 1, c, c_lens, B, Sq, Sc, D, cost, nullptr,
                                                 nullptr, (cudaStream_t)stream);
}

extern "C" size_t asp_l2max_workspace_bytes(int B, int Sq, int Sc, int D) {
    return ((Sq > 10 || Sc > 10) && asp::ot_varlen_supported(Sq, Sc, D)) ? asp::ot_varlen_workspace_bytes(B) : 0;
}

extern "C" int asp_l2max_ws(const float* q, const int32_t* q_lens, int q_broadcast, const float* c, const int32_t* c_lens,
                            int B, int Sq, int Sc, int D, float* best, int32_t* flat_idx, float* pair_sims, void* workspace,
                            size_t workspace_bytes, asp_stream_t stream) {
    int rc = asp::check_pair_args(q, q_lens, c, c_lens, B, Sq, Sc, D);
    if (rc) return rc;
    ASP_REQUIRE(best, "asp_l2max: best is NULL");
    if (B == 0) return ASP_OK;
    if ((Sq > 10 || Sc > 10) && asp::ot_varlen_supported(Sq, Sc, D))  // long documents: rows staged once, no re-reads
        return asp::l2max_varlen_launch(q, q_lens, q_broadcast ? B : 1, c, c_lens, B, Sq, Sc, D, best, flat_idx, pair_sims,
                                        workspace, workspace_bytes, (cudaStream_t)stream);
    return asp::launch_pair_cost<asp::MODE_L2MAX>(q, q_lens, q_broadcast ? B : 1, c, c_lens, B, Sq, Sc, D, pair_sims, best,
                                                  flat_idx, (cudaStream_t)stream);
}

extern "C" int asp_l2max(const float* q, const int32_t* q_lens, int q_broadcast, const float* c,
                         const int32_t* c_lens, int B, int Sq, int Sc, int D, float* best, int32_t* flat_idx,
                         float* pair_sims, asp_stream_t stream) {
    return asp_l2max_ws(q, q_lens, q_broadcast, c, c_lens, B, Sq, Sc, D, best, flat_idx, pair_sims, nullptr, 0, stream);
}
