// K2+K4 all-pairs: otAspire distances of EVERY query document against EVERY candidate document (BASELINE configs[3]:
// 1k queries x 1M candidates) -- the Gram matrices on tcgen05, the Sinkhorn solves on the MUFU pipe, one kernel.
//
// Replaces AllPairMaskedWasserstein.compute_distance (src/learning/facetid_models/pair_distances.py:21-92, + geomloss
// SamplesLoss) as reached for every query of a pool file against the whole corpus (src/pre_process/pp_gen_nearest.py:
// 131-204: one caching_score call per query, each re-reading all candidate encodings).  Here a candidate row that has
// reached shared memory serves 12 query documents at once, which is the regime where the tensor cores pay: the 1 x N
// kernel (ot_fused.cu) spends 1320 FFMA2 per pair on the Gram tile and is bound by fp32 issue; this kernel spends none.
//
// One persistent CTA per SM, warps specialised by role:
//   warp 0      TMA producer: per 32-wide K block the four operand tiles (query hi / lo, candidate hi / lo; bf16 halves
//               of the fp32 rows written once by split_rows_kernel) -> 2-stage mbarrier ring, 64B swizzle
//   warp 1      MMA issuer: 128 x 160 fp32 accumulator in TMEM = 12 query documents x 16 candidate documents of <= 10
//               sentences, hi.hi + hi.lo + lo.hi + lo.lo (fp32-equivalent dot products), double buffered (2 x 256 columns)
//               so tile t+1 is contracted while tile t is drained
//   warps 4-7   drain: one accumulator row (query sentence) per thread, tcgen05.ld 32 columns at a time,
//               C = sqrt(max(|q|^2 + |c|^2 - 2 q.c, 1e-8)) (geomloss' distance formula) scattered into per-pair 10 x 10 cost
//               tiles in shared memory: 32 pairs (2 query documents x 16 candidate documents) = one group = the work of
//               one Sinkhorn warp
//   warps 8..   Sinkhorn: one pair per thread, cost tile streamed from shared memory every step
//               (solve_pair_thread_stream: softmax marginals, eps-scaling loop, final extrapolation, dual value) -- the
//               same solver, instruction for instruction, as the 1 x N kernel's phase 2
// Groups go to the Sinkhorn warps round robin (every pair costs the same n_eps + 2 steps); a full / empty mbarrier pair per
// warp hands the 12.8 KB cost block over.  The MUFU pipe (100 ex2 + 20 lg2 per pair and step) is the bound: the Gram
// contraction of a 192-pair tile takes ~1/5 of the time its 192 solves take.
// Tiles are walked candidate-tile-major, so the query operands (NQ x Sq x D x 4 B, 31 MB for 1k queries) stay in L2 and
// every candidate tile is fetched from HBM once.
#include "allpairs_common.cuh"
#include "ot_pair.cuh"

namespace asp {

constexpr int kOaStages = 2;
constexpr int kOaFT = 10;                    // sentences per document handled by the per-thread solver
constexpr int kOaDocsM = 12, kOaDocsN = 16;  // documents per tile
constexpr int kOaGroups = kOaDocsM / 2;      // groups of 32 pairs per tile
constexpr int kOaLd = kOaFT * kOaFT;         // cost floats per pair
constexpr int kOaFrontWarps = 8;             // producer, MMA, 2 spare, 4 drain
constexpr int kOaAccCols = 256;              // TMEM columns per accumulator buffer
// Waits of the front warps are long (the Sinkhorn warps set the pace).  ns > 0: sleep-polling between barrier tests
// instead of the hinted try_wait -- measured to make no difference (3.82e8 vs 3.84e8 pairs/s, profiles/r02_r_polling_ab.txt)
#ifndef ASP_OA_IDLE_FRONT
#define ASP_OA_IDLE_FRONT 0
#endif
#ifndef ASP_OA_IDLE_DRAIN
#define ASP_OA_IDLE_DRAIN 0
#endif
// 1: a quarter of the exponentials on the FMA / ALU pipes (ot_pair.cuh: ex2_poly2) to relieve the MUFU pipe.  Measured
// SLOWER with 12 Sinkhorn warps (3.64e8 vs 3.83e8 pairs/s, profiles/r02_3d_otallpairs_poly_ab.txt): 15 issue slots per
// two exponentials instead of 2 -- at 0.75 of the MUFU bound the warps are as short of issue slots as of MUFU results.
#ifndef ASP_OA_POLY
#define ASP_OA_POLY 0
#endif
template <int NS>
__device__ __forceinline__ void oa_wait(uint64_t* bar, uint32_t parity) {
    if constexpr (NS == 0) mbar_wait_parked(bar, parity);
    else mbar_wait_sleep<NS>(bar, parity);
}

struct OtAllPairsArgs {
    const float* qn;        // [NQ*Sq] squared norms of the query sentence rows
    const float* cn;        // [NC*Sc]
    const int32_t* q_lens;  // [NQ]
    const int32_t* c_lens;  // [NC]
    int NQ, NC, Sq, Sc, D;
    int nqt, nct;           // query / candidate tiles
    float inv_temp;
    float* scores;          // [NQ, NC] dual values
};

template <int NW>
constexpr int oa_smem_bytes() { return kOaStages * kApStage + NW * 32 * kOaLd * 4 + 1024; }

__device__ __forceinline__ void oa_phase2(float* Cs, int ql, int cl, int b, int Sq, int Sc, const float* eps_s, int n_eps,
                                          float inv_temp, const OtOut& out) {
    solve_pair_thread_stream<kOaFT, kOaFT, false, ASP_OA_POLY != 0>(Cs, ql, cl, b, Sq, Sc, eps_s, n_eps, inv_temp, out);
}
__device__ __forceinline__ void oa_phase2_full(float* Cs, int b, const float* eps_s, int n_eps, float inv_temp,
                                               const OtOut& out) {
    solve_pair_thread_stream<kOaFT, kOaFT, true, ASP_OA_POLY != 0>(Cs, kOaFT, kOaFT, b, kOaFT, kOaFT, eps_s, n_eps, inv_temp, out);
}

template <int NW>
__global__ void __launch_bounds__((kOaFrontWarps + NW) * 32, 1)
ot_allpairs_kernel(const __grid_constant__ CUtensorMap tq_hi, const __grid_constant__ CUtensorMap tq_lo,
                   const __grid_constant__ CUtensorMap tc_hi, const __grid_constant__ CUtensorMap tc_lo,
                   const OtAllPairsArgs g, const EpsSched sched) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ float eps_s[ASP_MAX_EPS];
    __shared__ float cn_s[2][kApBlockN];
    __shared__ uint64_t full[kOaStages], empty[kOaStages], acc_full[2], acc_empty[2], cfull[NW], cempty[NW];
    __shared__ uint32_t tmem_slot;
    float* cost = reinterpret_cast<float*>(smem + kOaStages * kApStage);  // [NW][32 pairs][10 x 10]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Sq = g.Sq, Sc = g.Sc;
    const int rows_m = kOaDocsM * Sq, cols_n = kOaDocsN * Sc;
    const int kblocks = g.D / kApBlockK;
    const int ntiles = g.nqt * g.nct;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    for (int k = threadIdx.x; k < sched.n; k += blockDim.x) eps_s[k] = sched.eps[k];
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tq_hi);
        tma_prefetch_desc(&tq_lo);
        tma_prefetch_desc(&tc_hi);
        tma_prefetch_desc(&tc_lo);
        for (int s = 0; s < kOaStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 128);
        }
        for (int w = 0; w < NW; ++w) {
            mbar_init(&cfull[w], 2 * Sq);  // one arrival per accumulator row of the group's two query documents
            mbar_init(&cempty[w], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&tmem_slot, 2 * kOaAccCols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    // NW = 12: the 20 warps start at 96 registers (launch bound); the front warps hand theirs to the Sinkhorn warps:
    // 8 x 56 + 12 x 120 = 1888 <= 2048 per lane slot (warpgroup-aligned: warps 0-7 / 8-19)
    if (warp < kOaFrontWarps) {
    if (NW > 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ------------------------------ TMA producer ---------------------------------------------------------------
        if (lane == 0) {
            int it = 0;
            for (int k = 0; k < my_tiles; ++k) {
                const int t = blockIdx.x + k * gridDim.x;
                const int m0 = (t % g.nqt) * rows_m, n0 = (t / g.nqt) * cols_n;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % kOaStages, ph = (it / kOaStages) & 1;
                    oa_wait<ASP_OA_IDLE_FRONT>(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(rows_m + cols_n) * (kApBlockK * 2) * 2);
                    uint8_t* sa = smem + s * kApStage;
                    tma_load_2d(sa, &tq_hi, &full[s], kb * kApBlockK, m0);
                    tma_load_2d(sa + kApABytes, &tq_lo, &full[s], kb * kApBlockK, m0);
                    tma_load_2d(sa + 2 * kApABytes, &tc_hi, &full[s], kb * kApBlockK, n0);
                    tma_load_2d(sa + 2 * kApABytes + kApBBytes, &tc_lo, &full[s], kb * kApBlockK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer -----------------------------------------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kApBlockM, kApBlockN);
            int it = 0;
            for (int k = 0; k < my_tiles; ++k) {
                const int buf = k & 1;
                oa_wait<ASP_OA_IDLE_FRONT>(&acc_empty[buf], ((k >> 1) & 1) ^ 1);  // drain warps are done with this accumulator
                tc_fence_after_sync();
                const uint32_t acc = tmem_base + (uint32_t)(buf * kOaAccCols);
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % kOaStages, ph = (it / kOaStages) & 1;
                    oa_wait<ASP_OA_IDLE_FRONT>(&full[s], ph);
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(smem + s * kApStage);
                    const uint64_t a_hi = umma_desc_sw64(sa), a_lo = umma_desc_sw64(sa + kApABytes);
                    const uint64_t b_hi = umma_desc_sw64(sa + 2 * kApABytes), b_lo = umma_desc_sw64(sa + 2 * kApABytes + kApBBytes);
#pragma unroll
                    for (int kk = 0; kk < kApBlockK / 16; ++kk) {  // 16 bf16 = 32 bytes along K inside the swizzle atom
                        umma_bf16(acc, a_lo + 2 * kk, b_lo + 2 * kk, idesc, (kb | kk) != 0);  // small terms first
                        umma_bf16(acc, a_lo + 2 * kk, b_hi + 2 * kk, idesc, true);
                        umma_bf16(acc, a_hi + 2 * kk, b_lo + 2 * kk, idesc, true);
                        umma_bf16(acc, a_hi + 2 * kk, b_hi + 2 * kk, idesc, true);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else if (warp >= 4 && warp < kOaFrontWarps) {
        // ------------------------------ drain: TMEM accumulator -> per-pair cost tiles -----------------------------
        const int dt = threadIdx.x - 128;          // 0..127
        const int r = (warp & 3) * 32 + lane;      // accumulator row = TMEM lane (== dt)
        const bool in_tile = r < rows_m;
        const int qd = in_tile ? r / Sq : 0, i = r - qd * Sq;
        const int gi = qd >> 1;
        for (int k = 0; k < my_tiles; ++k) {
            const int t = blockIdx.x + k * gridDim.x;
            const int m0 = (t % g.nqt) * rows_m, n0 = (t / g.nqt) * cols_n;
            const int buf = k & 1;
            for (int c = dt; c < kApBlockN; c += 128)
                cn_s[buf][c] = (c < cols_n && n0 + c < g.NC * Sc) ? __ldg(g.cn + n0 + c) : 0.f;
            const float qn = (in_tile && m0 + r < g.NQ * Sq) ? __ldg(g.qn + m0 + r) : 0.f;
            const int G = k * kOaGroups + gi, w = G % NW, use = G / NW;
            if (in_tile) oa_wait<ASP_OA_IDLE_DRAIN>(&cempty[w], (use & 1) ^ 1);  // the Sinkhorn warp has finished its previous group
            asm volatile("bar.sync 1, 128;" ::: "memory");               // cn_s staged
            oa_wait<ASP_OA_IDLE_DRAIN>(&acc_full[buf], (k >> 1) & 1);
            tc_fence_after_sync();
            float* dst = cost + ((size_t)(w * 32 + (qd & 1) * kOaDocsN) * kOaLd + i * kOaFT);
            int j = 0, cdoc = 0;
#pragma unroll 1
            for (int c0 = 0; c0 < kApBlockN; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(buf * kOaAccCols + c0), v);
                if (in_tile) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        if (c0 + e < cols_n) {
                            const float d2 = qn + cn_s[buf][c0 + e] - 2.f * v[e];
                            dst[cdoc * kOaLd + j] = sqrtf(fmaxf(d2, 1e-8f));
                            if (++j == Sc) {
                                j = 0;
                                ++cdoc;
                            }
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before_sync();
            mbar_arrive(&acc_empty[buf]);
            if (in_tile) mbar_arrive(&cfull[w]);  // release: this row's 16 x Sc costs are visible to the Sinkhorn warp
        }
    }
    } else {
        // ------------------------------ Sinkhorn: one pair per thread ----------------------------------------------
        if (NW > 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
        const int w = warp - kOaFrontWarps;
        OtOut out = {};
        out.dual = g.scores;
        float* Cs = cost + (size_t)(w * 32 + lane) * kOaLd;
        const int total_groups = my_tiles * kOaGroups;
        for (int G = w, use = 0; G < total_groups; G += NW, ++use) {
            const int k = G / kOaGroups, gi = G - k * kOaGroups;
            const int t = blockIdx.x + k * gridDim.x;
            const int gq = (t % g.nqt) * kOaDocsM + 2 * gi + (lane >> 4), gc = (t / g.nqt) * kOaDocsN + (lane & 15);
            const bool valid = gq < g.NQ && gc < g.NC;
            int ql = 0, cl = 0;
            if (valid) {
                ql = min(max(__ldg(g.q_lens + gq), 0), Sq);
                cl = min(max(__ldg(g.c_lens + gc), 0), Sc);
            }
            const bool all_full = __all_sync(0xffffffffu, !valid || (ql == kOaFT && cl == kOaFT));
            mbar_wait_parked(&cfull[w], use & 1);
            if (valid) {
                const int b = gq * g.NC + gc;
                if (all_full) oa_phase2_full(Cs, b, eps_s, sched.n, g.inv_temp, out);
                else oa_phase2(Cs, ql, cl, b, Sq, Sc, eps_s, sched.n, g.inv_temp, out);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&cempty[w]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 2 * kOaAccCols);
}

bool ot_allpairs_supported(int Sq, int Sc, int D) {
    return Sq >= 1 && Sc >= 1 && Sq <= kOaFT && Sc <= kOaFT && D >= 64 && (D % 64) == 0;
}

size_t ot_allpairs_workspace_bytes(int NQ, int NC, int Sq, int Sc, int D) {
    return split_rows_bytes((size_t)NQ * Sq, D) + split_rows_bytes((size_t)NC * Sc, D);
}

int g_oa_warps = 12;  // asp_set_option("oa_warps"): Sinkhorn warps per CTA (12: three per scheduler, default; 8)

template <int NW>
static int oa_launch(const CUtensorMap& tq_hi, const CUtensorMap& tq_lo, const CUtensorMap& tc_hi, const CUtensorMap& tc_lo,
                     const OtAllPairsArgs& g, const EpsSched& sched, cudaStream_t stream) {
    constexpr int smem = oa_smem_bytes<NW>();
    static thread_local int attr_dev = -1;
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        ASP_CUDA(cudaFuncSetAttribute(ot_allpairs_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_dev = dev;
    }
    const long long ntiles = (long long)g.nqt * g.nct;
    const int ctas = (int)std::min<long long>(sm_count(), ntiles);
    ot_allpairs_kernel<NW><<<ctas, (kOaFrontWarps + NW) * 32, smem, stream>>>(tq_hi, tq_lo, tc_hi, tc_lo, g, sched);
    ASP_LAUNCH_CHECK("ot_allpairs_kernel");
    return ASP_OK;
}

int ot_allpairs_launch(const float* q, const int32_t* q_lens, int NQ, const float* c, const int32_t* c_lens, int NC, int Sq,
                       int Sc, int D, const EpsSched& sched, float temp, float* scores, void* workspace, cudaStream_t stream) {
    ASP_REQUIRE((long long)NQ * NC <= 0x7fffffffLL, "asp_ot_score_allpairs: NQ * NC = %lld exceeds 2^31-1; split the queries",
                (long long)NQ * NC);
    char* w = static_cast<char*>(workspace);
    SplitRows qs, cs;
    int rc;
    if ((rc = split_rows_launch(q, (size_t)NQ * Sq, D, &w, &qs, stream))) return rc;
    if ((rc = split_rows_launch(c, (size_t)NC * Sc, D, &w, &cs, stream))) return rc;
    CUtensorMap tq_hi, tq_lo, tc_hi, tc_lo;
    if ((rc = make_tmap_bf16_k32(&tq_hi, qs.hi, (uint64_t)NQ * Sq, D, kOaDocsM * Sq))) return rc;
    if ((rc = make_tmap_bf16_k32(&tq_lo, qs.lo, (uint64_t)NQ * Sq, D, kOaDocsM * Sq))) return rc;
    if ((rc = make_tmap_bf16_k32(&tc_hi, cs.hi, (uint64_t)NC * Sc, D, kOaDocsN * Sc))) return rc;
    if ((rc = make_tmap_bf16_k32(&tc_lo, cs.lo, (uint64_t)NC * Sc, D, kOaDocsN * Sc))) return rc;
    OtAllPairsArgs g{qs.norms, cs.norms, q_lens, c_lens, NQ, NC, Sq, Sc, D, (NQ + kOaDocsM - 1) / kOaDocsM,
                     (NC + kOaDocsN - 1) / kOaDocsN, 1.0f / temp, scores};
    return g_oa_warps == 12 ? oa_launch<12>(tq_hi, tq_lo, tc_hi, tc_lo, g, sched, stream)
                            : oa_launch<8>(tq_hi, tq_lo, tc_hi, tc_lo, g, sched, stream);
}

}  // namespace asp

extern "C" size_t asp_ot_score_allpairs_workspace_bytes(int NQ, int NC, int Sq, int Sc, int D) {
    if (NQ < 0 || NC < 0) return 0;
    if (asp::ot_allpairs_supported(Sq, Sc, D) && asp::g_ot_kernel == 0) return asp::ot_allpairs_workspace_bytes(NQ, NC, Sq, Sc, D);
    return asp_ot_score_workspace_bytes(NC, Sq, Sc, D);
}
