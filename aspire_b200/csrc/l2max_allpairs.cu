// K2+K3 all-pairs: tsAspire scores of EVERY query document against EVERY candidate document (config 3:
// 1k queries x 100k candidates) -- a dense [Q*S, D] x [D, C*S] contraction with a segmented max epilogue, on tcgen05.
//
// Replaces the numpy path of src/pre_process/pp_gen_nearest.py: `-scipy cdist(query_sents, pool_sents)` (:942; the
// all-queries x all-corpus variant :788-795) followed by the per-candidate column-slice `np.max` (:949-961), and
// allpair_masked_dist_l2max (pair_distances.py:138-186) applied to all Q x C pairs.
//
// Documents are rows-of-S sentence blocks ([N, S, D] zero padded), so a 128-row A tile holds floor(128/S) whole query
// documents and a 160-column B tile floor(160/S) whole candidate documents (12 x 16 documents per CTA at S = 10).
// Main loop = gemm.cu's (TMA 128B-swizzled boxes -> 4-stage mbarrier ring -> tcgen05.mma, fp32 accumulator in TMEM),
// with fp32-equivalent bf16x3 operands: the fp32 representations are split once (split_rows_kernel) into bf16 hi / lo
// halves + exact fp32 squared norms, and every K block accumulates hi.hi + hi.lo + lo.hi + lo.lo (the last term is
// coherent -- not noise -- when a query sentence nearly equals a candidate sentence, so it is kept here).
// Epilogue: each thread owns one query-sentence row of the accumulator: d^2 = |q|^2 + |c|^2 - 2 q.c, running
// (min d^2, first j) per candidate document -> shared memory -> min over the query document's rows (first i on ties,
// i.e. the flat index i*S+j of the first maximum, pair_distances.py:176) -> score = -sqrt(max(d^2, 1e-8)).
#include <algorithm>
#include "allpairs_common.cuh"

namespace asp {

using namespace tc;

constexpr int kApStages = 2;  // 94 KB per CTA: TWO CTAs per SM, so one tile's epilogue runs under the other's main loop
constexpr int kApEpi = 128 + kApBlockM * 17 * 8 + kApBlockN * 4 + 64;  // barriers, per-row minima, |c|^2 of the tile, lens
constexpr int kApSmem = kApStages * kApStage + kApEpi + 1024;

struct AllPairsArgs {
    const float* qn;        // [NQ*S] squared norms of the query sentence rows
    const float* cn;        // [NC*S]
    const int32_t* q_lens;  // [NQ]
    const int32_t* c_lens;  // [NC]
    int NQ, NC, S, D;
    int docs_m, docs_n;     // whole documents per tile
    const float* q;         // original fp32 rows (exact re-evaluation of strongly cancelling minima)
    const float* c;
    float* scores;          // [NQ, NC]
    int32_t* flat_idx;      // [NQ, NC] or NULL
};

__global__ void __launch_bounds__(128)
l2max_allpairs_kernel(const __grid_constant__ CUtensorMap tq_hi, const __grid_constant__ CUtensorMap tq_lo,
                      const __grid_constant__ CUtensorMap tc_hi, const __grid_constant__ CUtensorMap tc_lo,
                      const AllPairsArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kApStages * kApStage);
    uint64_t* empty = full + kApStages;
    uint64_t* accum = empty + kApStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
    float* ep_d2 = reinterpret_cast<float*>(smem + kApStages * kApStage + 128);  // [128][17] min d^2 per (row, cand doc)
    int* ep_j = reinterpret_cast<int*>(ep_d2 + kApBlockM * 17);                  // [128][17] its column within the doc
    float* cn_s = reinterpret_cast<float*>(ep_j + kApBlockM * 17);               // [160] |c|^2 of the tile's columns
    int* cl_s = reinterpret_cast<int*>(cn_s + kApBlockN);                        // [16] lengths of the tile's cand docs

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = g.S;
    const int qd0 = blockIdx.x * g.docs_m, cd0 = blockIdx.y * g.docs_n;  // first query / candidate document of the tile
    const int m0 = qd0 * S, n0 = cd0 * S;
    const int rows_m = g.docs_m * S, cols_n = g.docs_n * S;
    const int total = g.D / kApBlockK;  // K blocks; each carries hi.hi, hi.lo, lo.hi, lo.lo

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tq_hi);
        tma_prefetch_desc(&tc_hi);
        for (int s = 0; s < kApStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        for (int it = 0; it < total; ++it) {
            const int s = it % kApStages, ph = (it / kApStages) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], (uint32_t)(rows_m + cols_n) * (kApBlockK * 2) * 2);
            uint8_t* sa = smem + s * kApStage;
            tma_load_2d(sa, &tq_hi, &full[s], it * kApBlockK, m0);
            tma_load_2d(sa + kApABytes, &tq_lo, &full[s], it * kApBlockK, m0);
            tma_load_2d(sa + 2 * kApABytes, &tc_hi, &full[s], it * kApBlockK, n0);
            tma_load_2d(sa + 2 * kApABytes + kApBBytes, &tc_lo, &full[s], it * kApBlockK, n0);
        }
    } else if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(kApBlockM, kApBlockN);
        for (int it = 0; it < total; ++it) {
            const int s = it % kApStages, ph = (it / kApStages) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after_sync();
            const uint32_t sa = smem_u32(smem + s * kApStage);
            const uint64_t a_hi = umma_desc_sw64(sa), a_lo = umma_desc_sw64(sa + kApABytes);
            const uint64_t b_hi = umma_desc_sw64(sa + 2 * kApABytes), b_lo = umma_desc_sw64(sa + 2 * kApABytes + kApBBytes);
#pragma unroll
            for (int k = 0; k < kApBlockK / 16; ++k) {  // 16 bf16 = 32 bytes along K inside the swizzle atom
                umma_bf16(tmem_base, a_lo + 2 * k, b_lo + 2 * k, idesc, (it | k) != 0);  // small terms first
                umma_bf16(tmem_base, a_lo + 2 * k, b_hi + 2 * k, idesc, true);
                umma_bf16(tmem_base, a_hi + 2 * k, b_lo + 2 * k, idesc, true);
                umma_bf16(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, true);
            }
            umma_commit(&empty[s]);
        }
        umma_commit(accum);
    }
    __syncwarp();

    // ---------------- epilogue ----------------
    // (warps 2 and 3, idle during the main loop, stage the tile's |c|^2 and candidate lengths in shared memory)
    if (warp >= 2) {
        for (int t = threadIdx.x - 64; t < kApBlockN; t += 64)
            cn_s[t] = (t < cols_n && n0 + t < g.NC * S) ? __ldg(g.cn + n0 + t) : 0.f;
        for (int t = threadIdx.x - 64; t < 16; t += 64)
            cl_s[t] = (t < g.docs_n && cd0 + t < g.NC) ? min(max(__ldg(g.c_lens + cd0 + t), 0), S) : 0;
    }
    __syncthreads();
    mbar_wait(accum, 0);
    tc_fence_after_sync();
    const int r = warp * 32 + lane;  // accumulator row = query sentence row m0 + r
    const bool row_in = r < rows_m && (m0 + r) < g.NQ * S;
    const float qn = row_in ? __ldg(g.qn + m0 + r) : 0.f;
    float best = INFINITY;
    int best_j = 0, cdoc = 0, j = 0;
#pragma unroll 1
    for (int c0 = 0; c0 < kApBlockN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int col = c0 + e;
            if (col < cols_n) {
                const int cl = cl_s[cdoc & 15];
                const float d2 = qn + cn_s[col] - 2.f * v[e];
                if (j < cl && d2 < best) {
                    best = d2;
                    best_j = j;
                }
                if (++j == S) {  // candidate document finished
                    ep_d2[r * 17 + (cdoc & 15)] = best;
                    ep_j[r * 17 + (cdoc & 15)] = best_j;
                    best = INFINITY;
                    best_j = 0;
                    j = 0;
                    ++cdoc;
                }
            }
        }
        __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 256);
    // min over the rows of each query document (first row wins ties => first flat index i*S+j)
    for (int t = threadIdx.x; t < g.docs_m * g.docs_n; t += blockDim.x) {
        const int qd = t / g.docs_n, cd = t - qd * g.docs_n;
        const int gq = qd0 + qd, gc = cd0 + cd;
        if (gq >= g.NQ || gc >= g.NC) continue;
        const int ql = min(max(__ldg(g.q_lens + gq), 0), S), cl = min(max(__ldg(g.c_lens + gc), 0), S);
        float bd = INFINITY;
        int bi = 0;
        for (int i = 0; i < ql; ++i) {
            const float d = ep_d2[(qd * S + i) * 17 + cd];
            if (d < bd) {
                bd = d;
                bi = i * S + ep_j[(qd * S + i) * 17 + cd];
            }
        }
        const bool any = ql > 0 && cl > 0 && bd < INFINITY;
        if (any) {
            // |q|^2 + |c|^2 - 2 q.c loses its digits when q ~ c (near-duplicate sentences): re-evaluate such a
            // minimum as sum (q-c)^2 in fp32 from the original rows (rare; exact where the ranking is decided)
            // (torch.cdist / scipy cdist, which the reference uses for tsAspire, give exactly 0 for identical rows)
            const size_t qrow = (size_t)gq * S + bi / S, crow = (size_t)gc * S + bi % S;
            bool refined = false;
            if (bd < 0.01f * (__ldg(g.qn + qrow) + __ldg(g.cn + crow))) {
                refined = true;
                const float4* qr = reinterpret_cast<const float4*>(g.q + qrow * g.D);
                const float4* cr = reinterpret_cast<const float4*>(g.c + crow * g.D);
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                for (int k4 = 0; k4 < (g.D >> 2); ++k4) {
                    const float4 x = __ldg(qr + k4), y = __ldg(cr + k4);
                    const float d0 = x.x - y.x, d1 = x.y - y.y, d2_ = x.z - y.z, d3 = x.w - y.w;
                    s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2_, d2_, s2); s3 = fmaf(d3, d3, s3);
                }
                bd = (s0 + s1) + (s2 + s3);
            }
            bd = refined ? bd : fmaxf(bd, 1e-8f);
        }
        g.scores[(size_t)gq * g.NC + gc] = any ? -sqrtf(bd) : kPadNeg;
        if (g.flat_idx) g.flat_idx[(size_t)gq * g.NC + gc] = any ? bi : 0;
    }
}

}  // namespace asp

extern "C" size_t asp_l2max_allpairs_workspace_bytes(int NQ, int NC, int S, int D) {
    const size_t rows = (size_t)(NQ + NC) * S;
    return rows * D * 2 * 2 + rows * sizeof(float) + 1024;
}

extern "C" int asp_l2max_allpairs(const float* q, const int32_t* q_lens, int NQ, const float* c, const int32_t* c_lens,
                                  int NC, int S, int D, float* scores, int32_t* flat_idx, void* workspace,
                                  size_t workspace_bytes, asp_stream_t stream_) {
    using namespace asp;
    ASP_REQUIRE(q && c && q_lens && c_lens && scores, "asp_l2max_allpairs: NULL pointer");
    ASP_REQUIRE(NQ >= 0 && NC >= 0 && S >= 1 && S <= 64, "asp_l2max_allpairs: sentences per document must be in [1, 64] (got %d)", S);
    ASP_REQUIRE(D >= 64 && (D % 64) == 0, "asp_l2max_allpairs: D must be a multiple of 64 (got %d)", D);
    ASP_REQUIRE(aligned16(q) && aligned16(c), "asp_l2max_allpairs: q/c must be 16-byte aligned");
    ASP_REQUIRE(workspace && workspace_bytes >= asp_l2max_allpairs_workspace_bytes(NQ, NC, S, D),
                "asp_l2max_allpairs: workspace too small");
    if (NQ == 0 || NC == 0) return ASP_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t qrows = (size_t)NQ * S, crows = (size_t)NC * S;
    char* w = static_cast<char*>(workspace);
    w += (256 - (reinterpret_cast<uintptr_t>(w) & 255)) & 255;
    __nv_bfloat16* q_hi = reinterpret_cast<__nv_bfloat16*>(w); w += qrows * D * 2;
    __nv_bfloat16* q_lo = reinterpret_cast<__nv_bfloat16*>(w); w += qrows * D * 2;
    __nv_bfloat16* c_hi = reinterpret_cast<__nv_bfloat16*>(w); w += crows * D * 2;
    __nv_bfloat16* c_lo = reinterpret_cast<__nv_bfloat16*>(w); w += crows * D * 2;
    float* qn = reinterpret_cast<float*>(w); w += qrows * 4;
    float* cn = reinterpret_cast<float*>(w);
    split_rows_kernel<<<(unsigned)((qrows + 3) / 4), 128, 0, stream>>>(q, (long long)qrows, D, q_hi, q_lo, qn);
    ASP_LAUNCH_CHECK("split_rows_kernel");
    split_rows_kernel<<<(unsigned)((crows + 3) / 4), 128, 0, stream>>>(c, (long long)crows, D, c_hi, c_lo, cn);
    ASP_LAUNCH_CHECK("split_rows_kernel");
    // whole documents per tile; the epilogue keeps at most 16 candidate documents per tile (short documents, S < 10 --
    // CSFCube abstracts average 7 sentences -- simply leave the tile's last columns unused)
    const int docs_m = kApBlockM / S, docs_n = std::min(kApBlockN / S, 16);
    CUtensorMap tq_hi, tq_lo, tc_hi, tc_lo;
    int rc;
    if ((rc = make_tmap_bf16_k32(&tq_hi, q_hi, qrows, D, docs_m * S))) return rc;
    if ((rc = make_tmap_bf16_k32(&tq_lo, q_lo, qrows, D, docs_m * S))) return rc;
    if ((rc = make_tmap_bf16_k32(&tc_hi, c_hi, crows, D, docs_n * S))) return rc;
    if ((rc = make_tmap_bf16_k32(&tc_lo, c_lo, crows, D, docs_n * S))) return rc;
    AllPairsArgs g{qn, cn, q_lens, c_lens, NQ, NC, S, D, docs_m, docs_n, q, c, scores, flat_idx};
    static thread_local int attr_dev = -1;
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        ASP_CUDA(cudaFuncSetAttribute(l2max_allpairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kApSmem));
        attr_dev = dev;
    }
    dim3 grid((NQ + docs_m - 1) / docs_m, (NC + docs_n - 1) / docs_n);
    l2max_allpairs_kernel<<<grid, 128, kApSmem, stream>>>(tq_hi, tq_lo, tc_hi, tc_lo, g);
    ASP_LAUNCH_CHECK("l2max_allpairs_kernel");
    return ASP_OK;
}
