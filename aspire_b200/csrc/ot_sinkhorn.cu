// K4: masked Sinkhorn optimal transport (otAspire) on a precomputed cost tile.
//
// Restates, per (query, candidate) pair, what AllPairMaskedWasserstein.compute_distance
// (pair_distances.py:21-92) asks of geomloss 0.2.4 SamplesLoss("sinkhorn", p=1, debias=False):
//   marginals  alpha_i = softmax_i(max_j -C_ij / T), beta_j = softmax_j(max_i -C_ij / T)   (:56-60)
//   init       g = softmin(eps0, C^T, log alpha), f = softmin(eps0, C, log beta)
//   for eps in schedule:  g~ = softmin(eps, C^T, log alpha + f/eps), f~ = softmin(eps, C, log beta + g/eps)
//                         g, f = (g+g~)/2, (f+f~)/2            (both from the OLD f, g)
//   last extrapolation at eps_final, un-averaged
//   dual = <alpha,f>+<beta,g>;  plan P = exp((f+g-C)/eps_final) alpha x beta;  primal = sum P*(-C)
// with softmin(eps, C, h)_i = -eps * logsumexp_j(h_j - C_ij/eps).  Padded sentences carry zero mass and
// are skipped entirely (bit-identical to the reference's -1e5 log-weights).
//
// Two kernels:
//   sinkhorn_warp_kernel   -- one warp per pair, any Sq,Sc <= 128, max-stabilised logsumexp (robust path)
//   sinkhorn_thread_kernel -- one THREAD per pair for small tiles (Sq,Sc <= TS): the whole cost tile and
//                             both potentials live in registers, one shared exponential per (i,j) and
//                             step feeds both half-updates, no shuffles / shared memory in the loop.
#include "ot_pair.cuh"

namespace asp {

// ---------------------------------------------------------------------------------------------------
// Generic warp-per-pair kernel.  Shared memory per warp: C tile [Sq][ldc] + hf[Sq] + hg[Sc] + la[Sq] + lb[Sc].
// Lane l owns rows l, l+32, ... (as i) and columns l, l+32, ... (as j).
// All logsumexp work is done in base 2:  x2 = log2e * x.
// ---------------------------------------------------------------------------------------------------
// FAST (S <= 32 only): a step first tries the shared-exponential form of the thread solver -- lane i computes
// E_ij = 2^(u_i + v_j - C_ij t) once per entry, keeps the row sum, parks the row in shared memory, and lane j adds up
// column j -- half the exponentials and about half the instructions of the two max-stabilised softmins; when a sum
// leaves the fp32 range the step is redone with the stabilised code below, so results stay defined wherever the
// reference's are.
template <int RPL, bool FAST>  // rows (and columns) per lane: S <= 32*RPL
__global__ void __launch_bounds__(128)
sinkhorn_warp_kernel(const float* __restrict__ cost, const int32_t* __restrict__ q_lens, int q_group,
                     const int32_t* __restrict__ c_lens, int B, int Sq, int Sc, const EpsSched sched,
                     float inv_temp, OtOut out) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int ldc = Sc | 1;  // odd leading dimension: conflict-free column walks
    const int per_warp = (FAST ? 2 : 1) * Sq * ldc + 2 * Sq + 2 * Sc;
    float* Cs = smem + (size_t)warp * per_warp;
    float* Es = Cs + (FAST ? Sq * ldc : 0);  // FAST: the step's exponentials E_ij (same layout as Cs)
    float* hf = Es + Sq * ldc;  // log2-domain "h" built from f (indexed by i)
    float* hg = hf + Sq;        // ... from g (indexed by j)
    float* la = hg + Sc;        // log2(alpha_i)
    float* lb = la + Sq;        // log2(beta_j)

    for (int b = blockIdx.x * nwarps + warp; b < B; b += gridDim.x * nwarps) {
        const int ql = min(max(q_lens[b / q_group], 0), Sq), cl = min(max(c_lens[b], 0), Sc);
        const float* Cg = cost + (size_t)b * Sq * Sc;
        __syncwarp();
        for (int e = lane; e < ql * cl; e += 32) {
            const int i = e / cl, j = e - i * cl;
            Cs[i * ldc + j] = Cg[i * Sc + j];
        }
        __syncwarp();

        // ---- marginals (pair_distances.py:57-60): log_softmax of the best match, then exp ----------
        float fi[RPL], gj[RPL], ai[RPL], bj[RPL];  // potentials and weights of the owned rows/cols
        {
            float x[RPL], mx = -INFINITY;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int i = lane + 32 * r;
                float best = INFINITY;
                if (i < ql)
                    for (int j = 0; j < cl; ++j) best = fminf(best, Cs[i * ldc + j]);
                x[r] = (i < ql) ? -best * inv_temp : -INFINITY;
                mx = fmaxf(mx, x[r]);
            }
            mx = warp_max(mx);
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < RPL; ++r) s += (lane + 32 * r < ql) ? expf(x[r] - mx) : 0.f;
            const float lse = mx + logf(warp_sum(s));
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int i = lane + 32 * r;
                ai[r] = (i < ql) ? expf(x[r] - lse) : 0.f;
                if (i < ql) la[i] = (ai[r] > 0.f) ? log2f(ai[r]) : kLogZeroWeight * kLog2e;
            }
        }
        {
            float x[RPL], mx = -INFINITY;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int j = lane + 32 * r;
                float best = INFINITY;
                if (j < cl)
                    for (int i = 0; i < ql; ++i) best = fminf(best, Cs[i * ldc + j]);
                x[r] = (j < cl) ? -best * inv_temp : -INFINITY;
                mx = fmaxf(mx, x[r]);
            }
            mx = warp_max(mx);
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < RPL; ++r) s += (lane + 32 * r < cl) ? expf(x[r] - mx) : 0.f;
            const float lse = mx + logf(warp_sum(s));
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int j = lane + 32 * r;
                bj[r] = (j < cl) ? expf(x[r] - lse) : 0.f;
                if (j < cl) lb[j] = (bj[r] > 0.f) ? log2f(bj[r]) : kLogZeroWeight * kLog2e;
            }
        }
        __syncwarp();

        // softmin over j for the owned rows: -eps*ln2*log2 sum_j 2^(hg[j] - C_ij * t), t = log2e/eps
        auto softmin_rows = [&](float eps, float t, float (&res)[RPL]) {
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int i = lane + 32 * r;
                float m = -INFINITY, s = 0.f;
                if (i < ql) {
                    const float* row = Cs + i * ldc;
                    for (int j = 0; j < cl; ++j) m = fmaxf(m, fmaf(-row[j], t, hg[j]));
                    for (int j = 0; j < cl; ++j) s += ex2(fmaf(-row[j], t, hg[j]) - m);
                }
                res[r] = (i < ql) ? -eps * kLn2 * (m + lg2(s)) : 0.f;
            }
        };
        auto softmin_cols = [&](float eps, float t, float (&res)[RPL]) {
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int j = lane + 32 * r;
                float m = -INFINITY, s = 0.f;
                if (j < cl) {
                    const float* col = Cs + j;
                    for (int i = 0; i < ql; ++i) m = fmaxf(m, fmaf(-col[i * ldc], t, hf[i]));
                    for (int i = 0; i < ql; ++i) s += ex2(fmaf(-col[i * ldc], t, hf[i]) - m);
                }
                res[r] = (j < cl) ? -eps * kLn2 * (m + lg2(s)) : 0.f;
            }
        };
        // publish h vectors: hf[i] = la[i] + f_i * t ; hg[j] = lb[j] + g_j * t
        auto publish = [&](float t, bool with_potentials) {
            __syncwarp();
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int i = lane + 32 * r;
                if (i < ql) hf[i] = with_potentials ? fmaf(fi[r], t, la[i]) : la[i];
                if (i < cl) hg[i] = with_potentials ? fmaf(gj[r], t, lb[i]) : lb[i];
            }
            __syncwarp();
        };

        // one step in the shared-exponential form (RPL == 1): new un-averaged potentials in ft/gt; false = out of range
        auto fast_step = [&](float eps, float t, bool with_potentials, float (&ft)[RPL], float (&gt)[RPL]) -> bool {
            publish(t, with_potentials);
            float R = 0.f, S = 0.f;
            if (lane < ql) {
                const float u = hf[lane];
                const float* row = Cs + lane * ldc;
                float* erow = Es + lane * ldc;
                for (int j = 0; j < cl; ++j) {
                    const float e = ex2(fmaf(-row[j], t, u + hg[j]));
                    R += e;
                    erow[j] = e;
                }
            }
            __syncwarp();
            if (lane < cl)
                for (int i = 0; i < ql; ++i) S += Es[i * ldc + lane];
            const float lr = lg2(R), ls = lg2(S);
            const bool ok = (lane >= ql || fabsf(lr) < 1e30f) && (lane >= cl || fabsf(ls) < 1e30f);
            if (!__all_sync(0xffffffffu, ok)) return false;
            const float scale = eps * kLn2;
            ft[0] = (lane < ql) ? (with_potentials ? fi[0] : 0.f) - scale * (lr - la[lane]) : 0.f;
            gt[0] = (lane < cl) ? (with_potentials ? gj[0] : 0.f) - scale * (ls - lb[lane]) : 0.f;
            return true;
        };

        if (ql > 0 && cl > 0) {
            // k = -1: initialisation from f = g = 0 at eps[0]; k = 0..n-1: averaged steps; k = n: final un-averaged step
            for (int k = -1; k <= sched.n; ++k) {
                const float eps = sched.eps[min(max(k, 0), sched.n - 1)], t = kLog2e / eps;
                float gt[RPL], ft[RPL];
                bool done = false;
                if (FAST && RPL == 1) done = fast_step(eps, t, k >= 0, ft, gt);
                if (!done) {
                    publish(t, k >= 0);
                    softmin_cols(eps, t, gt);
                    softmin_rows(eps, t, ft);
                }
                const bool plain = (k < 0) || (k == sched.n);
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    gj[r] = plain ? gt[r] : 0.5f * (gj[r] + gt[r]);
                    fi[r] = plain ? ft[r] : 0.5f * (fi[r] + ft[r]);
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < RPL; ++r) fi[r] = gj[r] = 0.f;
        }

        // ---- outputs ---------------------------------------------------------------------------
        float dual = 0.f;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int i = lane + 32 * r;
            dual += ai[r] * fi[r] + bj[r] * gj[r];
            if (i < Sq) {
                if (out.f) out.f[(size_t)b * Sq + i] = (i < ql) ? fi[r] : 0.f;
                if (out.alpha) out.alpha[(size_t)b * Sq + i] = ai[r];
            }
            if (i < Sc) {
                if (out.g) out.g[(size_t)b * Sc + i] = (i < cl) ? gj[r] : 0.f;
                if (out.beta) out.beta[(size_t)b * Sc + i] = bj[r];
            }
        }
        dual = warp_sum(dual);
        if (out.dual && lane == 0) out.dual[b] = dual;

        if (out.primal || out.plan || out.weighted || out.neg_cost) {
            // plan (pair_distances.py:76-85): exp((f_i+g_j-C_ij)/blur) * alpha_i * beta_j, blur = eps_final
            const float tf = kLog2e / sched.eps[sched.n - 1];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < RPL; ++r) {  // reuse hf/hg as plain f, g ; la/lb as alpha, beta
                const int i = lane + 32 * r;
                if (i < ql) { hf[i] = fi[r]; la[i] = ai[r]; }
                if (i < cl) { hg[i] = gj[r]; lb[i] = bj[r]; }
            }
            __syncwarp();
            float primal = 0.f;
            for (int e = lane; e < Sq * Sc; e += 32) {
                const int i = e / Sc, j = e - i * Sc;
                float negc = 0.f, p = 0.f;
                if (i < ql && j < cl) {
                    const float cij = Cs[i * ldc + j];
                    negc = -cij;
                    p = ex2((hf[i] + hg[j] - cij) * tf) * (la[i] * lb[j]);
                }
                const float w = p * negc;
                primal += w;
                const size_t o = (size_t)b * Sq * Sc + e;
                if (out.neg_cost) out.neg_cost[o] = negc;
                if (out.plan) out.plan[o] = p;
                if (out.weighted) out.weighted[o] = w;
            }
            primal = warp_sum(primal);
            if (out.primal && lane == 0) out.primal[b] = primal;
        }
    }
}


// Thread-per-pair kernel for small tiles: the solver itself lives in ot_pair.cuh (shared with ot_fused.cu).
template <int TQ, int TC>
__global__ void __launch_bounds__(128)
sinkhorn_thread_kernel(const float* __restrict__ cost, const int32_t* __restrict__ q_lens, int q_group,
                       const int32_t* __restrict__ c_lens, int B, int Sq, int Sc, const EpsSched sched,
                       float inv_temp, OtOut out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int ql = min(max(q_lens[b / q_group], 0), Sq), cl = min(max(c_lens[b], 0), Sc);
    const float* Cg = cost + (size_t)b * Sq * Sc;
    solve_pair_thread<TQ, TC>([&](int i, int j) { return Cg[i * Sc + j]; }, ql, cl, b, Sq, Sc, sched.eps, sched.n, inv_temp,
                              out);
}

int g_ot_kernel = 0;  // 0 auto, 1 force warp-per-pair, 2 force thread-per-pair (asp_set_option "ot_kernel")

int launch_sinkhorn(const float* cost, const int32_t* q_lens, int q_group, const int32_t* c_lens, int B,
                           int Sq, int Sc, const EpsSched& sched, float temp, const OtOut& out, cudaStream_t stream) {
    const int smax = Sq > Sc ? Sq : Sc;
    const bool thread_ok = (Sq <= 10 && Sc <= 10);
    if (g_ot_kernel == 2 && !thread_ok) {
        set_error("ot: thread-per-pair kernel supports at most 10x10 sentences (got %dx%d)", Sq, Sc);
        return ASP_ERR_UNSUPPORTED;
    }
    // one thread per pair only pays once there are enough pairs to fill the machine
    if (thread_ok && (g_ot_kernel == 2 || (g_ot_kernel == 0 && B >= 64 * sm_count()))) {
        const int threads = 128, blocks = (B + threads - 1) / threads;
        sinkhorn_thread_kernel<10, 10><<<blocks, threads, 0, stream>>>(cost, q_lens, q_group, c_lens, B, Sq, Sc,
                                                                     sched, 1.0f / temp, out);
        ASP_LAUNCH_CHECK("sinkhorn_thread_kernel");
        return ASP_OK;
    }
    // the forced warp kernel (ot_kernel = 1) stays purely max-stabilised: it is the tests' independent cross-check
    const bool fast = (g_ot_kernel != 1) && smax <= 32;
    const int per_warp = ((fast ? 2 : 1) * Sq * (Sc | 1) + 2 * Sq + 2 * Sc) * (int)sizeof(float);
    int warps = 4;
    while (warps > 1 && warps * per_warp > 96 * 1024) warps >>= 1;
    const int smem = warps * per_warp;
    const int blocks = (B + warps - 1) / warps;
    const float inv_temp = 1.0f / temp;
#define ASP_LAUNCH_WARP(RPL, FAST)                                                                                      \
    do {                                                                                                                \
        if (smem > 48 * 1024)                                                                                           \
            ASP_CUDA(cudaFuncSetAttribute(sinkhorn_warp_kernel<RPL, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          smem));                                                                       \
        sinkhorn_warp_kernel<RPL, FAST><<<blocks, warps * 32, smem, stream>>>(cost, q_lens, q_group, c_lens, B, Sq,    \
                                                                              Sc, sched, inv_temp, out);                \
    } while (0)
    if (fast) ASP_LAUNCH_WARP(1, true);
    else if (smax <= 32) ASP_LAUNCH_WARP(1, false);
    else if (smax <= 64) ASP_LAUNCH_WARP(2, false);
    else ASP_LAUNCH_WARP(4, false);
#undef ASP_LAUNCH_WARP
    ASP_LAUNCH_CHECK("sinkhorn_warp_kernel");
    return ASP_OK;
}

int make_sched(const float* eps_host, int n_eps, EpsSched* s) {
    ASP_REQUIRE(eps_host && n_eps >= 1, "ot: eps schedule missing");
    if (n_eps > ASP_MAX_EPS) {
        set_error("ot: schedule of %d entries exceeds ASP_MAX_EPS=%d", n_eps, ASP_MAX_EPS);
        return ASP_ERR_UNSUPPORTED;
    }
    s->n = n_eps;
    for (int i = 0; i < n_eps; ++i) {
        ASP_REQUIRE(eps_host[i] > 0.f, "ot: eps[%d]=%g must be positive", i, (double)eps_host[i]);
        s->eps[i] = eps_host[i];
    }
    return ASP_OK;
}

OtOut to_out(const asp_ot_outputs* o) {
    OtOut r{o->dual, o->primal, o->f, o->g, o->alpha, o->beta, o->neg_cost, o->plan, o->weighted};
    return r;
}

}  // namespace asp

extern "C" int asp_ot_sinkhorn_from_cost(const float* cost, const int32_t* q_lens, int q_broadcast,
                                         const int32_t* c_lens, int B, int Sq, int Sc, const float* eps_host,
                                         int n_eps, float temp, const asp_ot_outputs* out, asp_stream_t stream) {
    ASP_REQUIRE(cost && q_lens && c_lens && out, "asp_ot_sinkhorn_from_cost: NULL pointer");
    ASP_REQUIRE(B >= 0 && Sq >= 1 && Sc >= 1, "asp_ot_sinkhorn_from_cost: bad shape");
    ASP_REQUIRE(temp > 0.f, "asp_ot_sinkhorn_from_cost: temp must be > 0");
    if (Sq > ASP_MAX_SENTS || Sc > ASP_MAX_SENTS) {
        asp::set_error("ot: %dx%d sentences exceeds ASP_MAX_SENTS=%d", Sq, Sc, ASP_MAX_SENTS);
        return ASP_ERR_UNSUPPORTED;
    }
    asp::EpsSched sched;
    int rc = asp::make_sched(eps_host, n_eps, &sched);
    if (rc) return rc;
    if (B == 0) return ASP_OK;
    return asp::launch_sinkhorn(cost, q_lens, q_broadcast ? B : 1, c_lens, B, Sq, Sc, sched, temp, asp::to_out(out),
                                (cudaStream_t)stream);
}
