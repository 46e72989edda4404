// Per-THREAD masked Sinkhorn solver for small tiles (Sq <= TQ, Sc <= TC), shared by the stand-alone
// thread-per-pair kernel (ot_sinkhorn.cu) and the fused cost+OT kernel (ot_fused.cu).
#pragma once
#include "common.cuh"

namespace asp {

struct OtOut {
    float *dual, *primal, *f, *g, *alpha, *beta, *neg_cost, *plan, *weighted;
};

// ---------------------------------------------------------------------------------------------------
// Thread-per-pair kernel for small tiles (Sq <= TQ, Sc <= TC).
//
// State per thread: C[TQ][TC], log2 weights, f, g -- all in registers.  One step at epsilon (t = log2e/eps):
//     u_i = la_i + f_i t,  v_j = lb_j + g_j t,  E_ij = 2^(u_i + v_j - C_ij t)      (ONE exponential per entry)
//     R_i = sum_j E_ij,    S_j = sum_i E_ij
//     f~_i = f_i - eps ln2 (log2 R_i - la_i),   g~_j = g_j - eps ln2 (log2 S_j - lb_j)
// which is algebraically the two geomloss softmins taken from the OLD (f, g); E is the current plan estimate
// (entries <= ~1 near feasibility), so no max-subtraction is needed.  A row/column whose sum leaves the
// fp32 range (0, inf, nan) is recomputed with the max-stabilised form, so the result stays defined wherever
// the reference's is.  init = un-averaged step from f=g=0 at eps[0]; loop = averaged steps; final =
// un-averaged step at eps[n-1].
// ---------------------------------------------------------------------------------------------------
template <int TQ, int TC>
struct PairState {
    float C[TQ][TC];
    float la[TQ], lb[TC], f[TQ], g[TC];
};

template <int TQ, int TC>
__device__ __forceinline__ void sinkhorn_step(PairState<TQ, TC>& st, int ql, int cl, float eps, float weight) {
    // weight = 1 (un-averaged) or 0.5 (averaged): new = old - weight * eps ln2 (log2 sum - logw)
    const float t = kLog2e / eps;
    const float scale = weight * eps * kLn2;
    float u[TQ], v[TC], S[TC];
#pragma unroll
    for (int i = 0; i < TQ; ++i) u[i] = fmaf(st.f[i], t, st.la[i]);
#pragma unroll
    for (int j = 0; j < TC; ++j) {
        v[j] = fmaf(st.g[j], t, st.lb[j]);
        S[j] = 0.f;
    }
    const float nt = -t;
    bool bad = false;
    float fnew[TQ];
#pragma unroll
    for (int i = 0; i < TQ; ++i) {
        float R = 0.f;
#pragma unroll
        for (int j = 0; j < TC; ++j) {
            const float e = ex2(fmaf(st.C[i][j], nt, u[i] + v[j]));
            R += e;
            S[j] += e;
        }
        const float l = lg2(R);
        fnew[i] = st.f[i] - scale * (l - st.la[i]);
        bad |= (i < ql) && !(fabsf(l) < 1e30f);
    }
    float gnew[TC];
#pragma unroll
    for (int j = 0; j < TC; ++j) {
        const float l = lg2(S[j]);
        gnew[j] = st.g[j] - scale * (l - st.lb[j]);
        bad |= (j < cl) && !(fabsf(l) < 1e30f);
    }
    if (__builtin_expect(bad, 0)) {
        // max-stabilised recomputation of both half-steps from the old potentials (rare).  Padded rows/columns
        // carry log-weight -1e5 and vanish from every sum, exactly as in the reference.
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            float m = -INFINITY, s = 0.f;
#pragma unroll
            for (int j = 0; j < TC; ++j) m = fmaxf(m, fmaf(st.C[i][j], nt, v[j]));
#pragma unroll
            for (int j = 0; j < TC; ++j) s += ex2(fmaf(st.C[i][j], nt, v[j]) - m);
            const float ft = -eps * kLn2 * (m + lg2(s));
            fnew[i] = st.f[i] + weight * (ft - st.f[i]);
        }
#pragma unroll
        for (int j = 0; j < TC; ++j) {
            float m = -INFINITY, s = 0.f;
#pragma unroll
            for (int i = 0; i < TQ; ++i) m = fmaxf(m, fmaf(st.C[i][j], nt, u[i]));
#pragma unroll
            for (int i = 0; i < TQ; ++i) s += ex2(fmaf(st.C[i][j], nt, u[i]) - m);
            const float gt = -eps * kLn2 * (m + lg2(s));
            gnew[j] = st.g[j] + weight * (gt - st.g[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < TQ; ++i) st.f[i] = (i < ql) ? fnew[i] : 0.f;
#pragma unroll
    for (int j = 0; j < TC; ++j) st.g[j] = (j < cl) ? gnew[j] : 0.f;
}


// Solve one pair in the calling thread.  load_cost(i, j) returns C_ij for i < ql, j < cl (never called outside).
// Writes every requested output of pair b.
template <int TQ, int TC, typename LoadCost>
__device__ __forceinline__ void solve_pair_thread(LoadCost load_cost, int ql, int cl, int b, int Sq, int Sc,
                                                  const EpsSched& sched, float inv_temp, const OtOut& out) {
    PairState<TQ, TC> st;
    const float kBig = 1.0e30f;
#pragma unroll
    for (int i = 0; i < TQ; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) st.C[i][j] = (i < ql && j < cl) ? load_cost(i, j) : kBig;

    // marginals: log_softmax over valid sentences of (-min dist)/T, exp, then log again as geomloss does
    float alpha[TQ], beta[TC];
    {
        float x[TQ], mx = -INFINITY, s = 0.f;
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            float best = kBig;
#pragma unroll
            for (int j = 0; j < TC; ++j) best = fminf(best, st.C[i][j]);
            x[i] = -best * inv_temp;
            if (i < ql) mx = fmaxf(mx, x[i]);
        }
#pragma unroll
        for (int i = 0; i < TQ; ++i) s += (i < ql) ? expf(x[i] - mx) : 0.f;
        const float lse = mx + logf(s);
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            alpha[i] = (i < ql) ? expf(x[i] - lse) : 0.f;
            st.la[i] = (alpha[i] > 0.f) ? log2f(alpha[i]) : kLogZeroWeight * kLog2e;
        }
    }
    {
        float x[TC], mx = -INFINITY, s = 0.f;
#pragma unroll
        for (int j = 0; j < TC; ++j) {
            float best = kBig;
#pragma unroll
            for (int i = 0; i < TQ; ++i) best = fminf(best, st.C[i][j]);
            x[j] = -best * inv_temp;
            if (j < cl) mx = fmaxf(mx, x[j]);
        }
#pragma unroll
        for (int j = 0; j < TC; ++j) s += (j < cl) ? expf(x[j] - mx) : 0.f;
        const float lse = mx + logf(s);
#pragma unroll
        for (int j = 0; j < TC; ++j) {
            beta[j] = (j < cl) ? expf(x[j] - lse) : 0.f;
            st.lb[j] = (beta[j] > 0.f) ? log2f(beta[j]) : kLogZeroWeight * kLog2e;
        }
    }
    // padded entries: any finite cost works (their weight is 2^-144269 = 0); keep them small and finite
#pragma unroll
    for (int i = 0; i < TQ; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j)
            if (!(i < ql && j < cl)) st.C[i][j] = 0.f;
#pragma unroll
    for (int i = 0; i < TQ; ++i) st.f[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TC; ++j) st.g[j] = 0.f;

    if (ql > 0 && cl > 0) {
        // k = -1: initialisation (un-averaged step from f = g = 0 at eps[0]); k = 0..n-1: averaged steps;
        // k = n: final un-averaged extrapolation at eps[n-1].  One loop => one copy of the step in the binary.
#pragma unroll 1
        for (int k = -1; k <= sched.n; ++k) {
            const bool plain = (k < 0) | (k == sched.n);
            sinkhorn_step<TQ, TC>(st, ql, cl, sched.eps[min(max(k, 0), sched.n - 1)], plain ? 1.0f : 0.5f);
        }
    }

    float dual = 0.f;
#pragma unroll
    for (int i = 0; i < TQ; ++i) dual = fmaf(alpha[i], st.f[i], dual);
#pragma unroll
    for (int j = 0; j < TC; ++j) dual = fmaf(beta[j], st.g[j], dual);
    if (out.dual) out.dual[b] = dual;
#pragma unroll
    for (int i = 0; i < TQ; ++i)
        if (i < Sq) {
            if (out.f) out.f[(size_t)b * Sq + i] = st.f[i];
            if (out.alpha) out.alpha[(size_t)b * Sq + i] = alpha[i];
        }
#pragma unroll
    for (int j = 0; j < TC; ++j)
        if (j < Sc) {
            if (out.g) out.g[(size_t)b * Sc + j] = st.g[j];
            if (out.beta) out.beta[(size_t)b * Sc + j] = beta[j];
        }
    if (out.primal || out.plan || out.weighted || out.neg_cost) {
        const float tf = kLog2e / sched.eps[sched.n - 1];
        float primal = 0.f;
#pragma unroll
        for (int i = 0; i < TQ; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) {
                if (i < Sq && j < Sc) {
                    const bool valid = (i < ql && j < cl);
                    const float cij = st.C[i][j];
                    const float p = valid ? ex2((st.f[i] + st.g[j] - cij) * tf) * (alpha[i] * beta[j]) : 0.f;
                    const float negc = valid ? -cij : 0.f;
                    const float w = p * negc;
                    primal += w;
                    const size_t o = (size_t)b * Sq * Sc + i * Sc + j;
                    if (out.neg_cost) out.neg_cost[o] = negc;
                    if (out.plan) out.plan[o] = p;
                    if (out.weighted) out.weighted[o] = w;
                }
            }
        if (out.primal) out.primal[b] = primal;
    }
}


}  // namespace asp
