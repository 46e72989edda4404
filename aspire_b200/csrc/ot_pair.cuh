// Per-THREAD masked Sinkhorn solver for small tiles (Sq <= TQ, Sc <= TC), shared by the stand-alone
// thread-per-pair kernel (ot_sinkhorn.cu) and the fused cost+OT kernel (ot_fused.cu).
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
//
// The column index is processed in PAIRS with the sm_100 packed fp32 instructions (FFMA2 / FADD2 through
// __ffma2_rn / __fadd2_rn): per two entries the step issues 4 packed ALU instructions + 2 MUFU.EX2 instead of
// 8 scalar ALU + 2 MUFU, which matters because the loop is issue-bound next to the MUFU pipe.
#pragma once
#include "common.cuh"

namespace asp {

struct OtOut {
    float *dual, *primal, *f, *g, *alpha, *beta, *neg_cost, *plan, *weighted;
};

template <int TQ, int TC>
struct PairState {
    static_assert(TC % 2 == 0, "columns are processed in pairs");
    float2 C[TQ][TC / 2];
    float la[TQ], f[TQ];
    float2 lb[TC / 2], g[TC / 2];
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }

template <int TQ, int TC, bool FULL = false>
__device__ __forceinline__ void sinkhorn_step(PairState<TQ, TC>& st, int ql_in, int cl_in, float eps, float weight) {
    // weight = 1 (un-averaged) or 0.5 (averaged): new = old - weight * eps ln2 (log2 sum - logw)
    // FULL: the pair is known to have TQ x TC valid sentences, so every length mask folds away at compile time.
    constexpr int TP = TC / 2;
    const int ql = FULL ? TQ : ql_in, cl = FULL ? TC : cl_in;
    const float t = kLog2e / eps;
    const float scale = weight * eps * kLn2;
    const float2 t2 = dup2(t), nt2 = dup2(-t), nscale2 = dup2(-scale);
    float2 v[TP], S[TP];
#pragma unroll
    for (int j = 0; j < TP; ++j) {
        v[j] = __ffma2_rn(st.g[j], t2, st.lb[j]);
        S[j] = f2(0.f, 0.f);
    }
    // `chk` = largest |log-sum| over the valid rows / columns: it leaves the finite range iff a sum over- or
    // underflowed (the sums are of non-negative terms, so +-inf is the only way out; max runs on the ALU pipe, the
    // FMA pipe being the busy one here)
    float chk = 0.f;
    float fnew[TQ];
#pragma unroll
    for (int i = 0; i < TQ; ++i) {
        const float2 uu = dup2(fmaf(st.f[i], t, st.la[i]));
        float2 R = f2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < TP; ++j) {
            const float2 x = __ffma2_rn(st.C[i][j], nt2, __fadd2_rn(uu, v[j]));
            const float2 e = f2(ex2(x.x), ex2(x.y));
            R = __fadd2_rn(R, e);
            S[j] = __fadd2_rn(S[j], e);
        }
        const float l = lg2(R.x + R.y);
        fnew[i] = fmaf(-scale, l - st.la[i], st.f[i]);
        chk = fmaxf(chk, (i < ql) ? fabsf(l) : 0.f);
    }
    float2 gnew[TP];
#pragma unroll
    for (int j = 0; j < TP; ++j) {
        const float2 l = f2(lg2(S[j].x), lg2(S[j].y));
        gnew[j] = __ffma2_rn(nscale2, __fadd2_rn(l, f2(-st.lb[j].x, -st.lb[j].y)), st.g[j]);
        chk = fmaxf(chk, fmaxf((2 * j < cl) ? fabsf(l.x) : 0.f, (2 * j + 1 < cl) ? fabsf(l.y) : 0.f));
    }
    const bool bad = !(chk < 1e30f);
    if (__builtin_expect(bad, 0)) {
        // max-stabilised recomputation of both half-steps from the old potentials (rare).  Padded rows/columns
        // carry log-weight -1e5 and vanish from every sum, exactly as in the reference.
        const float nt = -t;
        float u[TQ];
#pragma unroll
        for (int i = 0; i < TQ; ++i) u[i] = fmaf(st.f[i], t, st.la[i]);
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            float m = -INFINITY, s = 0.f;
#pragma unroll
            for (int j = 0; j < TP; ++j)
                m = fmaxf(m, fmaxf(fmaf(st.C[i][j].x, nt, v[j].x), fmaf(st.C[i][j].y, nt, v[j].y)));
#pragma unroll
            for (int j = 0; j < TP; ++j)
                s += ex2(fmaf(st.C[i][j].x, nt, v[j].x) - m) + ex2(fmaf(st.C[i][j].y, nt, v[j].y) - m);
            const float ft = -eps * kLn2 * (m + lg2(s));
            fnew[i] = st.f[i] + weight * (ft - st.f[i]);
        }
#pragma unroll
        for (int j = 0; j < TP; ++j) {
            float mx = -INFINITY, my = -INFINITY, sx = 0.f, sy = 0.f;
#pragma unroll
            for (int i = 0; i < TQ; ++i) {
                mx = fmaxf(mx, fmaf(st.C[i][j].x, nt, u[i]));
                my = fmaxf(my, fmaf(st.C[i][j].y, nt, u[i]));
            }
#pragma unroll
            for (int i = 0; i < TQ; ++i) {
                sx += ex2(fmaf(st.C[i][j].x, nt, u[i]) - mx);
                sy += ex2(fmaf(st.C[i][j].y, nt, u[i]) - my);
            }
            const float gx = -eps * kLn2 * (mx + lg2(sx)), gy = -eps * kLn2 * (my + lg2(sy));
            gnew[j] = f2(st.g[j].x + weight * (gx - st.g[j].x), st.g[j].y + weight * (gy - st.g[j].y));
        }
    }
#pragma unroll
    for (int i = 0; i < TQ; ++i) st.f[i] = (i < ql) ? fnew[i] : 0.f;
#pragma unroll
    for (int j = 0; j < TP; ++j) st.g[j] = f2((2 * j < cl) ? gnew[j].x : 0.f, (2 * j + 1 < cl) ? gnew[j].y : 0.f);
}

// Solve one pair in the calling thread.  load_cost(i, j) returns C_ij for i < ql, j < cl (never called outside).
// eps_sched[0..n_eps): the epsilon schedule (kernel-parameter / constant-bank or shared memory).
// Writes every requested output of pair b.
template <int TQ, int TC, bool FULL = false, typename LoadCost>
__device__ __forceinline__ void solve_pair_thread(LoadCost load_cost, int ql_in, int cl_in, int b, int Sq_in, int Sc_in,
                                                  const float* eps_sched, int n_eps, float inv_temp, const OtOut& out) {
    constexpr int TP = TC / 2;
    const int ql = FULL ? TQ : ql_in, cl = FULL ? TC : cl_in, Sq = FULL ? TQ : Sq_in, Sc = FULL ? TC : Sc_in;
    PairState<TQ, TC> st;
    const float kBig = 1.0e30f;
#pragma unroll
    for (int i = 0; i < TQ; ++i)
#pragma unroll
        for (int j = 0; j < TP; ++j)
            st.C[i][j] = f2((i < ql && 2 * j < cl) ? load_cost(i, 2 * j) : kBig,
                            (i < ql && 2 * j + 1 < cl) ? load_cost(i, 2 * j + 1) : kBig);

    // marginals (pair_distances.py:57-60): log_softmax over valid sentences of (-min dist)/T, exp, then log again as
    // geomloss does (zero weights -> log-weight -1e5).  Only the log2 weights stay live; alpha/beta are
    // re-materialised as 2^la at the end (keeps ~20 registers out of the hot loop).
    {
        float x[TQ], mx = -INFINITY, s = 0.f;
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            float best = kBig;
#pragma unroll
            for (int j = 0; j < TP; ++j) best = fminf(best, fminf(st.C[i][j].x, st.C[i][j].y));
            x[i] = -best * inv_temp;
            if (i < ql) mx = fmaxf(mx, x[i]);
        }
#pragma unroll
        for (int i = 0; i < TQ; ++i) s += (i < ql) ? expf(x[i] - mx) : 0.f;
        const float lse = mx + logf(s);
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            const float a = (i < ql) ? expf(x[i] - lse) : 0.f;
            st.la[i] = (a > 0.f) ? log2f(a) : kLogZeroWeight * kLog2e;
        }
    }
    {
        float x[TC], mx = -INFINITY, s = 0.f;
#pragma unroll
        for (int j = 0; j < TP; ++j) {
            float bx = kBig, by = kBig;
#pragma unroll
            for (int i = 0; i < TQ; ++i) {
                bx = fminf(bx, st.C[i][j].x);
                by = fminf(by, st.C[i][j].y);
            }
            x[2 * j] = -bx * inv_temp;
            x[2 * j + 1] = -by * inv_temp;
            if (2 * j < cl) mx = fmaxf(mx, x[2 * j]);
            if (2 * j + 1 < cl) mx = fmaxf(mx, x[2 * j + 1]);
        }
#pragma unroll
        for (int j = 0; j < TC; ++j) s += (j < cl) ? expf(x[j] - mx) : 0.f;
        const float lse = mx + logf(s);
#pragma unroll
        for (int j = 0; j < TP; ++j) {
            const float b0 = (2 * j < cl) ? expf(x[2 * j] - lse) : 0.f;
            const float b1 = (2 * j + 1 < cl) ? expf(x[2 * j + 1] - lse) : 0.f;
            st.lb[j] = f2((b0 > 0.f) ? log2f(b0) : kLogZeroWeight * kLog2e, (b1 > 0.f) ? log2f(b1) : kLogZeroWeight * kLog2e);
        }
    }
    // padded entries: any finite cost works (their weight is 2^-144269 = 0); keep them small and finite
#pragma unroll
    for (int i = 0; i < TQ; ++i)
#pragma unroll
        for (int j = 0; j < TP; ++j) {
            if (!(i < ql && 2 * j < cl)) st.C[i][j].x = 0.f;
            if (!(i < ql && 2 * j + 1 < cl)) st.C[i][j].y = 0.f;
        }
#pragma unroll
    for (int i = 0; i < TQ; ++i) st.f[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TP; ++j) st.g[j] = f2(0.f, 0.f);

    if (ql > 0 && cl > 0) {
        // k = -1: initialisation (un-averaged step from f = g = 0 at eps[0]); k = 0..n-1: averaged steps;
        // k = n: final un-averaged extrapolation at eps[n-1].  One loop => one copy of the step in the binary.
#pragma unroll 1
        for (int k = -1; k <= n_eps; ++k) {
            const bool plain = (k < 0) | (k == n_eps);
            sinkhorn_step<TQ, TC, FULL>(st, ql, cl, eps_sched[min(max(k, 0), n_eps - 1)], plain ? 1.0f : 0.5f);
        }
    }

    // weights back from their logs (exact zeros for padded sentences: 2^(-1e5 log2e) underflows to 0)
    float alpha[TQ], beta[TC];
#pragma unroll
    for (int i = 0; i < TQ; ++i) alpha[i] = (i < ql) ? exp2f(st.la[i]) : 0.f;
#pragma unroll
    for (int j = 0; j < TP; ++j) {
        beta[2 * j] = (2 * j < cl) ? exp2f(st.lb[j].x) : 0.f;
        beta[2 * j + 1] = (2 * j + 1 < cl) ? exp2f(st.lb[j].y) : 0.f;
    }
    float dual = 0.f;
#pragma unroll
    for (int i = 0; i < TQ; ++i) dual = fmaf(alpha[i], st.f[i], dual);
#pragma unroll
    for (int j = 0; j < TP; ++j) dual = fmaf(beta[2 * j + 1], st.g[j].y, fmaf(beta[2 * j], st.g[j].x, dual));
    if (out.dual) out.dual[b] = dual;
    if (out.f || out.alpha) {
#pragma unroll
        for (int i = 0; i < TQ; ++i)
            if (i < Sq) {
                if (out.f) out.f[(size_t)b * Sq + i] = st.f[i];
                if (out.alpha) out.alpha[(size_t)b * Sq + i] = alpha[i];
            }
    }
    if (out.g || out.beta) {
#pragma unroll
        for (int j = 0; j < TC; ++j)
            if (j < Sc) {
                if (out.g) out.g[(size_t)b * Sc + j] = (j & 1) ? st.g[j / 2].y : st.g[j / 2].x;
                if (out.beta) out.beta[(size_t)b * Sc + j] = beta[j];
            }
    }
    if (out.primal || out.plan || out.weighted || out.neg_cost) {
        // plan (pair_distances.py:76-85): exp((f_i+g_j-C_ij)/blur) * alpha_i * beta_j, blur = eps_final
        const float tf = kLog2e / eps_sched[n_eps - 1];
        float primal = 0.f;
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
#pragma unroll
            for (int j = 0; j < TC; ++j) {
                if (i < Sq && j < Sc) {
                    const bool valid = (i < ql && j < cl);
                    const float cij = (j & 1) ? st.C[i][j / 2].y : st.C[i][j / 2].x;
                    const float gj = (j & 1) ? st.g[j / 2].y : st.g[j / 2].x;
                    const float p = valid ? ex2((st.f[i] + gj - cij) * tf) * (alpha[i] * beta[j]) : 0.f;
                    const float negc = valid ? -cij : 0.f;
                    const float w = p * negc;
                    primal += w;
                    const size_t o = (size_t)b * Sq * Sc + i * Sc + j;
                    if (out.neg_cost) out.neg_cost[o] = negc;
                    if (out.plan) out.plan[o] = p;
                    if (out.weighted) out.weighted[o] = w;
                }
            }
        }
        if (out.primal) out.primal[b] = primal;
    }
}

}  // namespace asp

// ---------------------------------------------------------------------------------------------------------------------
// Register-light variant of the per-thread solver: the TQ x TC cost tile of the pair stays in SHARED memory
// (Cs[i * TC + j], 16-byte aligned, TC even) and is streamed through registers four entries at a time in every step,
// so the thread keeps only the potentials, the log-weights and the column sums (~100 registers instead of ~200).
// Same arithmetic, same order of operations per entry as sinkhorn_step / solve_pair_thread above.
namespace asp {

// 2^x for a pair of exponents on the FMA / ALU pipes instead of the MUFU pipe: round-to-nearest split x = n + f with the
// 1.5 * 2^23 trick, a degree-5 polynomial for 2^f on [-0.5, 0.5] (max relative error 2.5e-7, the class of ex2.approx),
// and n added into the exponent field.  Exponents are clamped to [-126, 126]: below that the true value is < 2^-126
// anyway; above it the caller must take the stabilised path (it tracks the largest exponent it passed in).
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
    x.x = fminf(fmaxf(x.x, -126.f), 126.f);
    x.y = fminf(fmaxf(x.y, -126.f), 126.f);
    const float2 t = __fadd2_rn(x, dup2(12582912.f));
    const float2 nf = __fadd2_rn(t, dup2(-12582912.f));
    const float2 f = __ffma2_rn(nf, dup2(-1.f), x);
    float2 p = __ffma2_rn(dup2(0.0013400432653725147f), f, dup2(0.009676037356257439f));
    p = __ffma2_rn(p, f, dup2(0.05550327152013779f));
    p = __ffma2_rn(p, f, dup2(0.2402210682630539f));
    p = __ffma2_rn(p, f, dup2(0.6931471824645996f));
    p = __ffma2_rn(p, f, dup2(1.0000001192092896f));
    return f2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
              __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

template <int TQ, int TC>
struct PairStateS {
    float la[TQ], f[TQ];
    float2 lb[TC / 2], g[TC / 2];
};

// Max-stabilised Sinkhorn half-steps on memory operands (the rare path of sinkhorn_step_stream when a plain sum left the
// fp32 range).  buf = la[TQ], f[TQ], fnew[TQ] | lb[TC], g[TC], gnew[TC]; rolled loops, a handful of registers.
static __device__ __noinline__ void stabilised_step_mem(const float* Cs, int TQ, int TC, float* buf, float eps, float weight) {
    const float t = kLog2e / eps, nt = -t;
    const float *la = buf, *f = buf + TQ, *lb = buf + 3 * TQ, *g = buf + 3 * TQ + TC;
    float *fnew = buf + 2 * TQ, *gnew = buf + 3 * TQ + 2 * TC;
#pragma unroll 1
    for (int i = 0; i < TQ; ++i) {
        float m = -INFINITY, s = 0.f;
#pragma unroll 1
        for (int j = 0; j < TC; ++j) m = fmaxf(m, fmaf(Cs[i * TC + j], nt, fmaf(g[j], t, lb[j])));
#pragma unroll 1
        for (int j = 0; j < TC; ++j) s += ex2(fmaf(Cs[i * TC + j], nt, fmaf(g[j], t, lb[j])) - m);
        const float ft = -eps * kLn2 * (m + lg2(s));
        fnew[i] = f[i] + weight * (ft - f[i]);
    }
#pragma unroll 1
    for (int j = 0; j < TC; ++j) {
        float m = -INFINITY, s = 0.f;
#pragma unroll 1
        for (int i = 0; i < TQ; ++i) m = fmaxf(m, fmaf(Cs[i * TC + j], nt, fmaf(f[i], t, la[i])));
#pragma unroll 1
        for (int i = 0; i < TQ; ++i) s += ex2(fmaf(Cs[i * TC + j], nt, fmaf(f[i], t, la[i])) - m);
        const float gt = -eps * kLn2 * (m + lg2(s));
        gnew[j] = g[j] + weight * (gt - g[j]);
    }
}

// POLY: every fourth pair of exponentials is evaluated by ex2_poly2 on the FMA / ALU pipes.  For a kernel whose Sinkhorn
// warps are bound by the MUFU pipe (the Q x C all-pairs kernel: 120 MUFU results per pair and step against ~200 packed
// FMA-pipe instructions) that moves a quarter of the exponentials to pipes with room; the 1 x N kernel, bound by fp32
// issue, keeps POLY = false.
template <int TQ, int TC, bool FULL, bool POLY = false>
__device__ __forceinline__ void sinkhorn_step_stream(PairStateS<TQ, TC>& st, const float* Cs, int ql_in, int cl_in, float eps,
                                                     float t, float weight) {
    // t = log2e / eps: from the caller's per-step table when it has one (saves the division per step), else computed there
    static_assert((TQ * TC) % 4 == 0 && TC % 2 == 0, "tile is read as float4 chunks holding whole column pairs");
    constexpr int TP = TC / 2;
    const int ql = FULL ? TQ : ql_in, cl = FULL ? TC : cl_in;
    const float scale = weight * eps * kLn2;
    const float2 t2 = dup2(t), nt2 = dup2(-t), nscale2 = dup2(-scale);
    float2 v[TP], S[TP];
#pragma unroll
    for (int j = 0; j < TP; ++j) {
        v[j] = __ffma2_rn(st.g[j], t2, st.lb[j]);
        S[j] = f2(0.f, 0.f);
    }
    float chk = 0.f;
    float xpoly = -INFINITY;  // largest exponent handed to ex2_poly2 (which saturates instead of overflowing)
    float fnew[TQ];
    float2 R = f2(0.f, 0.f);
    const float4* C4 = reinterpret_cast<const float4*>(Cs);
    constexpr int kGroup = (TC % 2 == 0 && (2 * TC) % 4 == 0) ? (2 * TC) / 4 : 1;  // float4 chunks per two rows
#pragma unroll
    for (int c = 0; c < TQ * TC / 4; ++c) {
        // keep the compiler from hoisting the whole tile into registers (that is what this variant exists to avoid):
        // loads may not move above the end of the previous two-row group
        if (c % kGroup == 0) asm volatile("" ::: "memory");
        const float4 cv = C4[c];
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
            const int e = 4 * c + 2 * hlf, i = e / TC, j = (e % TC) / 2;  // compile-time after unrolling
            const float2 cc = hlf ? f2(cv.z, cv.w) : f2(cv.x, cv.y);
            const float2 uu = dup2(fmaf(st.f[i], t, st.la[i]));
            const float2 x = __ffma2_rn(cc, nt2, __fadd2_rn(uu, v[j]));
            float2 ex;
            if (POLY && ((2 * c + hlf) & 3) == 3) {  // compile-time after unrolling
                xpoly = fmaxf(xpoly, fmaxf(x.x, x.y));
                ex = ex2_poly2(x);
            } else {
                ex = f2(ex2(x.x), ex2(x.y));
            }
            R = __fadd2_rn(R, ex);
            S[j] = __fadd2_rn(S[j], ex);
            if (j == TP - 1) {  // row i complete
                const float l = lg2(R.x + R.y);
                fnew[i] = fmaf(-scale, l - st.la[i], st.f[i]);
                chk = fmaxf(chk, (i < ql) ? fabsf(l) : 0.f);
                R = f2(0.f, 0.f);
            }
        }
    }
    float2 gnew[TP];
#pragma unroll
    for (int j = 0; j < TP; ++j) {
        const float2 l = f2(lg2(S[j].x), lg2(S[j].y));
        gnew[j] = __ffma2_rn(nscale2, __fadd2_rn(l, f2(-st.lb[j].x, -st.lb[j].y)), st.g[j]);
        chk = fmaxf(chk, fmaxf((2 * j < cl) ? fabsf(l.x) : 0.f, (2 * j + 1 < cl) ? fabsf(l.y) : 0.f));
    }
    const bool bad = !(chk < 1e30f) || (POLY && xpoly > 120.f);
    if (__builtin_expect(bad, 0)) {
        // max-stabilised recomputation of both half-steps from the old potentials (rare): out of line, rolled loops
        float buf[4 * TQ + 4 * TC];  // la, f, fnew | lb, g, gnew  (the helper works on memory)
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            buf[i] = st.la[i];
            buf[TQ + i] = st.f[i];
        }
#pragma unroll
        for (int j = 0; j < TP; ++j) {
            buf[3 * TQ + 2 * j] = st.lb[j].x;
            buf[3 * TQ + 2 * j + 1] = st.lb[j].y;
            buf[3 * TQ + TC + 2 * j] = st.g[j].x;
            buf[3 * TQ + TC + 2 * j + 1] = st.g[j].y;
        }
        stabilised_step_mem(Cs, TQ, TC, buf, eps, weight);
#pragma unroll
        for (int i = 0; i < TQ; ++i) fnew[i] = buf[2 * TQ + i];
#pragma unroll
        for (int j = 0; j < TP; ++j) gnew[j] = f2(buf[3 * TQ + 2 * TC + 2 * j], buf[3 * TQ + 2 * TC + 2 * j + 1]);
    }
#pragma unroll
    for (int i = 0; i < TQ; ++i) st.f[i] = (i < ql) ? fnew[i] : 0.f;
#pragma unroll
    for (int j = 0; j < TP; ++j) st.g[j] = f2((2 * j < cl) ? gnew[j].x : 0.f, (2 * j + 1 < cl) ? gnew[j].y : 0.f);
}

// Solve one pair in the calling thread with the cost tile in shared memory.  On entry Cs[i*TC+j] holds the distance for
// i < ql, j < cl and anything >= 1e30 elsewhere (what phase 1 of the fused kernel writes); padded entries are
// overwritten with 0 (any finite value works: their weight is exactly 0).
template <int TQ, int TC, bool FULL, bool POLY = false>
__device__ __forceinline__ void solve_pair_thread_stream(float* Cs, int ql_in, int cl_in, int b, int Sq_in, int Sc_in,
                                                         const float* eps_sched, int n_eps, float inv_temp,
                                                         const OtOut& out, const float* t_sched = nullptr) {
    // t_sched (optional): log2e / eps_sched[k], precomputed once per CTA with the same IEEE division the step would do
    constexpr int TP = TC / 2;
    const int ql = FULL ? TQ : ql_in, cl = FULL ? TC : cl_in, Sq = FULL ? TQ : Sq_in, Sc = FULL ? TC : Sc_in;
    PairStateS<TQ, TC> st;
    const float kBig = 1.0e30f;
    // marginals (pair_distances.py:57-60): row / column minima of the valid block, two small softmaxes
    {
        float x[TQ], mx = -INFINITY, s = 0.f;
        float cb[TC];
#pragma unroll
        for (int j = 0; j < TC; ++j) cb[j] = kBig;
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            float best = kBig;
#pragma unroll
            for (int j = 0; j < TC; ++j) {
                const float cij = (FULL || (i < ql && j < cl)) ? Cs[i * TC + j] : kBig;
                best = fminf(best, cij);
                cb[j] = fminf(cb[j], cij);
            }
            x[i] = -best * inv_temp;
            if (i < ql) mx = fmaxf(mx, x[i]);
        }
#pragma unroll
        for (int i = 0; i < TQ; ++i) s += (i < ql) ? expf(x[i] - mx) : 0.f;
        const float lse = mx + logf(s);
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            const float a = (i < ql) ? expf(x[i] - lse) : 0.f;
            st.la[i] = (a > 0.f) ? log2f(a) : kLogZeroWeight * kLog2e;
        }
        float y[TC], my = -INFINITY, sy = 0.f;
#pragma unroll
        for (int j = 0; j < TC; ++j) {
            y[j] = -cb[j] * inv_temp;
            if (j < cl) my = fmaxf(my, y[j]);
        }
#pragma unroll
        for (int j = 0; j < TC; ++j) sy += (j < cl) ? expf(y[j] - my) : 0.f;
        const float lsey = my + logf(sy);
#pragma unroll
        for (int j = 0; j < TP; ++j) {
            const float b0 = (2 * j < cl) ? expf(y[2 * j] - lsey) : 0.f;
            const float b1 = (2 * j + 1 < cl) ? expf(y[2 * j + 1] - lsey) : 0.f;
            st.lb[j] = f2((b0 > 0.f) ? log2f(b0) : kLogZeroWeight * kLog2e, (b1 > 0.f) ? log2f(b1) : kLogZeroWeight * kLog2e);
        }
    }
    if (!FULL) {
#pragma unroll
        for (int i = 0; i < TQ; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j)
                if (!(i < ql && j < cl)) Cs[i * TC + j] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < TQ; ++i) st.f[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TP; ++j) st.g[j] = f2(0.f, 0.f);

    if (ql > 0 && cl > 0) {
#pragma unroll 1
        for (int k = -1; k <= n_eps; ++k) {
            const bool plain = (k < 0) | (k == n_eps);
            const int ks = min(max(k, 0), n_eps - 1);
            const float eps = eps_sched[ks];
            sinkhorn_step_stream<TQ, TC, FULL, POLY>(st, Cs, ql, cl, eps, t_sched ? t_sched[ks] : kLog2e / eps, plain ? 1.0f : 0.5f);
        }
    }

    float dual = 0.f;
#pragma unroll
    for (int i = 0; i < TQ; ++i) dual = fmaf((i < ql) ? exp2f(st.la[i]) : 0.f, st.f[i], dual);
#pragma unroll
    for (int j = 0; j < TP; ++j) {
        const float b0 = (2 * j < cl) ? exp2f(st.lb[j].x) : 0.f, b1 = (2 * j + 1 < cl) ? exp2f(st.lb[j].y) : 0.f;
        dual = fmaf(b1, st.g[j].y, fmaf(b0, st.g[j].x, dual));
    }
    if (out.dual) out.dual[b] = dual;
    if (out.f || out.alpha) {
#pragma unroll
        for (int i = 0; i < TQ; ++i)
            if (i < Sq) {
                if (out.f) out.f[(size_t)b * Sq + i] = st.f[i];
                if (out.alpha) out.alpha[(size_t)b * Sq + i] = (i < ql) ? exp2f(st.la[i]) : 0.f;
            }
    }
    if (out.g || out.beta) {
#pragma unroll
        for (int j = 0; j < TC; ++j)
            if (j < Sc) {
                if (out.g) out.g[(size_t)b * Sc + j] = (j & 1) ? st.g[j / 2].y : st.g[j / 2].x;
                if (out.beta) out.beta[(size_t)b * Sc + j] = (j < cl) ? exp2f((j & 1) ? st.lb[j / 2].y : st.lb[j / 2].x) : 0.f;
            }
    }
    if (out.primal || out.plan || out.weighted || out.neg_cost) {
        const float tf = kLog2e / eps_sched[n_eps - 1];
        float primal = 0.f;
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            const float ai = (i < ql) ? exp2f(st.la[i]) : 0.f;
#pragma unroll
            for (int j = 0; j < TC; ++j) {
                if (i < Sq && j < Sc) {
                    const bool valid = (i < ql && j < cl);
                    const float cij = Cs[i * TC + j];
                    const float gj = (j & 1) ? st.g[j / 2].y : st.g[j / 2].x;
                    const float bj = (j < cl) ? exp2f((j & 1) ? st.lb[j / 2].y : st.lb[j / 2].x) : 0.f;
                    const float p = valid ? ex2((st.f[i] + gj - cij) * tf) * (ai * bj) : 0.f;
                    const float negc = valid ? -cij : 0.f;
                    const float w = p * negc;
                    primal += w;
                    const size_t o = (size_t)b * Sq * Sc + i * Sc + j;
                    if (out.neg_cost) out.neg_cost[o] = negc;
                    if (out.plan) out.plan[o] = p;
                    if (out.weighted) out.weighted[o] = w;
                }
            }
        }
        if (out.primal) out.primal[b] = primal;
    }
}

}  // namespace asp

// ---------------------------------------------------------------------------------------------------------------------
// Low-latency variant for passes that carry only a few pairs (a single query against a 1k pool leaves <= 2 pairs per
// Sinkhorn warp): TEN LANES cooperate on one pair instead of one thread solving it alone.  Lane 10p + r of the warp owns
// row r AND column r of pair slot p (p < 3): it keeps its cost row in registers, computes the ten shared exponentials
// of its row per step, leaves them in a small shared scratch tile, and sums its column from there; potentials of the
// other rows/columns travel by shuffle.  Same formulas as sinkhorn_step, ~12 MUFU per lane and step instead of 121, so
// a solve takes ~1/6 of the thread-per-pair time when the warp is otherwise empty.  Returns false (nothing written)
// if a sum left the fp32 range; the caller then falls back to the per-thread solver and its stabilised path.
namespace asp {

template <int T>  // tile is T x T (T = 10), cost stride T
__device__ __forceinline__ bool solve_pairs_rows(const float* Cs_pair, int ql, int cl, int b, int Sq, int Sc, bool active,
                                                 int group_base, int r, float* scratch, const float* eps_sched, int n_eps,
                                                 float inv_temp, const OtOut& out, const float* t_sched = nullptr) {
    // `active`: this lane belongs to a pair slot that holds a pair; group_base = first lane of the slot; r = lane - base.
    // scratch: T*T floats per slot (this slot's block), visible to the slot's ten lanes.
    const unsigned full = 0xffffffffu;
    const float kBig = 1.0e30f;
    const bool row_ok = active && r < ql, col_ok = active && r < cl;
    float C[T];
#pragma unroll
    for (int j = 0; j < T; ++j) C[j] = (row_ok && j < cl) ? Cs_pair[r * T + j] : kBig;
    // marginals: row minimum is local, column minimum is read from the tile
    float rmin = kBig, cmin = kBig;
#pragma unroll
    for (int j = 0; j < T; ++j) rmin = fminf(rmin, C[j]);
    if (col_ok)
#pragma unroll
        for (int i = 0; i < T; ++i)
            if (i < ql) cmin = fminf(cmin, Cs_pair[i * T + r]);
    const float xr = row_ok ? -rmin * inv_temp : -INFINITY, xc = col_ok ? -cmin * inv_temp : -INFINITY;
    float mr = -INFINITY, mc = -INFINITY;
#pragma unroll
    for (int k = 0; k < T; ++k) {
        mr = fmaxf(mr, __shfl_sync(full, xr, group_base + k));
        mc = fmaxf(mc, __shfl_sync(full, xc, group_base + k));
    }
    const float er = row_ok ? expf(xr - mr) : 0.f, ec = col_ok ? expf(xc - mc) : 0.f;
    float sr = 0.f, sc = 0.f;
#pragma unroll
    for (int k = 0; k < T; ++k) {
        sr += __shfl_sync(full, er, group_base + k);
        sc += __shfl_sync(full, ec, group_base + k);
    }
    const float alpha = row_ok ? expf(xr - (mr + logf(sr))) : 0.f, beta = col_ok ? expf(xc - (mc + logf(sc))) : 0.f;
    const float la = (alpha > 0.f) ? log2f(alpha) : kLogZeroWeight * kLog2e;
    const float lb = (beta > 0.f) ? log2f(beta) : kLogZeroWeight * kLog2e;
#pragma unroll
    for (int j = 0; j < T; ++j)
        if (!(row_ok && j < cl)) C[j] = 0.f;  // padded entries: any finite value (their weight is 0)
    float f = 0.f, g = 0.f;
    bool ok = true;
    if (__any_sync(full, active && ql > 0 && cl > 0)) {
#pragma unroll 1
        for (int k = -1; k <= n_eps; ++k) {
            const int ks = min(max(k, 0), n_eps - 1);
            const float eps = eps_sched[ks];
            const float t = t_sched ? t_sched[ks] : kLog2e / eps;
            const bool plain = (k < 0) | (k == n_eps);
            const float scale = (plain ? 1.0f : 0.5f) * eps * kLn2;
            const float u = fmaf(f, t, la), v_own = fmaf(g, t, lb);
            float Rx = 0.f, Ry = 0.f;  // even / odd columns, as the packed accumulator of the per-thread solver sums them
#pragma unroll
            for (int j = 0; j < T; ++j) {
                const float vj = __shfl_sync(full, v_own, group_base + j);
                const float e = ex2(fmaf(C[j], -t, u + vj));
                if (j & 1) Ry += e; else Rx += e;
                if (active) scratch[r * T + j] = e;
            }
            const float R = Rx + Ry;
            __syncwarp();
            float S = 0.f;
            if (active)
#pragma unroll
                for (int i = 0; i < T; ++i) S += scratch[i * T + r];
            __syncwarp();
            const float lr = lg2(R), ls = lg2(S);
            const bool bad = (row_ok && !(fabsf(lr) < 1e30f)) || (col_ok && !(fabsf(ls) < 1e30f));
            if (__any_sync(full, bad && ql > 0 && cl > 0)) {
                ok = false;
                break;
            }
            f = row_ok ? fmaf(-scale, lr - la, f) : 0.f;
            g = col_ok ? fmaf(-scale, ls - lb, g) : 0.f;
        }
    }
    if (!ok) return false;
    if (!(ql > 0 && cl > 0)) f = g = 0.f;
    // outputs
    // (same order of operations as the per-thread solver, so that a pair's values do not depend on the path taken)
    float dual = 0.f;
#pragma unroll
    for (int k = 0; k < T; ++k) dual = fmaf(__shfl_sync(full, alpha, group_base + k), __shfl_sync(full, f, group_base + k), dual);
#pragma unroll
    for (int k = 0; k < T; ++k) dual = fmaf(__shfl_sync(full, beta, group_base + k), __shfl_sync(full, g, group_base + k), dual);
    float primal = 0.f;
    const bool want_mat = out.primal || out.plan || out.weighted || out.neg_cost;
    if (want_mat) {
        const float tf = kLog2e / eps_sched[n_eps - 1];
        float w[T];
#pragma unroll
        for (int j = 0; j < T; ++j) {
            const float gj = __shfl_sync(full, g, group_base + j), bj = __shfl_sync(full, beta, group_base + j);
            const bool valid = row_ok && j < cl;
            const float p = valid ? ex2((f + gj - C[j]) * tf) * (alpha * bj) : 0.f;
            const float negc = valid ? -C[j] : 0.f;
            w[j] = p * negc;
            if (active && r < Sq && j < Sc) {
                const size_t o = (size_t)b * Sq * Sc + r * Sc + j;
                if (out.neg_cost) out.neg_cost[o] = negc;
                if (out.plan) out.plan[o] = p;
                if (out.weighted) out.weighted[o] = w[j];
            }
        }
        // row-major running sum handed from row to row (the per-thread solver adds the 100 products in that order)
        float run = 0.f;
#pragma unroll
        for (int i = 0; i < T; ++i) {
            const float prev = __shfl_sync(full, run, group_base + (i ? i - 1 : 0));
            if (r == i) {
                run = i ? prev : 0.f;
#pragma unroll
                for (int j = 0; j < T; ++j) run += w[j];
            }
        }
        primal = __shfl_sync(full, run, group_base + T - 1);
    }
    if (active) {
        if (r == 0) {
            if (out.dual) out.dual[b] = dual;
            if (out.primal) out.primal[b] = primal;
        }
        if (r < Sq) {
            if (out.f) out.f[(size_t)b * Sq + r] = f;
            if (out.alpha) out.alpha[(size_t)b * Sq + r] = alpha;
        }
        if (r < Sc) {
            if (out.g) out.g[(size_t)b * Sc + r] = g;
            if (out.beta) out.beta[(size_t)b * Sc + r] = beta;
        }
    }
    return true;
}

}  // namespace asp
