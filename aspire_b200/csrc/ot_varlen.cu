// K2+K4 fused for LONG / RAGGED documents (up to 32 x 32 sentences): otAspire score -- and the tsAspire masked max --
// straight from the sentence representations, one kernel, the cost tile never leaves the SM.
//
// Replaces, for documents the 10-sentence kernel (ot_fused.cu) does not cover,
//   pad mask + -cdist            src/learning/facetid_models/pair_distances.py:39-50
//   softmax marginals            pair_distances.py:56-60
//   geomloss SamplesLoss(...)    pair_distances.py:68-72 / 88-91   (eps-scaling Sinkhorn, as restated in ot_sinkhorn.cu)
//   plan + primal value          pair_distances.py:76-85
//   masked max / flat argmax     pair_distances.py:157-176         (tsAspire, mode 1)
// which used to be two kernels with the [B,Sq,Sc] cost tensor written to and re-read from HBM.
//
// One persistent CTA of 8 warps per SM; a WARP owns a pair from the first byte to the score:
//   * stream.  Only the VALID rows of a pair are read, each exactly once: (ql + cl) * D * 4 bytes.  A row arrives in
//     128-byte pieces (32 floats) through a per-warp shared-memory ring filled with cp.async (16 B per lane, L1
//     bypassed).  One "slice" = the same 32 floats of all ql + cl rows; slices are variable-sized (8-row units) and the
//     ring is a byte ring, so short documents keep as many bytes in flight as long ones, and the stream runs straight
//     across pair boundaries (pairs are claimed one ahead from a global counter).
//   * Gram.  The 32 lanes form an 8 x 4 grid over the (query row, candidate row) plane: lane (li, lj) owns rows
//     li + 8a and columns lj + 4b and keeps their dot products in packed fp32 accumulators (FFMA2).  All operand loads
//     are 128-bit shared-memory BROADCASTS (8 distinct rows for the query side, 4 for the candidate side; the 36-float
//     row pitch puts them in distinct banks), so one LDS.128 feeds 8-16 FFMA2 and no cross-lane reduction is needed.
//     Pairs are bucketed by (ceil(ql/8), ceil(cl/8)): a 9 x 12 pair pays 16 x 16, not 32 x 32.  Squared row norms are
//     taken from the pieces each lane copied itself (one extra LDS.128 + 2 FFMA2 per piece).
//   * cost tile.  sqrt(max(|q|^2 + |c|^2 - 2 q.c, 1e-8)) -> a 32 x 36 tile in shared memory (geomloss' formula).
//   * Sinkhorn.  Lane i owns row i AND column i.  Per step ONE exponential per entry (E_ij = 2^(u_i + v_j - C_ij t),
//     the current plan estimate): the row owner computes its row from registers, keeps the row sum and parks the row in
//     shared memory (the cost tile's buffer -- the costs live in registers by then); the column owner adds up its
//     column.  A sum that leaves the fp32 range redoes the step in the max-stabilised form.  Same schedule, averaging
//     and final extrapolation as the other solvers (ot_pair.cuh).
#include <algorithm>
#include <type_traits>
#include "ot_pair.cuh"

#ifndef ASP_VL_U_SMALL
#define ASP_VL_U_SMALL 8
#define ASP_VL_U_BIG 4
#endif
// Idle waits (a Sinkhorn warp waiting for its next cost tile, a producer waiting for ring space): nanoseconds of plain
// sleep between non-blocking barrier tests; 0 = mbarrier.try_wait with a suspend hint.  The hinted form wakes on every
// barrier event of the CTA (measured: one retry per ~17 ns per waiting warp, 45 % of the kernel's issued instructions),
// which takes issue slots from the Gram warp of the same scheduler; a plain sleep does not.
#ifndef ASP_VL_QUAD_MAP
#define ASP_VL_QUAD_MAP 1
#endif
#ifndef ASP_VL_IDLE_SINK
#define ASP_VL_IDLE_SINK 0
#endif
#ifndef ASP_VL_IDLE_PROD
#define ASP_VL_IDLE_PROD 0
#endif

namespace asp {

constexpr int kVlPitch = 32;                       // floats per staged row piece: 128 bytes, 128-byte aligned; its eight 16-byte
                                                   // chunks are stored XOR-swizzled by (row & 7) -- what TMA's 128B swizzle does
constexpr int kVlUnitFloats = 8 * kVlPitch;        // ring allocation unit: 8 row pieces = 1 KB
constexpr int kVlTileLd = 36;
constexpr int kVlTileFloats = 32 * kVlTileLd;      // cost tile / exponential scratch
constexpr int kVlMaxS = 32;
constexpr int kVlCounterSlots = 1024;

__device__ unsigned int g_vl_counter[kVlCounterSlots];

struct VlArgs {
    const float* q;
    const int32_t* q_lens;
    const float* c;
    const int32_t* c_lens;
    const int32_t* c_index;
    int q_group, B, Sq, Sc, D;
    float inv_temp;
    unsigned int* counter;
    const int32_t* order;  // optional: pairs in the order they are claimed (sorted by shape, largest first); NULL = 0..B-1
    int mode;  // 0: otAspire (outputs in OtOut), 1: tsAspire masked max
    int dev_flags;  // developer switches (asp_set_option "vl_flags"): 1 = every pair runs the 32 x 32 Gram variant,
                    // 2 = pairs are claimed in input order even when a workspace for the shape sort is given
    float* best;
    int32_t* flat_idx;
    float* pair_sims;
};

__device__ __forceinline__ void vl_cp_async16(uint32_t smem_dst, const float* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void vl_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most n groups are pending (n is warp-uniform; waiting for more than asked is always safe)
__device__ __forceinline__ void vl_wait_pending(int n) {
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    }
}
__device__ __forceinline__ float vl_sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t vl_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One slice (32 floats of every row of the pair) of the Gram tile.  Slot rows [0, ql) = query rows, [ql, ql+cl) =
// candidate rows.  Rows a lane touches beyond the valid ones hold stale bytes: they only ever reach accumulators of
// entries that the epilogue masks.
__device__ __forceinline__ float4 vl_lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
template <int OFF>
__device__ __forceinline__ float4 vl_lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
    return v;
}

// Slot layout: row r of the slice (query rows first, then candidate rows) is the 128 bytes at slot + 128 r; its 16-byte
// chunk k sits at chunk position k ^ (r & 7).  A piece therefore lands in ONE 128-byte-aligned line of shared memory (the
// four sectors of a cp.async'ed global line merge into one wavefront; with padded 144-byte rows every LDGSTS took 11
// wavefronts instead of 4), while the Gram loads -- 8 rows r = li + 8a at the same k, or 4 rows ql + lj + 4b -- still hit
// distinct bank groups.  Since r & 7 does not depend on a, and flips only bit 2 with the parity of b, a lane needs three
// swizzled base addresses per slice and one XOR with the compile-time (k << 4) per step.
template <int NA, int NB>
__device__ __forceinline__ void vl_slice(uint32_t slot, int ql, int li, int lj, int r0, int chunk, float2 (&acc)[4][8],
                                         float2 (&nq)[8], float2 (&nc)[8]) {
    {
        // squared norms from the pieces the producer lane of the same index copied: rows r0 + 4m (and ql + r0 + 4m)
        const uint32_t qe = slot + r0 * 128 + ((chunk ^ r0) << 4), qo = qe ^ 64;
#pragma unroll
        for (int m = 0; m < 2 * NA; ++m) {
            const float4 v = vl_lds128((m & 1 ? qo : qe) + m * 512);
            nq[m] = __ffma2_rn(make_float2(v.x, v.y), make_float2(v.x, v.y), nq[m]);
            nq[m] = __ffma2_rn(make_float2(v.z, v.w), make_float2(v.z, v.w), nq[m]);
        }
        const int rc = ql + r0;
        const uint32_t ce = slot + rc * 128 + ((chunk ^ (rc & 7)) << 4), co = ce ^ 64;
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            const float4 v = vl_lds128((m & 1 ? co : ce) + m * 512);
            nc[m] = __ffma2_rn(make_float2(v.x, v.y), make_float2(v.x, v.y), nc[m]);
            nc[m] = __ffma2_rn(make_float2(v.z, v.w), make_float2(v.z, v.w), nc[m]);
        }
    }
    const uint32_t pq = (slot + li * 128) ^ (li << 4);
    const int rcl = ql + lj;
    const uint32_t pc0 = (slot + rcl * 128) ^ ((rcl & 7) << 4), pc1 = pc0 ^ 64;
    // Unroll policy: a taken branch costs this warp (the only Gram warp of its scheduler) an instruction-fetch bubble
    // that nothing hides, so small tiles run the eight 4-float steps of a slice as straight-line code and large tiles in
    // two trips (with one or two steps per trip, 60 % of the Gram warps' samples were "no instruction").
    constexpr int U = (NA * NB <= 12) ? ASP_VL_U_SMALL : ASP_VL_U_BIG;
#pragma unroll U
    for (int kq = 0; kq < 8; ++kq) {
        const uint32_t aq = pq ^ (kq << 4), ac0 = pc0 ^ (kq << 4), ac1 = pc1 ^ (kq << 4);
        float4 qv[NA], cv[NB];
#pragma unroll
        for (int a = 0; a < NA; ++a) qv[a] = vl_lds128(aq + a * 1024);
#pragma unroll
        for (int b = 0; b < NB; ++b) cv[b] = vl_lds128((b & 1 ? ac1 : ac0) + b * 512);
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                acc[a][b] = __ffma2_rn(make_float2(qv[a].x, qv[a].y), make_float2(cv[b].x, cv[b].y), acc[a][b]);
                acc[a][b] = __ffma2_rn(make_float2(qv[a].z, qv[a].w), make_float2(cv[b].z, cv[b].w), acc[a][b]);
            }
    }
}

// Max-stabilised recomputation of both half-steps of one Sinkhorn step from the old potentials (the rare path of
// vl_sinkhorn_steps, when a plain sum left the fp32 range).  Works on the cost tile in shared memory (which the step
// loop never modifies) with rolled loops: one copy, out of line.
static __device__ __noinline__ void vl_stabilised_step(const float* tile, const float* us, const float* vs, int ql, int cl,
                                                       float t, float epsl, int lane, float& ft, float& gt) {
    const float nt = -t;
    float m = -INFINITY, sm = 0.f;
    if (lane < ql) {
        const float* row = tile + lane * kVlTileLd;
        for (int j = 0; j < cl; ++j) m = fmaxf(m, fmaf(row[j], nt, vs[j]));
        for (int j = 0; j < cl; ++j) sm += ex2(fmaf(row[j], nt, vs[j]) - m);
    }
    ft = -epsl * (m + lg2(sm));
    m = -INFINITY;
    sm = 0.f;
    if (lane < cl) {
        const float* col = tile + lane;
        for (int i = 0; i < ql; ++i) m = fmaxf(m, fmaf(col[i * kVlTileLd], nt, us[i]));
        for (int i = 0; i < ql; ++i) sm += ex2(fmaf(col[i * kVlTileLd], nt, us[i]) - m);
    }
    gt = -epsl * (m + lg2(sm));
}

// The eps-scaling loop of one pair, lane i owning row i AND column i.  Both sums of a step are computed from REGISTERS:
//   R_i = sum_j 2^(u_i + v_j - C_ij t) from the lane's row of C,   S_j = sum_i 2^(u_i + v_j - C_ij t) from its column,
// with u = la + f t and v = lb + g t of all lanes published through a double-buffered 2 x 32-float block of shared memory
// (ONE __syncwarp per step).  The first version computed every exponential once and transposed the 32 x 32 tile of
// exponentials through shared memory for the column sums: half the MUFU work, but two more barriers per step, 50 % more
// instructions, a serial chain row pass -> store -> barrier -> column pass, and 54 instead of ~21 shared-memory
// wavefronts per step on a kernel whose busiest unit is the shared-memory pipe.  Templated on the Gram bucket (8*NA rows,
// 4*NB columns) so that a step is straight-line code; padded rows/columns carry log-weight -1e5 and add exact zeros.
// A sum that leaves the fp32 range redoes the step in the max-stabilised form (rare).
template <int NA, int NB>
__device__ __forceinline__ void vl_sinkhorn_steps(const float2 (&Cr)[16], const float2 (&Cc)[16], float la, float lb, bool row_ok,
                                                  bool col_ok, int ql, int cl, const float2* step_s, int n_eps, float* uv, const float* tile,
                                                  int lane, float& f_out, float& g_out) {
    float f = 0.f, g = 0.f;
#pragma unroll 1
    for (int k = -1; k <= n_eps; ++k) {
        const float2 st = step_s[min(max(k, 0), n_eps - 1)];
        const float t = st.x, epsl = st.y;  // log2e / eps, eps * ln2
        const bool plain = (k < 0) | (k == n_eps);
        const float u = fmaf(f, t, la), v = fmaf(g, t, lb);  // f = g = 0 at k = -1
        float* us = uv + ((k & 1) ? 64 : 0);
        float* vs = us + 32;
        us[lane] = u;
        vs[lane] = v;
        __syncwarp();
        const float2 uu = dup2(u), vv = dup2(v), nt2 = dup2(-t);
        float2 R = f2(0.f, 0.f), S = f2(0.f, 0.f);
#pragma unroll
        for (int jc = 0; jc < NB; ++jc) {
            const float4 v4 = *reinterpret_cast<const float4*>(vs + 4 * jc);
            const float2 x0 = __ffma2_rn(Cr[2 * jc], nt2, __fadd2_rn(uu, f2(v4.x, v4.y)));
            const float2 x1 = __ffma2_rn(Cr[2 * jc + 1], nt2, __fadd2_rn(uu, f2(v4.z, v4.w)));
            R = __fadd2_rn(R, f2(ex2(x0.x), ex2(x0.y)));
            R = __fadd2_rn(R, f2(ex2(x1.x), ex2(x1.y)));
        }
#pragma unroll
        for (int ic = 0; ic < 2 * NA; ++ic) {
            const float4 u4 = *reinterpret_cast<const float4*>(us + 4 * ic);
            const float2 x0 = __ffma2_rn(Cc[2 * ic], nt2, __fadd2_rn(vv, f2(u4.x, u4.y)));
            const float2 x1 = __ffma2_rn(Cc[2 * ic + 1], nt2, __fadd2_rn(vv, f2(u4.z, u4.w)));
            S = __fadd2_rn(S, f2(ex2(x0.x), ex2(x0.y)));
            S = __fadd2_rn(S, f2(ex2(x1.x), ex2(x1.y)));
        }
        const float lr = lg2(R.x + R.y), ls = lg2(S.x + S.y);
        float ft = f - epsl * (lr - la), gt = g - epsl * (ls - lb);
        const bool ok = (!row_ok || fabsf(lr) < 1e30f) && (!col_ok || fabsf(ls) < 1e30f);
        if (!__all_sync(0xffffffffu, ok)) vl_stabilised_step(tile, us, vs, ql, cl, t, epsl, lane, ft, gt);
        f = row_ok ? (plain ? ft : 0.5f * (f + ft)) : 0.f;
        g = col_ok ? (plain ? gt : 0.5f * (g + gt)) : 0.f;
    }
    f_out = f;
    g_out = g;
}

// ---- the kernel: producer warp / Gram warps / Sinkhorn warps -------------------------------------------------------------
// 16 warps per SM in four warpgroups, each with its own register budget (setmaxnreg) and its own SMALL loop of code --
// the first version of this kernel ran all phases in every warp and spent 40-60 % of its cycles waiting for
// instructions (eight warps wandering through 140 KB of shape-specialised code):
//   warps 0-3   Gram warps (232 registers, one per scheduler): wait for a slice (mbarrier), multiply, release the ring
//               space; at the end of a pair turn the Gram values into distances and hand the cost tile to a Sinkhorn warp.
//   warps 4-11  Sinkhorn warps (96 registers, two per scheduler): Gram warp g feeds Sinkhorn warps 2g and 2g+1
//               alternately, each with two tile buffers (one being solved, one being filled).
//   warps 12-15 producer warps (80 registers), one per Gram warp: claim pairs from the global counter (one ahead) and
//               stream their slices into the Gram warp's byte ring with cp.async; completion is signalled per slice on
//               an mbarrier (cp.async.mbarrier.arrive.noinc), ring space comes back through two shared counters.
constexpr int kWsGram = 4, kWsSink = 8, kWsWarps = 16;
#ifndef ASP_VL_RING_UNITS
#define ASP_VL_RING_UNITS 28
#endif
constexpr int kWsRingUnits = ASP_VL_RING_UNITS;                // per Gram warp: 28 KB
constexpr int kWsRingFloats = kWsRingUnits * kVlUnitFloats;
constexpr int kWsSeq = 16;                                     // slice barriers per Gram warp (slices in flight < 16)
constexpr int kWsQueue = 8;                                    // pair descriptors per Gram warp
constexpr int kWsSmemFloats = kWsGram * kWsRingFloats + kWsSink * 2 * kVlTileFloats + kWsSink * 128 + kWsGram * 64 +
                              256;  // + slack to align the rings to 1 KB
constexpr int kWsGramRegs = 224, kWsSinkRegs = 120, kWsProdRegs = 48;  // 4*224 + 8*120 + 4*48 = 2048

struct WsShared {
    uint64_t slice_full[kWsGram][kWsSeq], slice_empty[kWsGram][kWsSeq];
    uint64_t tile_full[kWsSink][2], tile_empty[kWsSink][2];
    int dq[kWsGram][kWsQueue][4];  // b (-1 = end of stream), ql, cl
    volatile int dq_tail[kWsGram], dq_head[kWsGram], cons_vpos[kWsGram], cons_count[kWsGram];
    int tmeta[kWsSink][2][4];      // b (-1 = end), ql, cl
};

__device__ __forceinline__ void ws_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vl_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ws_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(vl_smem_u32(bar)) : "memory");
}
// arrive-on of this thread's earlier cp.async copies: fires when they have landed (does not change the pending count,
// so the barrier is initialised with one arrival per producer lane)
__device__ __forceinline__ void ws_cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(vl_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ws_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(vl_smem_u32(bar)), "r"(parity), "r"(1000000u)
            : "memory");
    } while (!ok);
}
template <int NS>
__device__ __forceinline__ void ws_mbar_wait_idle(uint64_t* bar, uint32_t parity) {
    if constexpr (NS == 0) {
        ws_mbar_wait(bar, parity);
    } else {
        for (;;) {
            uint32_t ok;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(vl_smem_u32(bar)), "r"(parity)
                : "memory");
            if (ok) break;
            __nanosleep(NS);
        }
    }
}
template <int N>
__device__ __forceinline__ void ws_setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void ws_setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__global__ void __launch_bounds__(kWsWarps * 32, 1)
ot_varlen_kernel(const VlArgs a, const EpsSched sched, const OtOut out) {
    extern __shared__ __align__(16) float smem_raw[];
    float* smem = smem_raw + (((1024u - (vl_smem_u32(smem_raw) & 1023u)) & 1023u) >> 2);  // rings start on a 1 KB boundary
    __shared__ float2 step_s[ASP_MAX_EPS];  // per schedule entry: log2e / eps, eps * ln2
    __shared__ WsShared sh;
    for (int k = threadIdx.x; k < sched.n; k += blockDim.x) step_s[k] = make_float2(kLog2e / sched.eps[k], sched.eps[k] * kLn2);
    if (threadIdx.x < kWsGram * kWsSeq) {
        ws_mbar_init(&sh.slice_full[0][0] + threadIdx.x, 32);
        ws_mbar_init(&sh.slice_empty[0][0] + threadIdx.x, 1);
    }
    if (threadIdx.x < kWsSink * 2) {
        ws_mbar_init(&sh.tile_full[0][0] + threadIdx.x, 1);
        ws_mbar_init(&sh.tile_empty[0][0] + threadIdx.x, 1);
    }
    if (threadIdx.x < kWsGram) {
        sh.dq_tail[threadIdx.x] = 0;
        sh.dq_head[threadIdx.x] = 0;
        sh.cons_vpos[threadIdx.x] = 0;
        sh.cons_count[threadIdx.x] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tiles = smem + (size_t)kWsGram * kWsRingFloats;                       // [kWsSink][2][kVlTileFloats]
    float* sink_misc = tiles + (size_t)kWsSink * 2 * kVlTileFloats;              // [kWsSink][2][u | v]
    float* gram_misc = sink_misc + (size_t)kWsSink * 128;                        // [kWsGram][64]: row norms
    const int D = a.D, nsl = D >> 5;

    if (warp >= kWsGram + kWsSink) {
        // ============================== producer warp of Gram warp g ==========================================
        ws_setmaxnreg_dec<kWsProdRegs>();
        const int g = warp - (kWsGram + kWsSink);
        const int r0 = lane >> 3, chunk = lane & 7;  // this lane copies rows r0 + 4m, 16-byte chunk `chunk` of the piece
        const unsigned rstride = 16u * (unsigned)D;  // bytes between the rows r and r + 4
        const uint32_t ring_u32 = vl_smem_u32(smem + (size_t)g * kWsRingFloats);
        int nx_b = 0, nx_ql = 0, nx_cl = 0, nx_ci = 0;  // the pair claimed one ahead (valid in lane 0 until broadcast)
        auto claim = [&]() {
            if (lane == 0) {
                nx_b = (int)atomicAdd(a.counter, 1u);
                if (nx_b < a.B) {
                    if (a.order) nx_b = a.order[nx_b];
                    nx_ci = a.c_index ? a.c_index[nx_b] : nx_b;
                    nx_ql = min(max(a.q_lens[nx_b / a.q_group], 0), a.Sq);
                    nx_cl = min(max(a.c_lens[nx_ci], 0), a.Sc);
                }
            }
        };
        claim();
        int npush = 0, off = 0, vpos = 0, issued = 0;
        for (;;) {
            // ---- next pair of this stream ----
            while (npush - sh.dq_head[g] >= kWsQueue) __nanosleep(64);
            const int nb = __shfl_sync(0xffffffffu, nx_b, 0);
            int* d = sh.dq[g][npush & (kWsQueue - 1)];
            if (nb >= a.B) {  // no pair left: end marker
                if (lane == 0) {
                    d[0] = -1;
                    __threadfence_block();
                    sh.dq_tail[g] = npush + 1;
                }
                break;
            }
            const int pql = __shfl_sync(0xffffffffu, nx_ql, 0), pcl = __shfl_sync(0xffffffffu, nx_cl, 0);
            const int ci = __shfl_sync(0xffffffffu, nx_ci, 0);
            if (lane == 0) {
                d[0] = nb;
                d[1] = pql;
                d[2] = pcl;
                __threadfence_block();
                sh.dq_tail[g] = npush + 1;
            }
            ++npush;
            const char* qsrc = reinterpret_cast<const char*>(a.q + ((size_t)(nb / a.q_group) * a.Sq + r0) * D + chunk * 4);
            const char* csrc = reinterpret_cast<const char*>(a.c + ((size_t)ci * a.Sc + r0) * D + chunk * 4);
            claim();  // its latency hides under this pair's slices
            if (pql + pcl == 0) continue;  // nothing to stream for an empty pair
            const int pu = (pql + pcl + 7) >> 3;
            const int mq = (pql + 3) >> 2, mc = (pcl + 3) >> 2;  // row groups of 4 on each side
            const bool last_q = r0 + 4 * (mq - 1) < pql, last_c = r0 + 4 * (mc - 1) < pcl;  // this lane has a row in the last group
#pragma unroll 1
            for (int s = 0; s < nsl; ++s) {
                int o = off, waste = 0;
                if (o + pu > kWsRingUnits) {
                    waste = kWsRingUnits - o;
                    o = 0;
                }
                // room in the ring, and the barriers of slice (issued - 16) free again?  If not, park on the "consumed"
                // barrier of the oldest slice in flight (no polling: the Gram warp of this scheduler needs the issue slots)
                for (;;) {
                    const int cc = sh.cons_count[g];
                    const int cv = sh.cons_vpos[g];  // read after the count: never older than it
                    if (vpos + waste + pu - cv <= kWsRingUnits && issued - cc < kWsSeq - 1) break;
                    ws_mbar_wait_idle<ASP_VL_IDLE_PROD>(&sh.slice_empty[g][cc & (kWsSeq - 1)], (cc >> 4) & 1);
                }
                // row r0 + 4m of each side; chunk position = chunk ^ (row & 7), which alternates with the parity of m
                const uint32_t slot = ring_u32 + (uint32_t)o * 1024u;
                const uint32_t dqe = slot + r0 * 128 + ((chunk ^ r0) << 4), dqo = dqe ^ 64;
                const int rcp = pql + r0;
                const uint32_t dce = slot + rcp * 128 + ((chunk ^ (rcp & 7)) << 4), dco = dce ^ 64;
                // every group but the last is complete for every lane: no per-lane predicate
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    if (m >= mq - 1) break;
                    vl_cp_async16((m & 1 ? dqo : dqe) + m * 512, reinterpret_cast<const float*>(qsrc + (size_t)(m * rstride)));
                }
                if (mq > 0 && last_q)
                    vl_cp_async16(((mq - 1) & 1 ? dqo : dqe) + (mq - 1) * 512,
                                  reinterpret_cast<const float*>(qsrc + (size_t)((mq - 1) * rstride)));
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    if (m >= mc - 1) break;
                    vl_cp_async16((m & 1 ? dco : dce) + m * 512, reinterpret_cast<const float*>(csrc + (size_t)(m * rstride)));
                }
                if (mc > 0 && last_c)
                    vl_cp_async16(((mc - 1) & 1 ? dco : dce) + (mc - 1) * 512,
                                  reinterpret_cast<const float*>(csrc + (size_t)((mc - 1) * rstride)));
                ws_cp_async_arrive(&sh.slice_full[g][issued & (kWsSeq - 1)]);
                ++issued;
                qsrc += 128;
                csrc += 128;
                off = o + pu;
                vpos += waste + pu;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        return;
    }

    if (warp < kWsGram) {
        // ============================== Gram warp =============================================================
        ws_setmaxnreg_inc<kWsGramRegs>();
        const int g = warp;
        const uint32_t ring_s = vl_smem_u32(smem + (size_t)g * kWsRingFloats);
        float* nrm = gram_misc + g * 64;
        // lane grid: 8 query-row groups x 4 candidate-row groups.  Which lanes share an address decides what a 128-bit
        // broadcast load costs: 2 passes of the shared-memory pipe when every aligned group of four lanes holds at most
        // two distinct addresses, 4 otherwise (tools/ubench3.cu, profiles/r02_3q_lds_wavefronts.txt).  li takes lane bits
        // 0, 2, 3 and lj bits 1, 4: a quad then sees two query rows and two candidate rows (li = lane & 7 made every
        // query-side load a 4-pass one).
#if ASP_VL_QUAD_MAP
        const int li = (lane & 1) | ((lane >> 1) & 6), lj = ((lane >> 1) & 1) | ((lane >> 3) & 2);
#else
        const int li = lane & 7, lj = lane >> 3;
#endif
        const int r0 = lane >> 3, chunk = lane & 7;  // norm duty: the pieces the producer lane of the same index copied
        int off = 0, vpos = 0, n = 0, popped = 0, t = 0;
        for (;;) {
            while (sh.dq_tail[g] <= popped) __nanosleep(64);
            const int* d = sh.dq[g][popped & (kWsQueue - 1)];
            const int b = d[0], ql = d[1], cl = d[2];
            __syncwarp();
            ++popped;
            if (lane == 0) sh.dq_head[g] = popped;
            if (b < 0) break;
            const int NA = (ql + 7) >> 3, NB = (cl + 3) >> 2;  // Gram buckets: 8 query rows x 4 candidate rows
            const int units = (ql + cl + 7) >> 3;
            const int shape = (ql > 0 && cl > 0) ? ((a.dev_flags & 1) ? 31 : (NA - 1) * 8 + (NB - 1)) : -1;
            float2 acc[4][8], nq[8], nc[8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);
#pragma unroll
            for (int m = 0; m < 8; ++m) nq[m] = nc[m] = make_float2(0.f, 0.f);
            const int ns = (ql + cl > 0) ? nsl : 0;
            // the slice loop lives INSIDE each shape variant: no indirect branch per slice
            auto run_pair = [&](auto na_c, auto nb_c) {
                constexpr int NAc = decltype(na_c)::value, NBc = decltype(nb_c)::value;
#pragma unroll 1
                for (int s = 0; s < ns; ++s) {
                    ws_mbar_wait(&sh.slice_full[g][n & (kWsSeq - 1)], (n >> 4) & 1);
                    int o = off, waste = 0;
                    if (o + units > kWsRingUnits) {
                        waste = kWsRingUnits - o;
                        o = 0;
                    }
                    if constexpr (NAc > 0) vl_slice<NAc, NBc>(ring_s + (uint32_t)o * 1024u, ql, li, lj, r0, chunk, acc, nq, nc);
                    __syncwarp();  // every lane is done with the slot
                    off = o + units;
                    vpos += waste + units;
                    if (lane == 0) {
                        sh.cons_vpos[g] = vpos;
                        sh.cons_count[g] = n + 1;
                        ws_mbar_arrive(&sh.slice_empty[g][n & (kWsSeq - 1)]);  // release; wakes a parked producer
                    }
                    ++n;
                }
            };
            switch (shape) {
#define ASP_VL_CASE(na, nb) \
    case (na - 1) * 8 + (nb - 1): run_pair(std::integral_constant<int, na>{}, std::integral_constant<int, nb>{}); break;
#define ASP_VL_ROW(na)                                                                                        \
    ASP_VL_CASE(na, 1) ASP_VL_CASE(na, 2) ASP_VL_CASE(na, 3) ASP_VL_CASE(na, 4) ASP_VL_CASE(na, 5) ASP_VL_CASE(na, 6) \
    ASP_VL_CASE(na, 7) ASP_VL_CASE(na, 8)
                ASP_VL_ROW(1) ASP_VL_ROW(2) ASP_VL_ROW(3) ASP_VL_ROW(4)
#undef ASP_VL_ROW
#undef ASP_VL_CASE
                default:  // a pair with an empty side: its rows are drained unread
                    run_pair(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
                    break;
            }
            // ---- squared norms: sum the 8 chunk lanes of each row ----
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                float tq = nq[m].x + nq[m].y, tc = nc[m].x + nc[m].y;
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    tq += __shfl_xor_sync(0xffffffffu, tq, o);
                    tc += __shfl_xor_sync(0xffffffffu, tc, o);
                }
                if (chunk == 0) {
                    nrm[r0 + 4 * m] = tq;
                    nrm[32 + r0 + 4 * m] = tc;
                }
            }
            // ---- the tile buffer this pair goes to ----
            const int sw = 2 * g + (t & 1), buf = (t >> 1) & 1;
            float* tile = tiles + (size_t)(sw * 2 + buf) * kVlTileFloats;
            if (a.mode == 0) ws_mbar_wait(&sh.tile_empty[sw][buf], ((t >> 2) & 1) ^ 1);
            __syncwarp();
            // ---- distances -> cost tile (0 outside the valid block) ----
            float run_best = kPadNeg;
            int run_idx = 0x7fffffff;
#pragma unroll
            for (int ai = 0; ai < 4; ++ai) {
                if (ai < NA) {
                    const int i = li + 8 * ai;
                    const float qn = nrm[i];
#pragma unroll
                    for (int bj = 0; bj < 8; ++bj) {
                        if (bj < NB) {
                            const int j = lj + 4 * bj;
                            const float d2 = qn + nrm[32 + j] - 2.f * (acc[ai][bj].x + acc[ai][bj].y);
                            const float dist = vl_sqrt_approx(fmaxf(d2, 1e-8f));
                            const bool valid = i < ql && j < cl;
                            tile[i * kVlTileLd + j] = valid ? dist : 0.f;
                            if (valid) {
                                const float sim = -dist;
                                const int idx = i * a.Sc + j;
                                if (sim > run_best || (sim == run_best && idx < run_idx)) {
                                    run_best = sim;
                                    run_idx = idx;
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (a.mode == 0) {
                if (lane == 0) {
                    int* tm = sh.tmeta[sw][buf];
                    tm[0] = b;
                    tm[1] = ql;
                    tm[2] = cl;
                    ws_mbar_arrive(&sh.tile_full[sw][buf]);  // release: tile + descriptor visible to the Sinkhorn warp
                }
                ++t;
                continue;
            }
            // ---- tsAspire: masked max, first flat index on ties (pair_distances.py:173-176) ----
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, run_best, s);
                const int oi = __shfl_xor_sync(0xffffffffu, run_idx, s);
                if (ob > run_best || (ob == run_best && oi < run_idx)) {
                    run_best = ob;
                    run_idx = oi;
                }
            }
            if (run_idx != 0x7fffffff) {  // near-duplicate winner: re-evaluate it directly (common.cuh: l2max_refine)
                const int wi = run_idx / a.Sc, wj = run_idx - wi * a.Sc;
                if (l2max_needs_refine(run_best, nrm[wi] + nrm[32 + wj]))
                    run_best = l2max_refine(a.q + ((size_t)(b / a.q_group) * a.Sq + wi) * D, a.c + ((size_t)b * a.Sc + wj) * D, D, lane);
            }
            if (lane == 0) {
                a.best[b] = run_best;
                if (a.flat_idx) a.flat_idx[b] = (run_idx == 0x7fffffff) ? 0 : run_idx;
            }
            if (a.pair_sims) {
                float* ps = a.pair_sims + (size_t)b * a.Sq * a.Sc;
                for (int e = lane; e < a.Sq * a.Sc; e += 32) {
                    const int i = e / a.Sc, j = e - i * a.Sc;
                    ps[e] = (i < ql && j < cl) ? -tile[i * kVlTileLd + j] : kPadNeg;
                }
            }
            __syncwarp();
        }
        if (a.mode == 0) {  // end markers for the two Sinkhorn warps of this Gram warp
#pragma unroll 1
            for (int e = 0; e < 2; ++e, ++t) {
                const int sw = 2 * g + (t & 1), buf = (t >> 1) & 1;
                ws_mbar_wait(&sh.tile_empty[sw][buf], ((t >> 2) & 1) ^ 1);
                if (lane == 0) {
                    sh.tmeta[sw][buf][0] = -1;
                    ws_mbar_arrive(&sh.tile_full[sw][buf]);
                }
            }
        }
        return;
    }

    // ============================== Sinkhorn warp ==========================================================
    ws_setmaxnreg_dec<kWsSinkRegs>();
    if (a.mode != 0) return;
    const int sw = warp - kWsGram;
    float* uv = sink_misc + sw * 128;  // [2][u 32 | v 32]: the potentials of a step, double buffered
    float* vs = uv;                    // (afterwards: g and beta for the plan)
    float* bs = uv + 32;
#pragma unroll 1
    for (int t = 0;; ++t) {
        const int buf = t & 1;
        ws_mbar_wait_idle<ASP_VL_IDLE_SINK>(&sh.tile_full[sw][buf], (t >> 1) & 1);
        const int* tm = sh.tmeta[sw][buf];
        const int b = tm[0], ql = tm[1], cl = tm[2];
        if (b < 0) break;
        float* tile = tiles + (size_t)(sw * 2 + buf) * kVlTileFloats;
        const int NA = (ql + 7) >> 3, NB = (cl + 3) >> 2;
        const bool row_ok = lane < ql, col_ok = lane < cl;
        float2 Cr[16];
#pragma unroll
        for (int jc = 0; jc < 8; ++jc) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (jc < NB && row_ok) v = *reinterpret_cast<const float4*>(tile + lane * kVlTileLd + 4 * jc);
            Cr[2 * jc] = f2(v.x, v.y);
            Cr[2 * jc + 1] = f2(v.z, v.w);
        }
        float alpha, beta, la, lb;
        {
            // pair_distances.py:57-60: log_softmax over the valid sentences of (-min dist) / T, exp; geomloss then takes
            // the log again with -1e5 for zero weights
            float rmin = INFINITY, cmin = INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < cl) rmin = fminf(rmin, (j & 1) ? Cr[j >> 1].y : Cr[j >> 1].x);
            if (col_ok)
                for (int i = 0; i < ql; ++i) cmin = fminf(cmin, tile[i * kVlTileLd + lane]);
            const float xr = row_ok ? -rmin * a.inv_temp : -INFINITY, xc = col_ok ? -cmin * a.inv_temp : -INFINITY;
            const float mr = warp_max(xr), mc = warp_max(xc);
            const float sr = warp_sum(row_ok ? expf(xr - mr) : 0.f), sc = warp_sum(col_ok ? expf(xc - mc) : 0.f);
            const bool both = ql > 0 && cl > 0;  // a pair with an empty side carries no mass at all
            alpha = (row_ok && both) ? expf(xr - (mr + logf(sr))) : 0.f;
            beta = (col_ok && both) ? expf(xc - (mc + logf(sc))) : 0.f;
            la = (alpha > 0.f) ? log2f(alpha) : kLogZeroWeight * kLog2e;
            lb = (beta > 0.f) ? log2f(beta) : kLogZeroWeight * kLog2e;
        }
        // this lane's COLUMN of the tile as well (rows beyond the bucket / the document: 0, like the padded row entries)
        float2 Cc[16];
#pragma unroll
        for (int ic = 0; ic < 16; ++ic) {
            float c0 = 0.f, c1 = 0.f;
            if (2 * ic < 8 * NA && col_ok) {
                if (2 * ic < ql) c0 = tile[(2 * ic) * kVlTileLd + lane];
                if (2 * ic + 1 < ql) c1 = tile[(2 * ic + 1) * kVlTileLd + lane];
            }
            Cc[ic] = f2(c0, c1);
        }
        float f = 0.f, g = 0.f;
        float* erow = tile + lane * kVlTileLd;
        if (ql > 0 && cl > 0) {
            switch ((NA - 1) * 8 + (NB - 1)) {
#define ASP_VL_CASE(na, nb)                                                                                               \
    case (na - 1) * 8 + (nb - 1):                                                                                         \
        vl_sinkhorn_steps<na, nb>(Cr, Cc, la, lb, row_ok, col_ok, ql, cl, step_s, sched.n, uv, tile, lane, f, g);         \
        break;
#define ASP_VL_ROW(na)                                                                                        \
    ASP_VL_CASE(na, 1) ASP_VL_CASE(na, 2) ASP_VL_CASE(na, 3) ASP_VL_CASE(na, 4) ASP_VL_CASE(na, 5) ASP_VL_CASE(na, 6) \
    ASP_VL_CASE(na, 7) ASP_VL_CASE(na, 8)
                ASP_VL_ROW(1) ASP_VL_ROW(2) ASP_VL_ROW(3) ASP_VL_ROW(4)
#undef ASP_VL_ROW
#undef ASP_VL_CASE
                default: break;
            }
        }
        __syncwarp();

        const float dual = warp_sum(alpha * f + beta * g);
        if (out.dual && lane == 0) out.dual[b] = dual;
        const int Sq = a.Sq, Sc = a.Sc;
        if (lane < Sq) {
            if (out.f) out.f[(size_t)b * Sq + lane] = f;
            if (out.alpha) out.alpha[(size_t)b * Sq + lane] = alpha;
        }
        if (lane < Sc) {
            if (out.g) out.g[(size_t)b * Sc + lane] = g;
            if (out.beta) out.beta[(size_t)b * Sc + lane] = beta;
        }
        if (out.primal || out.plan || out.weighted || out.neg_cost) {
            // plan (pair_distances.py:76-85): exp((f_i + g_j - C_ij) / blur) * alpha_i * beta_j, blur = eps_final
            const float tf = kLog2e / sched.eps[sched.n - 1];
            vs[lane] = g;
            bs[lane] = beta;
#pragma unroll
            for (int jc = 0; jc < 8; ++jc)  // the costs go back to the tile (this lane's row): the loop below is rolled
                *reinterpret_cast<float4*>(erow + 4 * jc) = make_float4(Cr[2 * jc].x, Cr[2 * jc].y, Cr[2 * jc + 1].x, Cr[2 * jc + 1].y);
            __syncwarp();
            float primal = 0.f;
#pragma unroll 4
            for (int j = 0; j < Sc; ++j) {
                const bool valid = row_ok && j < cl;
                const float cij = valid ? erow[j] : 0.f;
                const float p = valid ? ex2((f + vs[j] - cij) * tf) * (alpha * bs[j]) : 0.f;
                const float negc = -cij;
                const float w = p * negc;
                primal += w;
                if (lane < Sq) {
                    const size_t o = (size_t)b * Sq * Sc + (size_t)lane * Sc + j;
                    if (out.neg_cost) out.neg_cost[o] = valid ? negc : 0.f;
                    if (out.plan) out.plan[o] = p;
                    if (out.weighted) out.weighted[o] = valid ? w : 0.f;
                }
            }
            primal = warp_sum(primal);
            if (out.primal && lane == 0) out.primal[b] = primal;
        }
        __syncwarp();
        if (lane == 0) ws_mbar_arrive(&sh.tile_empty[sw][buf]);
    }
}

// ---- pre-pass: claim order = pairs sorted by Gram shape, largest first ------------------------------------------------
// The Gram code is specialised per shape (32 variants, ~250 KB of instructions in all).  With pairs in arbitrary order
// the four Gram warps of an SM are in four different variants and change variant with every pair: the instruction
// caches miss all the time (measured: 16 ms per 100 k pairs against 5.6 ms with every pair forced onto ONE variant).
// Sorted by shape, all warps of the GPU sit in the same one or two variants for long stretches, and the largest pairs
// go first, which also evens out the tail.  Two tiny kernels (histogram, scatter); the order inside a shape is
// arbitrary, which never shows: a pair's result does not depend on where or when it is computed.
constexpr int kVlBins = 34;  // 32 shapes + pairs with an empty side + spare
__device__ __forceinline__ int vl_shape_bin(int ql, int cl) {
    if (ql <= 0 || cl <= 0) return 32;
    const int NA = (ql + 7) >> 3, NB = (cl + 3) >> 2;
    return 31 - ((NA - 1) * 8 + (NB - 1));  // bin 0 = 32 x 32 sentences
}
__global__ void vl_hist_kernel(const VlArgs a, unsigned int* hist) {
    __shared__ unsigned int h[kVlBins];
    if (threadIdx.x < kVlBins) h[threadIdx.x] = 0;
    __syncthreads();
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < a.B; b += gridDim.x * blockDim.x) {
        const int ci = a.c_index ? a.c_index[b] : b;
        atomicAdd(&h[vl_shape_bin(min(a.q_lens[b / a.q_group], a.Sq), min(a.c_lens[ci], a.Sc))], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kVlBins && h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}
__global__ void vl_scatter_kernel(const VlArgs a, const unsigned int* hist, unsigned int* cursor, int32_t* order) {
    __shared__ unsigned int base[kVlBins], cnt[kVlBins], off[kVlBins];
    if (threadIdx.x < kVlBins) cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int k = 0; k < kVlBins; ++k) {
            base[k] = run;
            run += hist[k];
        }
    }
    __syncthreads();
    // block-local ranks first, one global atomic per (block, bin)
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    int bin = -1;
    unsigned int rank = 0;
    if (b < a.B) {
        const int ci = a.c_index ? a.c_index[b] : b;
        bin = vl_shape_bin(min(a.q_lens[b / a.q_group], a.Sq), min(a.c_lens[ci], a.Sc));
        rank = atomicAdd(&cnt[bin], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kVlBins && cnt[threadIdx.x]) off[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], cnt[threadIdx.x]);
    __syncthreads();
    if (bin >= 0) order[base[bin] + off[bin] + rank] = b;
}

int g_vl_flags = 0;
int g_ot_varlen = 1;  // asp_set_option("ot_varlen"): 0 = long documents take the two-kernel path (cost tensor through HBM)

bool ot_varlen_supported(int Sq, int Sc, int D) {
    return g_ot_varlen && Sq <= kVlMaxS && Sc <= kVlMaxS && D >= 32 && (D % 32) == 0;
}
constexpr int kVlSortMin = 4096;  // below this many pairs every Gram warp sees only a handful: not worth two launches
size_t ot_varlen_workspace_bytes(int B) { return B >= kVlSortMin ? (size_t)B * sizeof(int32_t) + 2 * kVlBins * sizeof(unsigned int) : 0; }

static int vl_launch(VlArgs a, const EpsSched& sched, const OtOut& out, void* workspace, size_t workspace_bytes,
                     cudaStream_t stream) {
    static std::atomic<unsigned int> next_slot{0};
    static thread_local int attr_dev = -1;
    static thread_local unsigned int* counters = nullptr;
    const int smem = kWsSmemFloats * (int)sizeof(float);
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        ASP_CUDA(cudaFuncSetAttribute(ot_varlen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ASP_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&counters), g_vl_counter));
        attr_dev = dev;
    }
    // launch-owned pair counter: zeroed on the launch's own stream, so an aborted earlier launch cannot leave it armed
    a.counter = counters + (next_slot.fetch_add(1) % kVlCounterSlots);
    ASP_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), stream));
    a.order = nullptr;
    const size_t need = ot_varlen_workspace_bytes(a.B);
    if (need && workspace && workspace_bytes >= need && !(g_vl_flags & 2)) {
        unsigned int* hist = static_cast<unsigned int*>(workspace);
        int32_t* order = reinterpret_cast<int32_t*>(hist + 2 * kVlBins);
        ASP_CUDA(cudaMemsetAsync(hist, 0, 2 * kVlBins * sizeof(unsigned int), stream));
        const int threads = 256, blocks = (a.B + threads - 1) / threads;
        vl_hist_kernel<<<std::min(blocks, 4 * sm_count()), threads, 0, stream>>>(a, hist);
        ASP_LAUNCH_CHECK("vl_hist_kernel");
        vl_scatter_kernel<<<blocks, threads, 0, stream>>>(a, hist, hist + kVlBins, order);
        ASP_LAUNCH_CHECK("vl_scatter_kernel");
        a.order = order;
    }
    const int ctas = std::min(sm_count(), (a.B + kWsGram - 1) / kWsGram);
    ot_varlen_kernel<<<ctas, kWsWarps * 32, smem, stream>>>(a, sched, out);
    ASP_LAUNCH_CHECK("ot_varlen_kernel");
    return ASP_OK;
}

int ot_varlen_launch(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens,
                     const int32_t* c_index, int B, int Sq, int Sc, int D, const EpsSched& sched, float temp, const OtOut& out,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    VlArgs a{q, q_lens, c, c_lens, c_index, q_group, B, Sq, Sc, D, 1.0f / temp, nullptr, nullptr, 0, g_vl_flags, nullptr, nullptr, nullptr};
    return vl_launch(a, sched, out, workspace, workspace_bytes, stream);
}

int l2max_varlen_launch(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens, int B,
                        int Sq, int Sc, int D, float* best, int32_t* flat_idx, float* pair_sims, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream) {
    VlArgs a{q, q_lens, c, c_lens, nullptr, q_group, B, Sq, Sc, D, 1.0f, nullptr, nullptr, 1, g_vl_flags, best, flat_idx, pair_sims};
    EpsSched sched;
    sched.n = 1;
    sched.eps[0] = 1.f;
    OtOut none{};
    return vl_launch(a, sched, none, workspace, workspace_bytes, stream);
}

}  // namespace asp
