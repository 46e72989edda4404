// Shared device/host helpers for the aspire_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/aspire_b200.h"

namespace asp {

// ---- error plumbing (thread-local message, returned through asp_last_error) -------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch();  // every kernel launch of this library is counted (asp_launch_count)

#define ASP_REQUIRE(cond, ...)                       \
    do {                                             \
        if (!(cond)) {                               \
            ::asp::set_error(__VA_ARGS__);           \
            return ASP_ERR_INVALID;                  \
        }                                            \
    } while (0)

#define ASP_CUDA(call)                                                  \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return ::asp::cuda_fail(e__, #call);    \
    } while (0)

#define ASP_LAUNCH_CHECK(name)                                               \
    do {                                                                     \
        cudaError_t e__ = cudaGetLastError();                                \
        if (e__ != cudaSuccess) return ::asp::cuda_fail(e__, "launch " name); \
        ::asp::count_launch();                                               \
    } while (0)

// ---- programmatic dependent launch ---------------------------------------------------------------------------------
// The encoder is ~85 dependent kernels on one stream.  Launched with the programmatic-serialization attribute, kernel
// n+1 is placed on the SMs while kernel n drains, runs its prologue (barrier set-up, TMEM allocation, descriptor
// prefetch) and then blocks in pdl_wait() until kernel n has completed and its writes are visible.  Every kernel
// launched through launch_pdl() MUST call pdl_wait() before its first access to global data another kernel writes or
// reads; pdl_trigger() (all threads, at the top) lets the next kernel be placed as soon as every CTA of this one has
// started.  A kernel that follows a plain launch simply starts after it, as usual.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
extern int g_pdl;  // asp_set_option("pdl"): 1 (default) on, 0 plain stream order
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
int sm_count();
extern int g_ot_kernel;  // asp_set_option("ot_kernel")
extern int g_gemm_kernel;    // asp_set_option("gemm_kernel")
extern int g_gemm_cluster;   // asp_set_option("gemm_cluster")
extern int g_gemm_pair;      // asp_set_option("gemm_pair")
extern int g_attn_tc;        // asp_set_option("attn_tc")
// asp_set_option("ln_on_read"): 1 (default) = the encoder's inner LayerNorms write only the bf16 GEMM operand and per-row
// (mean, rstd); the residual epilogue of the next GEMM normalises the pre-LayerNorm rows as it reads them.  0 = every
// LayerNorm writes the fp32 residual stream (one 4-byte write and one read less per element and LayerNorm with 1).
extern int g_ln_on_read;
struct LnOnRead {
    const void* stats;  // float2 (mean, rstd) per row
    const float* gamma;
    const float* beta;
};
extern int g_span_tma;       // asp_set_option("span_tma")

// Epsilon schedule passed by value in kernel parameter space (uniform, read through the constant bank).
struct EpsSched {
    int n;
    float eps[ASP_MAX_EPS];
};

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kPadNeg = -1.0e9f;        // pair_distances.py:39 mask constant
constexpr float kLogZeroWeight = -100000.0f;  // geomloss log_weights() value for zero-mass points

// ---- cross-file launch helpers (q_group = consecutive candidates sharing one query; 1 = paired) ----------
struct OtOut;
int check_pair_args(const float* q, const int32_t* q_lens, const float* c, const int32_t* c_lens, int B, int Sq, int Sc,
                    int D);
int pair_cost_launch(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens, int B,
                     int Sq, int Sc, int D, float* cost, cudaStream_t stream);
int launch_sinkhorn(const float* cost, const int32_t* q_lens, int q_group, const int32_t* c_lens, int B, int Sq, int Sc,
                    const EpsSched& sched, float temp, const OtOut& out, cudaStream_t stream);
int make_sched(const float* eps_host, int n_eps, EpsSched* s);
// long / ragged documents (<= 32 x 32 sentences): cost tile + Sinkhorn (or the tsAspire max) in one kernel (ot_varlen.cu)
extern int g_ot_varlen;  // asp_set_option("ot_varlen")
extern int g_vl_flags;   // asp_set_option("vl_flags")
bool ot_varlen_supported(int Sq, int Sc, int D);
int ot_varlen_launch(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens,
                     const int32_t* c_index, int B, int Sq, int Sc, int D, const EpsSched& sched, float temp, const OtOut& out,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream);
int l2max_varlen_launch(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens, int B,
                        int Sq, int Sc, int D, float* best, int32_t* flat_idx, float* pair_sims, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream);
size_t ot_varlen_workspace_bytes(int B);  // scratch for the shape sort (0: batch too small to bother)
OtOut to_out(const asp_ot_outputs* o);
// Q x C all-pairs otAspire on tcgen05 (ot_allpairs.cu): documents of <= 10 sentences, D % 64 == 0
// pools (one query against >= 128 candidates): Gram tiles on tcgen05, candidates split to bf16 hi/lo by producer warps
extern int g_ot_fused_tc;  // asp_set_option("ot_fused_tc")
bool ot_fused_tc_supported(int q_group, int B, int Sq, int Sc, int D);
int ot_fused_tc_launch(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens,
                       const int32_t* c_index, int B, int Sq, int Sc, int D, const EpsSched& sched, float temp, const OtOut& out,
                       cudaStream_t stream);
extern int g_oa_warps;  // asp_set_option("oa_warps")
bool ot_allpairs_supported(int Sq, int Sc, int D);
size_t ot_allpairs_workspace_bytes(int NQ, int NC, int Sq, int Sc, int D);
int ot_allpairs_launch(const float* q, const int32_t* q_lens, int NQ, const float* c, const int32_t* c_lens, int NC, int Sq,
                       int Sc, int D, const EpsSched& sched, float temp, float* scores, void* workspace, cudaStream_t stream);

// ---- device math ------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, s));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// tsAspire's reference distance is torch.cdist / scipy cdist (pair_distances.py:167, pp_gen_nearest.py:942), which for
// abstracts (<= 25 sentences) is the DIRECT sum (q - c)^2: identical sentences are at distance exactly 0, where
// |q|^2 + |c|^2 - 2 q.c leaves cancellation noise (and the 1e-8 clamp of geomloss' formula leaves 1e-4).  The winning
// sentence pair of a document pair is therefore re-evaluated directly when its squared distance is small against the
// norms (near-duplicates: rare, and exactly where the ranking is decided).  One warp, rows from L1/L2.
#ifdef __CUDACC__
__device__ __forceinline__ float l2max_refine(const float* qrow, const float* crow, int D, int lane) {
    float s0 = 0.f, s1 = 0.f;
    for (int k4 = lane; k4 < (D >> 2); k4 += 32) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(qrow) + k4), y = __ldg(reinterpret_cast<const float4*>(crow) + k4);
        const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s0 = fmaf(d2, d2, s0); s1 = fmaf(d3, d3, s1);
    }
    return -sqrtf(warp_sum(s0 + s1));
}
__device__ __forceinline__ bool l2max_needs_refine(float best, float norm_sum) {
    return best > kPadNeg && best * best < 0.01f * norm_sum;
}
#endif

// Streaming 128-bit load that does not allocate in L1 (data read exactly once).
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace asp
