// C-ABI glue: error state, version, options, and the composed otAspire entry point.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "common.cuh"

namespace asp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return ASP_ERR_CUDA;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace asp

extern "C" int asp_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* asp_last_error(void) { return asp::g_err; }

extern "C" int asp_sm_count(void) {
    int dev = 0, n = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    ASP_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    return n;
}

extern "C" long long asp_launch_count(void) { return asp::g_launches.load(std::memory_order_relaxed); }

extern "C" int asp_set_option(const char* key, int value) {
    ASP_REQUIRE(key, "asp_set_option: NULL key");
    if (strcmp(key, "ot_kernel") == 0) {
        ASP_REQUIRE(value >= 0 && value <= 2, "asp_set_option: ot_kernel must be 0 (auto), 1 (warp) or 2 (thread)");
        asp::g_ot_kernel = value;
        return ASP_OK;
    }
    if (strcmp(key, "ot_varlen") == 0) {  // developer switch: 1 one-kernel path for 11..32-sentence documents, 0 cost + Sinkhorn kernels
        asp::g_ot_varlen = value != 0;
        return ASP_OK;
    }
    if (strcmp(key, "vl_flags") == 0) {  // developer switches of ot_varlen.cu
        asp::g_vl_flags = value;
        return ASP_OK;
    }
    if (strcmp(key, "ot_fused_tc") == 0) {  // developer switch: 1 pools take the tcgen05 fused kernel, 0 the FFMA2 one
        asp::g_ot_fused_tc = value != 0;
        return ASP_OK;
    }
    if (strcmp(key, "span_tma") == 0) {  // developer switch: 1 span pooling staged by cp.async.bulk, 0 streaming loads
        ASP_REQUIRE(value == 0 || value == 1, "asp_set_option: span_tma must be 0 or 1");
        asp::g_span_tma = value;
        return ASP_OK;
    }
    if (strcmp(key, "ln_on_read") == 0) {  // developer switch: 1 inner LayerNorms leave statistics, the next residual epilogue normalises
        ASP_REQUIRE(value == 0 || value == 1, "asp_set_option: ln_on_read must be 0 or 1");
        asp::g_ln_on_read = value;
        return ASP_OK;
    }
    if (strcmp(key, "attn_tc") == 0) {  // developer switch: tcgen05 attention (plain bf16, L <= 256) 3 / 1 / 2 (see attention_tc.cu), 0 mma.sync attention
        ASP_REQUIRE(value >= 0 && value <= 5, "asp_set_option: attn_tc must be 0..5");
        asp::g_attn_tc = value;
        return ASP_OK;
    }
    if (strcmp(key, "oa_warps") == 0) {  // developer switch: Sinkhorn warps per CTA of ot_allpairs.cu
        ASP_REQUIRE(value == 8 || value == 12, "asp_set_option: oa_warps must be 8 or 12");
        asp::g_oa_warps = value;
        return ASP_OK;
    }
    if (strcmp(key, "gemm_kernel") == 0) {  // developer switch: see bert/gemm.cu
        ASP_REQUIRE(value >= 0 && value <= 4, "asp_set_option: gemm_kernel must be 0..4");
        asp::g_gemm_kernel = value;
        return ASP_OK;
    }
    if (strcmp(key, "pdl") == 0) {  // developer switch: programmatic dependent launch between the encoder's kernels
        asp::g_pdl = value != 0;
        return ASP_OK;
    }
    if (strcmp(key, "gemm_pair") == 0) {
        ASP_REQUIRE(value >= -1 && value <= 2, "asp_set_option: gemm_pair must be -1 (auto), 0, 1 or 2");
        asp::g_gemm_pair = value;
        return ASP_OK;
    }
    if (strcmp(key, "gemm_cluster") == 0) {
        ASP_REQUIRE(value == 1 || value == 2 || value == 4, "asp_set_option: gemm_cluster must be 1, 2 or 4");
        asp::g_gemm_cluster = value;
        return ASP_OK;
    }
    asp::set_error("asp_set_option: unknown key '%s'", key);
    return ASP_ERR_INVALID;
}

extern "C" int asp_ot_sinkhorn(const float* q, const int32_t* q_lens, int q_broadcast, const float* c,
                               const int32_t* c_lens, int B, int Sq, int Sc, int D, const float* eps_host, int n_eps,
                               float temp, float* cost_workspace, const asp_ot_outputs* out, asp_stream_t stream) {
    if (B == 0) return ASP_OK;
    ASP_REQUIRE(cost_workspace, "asp_ot_sinkhorn: cost_workspace is NULL (needs B*Sq*Sc floats)");
    int rc = asp_pair_cost(q, q_lens, q_broadcast, c, c_lens, B, Sq, Sc, D, cost_workspace, stream);
    if (rc) return rc;
    return asp_ot_sinkhorn_from_cost(cost_workspace, q_lens, q_broadcast, c_lens, B, Sq, Sc, eps_host, n_eps, temp,
                                     out, stream);
}
