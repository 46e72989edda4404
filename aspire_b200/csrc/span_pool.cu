// K1: per-sentence token-span mean pooling + CLS read-out, and geomloss' bounding-box diameter.
//
// span pool replaces the Python loop at examples/ex_aspire_consent.py:75-100: for every sentence slot it
// builds a dense float64 [B,L,768] mask on the host and re-reads the whole hidden state (Smax passes over
// [B,L,D]).  Sentence spans are contiguous token ranges, so here each hidden row is
// read at most once: one CTA per (document, sentence), threads own 128-bit column slices, token rows are
// streamed with 4 independent loads in flight per thread.
#include "common.cuh"

namespace asp {

__global__ void __launch_bounds__(192)
span_mean_pool_kernel(const float* __restrict__ hidden, const int32_t* __restrict__ spans, int L, int D, int Smax,
                      float* __restrict__ sent_reps, float* __restrict__ cls_reps) {
    const int b = blockIdx.y, s = blockIdx.x;
    const int d4 = D >> 2;
    const float4* hb = reinterpret_cast<const float4*>(hidden + (size_t)b * L * D);
    if (s == Smax) {  // CLS row
        if (cls_reps) {
            float4* o = reinterpret_cast<float4*>(cls_reps + (size_t)b * D);
            for (int k = threadIdx.x; k < d4; k += blockDim.x) o[k] = hb[k];
        }
        return;
    }
    int start = spans[((size_t)b * Smax + s) * 2], end = spans[((size_t)b * Smax + s) * 2 + 1];
    start = max(start, 0);
    end = min(end, L);
    const int n = end - start;
    float4* o = reinterpret_cast<float4*>(sent_reps + ((size_t)b * Smax + s) * D);
    for (int k = threadIdx.x; k < d4; k += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n > 0) {
            const float4* p = hb + (size_t)start * d4 + k;
            int t = 0;
            for (; t + 4 <= n; t += 4) {
                const float4 a0 = ldg_stream(p + (size_t)(t + 0) * d4), a1 = ldg_stream(p + (size_t)(t + 1) * d4);
                const float4 a2 = ldg_stream(p + (size_t)(t + 2) * d4), a3 = ldg_stream(p + (size_t)(t + 3) * d4);
                acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
                acc.x += a1.x; acc.y += a1.y; acc.z += a1.z; acc.w += a1.w;
                acc.x += a2.x; acc.y += a2.y; acc.z += a2.z; acc.w += a2.w;
                acc.x += a3.x; acc.y += a3.y; acc.z += a3.z; acc.w += a3.w;
            }
            for (; t < n; ++t) {
                const float4 a = ldg_stream(p + (size_t)t * d4);
                acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
            }
            const float cnt = (float)n;  // reference: sum / clamp(count, 1)
            acc.x /= cnt; acc.y /= cnt; acc.z /= cnt; acc.w /= cnt;
        }
        o[k] = acc;
    }
}

// TMA variant.  A sentence span is a run of whole token rows, i.e. ONE contiguous byte range of the hidden state
// ((end - start) * D * 4 bytes), so it is staged by the bulk-copy engine (cp.async.bulk, no tensor map needed): one thread
// issues kSpRows rows per stage into a kSpStages-deep shared-memory ring and arms the stage's mbarrier with the byte count;
// all threads then add up their 128-bit column slices from shared memory (conflict-free).  48 KB per CTA, four CTAs per SM.
constexpr int kSpRows = 4, kSpStages = 4, kSpThreads = 192, kSpMaxSlices = 4;  // D <= 4 * 192 * 4 = 3072

__device__ __forceinline__ void sp_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

__global__ void __launch_bounds__(kSpThreads)
span_mean_pool_tma_kernel(const float* __restrict__ hidden, const int32_t* __restrict__ spans, int L, int D, int Smax,
                          float* __restrict__ sent_reps, float* __restrict__ cls_reps) {
    extern __shared__ __align__(128) float sp_ring[];  // [kSpStages][kSpRows * D]
    __shared__ uint64_t full[kSpStages];
    const int b = blockIdx.y, s = blockIdx.x, tid = threadIdx.x;
    const int d4 = D >> 2;
    const float* hb = hidden + (size_t)b * L * D;
    if (s == Smax) {  // CLS row
        if (cls_reps) {
            float4* o = reinterpret_cast<float4*>(cls_reps + (size_t)b * D);
            for (int k = tid; k < d4; k += kSpThreads) o[k] = reinterpret_cast<const float4*>(hb)[k];
        }
        return;
    }
    int start = spans[((size_t)b * Smax + s) * 2], end = spans[((size_t)b * Smax + s) * 2 + 1];
    start = max(start, 0);
    end = min(end, L);
    const int n = max(end - start, 0);
    const int nchunks = (n + kSpRows - 1) / kSpRows;
    const uint32_t ring_u = (uint32_t)__cvta_generic_to_shared(sp_ring), bar_u = (uint32_t)__cvta_generic_to_shared(full);
    if (tid == 0) {
        for (int i = 0; i < kSpStages; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u + 8u * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int c) {  // thread 0: rows [c * kSpRows, ...) of the span into stage c % kSpStages
        const int rows = min(kSpRows, n - c * kSpRows);
        const uint32_t bytes = (uint32_t)rows * (uint32_t)D * 4u, st = (uint32_t)(c % kSpStages);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_u + 8u * st), "r"(bytes) : "memory");
        sp_bulk_load(ring_u + st * (uint32_t)(kSpRows * D * 4), hb + (size_t)(start + c * kSpRows) * D, bytes, bar_u + 8u * st);
    };
    if (tid == 0)
        for (int c = 0; c < min(nchunks, kSpStages); ++c) issue(c);
    float4 acc[kSpMaxSlices];
#pragma unroll
    for (int m = 0; m < kSpMaxSlices; ++m) acc[m] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < nchunks; ++c) {
        const uint32_t st = (uint32_t)(c % kSpStages), parity = (uint32_t)((c / kSpStages) & 1);
        uint32_t ok = 0;
        do {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(bar_u + 8u * st), "r"(parity)
                : "memory");
        } while (!ok);
        const int rows = min(kSpRows, n - c * kSpRows);
        const float4* stage = reinterpret_cast<const float4*>(sp_ring + (size_t)st * kSpRows * D);
#pragma unroll
        for (int m = 0; m < kSpMaxSlices; ++m) {
            const int k = tid + m * kSpThreads;
            if (k < d4)
                for (int r = 0; r < rows; ++r) {  // rows in span order: the same summation order as the streaming kernel
                    const float4 a = stage[(size_t)r * d4 + k];
                    acc[m].x += a.x; acc[m].y += a.y; acc[m].z += a.z; acc[m].w += a.w;
                }
        }
        if (c + kSpStages < nchunks) {  // block-uniform: the stage is refilled once every thread has read it
            __syncthreads();
            if (tid == 0) issue(c + kSpStages);
        }
    }
    float4* o = reinterpret_cast<float4*>(sent_reps + ((size_t)b * Smax + s) * D);
    const float cnt = (float)max(n, 1);  // reference: sum / clamp(count, 1)
#pragma unroll
    for (int m = 0; m < kSpMaxSlices; ++m) {
        const int k = tid + m * kSpThreads;
        if (k < d4) {
            float4 v = acc[m];
            if (n > 0) { v.x /= cnt; v.y /= cnt; v.z /= cnt; v.w /= cnt; }
            o[k] = v;
        }
    }
}

int g_span_tma = 1;  // asp_set_option("span_tma"): 1 = bulk-copy staging (D <= 3072), 0 = streaming 128-bit loads

// ---- bounding-box diameter (geomloss max_diameter) ---------------------------------------------------
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
    int old = __float_as_int(*a);
    while (v < __int_as_float(old)) {
        const int assumed = old;
        old = atomicCAS(reinterpret_cast<int*>(a), assumed, __float_as_int(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
    int old = __float_as_int(*a);
    while (v > __int_as_float(old)) {
        const int assumed = old;
        old = atomicCAS(reinterpret_cast<int*>(a), assumed, __float_as_int(v));
        if (old == assumed) break;
    }
}

__global__ void bbox_init_kernel(float* ws, int D) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < D; k += gridDim.x * blockDim.x) {
        ws[k] = INFINITY;
        ws[D + k] = -INFINITY;
    }
}

// rows [0,nx) come from x, rows [nx, nx+ny) from y.  Block = 256 threads; thread t owns float4 column
// slice (t % d4) and walks rows (t / d4) + m * rows_per_iter within the block's row range.
__global__ void __launch_bounds__(256)
bbox_minmax_kernel(const float* __restrict__ x, long long nx, const float* __restrict__ y, long long ny, int D,
                   float* __restrict__ ws, long long rows_per_block) {
    const int d4 = D >> 2;
    const long long total = nx + ny;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(r0 + rows_per_block, total);
    const int rstep = max((int)blockDim.x / d4, 1);  // rows walked concurrently by the block
    for (int k = threadIdx.x; k < d4 * rstep; k += blockDim.x) {
        const int col = k % d4, rsub = k / d4;
        float4 lo = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
        float4 hi = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (long long r = r0 + rsub; r < r1; r += rstep) {
            const float* row = (r < nx) ? x + (size_t)r * D : y + (size_t)(r - nx) * D;
            const float4 v = __ldg(reinterpret_cast<const float4*>(row) + col);
            lo.x = fminf(lo.x, v.x); lo.y = fminf(lo.y, v.y); lo.z = fminf(lo.z, v.z); lo.w = fminf(lo.w, v.w);
            hi.x = fmaxf(hi.x, v.x); hi.y = fmaxf(hi.y, v.y); hi.z = fmaxf(hi.z, v.z); hi.w = fmaxf(hi.w, v.w);
        }
        float* wl = ws + 4 * col;
        float* wh = ws + D + 4 * col;
        atomic_min_f(wl + 0, lo.x); atomic_min_f(wl + 1, lo.y); atomic_min_f(wl + 2, lo.z); atomic_min_f(wl + 3, lo.w);
        atomic_max_f(wh + 0, hi.x); atomic_max_f(wh + 1, hi.y); atomic_max_f(wh + 2, hi.z); atomic_max_f(wh + 3, hi.w);
    }
}

__global__ void __launch_bounds__(256) bbox_norm_kernel(const float* __restrict__ ws, int D, float* out) {
    __shared__ float part[8];
    float s = 0.f;
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
        const float d = ws[D + k] - ws[k];
        s = fmaf(d, d, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
        *out = sqrtf(t);
    }
}

}  // namespace asp

extern "C" int asp_span_mean_pool(const float* hidden, const int32_t* spans, int B, int L, int D, int Smax,
                                  float* sent_reps, float* cls_reps, asp_stream_t stream) {
    ASP_REQUIRE(hidden && spans && sent_reps, "asp_span_mean_pool: NULL pointer");
    ASP_REQUIRE(B >= 0 && L >= 1 && Smax >= 1, "asp_span_mean_pool: bad shape B=%d L=%d Smax=%d", B, L, Smax);
    ASP_REQUIRE(D >= 4 && D % 4 == 0, "asp_span_mean_pool: D=%d must be a positive multiple of 4", D);
    ASP_REQUIRE(asp::aligned16(hidden) && asp::aligned16(sent_reps) && (!cls_reps || asp::aligned16(cls_reps)),
                "asp_span_mean_pool: buffers must be 16-byte aligned");
    if (B == 0) return ASP_OK;
    ASP_REQUIRE(B <= 65535, "asp_span_mean_pool: B=%d exceeds 65535 documents per call", B);
    dim3 grid(Smax + 1, B);
    if (asp::g_span_tma && D <= asp::kSpMaxSlices * asp::kSpThreads * 4) {
        const int smem = asp::kSpStages * asp::kSpRows * D * 4;
        static thread_local int attr_dev = -1, attr_smem = 0;
        int dev = 0;
        ASP_CUDA(cudaGetDevice(&dev));
        if (attr_dev != dev || attr_smem < smem) {
            ASP_CUDA(cudaFuncSetAttribute(asp::span_mean_pool_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            ASP_CUDA(cudaFuncSetAttribute(asp::span_mean_pool_tma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            attr_dev = dev;
            attr_smem = smem;
        }
        asp::span_mean_pool_tma_kernel<<<grid, asp::kSpThreads, smem, (cudaStream_t)stream>>>(hidden, spans, L, D, Smax,
                                                                                           sent_reps, cls_reps);
        ASP_LAUNCH_CHECK("span_mean_pool_tma_kernel");
        return ASP_OK;
    }
    asp::span_mean_pool_kernel<<<grid, 192, 0, (cudaStream_t)stream>>>(hidden, spans, L, D, Smax, sent_reps,
                                                                       cls_reps);
    ASP_LAUNCH_CHECK("span_mean_pool_kernel");
    return ASP_OK;
}

extern "C" int asp_bbox_diameter(const float* x, long long nx, const float* y, long long ny, int D,
                                 float* workspace, float* diameter_out, asp_stream_t stream) {
    ASP_REQUIRE(workspace && diameter_out, "asp_bbox_diameter: NULL workspace/output");
    ASP_REQUIRE(nx >= 0 && ny >= 0 && nx + ny >= 1, "asp_bbox_diameter: no points");
    ASP_REQUIRE((nx == 0 || x) && (ny == 0 || y), "asp_bbox_diameter: NULL input");
    ASP_REQUIRE(D >= 4 && D % 4 == 0 && D <= 4096, "asp_bbox_diameter: D=%d must be a multiple of 4 in [4,4096]", D);
    ASP_REQUIRE((nx == 0 || asp::aligned16(x)) && (ny == 0 || asp::aligned16(y)), "asp_bbox_diameter: unaligned input");
    cudaStream_t st = (cudaStream_t)stream;
    asp::bbox_init_kernel<<<(D + 255) / 256, 256, 0, st>>>(workspace, D);
    ASP_LAUNCH_CHECK("bbox_init_kernel");
    const long long total = nx + ny;
    long long blocks = 4LL * asp::sm_count();
    if (blocks > total) blocks = total;
    const long long rows_per_block = (total + blocks - 1) / blocks;
    blocks = (total + rows_per_block - 1) / rows_per_block;
    asp::bbox_minmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, nx, y, ny, D, workspace, rows_per_block);
    ASP_LAUNCH_CHECK("bbox_minmax_kernel");
    asp::bbox_norm_kernel<<<1, 256, 0, st>>>(workspace, D, diameter_out);
    ASP_LAUNCH_CHECK("bbox_norm_kernel");
    return ASP_OK;
}
