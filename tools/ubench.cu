// Developer micro-benchmarks (not product code): per-SM issue rates of the instructions the OT kernels lean on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

__global__ void k_ffma(float* out, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
    float2 x[ILP];
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = __ffma2_rn(x[i], a2, b2);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ex2(float* out, float a) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = -1.0f - 0.01f * (threadIdx.x + i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            float y;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
            x[i] = y - a;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ex2 + 4 FFMA per ex2 (the Sinkhorn step's mix): do they overlap?
__global__ void k_mix(float* out, float a, float b) {
    float x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = -1.0f - 0.01f * (threadIdx.x + i); y[i] = i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            float e;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x[i]));
            y[i] = fmaf(y[i], a, e);
            y[i] = fmaf(y[i], a, b);
            y[i] = fmaf(y[i], a, b);
            x[i] = fmaf(e, a, -1.5f);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float run(F f, int blocks, int threads) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(blocks, threads);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f(blocks, threads);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, clock attr %d kHz\n", p.name, sms, khz);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
    for (int warps_per_sm : {4, 8, 16, 32}) {
        int threads = 128, blocks = sms * warps_per_sm / 4;
        double n = (double)blocks * threads * ITERS * ILP;
        float t1 = run([&](int b, int t) { k_ffma<<<b, t>>>(out, 1.0001f, 0.5f); }, blocks, threads);
        float t2 = run([&](int b, int t) { k_ffma2<<<b, t>>>(out, 1.0001f, 0.5f); }, blocks, threads);
        float t3 = run([&](int b, int t) { k_ex2<<<b, t>>>(out, 1.5f); }, blocks, threads);
        float t4 = run([&](int b, int t) { k_mix<<<b, t>>>(out, 0.5f, 0.25f); }, blocks, threads);
        printf("warps/SM %2d: FFMA %.1f Gop/s/SM (%.3f ms) | FFMA2 %.1f Gfma/s/SM (%.3f ms) | EX2 %.1f Gop/s/SM (%.3f ms) | "
               "mix(1 ex2+4 ffma) %.1f Gex2/s/SM (%.3f ms)\n",
               warps_per_sm, n / t1 / 1e6 / sms, t1, 2 * n / t2 / 1e6 / sms, t2, n / t3 / 1e6 / sms, t3, n / t4 / 1e6 / sms, t4);
    }
    printf("(divide Gop/s/SM by the SM clock in GHz to get ops/clk/SM)\n");
    return 0;
}
