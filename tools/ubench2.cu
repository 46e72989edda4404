// Developer micro-benchmark (not product code): do the FMA pipe (FFMA2 / FFMA) and the MUFU pipe (ex2) overlap when
// DIFFERENT warps of one scheduler use them (the fused OT kernel's phase 1 / phase 2 situation), and does FFMA2 hold
// the issue port for its second cycle?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench2 tools/ubench2.cu && tools/ubench2
#include <cstdio>
#include <cuda_runtime.h>
#define ILP 8

__device__ __forceinline__ float ffma2_chain(int iters, float a, float b, int seed) {
    float2 x[ILP];
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = make_float2(seed + i, seed - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = __ffma2_rn(x[i], a2, b2);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
    return s;
}
__device__ __forceinline__ float ffma_chain(int iters, float a, float b, int seed) {
    float x[2 * ILP];
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) x[i] = seed + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 2 * ILP; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) s += x[i];
    return s;
}
__device__ __forceinline__ float ex2_chain(int iters, float a, int seed) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = -1.0f - 0.01f * (seed + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            float y;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
            x[i] = y;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    return s;
}
// FFMA2 with an independent integer (ALU pipe) instruction after each one
__device__ __forceinline__ float ffma2_alu_chain(int iters, float a, float b, int seed) {
    float2 x[ILP];
    unsigned u[ILP];
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = make_float2(seed + i, seed - i); u[i] = seed * 7 + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            x[i] = __ffma2_rn(x[i], a2, b2);
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % ILP]), "r"(seed));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y + (float)u[i];
    return s;
}

// FFMA2 with THREE register-pair operands and 40 independent accumulators (the Gram tile's shape: acc[i][j] += q[i]*c[j])
__global__ void __launch_bounds__(256, 1) k_ffma2_3reg(float* out, const float2* in, int iters, int warps_active) {
    if ((int)(threadIdx.x >> 5) >= warps_active) return;
    float2 a[5], b[8], acc[5][8];
#pragma unroll
    for (int i = 0; i < 5; ++i) a[i] = in[threadIdx.x + 32 * i];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = in[threadIdx.x + 32 * (5 + j)];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 5; ++i) acc[i][j] = __ffma2_rn(a[i], b[j], acc[i][j]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// legacy tensor path: mma.sync.m16n8k8 tf32 (register operands), 8 independent accumulator tiles per warp
__global__ void __launch_bounds__(512, 1) k_mma_tf32(float* out, int iters, int warps_active) {
    if ((int)(threadIdx.x >> 5) >= warps_active) return;
    float acc[8][4];
    unsigned a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = 0x3f800000u + threadIdx.x * 8192u * i;
    b[0] = 0x3f000000u + threadIdx.x * 8192u; b[1] = 0x3e800000u;
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[t][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 8; ++t)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[t][0]), "+f"(acc[t][1]), "+f"(acc[t][2]), "+f"(acc[t][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) s += acc[t][0] + acc[t][1] + acc[t][2] + acc[t][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// how long does nanosleep(t) really suspend a warp?  (idle warps of the fused kernel poll an mbarrier)
__global__ void k_nanosleep(unsigned long long* out, unsigned t, int reps) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = 0; i < reps; ++i) __nanosleep(t);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

// mode bits: 1 = warps 0-3 run the FMA chain, 2 = warps 4-7 run the ex2 chain; kind 0 = FFMA2, 1 = FFMA, 2 = FFMA2+LOP3
__global__ void __launch_bounds__(256, 1) k_spec(float* out, int mode, int kind, int it_fma, int it_ex2, float a, float b) {
    const int warp = threadIdx.x >> 5;
    float s = 0.f;
    if (warp < 4) {
        if (mode & 1) {
            if (kind == 0) s = ffma2_chain(it_fma, a, b, threadIdx.x);
            else if (kind == 1) s = ffma_chain(it_fma, a, b, threadIdx.x);
            else s = ffma2_alu_chain(it_fma, a, b, threadIdx.x);
        }
    } else if (mode & 2) {
        s = ex2_chain(it_ex2, a, threadIdx.x);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

float timeit(int sms, float* out, int mode, int kind, int it_fma, int it_ex2) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_spec<<<sms, 256>>>(out, mode, kind, it_fma, it_ex2, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_spec<<<sms, 256>>>(out, mode, kind, it_fma, it_ex2, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, sizeof(float) * sms * 256);
    const int it_fma = 1 << 16, it_ex2 = 1 << 14;  // 8 FFMA2 (2 clk each) vs 8 ex2 (8 clk each) per iteration: equal pipe time
    const char* names[3] = {"FFMA2", "FFMA x2", "FFMA2+LOP3"};
    for (int kind = 0; kind < 3; ++kind) {
        float tf = timeit(sms, out, 1, kind, it_fma, it_ex2);
        float te = timeit(sms, out, 2, kind, it_fma, it_ex2);
        float tb = timeit(sms, out, 3, kind, it_fma, it_ex2);
        printf("%-11s warps 0-3 alone %.3f ms | ex2 warps 4-7 alone %.3f ms | both %.3f ms  (perfect overlap %.3f, serial %.3f)\n",
               names[kind], tf, te, tb, tf > te ? tf : te, tf + te);
    }
    {
        float2* in; cudaMalloc(&in, sizeof(float2) * 32 * 13 * 8); cudaMemset(in, 0, sizeof(float2) * 32 * 13 * 8);
        int khz2 = 0; cudaDeviceGetAttribute(&khz2, cudaDevAttrClockRate, 0);
        const int iters = 1 << 14;
        for (int wa : {4, 8}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_ffma2_3reg<<<sms, 256>>>(out, in, iters, wa); cudaDeviceSynchronize();
            cudaEventRecord(e0); k_ffma2_3reg<<<sms, 256>>>(out, in, iters, wa); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double instr_per_smsp = (double)iters * 40 * (wa / 4);
            printf("FFMA2 3-register-operand form, %d warp(s) per scheduler: %.3f ms -> %.2f clk per FFMA2 per scheduler at %d kHz\n",
                   wa / 4, ms, ms * 1e-3 * khz2 * 1e3 / instr_per_smsp, khz2);
        }
    }
    {
        int khz3 = 0; cudaDeviceGetAttribute(&khz3, cudaDevAttrClockRate, 0);
        float* out2; cudaMalloc(&out2, sizeof(float) * sms * 512);
        const int iters = 1 << 14;
        for (int wa : {4, 8, 16}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_mma_tf32<<<sms, 512>>>(out2, iters, wa); cudaDeviceSynchronize();
            cudaEventRecord(e0); k_mma_tf32<<<sms, 512>>>(out2, iters, wa); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double mma_per_smsp = (double)iters * 8 * (wa / 4);
            double clk = ms * 1e-3 * khz3 * 1e3 / mma_per_smsp;
            printf("mma.sync m16n8k8 tf32, %d warp(s) per scheduler: %.3f ms -> %.2f clk per MMA per scheduler = %.0f FMA/clk/SM\n",
                   wa / 4, ms, clk, 4.0 * 1024.0 / clk);
        }
    }
    {
        unsigned long long* o; cudaMalloc(&o, sizeof(unsigned long long) * sms);
        for (unsigned t : {100u, 1000u, 10000u, 100000u}) {
            k_nanosleep<<<sms, 32>>>(o, t, 200); cudaDeviceSynchronize();
            unsigned long long h[4]; cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
            printf("nanosleep(%u) x200: %.0f ns per call (globaltimer)\n", t, (double)h[0] / 200.0);
        }
    }
    // clocks per instruction for the single-warp-per-scheduler chains
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("(clock attr %d kHz: FFMA2 chain = %d instr per warp)\n", khz, it_fma * ILP);
    return 0;
}
