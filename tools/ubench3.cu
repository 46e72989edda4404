// Developer micro-benchmark (not product code): what a warp-wide 128-bit shared-memory load costs as a function of WHICH
// lanes share an address.  The var-len Gram (ot_varlen.cu) is built from LDS.128 broadcasts and is paced by the
// shared-memory pipe; ncu counts ~4 wavefronts per "8 distinct rows, same in every quarter warp" load and ~2 per
// "one row per quarter warp" load.  This prints clk per LDS.128 per SM for a set of lane -> row maps (8 warps per SM
// streaming independent loads; one wavefront = one clk of the pipe).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench3 tools/ubench3.cu && tools/ubench3
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds64(unsigned addr) {
    float2 v;
    asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

// row(lane) for each pattern; the address is row * 128 + ((chunk ^ row) & 7) * 16 (the kernel's swizzle: conflict-free)
__device__ __forceinline__ int row_of(int pat, int lane) {
    switch (pat) {
        case 0: return 0;                 // one address for the whole warp
        case 1: return lane & 7;          // 8 rows, every quarter warp sees all 8            (query side of the Gram)
        case 2: return lane >> 3;         // 4 rows, one per quarter warp                     (candidate side)
        case 3: return lane >> 2;         // 8 rows, two per quarter warp
        case 4: return lane >> 1;         // 16 rows, four per quarter warp
        case 5: return lane;              // 32 rows
        case 6: return lane & 3;          // 4 rows, every quarter warp sees all 4
        case 7: return lane & 1;          // 2 rows
        case 8: return lane & 15;         // 16 rows, halves repeat
        case 9: return (lane >> 4);       // 2 rows, one per half warp
        case 10: return (lane & 3) + 4 * (lane >> 4);  // 8 rows: 4 per half warp, every quarter of a half sees all 4
        default: return (lane & 1) + 2 * (lane >> 3);  // 8 rows: 2 per quarter warp, lanes alternate
    }
}

template <int WIDTH>
__global__ void __launch_bounds__(256) k(int pat, int iters, long long* clk, float* sink) {
    extern __shared__ __align__(1024) float sm[];
    for (int i = threadIdx.x; i < 32 * 32 * 4; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = row_of(pat, lane);
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + warp * 4096 * 0 + row * 128;
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const unsigned a = base + (((c ^ row) & 7) << 4);
            if (WIDTH == 16) {
                const float4 v = lds128(a);
                acc += v.x + v.y + v.z + v.w;
            } else {
                const float2 v = lds64(a);
                acc += v.x + v.y;
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    long long* clk;
    float* sink;
    cudaMalloc(&clk, 148 * 8);
    cudaMalloc(&sink, 148 * 256 * 4);
    const int iters = 2000;
    const char* names[] = {"1 row (warp broadcast)", "8 rows, all 8 in every quarter warp (q side)", "4 rows, one per quarter warp (c side)",
                           "8 rows, two per quarter warp", "16 rows, four per quarter warp", "32 rows", "4 rows, all 4 in every quarter",
                           "2 rows, alternating lanes", "16 rows (lane & 15)", "2 rows, one per half warp",
                           "8 rows: 4 per half warp, all 4 in each quarter", "8 rows: 2 per quarter, lanes alternate"};
    for (int width = 16; width >= 8; width -= 8)
        for (int pat = 0; pat < 12; ++pat) {
            for (int rep = 0; rep < 2; ++rep) {
                if (width == 16)
                    k<16><<<148, 256, 16384>>>(pat, iters, clk, sink);
                else
                    k<8><<<148, 256, 16384>>>(pat, iters, clk, sink);
                cudaDeviceSynchronize();
            }
            long long h[148];
            cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < 148; ++i) avg += h[i];
            avg /= 148;
            printf("LDS.%d  %-52s %6.2f clk per warp-wide load per SM\n", width * 8, names[pat], avg / (8.0 * iters * 8));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
