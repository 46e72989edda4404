// Developer probe (not product code): tcgen05.mma with the A operand in TENSOR MEMORY (kind::f16, M = 128, N = 64, K = 16).
// Which TMEM lane / column / half-word holds A[m][k]?  A one-hot A (A[m][k] = [k == k0], written by tcgen05.st under a
// candidate layout) times a known B gives D[m][n] = B[n][k0] iff the layout is right; the probe prints, for every k0 and
// column offset, whether D matches and otherwise which k it behaved like.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I aspire_b200/csrc -o tools/ubench4 tools/ubench4.cu -lcuda && tools/ubench4
#include <cstdio>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "bert/tc05.cuh"

using namespace asp::tc;

// B[n][k] for n < 64, k < 64 (only k < 16 used): small integers, exactly representable
__host__ __device__ inline float bval(int n, int k) { return (float)(((3 * n + 5 * k) % 13) - 6); }

__global__ void __launch_bounds__(128) probe(int k0, int a_col, int d_col, float* out, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // B tile, K-major SW128: row n = 128 bytes (64 bf16), 16-byte chunk c at position c ^ (n & 7)
    for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
        const int n = e >> 6, k = e & 63;
        __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(smem + n * 128);
        row[(((k >> 3) ^ (n & 7)) << 3) + (k & 7)] = __float2bfloat16(k < 16 ? bval(n, k) : 0.f);
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&slot, 256);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tb = slot;
    const int m = threadIdx.x;  // row = TMEM lane
    const uint32_t trow = tb + ((uint32_t)(warp * 32) << 16);
    // A[m][k] = (k == k0) * (1 + (m & 3)): packed two per column, element 2c in the low half (mode 0) or high half (mode 1)
    uint32_t pk[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) pk[c] = 0u;
    {
        const __nv_bfloat16 v = __float2bfloat16(1.f + (m & 3));
        const uint16_t bits = *reinterpret_cast<const uint16_t*>(&v);
        const int c = k0 >> 1, hi = (k0 & 1) ^ mode;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
            if (cc == c) pk[cc] = hi ? ((uint32_t)bits << 16) : (uint32_t)bits;
    }
    // pre-fill D with a marker through tcgen05.st as well
    uint32_t mark[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) mark[c] = __float_as_uint(-777.f);
#pragma unroll
    for (int c = 0; c < 64; c += 16) tmem_st16(trow + d_col + c, mark);
    tmem_st16(trow + a_col, pk);
    tmem_wait_st();
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after_sync();
        constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
        umma_bf16_ts(tb + d_col, tb + a_col, umma_desc_sw128(smem_u32(smem)), idesc, false);
        umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    float v[32];
    for (int h = 0; h < 2; ++h) {
        tmem_ld32(trow + d_col + 32 * h, v);
#pragma unroll
        for (int e = 0; e < 32; ++e) out[m * 64 + 32 * h + e] = v[e];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 256);
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * 64 * 4);
    static float h[128 * 64];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    const int cols[][2] = {{0, 64}, {128, 64}, {32, 64}, {8, 192}};
    for (int mode = 0; mode < 2; ++mode)
        for (auto& cc : cols)
            for (int k0 = 0; k0 < 16; k0 += (cc[0] == 0 ? 1 : 5)) {
                probe<<<1, 128, 16384>>>(k0, cc[0], cc[1], d, mode);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("mode %d a_col %d k0 %d: %s\n", mode, cc[0], k0, cudaGetErrorString(e));
                    return 1;
                }
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                // which k does the result look like?  D[m][n] == (1 + (m & 3)) * B[n][k]
                int match = -1, nm = 0;
                for (int k = 0; k < 16; ++k) {
                    bool ok = true;
                    for (int m = 0; m < 128 && ok; ++m)
                        for (int n = 0; n < 64; ++n)
                            if (h[m * 64 + n] != (1.f + (m & 3)) * bval(n, k)) { ok = false; break; }
                    if (ok) { match = k; ++nm; }
                }
                printf("mode %d (element 2c in the %s half)  A at column %3d, D at %3d, one-hot k0 = %2d  ->  behaves like k = %2d%s   D[0][0..3] = %g %g %g %g, D[5][0] = %g\n",
                       mode, mode ? "HIGH" : "low", cc[0], cc[1], k0, match, match == k0 ? "  OK" : "  MISMATCH", h[0], h[1], h[2], h[3], h[5 * 64]);
            }
    return 0;
}
