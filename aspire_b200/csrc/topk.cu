// K5: per-query top-k of a score matrix and the k-way merge of per-shard lists.
//
// Replaces the head of Python's `sorted(..., reverse=True)` at src/evaluation/evaluate.py:76 and
// src/pre_process/pp_gen_nearest.py:339.  Order is (score descending, id ascending) -- what a stable
// descending sort over a pool listed in id order gives -- so results do not depend on the shard count.
//
// asp_topk: one CTA per query.  4-pass 8-bit radix select finds the exact k-th largest key, an in-order
// compaction gathers the k winners (ties on the k-th key resolved by smallest id), a bitonic sort in shared
// memory orders them.  The row is read 5 times; at 4 B/pair against ~30 KB/pair of scoring traffic this is
// noise, and rows of a 1kx1M problem (4 MB) sit in the 126 MB L2 between passes.
#include "common.cuh"

namespace asp {

__device__ __forceinline__ uint32_t score_key(float x) {
    if (x != x) return 0u;  // NaN ranks last
    if (x == 0.f) x = 0.f;  // -0 == +0
    const uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

constexpr int kTopkThreads = 1024;
constexpr int kTopkMaxK = 2048;

// bitonic sort, descending, of n (power of two) 64-bit keys in shared memory
__device__ void bitonic_desc_u64(unsigned long long* a, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long x = a[lo], y = a[hi];
                if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kTopkThreads)
topk_kernel(const float* __restrict__ scores, long long N, int k, long long base_id, float* __restrict__ out_scores,
            long long* __restrict__ out_ids, int kpad) {
    extern __shared__ unsigned long long sel[];  // kpad composite keys
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_need, s_gt_base, s_eq_base;
    __shared__ unsigned int warp_cnt[2][32];
    const float* row = scores + (size_t)blockIdx.x * N;
    const int keff = (int)min((long long)k, N);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- radix select: exact key of the keff-th largest element ------------------------------------
    if (tid == 0) { s_prefix = 0u; s_need = (unsigned)keff; }
    unsigned int mask = 0u;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
        const unsigned int prefix = s_prefix;
        for (long long i = tid; i < N; i += blockDim.x) {
            const uint32_t key = score_key(row[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int need = s_need, cum = 0u;
            int bin = 255;
            for (; bin > 0; --bin) {
                if (cum + hist[bin] >= need) break;
                cum += hist[bin];
            }
            s_need = need - cum;
            s_prefix = prefix | ((unsigned)bin << shift);
        }
        mask |= 255u << shift;
        __syncthreads();
    }
    const uint32_t kth = s_prefix;
    const unsigned int need_eq = s_need;            // how many elements equal to kth are taken
    const unsigned int n_gt = (unsigned)keff - need_eq;  // all strictly greater elements are taken
    for (int i = tid; i < kpad; i += blockDim.x) sel[i] = 0ull;
    if (tid == 0) { s_gt_base = 0u; s_eq_base = 0u; }
    __syncthreads();

    // ---- in-order compaction ---------------------------------------------------------------------
    for (long long c0 = 0; c0 < N; c0 += blockDim.x) {
        const long long i = c0 + tid;
        uint32_t key = 0u;
        bool gt = false, eq = false;
        if (i < N) {
            key = score_key(row[i]);
            gt = key > kth;
            eq = key == kth;
        }
        const unsigned bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
        if (!__syncthreads_or(bg | be)) continue;
        if (lane == 0) { warp_cnt[0][warp] = __popc(bg); warp_cnt[1][warp] = __popc(be); }
        __syncthreads();
        unsigned int og = 0, oe = 0, tg = 0, te = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            if (w < warp) { og += warp_cnt[0][w]; oe += warp_cnt[1][w]; }
            tg += warp_cnt[0][w];
            te += warp_cnt[1][w];
        }
        const unsigned lm = (1u << lane) - 1u;
        const unsigned long long comp = ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
        if (gt) sel[s_gt_base + og + __popc(bg & lm)] = comp;
        if (eq) {
            const unsigned int pos = s_eq_base + oe + __popc(be & lm);
            if (pos < need_eq) sel[n_gt + pos] = comp;
        }
        __syncthreads();
        if (tid == 0) { s_gt_base += tg; s_eq_base += te; }
        __syncthreads();
    }
    bitonic_desc_u64(sel, kpad);
    for (int i = tid; i < k; i += blockDim.x) {
        float sc = -INFINITY;
        long long id = -1;
        if (i < keff) {
            const uint32_t idx = 0xffffffffu - (uint32_t)(sel[i] & 0xffffffffull);
            sc = row[idx];
            id = base_id + (long long)idx;
        }
        out_scores[(size_t)blockIdx.x * k + i] = sc;
        out_ids[(size_t)blockIdx.x * k + i] = id;
    }
}

// ---- merge of R sorted/unsorted lists: bitonic sort of (key, id) pairs, (score desc, id asc) ----------
__global__ void __launch_bounds__(1024)
topk_merge_kernel(const float* __restrict__ in_scores, const long long* __restrict__ in_ids, int n_in, int k,
                  float* __restrict__ out_scores, long long* __restrict__ out_ids, int npad) {
    extern __shared__ unsigned long long sm[];
    unsigned long long* keys = sm;                               // (score key << 32) | slot
    long long* ids = reinterpret_cast<long long*>(sm + npad);   // ids by slot
    const float* rs = in_scores + (size_t)blockIdx.x * n_in;
    const long long* ri = in_ids + (size_t)blockIdx.x * n_in;
    // Order by (key desc, id asc): rank ids first so that they fit in the low 32 bits -- ids inside one
    // query row are distinct non-negative candidate indices; entries with id < 0 are fillers.
    for (int i = threadIdx.x; i < npad; i += blockDim.x) ids[i] = (i < n_in) ? ri[i] : -1;
    __syncthreads();
    // sort slots by id descending-composite trick needs 64-bit ids; do a two-key bitonic instead
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        const bool valid = (i < n_in) && (ri[i] >= 0);
        keys[i] = valid ? (((unsigned long long)score_key(rs[i]) << 32) | (unsigned)i) : 0ull;
    }
    // bitonic with comparator: larger score key first; equal score -> smaller id first
    for (int size = 2; size <= npad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < npad / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long x = keys[lo], y = keys[hi];
                const uint32_t kx = (uint32_t)(x >> 32), ky = (uint32_t)(y >> 32);
                bool x_before_y;  // should x rank ahead of y?
                if (kx != ky) x_before_y = kx > ky;
                else if (x == 0ull || y == 0ull) x_before_y = (y == 0ull) && (x != 0ull);
                else x_before_y = ids[(uint32_t)x] < ids[(uint32_t)y];
                const bool equal = (x == y);
                if (!equal && (x_before_y != desc)) { keys[lo] = y; keys[hi] = x; }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const unsigned long long x = (i < npad) ? keys[i] : 0ull;
        float sc = -INFINITY;
        long long id = -1;
        if (x != 0ull) {
            const uint32_t slot = (uint32_t)x;
            sc = rs[slot];
            id = ids[slot];
        }
        out_scores[(size_t)blockIdx.x * k + i] = sc;
        out_ids[(size_t)blockIdx.x * k + i] = id;
    }
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace asp

extern "C" int asp_topk(const float* scores, int Q, long long N, int k, long long base_id, float* out_scores,
                        long long* out_ids, asp_stream_t stream) {
    ASP_REQUIRE(scores && out_scores && out_ids, "asp_topk: NULL pointer");
    ASP_REQUIRE(Q >= 0 && N >= 1 && k >= 1, "asp_topk: bad shape Q=%d N=%lld k=%d", Q, N, k);
    if (k > asp::kTopkMaxK) {
        asp::set_error("asp_topk: k=%d exceeds %d", k, asp::kTopkMaxK);
        return ASP_ERR_UNSUPPORTED;
    }
    ASP_REQUIRE(N < 0xffffffffLL, "asp_topk: N=%lld must be below 2^32", N);
    if (Q == 0) return ASP_OK;
    const int kpad = asp::next_pow2(k < 2 ? 2 : k);
    asp::topk_kernel<<<Q, asp::kTopkThreads, kpad * sizeof(unsigned long long), (cudaStream_t)stream>>>(
        scores, N, k, base_id, out_scores, out_ids, kpad);
    ASP_LAUNCH_CHECK("topk_kernel");
    return ASP_OK;
}

extern "C" int asp_topk_merge(const float* in_scores, const long long* in_ids, int Q, int R, int k,
                              float* out_scores, long long* out_ids, asp_stream_t stream) {
    ASP_REQUIRE(in_scores && in_ids && out_scores && out_ids, "asp_topk_merge: NULL pointer");
    ASP_REQUIRE(Q >= 0 && R >= 1 && k >= 1, "asp_topk_merge: bad shape Q=%d R=%d k=%d", Q, R, k);
    const long long n_in = (long long)R * k;
    if (n_in > 8192) {
        asp::set_error("asp_topk_merge: R*k=%lld exceeds 8192", n_in);
        return ASP_ERR_UNSUPPORTED;
    }
    if (Q == 0) return ASP_OK;
    const int npad = asp::next_pow2((int)(n_in < 2 ? 2 : n_in));
    const size_t smem = (size_t)npad * 16;
    if (smem > 48 * 1024)
        ASP_CUDA(cudaFuncSetAttribute(asp::topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    asp::topk_merge_kernel<<<Q, 1024, smem, (cudaStream_t)stream>>>(in_scores, in_ids, (int)n_in, k, out_scores,
                                                                    out_ids, npad);
    ASP_LAUNCH_CHECK("topk_merge_kernel");
    return ASP_OK;
}
