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

// ---------------------------------------------------------------------------------------------------------------------
// Single-pass top-k (k <= 128): the score matrix is read from HBM ONCE.
//
//   topk_chunk_kernel  -- one CTA per (query, chunk of 8192 scores).  Every thread keeps its 32 scores in registers as
//     order-preserving 32-bit keys.  The k-th largest of the 256 per-thread maxima, tau, is a lower bound of the chunk's
//     k-th largest score (k threads hold an element >= tau), so the chunk's top-k are among the elements >= tau:
//     the ones strictly above it (at most 32 (k-1): only k-1 thread maxima exceed tau) are all collected, and of the
//     ones equal to it the first (k - #greater) in index order (ties rank by smallest id).  For continuous scores that
//     is ~250 survivors per chunk; they are sorted in shared memory (composite key: score key << 32 | ~global id) and the
//     best k go to a scratch list.
//   topk_merge_packed_kernel -- one CTA per query sorts its nlists x k composite keys and writes the best k, decoded
//     (score, id) and / or still packed.  The same kernel merges the lists of R ranks after the all-gather of packed
//     keys ([R][Q][k] as gathered, no concatenation on the host side).
// The 5-pass radix select above stays for k > 128.
constexpr int kTkThreads = 256, kTkPer = 32, kTkChunk = kTkThreads * kTkPer;
constexpr int kTkMaxK = 128;

__device__ __forceinline__ float key_score(uint32_t key) {  // inverse of score_key (NaN / -0 aside)
    const uint32_t b = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
    return __uint_as_float(b);
}

__device__ void bitonic_desc_u32(uint32_t* a, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const uint32_t x = a[lo], y = a[hi];
                if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
            }
        }
    }
    __syncthreads();
}

// exclusive prefix sum of one value per thread over the block (256 threads); returns the block total in `total`
__device__ __forceinline__ unsigned int block_scan_excl(unsigned int v, unsigned int* warp_tot, unsigned int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();  // warp_tot may still be read from the previous scan
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    unsigned int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kTkThreads / 32; ++w) {
        const unsigned int x = warp_tot[w];
        if (w < warp) base += x;
        tot += x;
    }
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(kTkThreads)
topk_chunk_kernel(const float* __restrict__ scores, long long N, int k, int negate, long long base_id, int nchunks,
                  unsigned long long* __restrict__ lists, int cap) {
    extern __shared__ unsigned long long surv[];  // cap composite keys
    __shared__ uint32_t tmax[kTkThreads];
    __shared__ unsigned int warp_tot[kTkThreads / 32];
    const int tid = threadIdx.x;
    const long long row = blockIdx.y, start = (long long)blockIdx.x * kTkChunk;
    const int n = (int)min((long long)kTkChunk, N - start);
    const float* src = scores + row * N + start;
    // element e = 4 * (tid + 256 j) + c  (j < 8, c < 4): coalesced 16-byte loads when the chunk start is 16-byte aligned
    uint32_t key[kTkPer];
    const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int e0 = 4 * (tid + kTkThreads * j);
        float v[4];
        if (vec && e0 + 3 < n) {
            const float4 f = ldg_stream(reinterpret_cast<const float4*>(src + e0));
            v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = (e0 + c < n) ? __ldg(src + e0 + c) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) key[4 * j + c] = (e0 + c < n) ? score_key(negate ? -v[c] : v[c]) : 0u;
    }
    uint32_t m = 0u;
#pragma unroll
    for (int i = 0; i < kTkPer; ++i) m = max(m, key[i]);
    tmax[tid] = m;
    bitonic_desc_u32(tmax, kTkThreads);
    const int keff = min(k, n);
    const uint32_t tau = (keff >= 1 && keff <= kTkThreads) ? tmax[keff - 1] : 0u;
    // ---- survivors strictly above tau: all of them ----
    unsigned int cg = 0, ce = 0;
#pragma unroll
    for (int i = 0; i < kTkPer; ++i) {
        const int e = 4 * (tid + kTkThreads * (i >> 2)) + (i & 3);
        cg += (e < n && key[i] > tau);
        ce += (e < n && key[i] == tau);
    }
    unsigned int G, E;
    unsigned int og = block_scan_excl(cg, warp_tot, G);
    for (int i = tid; i < cap; i += kTkThreads) surv[i] = 0ull;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kTkPer; ++i) {
        const int e = 4 * (tid + kTkThreads * (i >> 2)) + (i & 3);
        if (e < n && key[i] > tau)
            surv[og++] = ((unsigned long long)key[i] << 32) | (unsigned long long)(0xffffffffu - (uint32_t)(base_id + start + e));
    }
    // ---- of the elements equal to tau: the first (keff - G) in index order.  Index order = j slab by slab (1024
    //      consecutive elements each), inside a slab thread by thread, inside a thread component by component ----
    int need = (int)keff - (int)min(G, (unsigned)keff);
    unsigned int pos = G;
    (void)block_scan_excl(ce, warp_tot, E);
    if (need > 0 && E > 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (need <= 0) break;  // block-uniform
            unsigned int c_slab = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int e = 4 * (tid + kTkThreads * j) + c;
                c_slab += (e < n && key[4 * j + c] == tau);
            }
            unsigned int tot;
            unsigned int off = block_scan_excl(c_slab, warp_tot, tot);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int e = 4 * (tid + kTkThreads * j) + c;
                if (e < n && key[4 * j + c] == tau) {
                    if ((int)off < need)
                        surv[pos + off] = ((unsigned long long)tau << 32) |
                                          (unsigned long long)(0xffffffffu - (uint32_t)(base_id + start + e));
                    ++off;
                }
            }
            const int took = min((int)tot, need);
            pos += took;
            need -= took;
        }
    }
    __syncthreads();
    // ---- sort the survivors, best k to the scratch list of this chunk ----
    int npad = 2;
    while (npad < (int)pos) npad <<= 1;
    bitonic_desc_u64(surv, min(npad, cap));
    unsigned long long* out = lists + ((size_t)row * nchunks + blockIdx.x) * k;
    for (int i = tid; i < k; i += kTkThreads) out[i] = (i < (int)pos) ? surv[i] : 0ull;
}

// lists: nlists lists of k composite keys per query; list l of query q starts at lists[l * l_stride + q * q_stride].
// CTA (q, grp) merges lists [grp * group, min(nlists, (grp + 1) * group)) into the grp-th output list of query q
// (gridDim.y output lists per query; the decoded outputs are only meaningful when gridDim.y == 1).
__global__ void __launch_bounds__(1024)
topk_merge_packed_kernel(const unsigned long long* __restrict__ lists, int nlists, size_t l_stride, size_t q_stride, int k,
                         int group, float* __restrict__ out_scores, long long* __restrict__ out_ids,
                         unsigned long long* __restrict__ out_packed, int npad) {
    extern __shared__ unsigned long long mk[];
    const int q = blockIdx.x, l0 = blockIdx.y * group;
    const int n_in = min(group, nlists - l0) * k;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        unsigned long long v = 0ull;
        if (i < n_in) v = lists[(size_t)(l0 + i / k) * l_stride + (size_t)q * q_stride + (i % k)];
        mk[i] = v;
    }
    bitonic_desc_u64(mk, npad);
    const size_t o = ((size_t)q * gridDim.y + blockIdx.y) * k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const unsigned long long x = (i < npad) ? mk[i] : 0ull;
        if (out_packed) out_packed[o + i] = x;
        if (out_scores) out_scores[o + i] = x ? key_score((uint32_t)(x >> 32)) : -INFINITY;
        if (out_ids) out_ids[o + i] = x ? (long long)(0xffffffffu - (uint32_t)(x & 0xffffffffull)) : -1;
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

// ---- single-pass entry points ---------------------------------------------------------------------------------------
namespace asp {
constexpr int kTkMergeCap = 8192;  // composite keys one merge CTA sorts in shared memory (64 KB)
static long long tk_chunks(long long N) { return (N + kTkChunk - 1) / kTkChunk; }
static int tk_group(int k) { return kTkMergeCap / k; }  // lists one merge CTA takes (k <= 128 -> >= 64)
}  // namespace asp

extern "C" size_t asp_topk_workspace_bytes(int Q, long long N, int k) {
    if (k < 1 || k > asp::kTkMaxK || N < 1 || Q < 0) return 0;
    // chunk lists + the lists of the intermediate merge levels (each level shrinks the list count by >= 64x)
    long long lists = asp::tk_chunks(N), total = lists;
    while (lists > asp::tk_group(k)) {
        lists = (lists + asp::tk_group(k) - 1) / asp::tk_group(k);
        total += lists;
    }
    return (size_t)(Q > 0 ? Q : 1) * (size_t)total * k * sizeof(unsigned long long);
}

extern "C" int asp_topk_ws(const float* scores, int Q, long long N, int k, long long base_id, int negate, float* out_scores,
                           long long* out_ids, unsigned long long* out_packed, void* workspace, size_t workspace_bytes,
                           asp_stream_t stream_) {
    using namespace asp;
    ASP_REQUIRE(scores && (out_scores || out_ids || out_packed), "asp_topk_ws: NULL pointer");
    ASP_REQUIRE(Q >= 0 && N >= 1 && k >= 1, "asp_topk_ws: bad shape Q=%d N=%lld k=%d", Q, N, k);
    if (Q == 0) return ASP_OK;
    const size_t need = asp_topk_workspace_bytes(Q, N, k);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (need == 0 || !workspace || workspace_bytes < need || base_id < 0 || base_id + N >= 0xffffffffLL) {
        // outside the single-pass kernels' limits (k > 128, ids beyond 32 bits, or no scratch): the radix-select kernel
        ASP_REQUIRE(!negate && !out_packed && out_scores && out_ids,
                    "asp_topk_ws: negate / packed output need k <= %d, ids < 2^32 and a workspace of asp_topk_workspace_bytes()",
                    kTkMaxK);
        return asp_topk(scores, Q, N, k, base_id, out_scores, out_ids, stream_);
    }
    int nlists = (int)tk_chunks(N);
    int cap = 2;
    while (cap < 32 * (k - 1) + k) cap <<= 1;
    const size_t smem = (size_t)cap * sizeof(unsigned long long);
    static thread_local int attr_dev = -1;
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        ASP_CUDA(cudaFuncSetAttribute(topk_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        ASP_CUDA(cudaFuncSetAttribute(topk_merge_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_dev = dev;
    }
    unsigned long long* lists = static_cast<unsigned long long*>(workspace);
    topk_chunk_kernel<<<dim3(nlists, Q), kTkThreads, smem, stream>>>(scores, N, k, negate, base_id, nlists, lists, cap);
    ASP_LAUNCH_CHECK("topk_chunk_kernel");
    const int group = tk_group(k);
    for (;;) {  // lists of query q: lists[(q * nlists + l) * k]
        const bool last = nlists <= group;
        const int ngroups = (nlists + group - 1) / group;
        const int n_in = (last ? nlists : group) * k;
        const int npad = next_pow2(n_in < 2 ? 2 : n_in);
        unsigned long long* next = lists + (size_t)Q * nlists * k;
        topk_merge_packed_kernel<<<dim3(Q, ngroups), 1024, (size_t)npad * 8, stream>>>(
            lists, nlists, (size_t)k, (size_t)nlists * k, k, group, last ? out_scores : nullptr, last ? out_ids : nullptr,
            last ? out_packed : next, npad);
        ASP_LAUNCH_CHECK("topk_merge_packed_kernel");
        if (last) break;
        lists = next;
        nlists = ngroups;
    }
    return ASP_OK;
}

extern "C" int asp_topk_merge_packed(const unsigned long long* gathered, int R, int Q, int k, float* out_scores,
                                     long long* out_ids, asp_stream_t stream) {
    using namespace asp;
    ASP_REQUIRE(gathered && (out_scores || out_ids), "asp_topk_merge_packed: NULL pointer");
    ASP_REQUIRE(R >= 1 && Q >= 0 && k >= 1, "asp_topk_merge_packed: bad shape R=%d Q=%d k=%d", R, Q, k);
    if ((long long)R * k > kTkMergeCap) {
        set_error("asp_topk_merge_packed: R*k=%lld exceeds %d", (long long)R * k, kTkMergeCap);
        return ASP_ERR_UNSUPPORTED;
    }
    if (Q == 0) return ASP_OK;
    const int npad = next_pow2(R * k < 2 ? 2 : R * k);
    if ((size_t)npad * 8 > 48 * 1024)
        ASP_CUDA(cudaFuncSetAttribute(topk_merge_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    // gathered[r][q][k]: list r of query q
    topk_merge_packed_kernel<<<Q, 1024, (size_t)npad * 8, (cudaStream_t)stream>>>(gathered, R, (size_t)Q * k, (size_t)k, k, R,
                                                                                 out_scores, out_ids, nullptr, npad);
    ASP_LAUNCH_CHECK("topk_merge_packed_kernel");
    return ASP_OK;
}
