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

// the same key in five instructions (chunk kernels): x + 0 folds -0 into +0, the sign mask comes from an arithmetic shift
__device__ __forceinline__ uint32_t score_key_fast(float x, uint32_t negmask) {
    const float y = __uint_as_float(__float_as_uint(x) ^ negmask) + 0.f;
    const uint32_t b = __float_as_uint(y);
    const uint32_t key = b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
    return (y != y) ? 0u : key;
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
//   topk_stream_kernel -- a CTA walks 4 consecutive 8192-score chunks of one row through a 3-stage cp.async ring (64 KB in
//     flight per CTA, two CTAs per SM); per chunk every thread turns its 32 scores into order-preserving 32-bit keys, a
//     sort-free bound tau <= (chunk's k-th largest) comes from the thread maxima (warp shuffles only), and every score
//     >= tau (~190 of 8192) is appended unsorted to the chunk's list as a composite key (score key << 32 | ~global id).
//   topk_exact_kernel  -- redoes, exactly and sorted, the rare chunk with more than L scores at or above its bound
//     (heavy ties): everything above the k-th largest thread maximum plus the first of its equals in index order.
//   topk_merge_packed_kernel -- one CTA per query gathers its chunk lists, narrows them with the same bound and sorts a
//     few hundred keys; writes the best k decoded (score, id) and / or still packed.  The same kernel merges the lists
//     of R ranks after the all-gather of packed keys ([R][Q][k] as gathered, no concatenation on the host side).
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

// descending bitonic sort of one value per lane inside a warp (registers + shuffles)
__device__ __forceinline__ uint32_t warp_sort_desc(uint32_t v, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const uint32_t o = __shfl_xor_sync(0xffffffffu, v, stride);
            const bool lower = (lane & stride) == 0;           // this lane keeps the larger of the pair in a descending run
            const bool desc = (lane & size) == 0 || size == 32;
            v = (lower == desc) ? max(v, o) : min(v, o);
        }
    }
    return v;
}

// The chunk step.  FAST PATH: a lower bound tau of the chunk's k-th largest score is taken from the thread maxima without
// any block-wide sort -- every warp sorts its 32 thread maxima with shuffles, tau = the minimum over the 8 warps of each
// warp's ceil(k/8)-th largest, so at least k thread maxima (hence k scores) are >= tau -- and EVERY score >= tau is
// appended, unsorted, to the chunk's list (typically 150-250 of 8192; ties included, so the list is a superset of the
// chunk's top-k under the (score desc, id asc) order).  The merge kernel does the only sort.  A chunk with more than L
// such scores (heavy ties) is flagged (count = kTkOverflow) and redone by topk_exact_kernel: k-th largest thread maximum
// by a block-wide sort, everything above it, the first of its equals in index order, sorted, k entries.
constexpr unsigned int kTkOverflow = 0xffffffffu;

// keys of the thread's 32 elements e = 4 * (tid + 256 j) + c.  STAGED: from the shared-memory stage the thread filled
// itself with cp.async; else straight from global memory (FULL: 8192 scores, 16-byte aligned, no bounds checks).
template <bool FULL, bool STAGED>
__device__ __forceinline__ void topk_load_keys(const float* __restrict__ src, const float4* staged, int n, int negate, int tid,
                                               uint32_t (&key)[kTkPer]) {
    const bool vec = FULL || ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int e0 = 4 * (tid + kTkThreads * j);
        float v[4];
        if (STAGED) {
            const float4 f = staged[tid + kTkThreads * j];
            v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else if (FULL || (vec && e0 + 3 < n)) {
            const float4 f = ldg_stream(reinterpret_cast<const float4*>(src + e0));
            v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = (e0 + c < n) ? __ldg(src + e0 + c) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) key[4 * j + c] = (FULL || e0 + c < n) ? score_key_fast(v[c], negate ? 0x80000000u : 0u) : 0u;
    }
}

// stage_f: the chunk's raw scores in shared memory (staged chunks) or NULL.  With it the few survivors of a thread are
// emitted by a loop over the set bits of its 32-bit survivor mask (the key is rebuilt from the staged score), instead of 32
// predicated store sequences.
template <bool FULL>
__device__ __forceinline__ void topk_chunk_fast(const uint32_t (&key)[kTkPer], int n, int k, unsigned int id0, int negate,
                                                const float* stage_f, unsigned long long* __restrict__ out,
                                                unsigned int* __restrict__ count_out, int L) {
    __shared__ uint32_t wsel[kTkThreads / 32];
    __shared__ unsigned int n_ge, n_out;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t m = 0u;
#pragma unroll
    for (int i = 0; i < kTkPer; ++i) m = max(m, key[i]);
    __syncthreads();  // the previous chunk of this CTA is done with wsel / n_ge / n_out
    if (tid == 0) { n_ge = 0u; n_out = 0u; }
    const int keff = min(k, n);
    const int jsel = (keff + 7) / 8;                       // every warp contributes its jsel-th largest thread maximum
    const uint32_t sorted = warp_sort_desc(m, lane);
    if (lane == min(jsel, 32) - 1) wsel[warp] = sorted;
    __syncthreads();
    uint32_t tau = wsel[0];
#pragma unroll
    for (int w = 1; w < kTkThreads / 32; ++w) tau = min(tau, wsel[w]);
    uint32_t gm = 0u;  // bit i: element i of this thread is at or above the bound
#pragma unroll
    for (int i = 0; i < kTkPer; ++i) {
        const int e = 4 * (tid + kTkThreads * (i >> 2)) + (i & 3);
        if ((FULL || e < n) && key[i] >= tau) gm |= 1u << i;
    }
    const unsigned int cge = __popc(gm);
    unsigned int wsum = cge;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    if (lane == 0 && wsum) atomicAdd(&n_ge, wsum);
    __syncthreads();
    const unsigned int total = n_ge;
    if (total <= (unsigned)L) {
        unsigned int pos = cge ? atomicAdd(&n_out, cge) : 0u;
        if (stage_f) {
            const uint32_t negmask = negate ? 0x80000000u : 0u;
            while (gm) {
                const int i = __ffs(gm) - 1;
                gm &= gm - 1;
                const int e = 4 * (tid + kTkThreads * (i >> 2)) + (i & 3);
                const uint32_t ki = score_key_fast(stage_f[e], negmask);
                out[pos++] = ((unsigned long long)ki << 32) | (unsigned long long)(0xffffffffu - (id0 + (unsigned)e));
            }
        } else {
#pragma unroll
            for (int i = 0; i < kTkPer; ++i) {
                const int e = 4 * (tid + kTkThreads * (i >> 2)) + (i & 3);
                if (gm & (1u << i))
                    out[pos++] = ((unsigned long long)key[i] << 32) | (unsigned long long)(0xffffffffu - (id0 + (unsigned)e));
            }
        }
    }
    if (tid == 0) *count_out = (total <= (unsigned)L) ? total : kTkOverflow;
}

// The chunk step for STAGED chunks (8192 scores in shared memory), in the float domain: thread maxima and the survivor
// test are one FMNMX and one FSETP per score -- only the 8 x 32 thread maxima and the ~190 survivors are turned into keys.
// NEG: rank by -score.  Same bound, same survivors, same lists as topk_chunk_fast on the keys (v >= key_score(tau) <=>
// score_key(v) >= tau for every non-NaN v; NaN never survives unless the bound is 0, i.e. everything does).
template <bool NEG>
__device__ __forceinline__ void topk_chunk_staged(const float* stage_f, int k, unsigned int id0, unsigned long long* __restrict__ out,
                                                  unsigned int* __restrict__ count_out, int L) {
    __shared__ uint32_t wsel[kTkThreads / 32];
    __shared__ unsigned int n_ge, n_out;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float4* st4 = reinterpret_cast<const float4*>(stage_f);
    float v[kTkPer];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 f = st4[tid + kTkThreads * j];
        v[4 * j] = NEG ? -f.x : f.x; v[4 * j + 1] = NEG ? -f.y : f.y; v[4 * j + 2] = NEG ? -f.z : f.z; v[4 * j + 3] = NEG ? -f.w : f.w;
    }
    float mf = v[0];
#pragma unroll
    for (int i = 1; i < kTkPer; ++i) mf = fmaxf(mf, v[i]);  // NaNs are ignored unless every score is one
    const uint32_t m = score_key_fast(mf, 0u);
    __syncthreads();  // the previous chunk of this CTA is done with wsel / n_ge / n_out
    if (tid == 0) { n_ge = 0u; n_out = 0u; }
    const int jsel = (min(k, kTkChunk) + 7) / 8;
    const uint32_t sorted = warp_sort_desc(m, lane);
    if (lane == min(jsel, 32) - 1) wsel[warp] = sorted;
    __syncthreads();
    uint32_t tau = wsel[0];
#pragma unroll
    for (int w = 1; w < kTkThreads / 32; ++w) tau = min(tau, wsel[w]);
    uint32_t gm = 0xffffffffu;
    if (tau != 0u) {
        const float tf = key_score(tau);
        gm = 0u;
#pragma unroll
        for (int i = 0; i < kTkPer; ++i)
            if (v[i] >= tf) gm |= 1u << i;
    }
    const unsigned int cge = __popc(gm);
    unsigned int wsum = cge;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    if (lane == 0 && wsum) atomicAdd(&n_ge, wsum);
    __syncthreads();
    const unsigned int total = n_ge;
    if (total <= (unsigned)L) {
        unsigned int pos = cge ? atomicAdd(&n_out, cge) : 0u;
        while (gm) {
            const int i = __ffs(gm) - 1;
            gm &= gm - 1;
            const int e = 4 * (tid + kTkThreads * (i >> 2)) + (i & 3);
            const uint32_t ki = score_key_fast(stage_f[e], NEG ? 0x80000000u : 0u);
            out[pos++] = ((unsigned long long)ki << 32) | (unsigned long long)(0xffffffffu - (id0 + (unsigned)e));
        }
    }
    if (tid == 0) *count_out = (total <= (unsigned)L) ? total : kTkOverflow;
}

// Streaming kernel: CTA (x, row) walks `span` consecutive chunks of one row.  Whole, 16-byte-aligned chunks come in through
// a 3-stage cp.async ring in shared memory (each thread copies exactly the 128 bytes it will read back, so the ring needs
// no block-wide synchronisation; two chunks = 64 KB are in flight per CTA while a third is being selected from, two CTAs
// per SM); a ragged or unaligned chunk is read directly.
constexpr int kTkStages = 3;
#ifndef ASP_TK_SPAN
#define ASP_TK_SPAN 4
#endif
constexpr int kTkSpan = ASP_TK_SPAN;
__device__ __forceinline__ void tk_cp_async16(uint32_t dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(kTkThreads, 2)
topk_stream_kernel(const float* __restrict__ scores, long long N, int k, int negate, long long base_id, int nchunks,
                   unsigned long long* __restrict__ lists, unsigned int* __restrict__ counts, int L) {
    extern __shared__ __align__(16) float ring[];  // [kTkStages][8192]
    const int tid = threadIdx.x;
    const long long row = blockIdx.x;  // rows on grid.x (up to 2^31 - 1 queries), chunk groups on grid.y
    const int c0 = blockIdx.y * kTkSpan, c1 = min(c0 + kTkSpan, nchunks);
    const float* rowp = scores + row * N;
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
    auto staged_ok = [&](int c) {
        return (N - (long long)c * kTkChunk >= kTkChunk) && ((reinterpret_cast<uintptr_t>(rowp + (long long)c * kTkChunk) & 15u) == 0);
    };
    auto issue = [&](int c) {  // this thread's 8 x 16 bytes of chunk c into stage c % kTkStages (empty group when not staged)
        if (c < c1 && staged_ok(c)) {
            const float* src = rowp + (long long)c * kTkChunk;
            const uint32_t dst = ring_u32 + (uint32_t)((c % kTkStages) * kTkChunk * 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) tk_cp_async16(dst + 16u * (tid + kTkThreads * j), src + 4 * (tid + kTkThreads * j));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(c0);
    issue(c0 + 1);
    for (int c = c0; c < c1; ++c) {
        issue(c + 2);                                              // its stage was read by this thread one iteration ago
        asm volatile("cp.async.wait_group 2;" ::: "memory");      // chunk c has landed (groups complete in order)
        const long long start = (long long)c * kTkChunk;
        unsigned long long* out = lists + ((size_t)row * nchunks + c) * L;
        unsigned int* cnt = counts + (size_t)row * nchunks + c;
        if (staged_ok(c)) {
            const float* stage_f = ring + (size_t)(c % kTkStages) * kTkChunk;
            if (negate) topk_chunk_staged<true>(stage_f, k, (unsigned)(base_id + start), out, cnt, L);
            else topk_chunk_staged<false>(stage_f, k, (unsigned)(base_id + start), out, cnt, L);
        } else {
            uint32_t key[kTkPer];
            const int n = (int)min((long long)kTkChunk, N - start);
            topk_load_keys<false, false>(rowp + start, nullptr, n, negate, tid, key);
            topk_chunk_fast<false>(key, n, k, (unsigned)(base_id + start), negate, nullptr, out, cnt, L);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// Exact redo of the chunks the streaming kernel flagged (more than L scores at or above its bound: heavy ties).  A small
// persistent grid scans the counts; a flagged chunk gets the k-th largest thread maximum by a block-wide sort, everything
// above it, the first of its equals in index order, sorted: k entries.
__global__ void __launch_bounds__(kTkThreads)
topk_exact_kernel(const float* __restrict__ scores, long long N, int k, int negate, long long base_id, int nchunks, int Q,
                  unsigned long long* __restrict__ lists, unsigned int* __restrict__ counts, int L, int cap) {
    extern __shared__ unsigned long long surv[];  // cap composite keys
    __shared__ uint32_t tmax[kTkThreads];
    __shared__ unsigned int warp_tot[kTkThreads / 32];
    const int tid = threadIdx.x;
    const long long nl = (long long)Q * nchunks;
    // stripe of this CTA: lists [lo, hi); 256 flags are tested per step, flagged lists are then redone one by one
    const long long per = (nl + gridDim.x - 1) / gridDim.x, lo = per * blockIdx.x, hi = min(nl, lo + per);
    for (long long base = lo; base < hi; base += kTkThreads) {
        const bool mine = base + tid < hi && counts[base + tid] == kTkOverflow;
        if (!__syncthreads_or(mine)) continue;
        for (long long list = base; list < min(hi, base + kTkThreads); ++list) {
        if (counts[list] != kTkOverflow) continue;  // block-uniform
        const long long row = list / nchunks, start = (list - row * nchunks) * kTkChunk;
        const int n = (int)min((long long)kTkChunk, N - start);
        uint32_t key[kTkPer];
        topk_load_keys<false, false>(scores + row * N + start, nullptr, n, negate, tid, key);
        uint32_t m = 0u;
#pragma unroll
        for (int i = 0; i < kTkPer; ++i) m = max(m, key[i]);
        __syncthreads();
        tmax[tid] = m;
        bitonic_desc_u32(tmax, kTkThreads);
        const int keff = min(k, n);
        const uint32_t tau = (keff >= 1 && keff <= kTkThreads) ? tmax[keff - 1] : 0u;
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
        // of the elements equal to tau: the first (keff - G) in index order.  Index order = j slab by slab (1024
        // consecutive elements each), inside a slab thread by thread, inside a thread component by component
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
        int npad = 2;
        while (npad < (int)pos) npad <<= 1;
        bitonic_desc_u64(surv, min(npad, cap));
        const int nout = min((int)pos, k);
        unsigned long long* out = lists + (size_t)list * L;
        for (int i = tid; i < nout; i += kTkThreads) out[i] = surv[i];
        if (tid == 0) counts[list] = (unsigned)nout;
        __syncthreads();
        }
    }
}

// descending bitonic sort of one 64-bit value per lane inside a warp
__device__ __forceinline__ unsigned long long warp_sort_desc64(unsigned long long v, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, stride);
            const bool lower = (lane & stride) == 0, desc = (lane & size) == 0;
            v = (lower == desc) ? max(v, o) : min(v, o);
        }
    }
    return v;
}

constexpr int kTkSel = 1024;  // survivors the merge sorts after its selection step
#ifndef ASP_TK_MERGE_THREADS
#define ASP_TK_MERGE_THREADS 256
#endif
constexpr int kTkMergeThreads = ASP_TK_MERGE_THREADS;  // 256: cheap block barriers in the 45-stage sort, three CTAs per SM

// Merge: CTA (q, grp) gathers lists [grp * group, min(nlists, (grp + 1) * group)) of query q -- list l starts at
// lists[l * l_stride + q * q_stride] and holds counts[q * nlists + l] entries (counts == NULL: exactly k, zeros = fillers)
// -- and writes the best k as the grp-th output list of query q (gridDim.y lists per query; the decoded outputs are
// meaningful when gridDim.y == 1).  Composite keys are distinct, so the best k are simply the k largest: the same bound
// as in the chunk kernel (minimum over the warps of each warp's ceil(k / warps)-th largest thread maximum) leaves a few
// hundred of the ~3 000 gathered keys, and only those are sorted (a 45-stage sort of 512 instead of a 78-stage sort of
// 4 096); more than kTkSel survivors, or a k the bound does not cover, sort everything.
__global__ void __launch_bounds__(1024)
topk_merge_packed_kernel(const unsigned long long* __restrict__ lists, const unsigned int* __restrict__ counts, int nlists,
                         size_t l_stride, size_t q_stride, int k, int group, float* __restrict__ out_scores,
                         long long* __restrict__ out_ids, unsigned long long* __restrict__ out_packed, int cap) {
    extern __shared__ unsigned long long mk[];
    __shared__ unsigned long long sel[kTkSel];
    __shared__ unsigned long long wsel[32];
    __shared__ int off_s[129];
    __shared__ unsigned int n_ge, n_out;
    const int q = blockIdx.x, l0 = blockIdx.y * group, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int nl = min(group, nlists - l0);
    // list lengths: loaded in parallel (one dependent global load per CTA instead of nl), prefix by one thread from shared memory
    if ((int)threadIdx.x < nl) off_s[threadIdx.x + 1] = counts ? (int)counts[(size_t)q * nlists + l0 + threadIdx.x] : k;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        off_s[0] = 0;
        for (int l = 0; l < nl; ++l) {
            run += off_s[l + 1];
            off_s[l + 1] = min(run, cap);
        }
        n_ge = 0u;
        n_out = 0u;
    }
    __syncthreads();
    const int total = off_s[nl];
    // gather: one flat loop, every element's load independent of the others
    for (int j = threadIdx.x; j < total; j += blockDim.x) {
        int l = 0;
        while (l + 1 < nl && j >= off_s[l + 1]) ++l;
        mk[j] = lists[(size_t)(l0 + l) * l_stride + (size_t)q * q_stride + (j - off_s[l])];
    }
    __syncthreads();
    unsigned long long* sorted = mk;
    int npad = 2;
    const int jsel = (k + nwarps - 1) / nwarps;
    bool selected = false;
    if (total > kTkSel && jsel <= 32) {
        unsigned long long m = 0ull;
        for (int i = threadIdx.x; i < total; i += blockDim.x) m = max(m, mk[i]);
        const unsigned long long srt = warp_sort_desc64(m, lane);
        if (lane == jsel - 1) wsel[warp] = srt;
        __syncthreads();
        unsigned long long tau = wsel[0];
        for (int w = 1; w < nwarps; ++w) tau = min(tau, wsel[w]);
        unsigned int c = 0;
        if (tau != 0ull)
            for (int i = threadIdx.x; i < total; i += blockDim.x) c += (mk[i] >= tau);
        unsigned int wsum = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        if (lane == 0 && wsum) atomicAdd(&n_ge, wsum);
        __syncthreads();
        if (tau != 0ull && n_ge <= (unsigned)kTkSel) {  // block-uniform
            unsigned int pos = c ? atomicAdd(&n_out, c) : 0u;
            for (int i = threadIdx.x; i < total; i += blockDim.x) {
                const unsigned long long x = mk[i];
                if (x >= tau) sel[pos++] = x;
            }
            while (npad < (int)n_ge) npad <<= 1;
            for (int i = (int)n_ge + threadIdx.x; i < npad; i += blockDim.x) sel[i] = 0ull;
            sorted = sel;
            selected = true;
        }
    }
    if (!selected) {
        while (npad < total) npad <<= 1;
        for (int i = total + threadIdx.x; i < npad; i += blockDim.x) mk[i] = 0ull;
    }
    bitonic_desc_u64(sorted, npad);
    const size_t o = ((size_t)q * gridDim.y + blockIdx.y) * k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const unsigned long long x = (i < npad) ? sorted[i] : 0ull;
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
static int tk_list_len(int k) {  // entries of a chunk list: room for every score at or above the chunk's bound
    int L = 64;
    while (L < 4 * k) L <<= 1;
    return L;  // <= 512 for k <= 128
}
struct TkPlan {
    long long nchunks;
    int L, group0;                 // chunk lists: length, lists per merge CTA of the first level
    size_t lists0, counts0, rest;  // bytes: chunk lists, their counts, lists of the later merge levels
};
static TkPlan tk_plan(int Q, long long N, int k) {
    TkPlan p;
    p.nchunks = tk_chunks(N);
    p.L = tk_list_len(k);
    p.group0 = std::min(kTkMergeCap / p.L, 128);
    const size_t q = (size_t)(Q > 0 ? Q : 1);
    p.lists0 = q * p.nchunks * p.L * sizeof(unsigned long long);
    p.counts0 = (q * p.nchunks * sizeof(unsigned int) + 255) & ~(size_t)255;
    long long lists = (p.nchunks + p.group0 - 1) / p.group0, total = 0;
    const int group = std::min(kTkMergeCap / k, 128);
    while (lists > 1) {  // level outputs of k entries each, merged `group` at a time
        total += lists;
        lists = (lists + group - 1) / group;
    }
    p.rest = q * (size_t)total * k * sizeof(unsigned long long);
    return p;
}
}  // namespace asp

extern "C" size_t asp_topk_workspace_bytes(int Q, long long N, int k) {
    if (k < 1 || k > asp::kTkMaxK || N < 1 || Q < 0) return 0;
    const asp::TkPlan p = asp::tk_plan(Q, N, k);
    return p.lists0 + p.counts0 + p.rest;
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
    const TkPlan p = tk_plan(Q, N, k);
    int nlists = (int)p.nchunks;
    int cap = 2;
    while (cap < 32 * (k - 1) + k) cap <<= 1;
    const size_t smem = (size_t)cap * sizeof(unsigned long long);
    static thread_local int attr_dev = -1;
    int dev = 0;
    ASP_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        ASP_CUDA(cudaFuncSetAttribute(topk_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTkStages * kTkChunk * 4));
        ASP_CUDA(cudaFuncSetAttribute(topk_stream_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ASP_CUDA(cudaFuncSetAttribute(topk_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        ASP_CUDA(cudaFuncSetAttribute(topk_merge_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        ASP_CUDA(cudaFuncSetAttribute(topk_merge_packed_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_dev = dev;
    }
    char* w = static_cast<char*>(workspace);
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(w);
    unsigned int* counts = reinterpret_cast<unsigned int*>(w + p.lists0);
    unsigned long long* next = reinterpret_cast<unsigned long long*>(w + p.lists0 + p.counts0);
    ASP_REQUIRE((nlists + kTkSpan - 1) / kTkSpan <= 65535, "asp_topk_ws: N=%lld exceeds %lld scores per row", N,
                65535LL * kTkSpan * kTkChunk);
    topk_stream_kernel<<<dim3(Q, (nlists + kTkSpan - 1) / kTkSpan), kTkThreads, kTkStages * kTkChunk * 4, stream>>>(
        scores, N, k, negate, base_id, nlists, lists, counts, p.L);
    ASP_LAUNCH_CHECK("topk_stream_kernel");
    const long long nl_all = (long long)Q * nlists;
    topk_exact_kernel<<<(unsigned)std::min<long long>(nl_all, 2LL * sm_count()), kTkThreads, smem, stream>>>(
        scores, N, k, negate, base_id, nlists, Q, lists, counts, p.L, cap);
    ASP_LAUNCH_CHECK("topk_exact_kernel");
    // level 0: variable-length chunk lists; later levels: lists of exactly k sorted entries
    int group = p.group0, stride = p.L;
    const unsigned int* cnt = counts;
    for (;;) {  // list l of query q: lists[(q * nlists + l) * stride]
        const bool last = nlists <= group;
        const int ngroups = (nlists + group - 1) / group;
        const size_t cap_keys = std::min<size_t>((size_t)next_pow2(std::min(nlists, group) * stride), kTkMergeCap);
        topk_merge_packed_kernel<<<dim3(Q, ngroups), kTkMergeThreads, cap_keys * 8, stream>>>(
            lists, cnt, nlists, (size_t)stride, (size_t)nlists * stride, k, group, last ? out_scores : nullptr,
            last ? out_ids : nullptr, last ? out_packed : next, (int)cap_keys);
        ASP_LAUNCH_CHECK("topk_merge_packed_kernel");
        if (last) break;
        lists = next;
        next += (size_t)Q * ngroups * k;
        nlists = ngroups;
        stride = k;
        cnt = nullptr;
        group = std::min(kTkMergeCap / k, 128);
    }
    return ASP_OK;
}

extern "C" int asp_topk_merge_packed(const unsigned long long* gathered, int R, int Q, int k, float* out_scores,
                                     long long* out_ids, asp_stream_t stream) {
    using namespace asp;
    ASP_REQUIRE(gathered && (out_scores || out_ids), "asp_topk_merge_packed: NULL pointer");
    ASP_REQUIRE(R >= 1 && Q >= 0 && k >= 1, "asp_topk_merge_packed: bad shape R=%d Q=%d k=%d", R, Q, k);
    if ((long long)R * k > kTkMergeCap || R > 128) {
        set_error("asp_topk_merge_packed: R=%d lists of k=%d exceed one merge (R <= 128, R*k <= %d)", R, k, kTkMergeCap);
        return ASP_ERR_UNSUPPORTED;
    }
    if (Q == 0) return ASP_OK;
    const size_t cap_keys = (size_t)next_pow2(R * k < 2 ? 2 : R * k);
    if (cap_keys * 8 > 48 * 1024)
        ASP_CUDA(cudaFuncSetAttribute(topk_merge_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    // gathered[r][q][k]: list r of query q
    topk_merge_packed_kernel<<<Q, kTkMergeThreads, cap_keys * 8, (cudaStream_t)stream>>>(gathered, nullptr, R, (size_t)Q * k, (size_t)k, k, R,
                                                                             out_scores, out_ids, nullptr, (int)cap_keys);
    ASP_LAUNCH_CHECK("topk_merge_packed_kernel");
    return ASP_OK;
}
