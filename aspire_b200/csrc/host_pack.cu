// Host-side packing of a candidate pool: n per-paper encodings [S_j, D] (fp32, each contiguous, anywhere in host
// memory) -> one zero-padded [n, Smax, D] staging buffer, the layout every scoring entry point takes.
//
// Replaces the Python loop of WordSentAlignBiEnc.caching_score that pads the candidates of a pool one by one
// (src/learning/facetid_models/disent_models.py:274-290) and its twin behind AspireModel.score_pool: at 1 000
// candidates that loop costs 13-17 ms of host time (3-4 ms here) in front of a 0.05 ms kernel.  Plain memcpy / memset on a few
// threads; no CUDA (the destination is normally pinned memory the caller then hands to cudaMemcpyAsync).
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>
#include "common.cuh"

extern "C" int asp_pack_pool(const float* const* srcs, const int32_t* lens, int n, int smax, int D, float* dst, int threads) {
    ASP_REQUIRE(n >= 0 && smax >= 1 && D >= 1, "asp_pack_pool: bad sizes (n %d, smax %d, D %d)", n, smax, D);
    if (n == 0) return ASP_OK;
    ASP_REQUIRE(srcs && lens && dst, "asp_pack_pool: NULL argument");
    for (int j = 0; j < n; ++j) {
        ASP_REQUIRE(lens[j] >= 0 && lens[j] <= smax, "asp_pack_pool: block %d has %d rows (smax %d)", j, lens[j], smax);
        ASP_REQUIRE(lens[j] == 0 || srcs[j], "asp_pack_pool: block %d is NULL", j);
    }
    const size_t row = (size_t)D * sizeof(float), slot = (size_t)smax * row;
    auto work = [&](int lo, int hi) {
        for (int j = lo; j < hi; ++j) {
            char* out = reinterpret_cast<char*>(dst) + (size_t)j * slot;
            const size_t used = (size_t)lens[j] * row;
            if (used) memcpy(out, srcs[j], used);
            if (used < slot) memset(out + used, 0, slot - used);
        }
    };
    const int nt = std::max(1, std::min({threads, n / 32 + 1, 32}));
    if (nt == 1) {
        work(0, n);
        return ASP_OK;
    }
    std::vector<std::thread> pool;
    const int per = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) pool.emplace_back(work, std::min(n, t * per), std::min(n, (t + 1) * per));
    for (auto& th : pool) th.join();
    return ASP_OK;
}
