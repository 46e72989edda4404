// Other aggregation heads on the sentence-sentence cost matrix: cheap epilogues of K2.
//
//   * l2top2   -- allpair_masked_dist_l2topk, src/learning/facetid_models/pair_distances.py:295-345: sum of the two
//                 largest entries of (-cdist + pad mask) -- with fewer than two valid sentence pairs the runner-up
//                 is a masked entry, i.e. the mask constant -1e9 absorbs the sum exactly as in the reference;
//   * attention -- AllPairMaskedAttention.compute_distance, pair_distances.py:95-135 with
//                 activations.masked_2d_softmax (models_common/activations.py:35-61): softmax over the valid block of
//                 -cdist / T, document similarity = sum softmax * (-cdist).
// Input: the distance tensor [B,Sq,Sc] written by asp_pair_cost (valid block; padding ignored).  One warp per pair;
// Sq*Sc <= 1024 entries, read once (4 B per entry: HBM-trivial next to the 3 KB per sentence the cost kernel reads).
#include "common.cuh"

namespace asp {

__global__ void __launch_bounds__(128)
pair_heads_kernel(const float* __restrict__ cost, const int32_t* __restrict__ q_lens, int q_group,
                  const int32_t* __restrict__ c_lens, int B, int Sq, int Sc, float inv_temp, float* __restrict__ top2,
                  float* __restrict__ att, float* __restrict__ att_probs) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int ql = min(max(q_lens[b / q_group], 0), Sq), cl = min(max(c_lens[b], 0), Sc);
    const float* C = cost + (size_t)b * Sq * Sc;
    const int n = Sq * Sc;
    // pass 1: two smallest distances (= two largest similarities) and the largest logit
    float d1 = INFINITY, d2 = INFINITY;
    for (int e = lane; e < n; e += 32) {
        const int i = e / Sc, j = e - i * Sc;
        if (i < ql && j < cl) {
            const float d = C[e];
            if (d < d1) { d2 = d1; d1 = d; } else if (d < d2) { d2 = d; }
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const float o1 = __shfl_xor_sync(0xffffffffu, d1, s), o2 = __shfl_xor_sync(0xffffffffu, d2, s);
        // merge two sorted pairs (d1<=d2), (o1<=o2) keeping the two smallest
        const float m1 = fminf(d1, o1);
        const float m2 = fminf(fmaxf(d1, o1), fminf(d2, o2));
        d1 = m1;
        d2 = m2;
    }
    if (top2 && lane == 0) {
        // similarities -d; a missing runner-up is a masked entry of the reference (-cdist - 1e9 == -1e9 in fp32)
        const float s1 = (d1 < INFINITY) ? -d1 : kPadNeg, s2 = (d2 < INFINITY) ? -d2 : kPadNeg;
        top2[b] = s1 + s2;
    }
    if (att || att_probs) {
        // log-softmax over the valid block of (-d * inv_temp): max logit = -d1 * inv_temp
        const float mx = -d1 * inv_temp;
        float sum = 0.f;
        for (int e = lane; e < n; e += 32) {
            const int i = e / Sc, j = e - i * Sc;
            if (i < ql && j < cl) sum += expf(-C[e] * inv_temp - mx);
        }
        sum = warp_sum(sum);
        const float lse = mx + logf(sum);
        float acc = 0.f;
        for (int e = lane; e < n; e += 32) {
            const int i = e / Sc, j = e - i * Sc;
            float p = 0.f;
            if (i < ql && j < cl) {
                p = expf(-C[e] * inv_temp - lse);
                acc += p * (-C[e]);
            }
            if (att_probs) att_probs[(size_t)b * n + e] = p;
        }
        acc = warp_sum(acc);
        if (att && lane == 0) att[b] = (ql > 0 && cl > 0) ? acc : 0.f;
    }
}

// caching_score's score mixing (disent_models.py:298-307): scores[b] = sent_prop * scores[b] +
// abs_prop * (-|| q_cls - c_cls[b] + 1e-6 ||_2), the second term being -functional.pairwise_distance(p=2) of the CLS
// vectors (eps = 1e-6 added to the difference, as torch does).  One warp per candidate.
__global__ void __launch_bounds__(128)
mix_cls_kernel(float* __restrict__ scores, const float* __restrict__ q_cls, int q_group, const float* __restrict__ c_cls,
               int B, int D, float sent_prop, float abs_prop) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const float* qv = q_cls + (size_t)(b / q_group) * D;
    const float* cv = c_cls + (size_t)b * D;
    float acc = 0.f;
    for (int k = lane; k < D; k += 32) {
        const float d = qv[k] - cv[k] + 1e-6f;
        acc = fmaf(d, d, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) scores[b] = sent_prop * scores[b] + abs_prop * (-sqrtf(acc));
}

}  // namespace asp

extern "C" int asp_mix_cls_scores(float* scores, const float* q_cls, int q_group, const float* c_cls, int B, int D,
                                  float sent_prop, float abs_prop, asp_stream_t stream) {
    ASP_REQUIRE(scores && q_cls && c_cls, "asp_mix_cls_scores: NULL pointer");
    ASP_REQUIRE(B >= 0 && D >= 1 && q_group >= 1, "asp_mix_cls_scores: bad shape B=%d D=%d q_group=%d", B, D, q_group);
    if (B == 0) return ASP_OK;
    asp::mix_cls_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(scores, q_cls, q_group, c_cls, B, D, sent_prop,
                                                                      abs_prop);
    ASP_LAUNCH_CHECK("mix_cls_kernel");
    return ASP_OK;
}

extern "C" int asp_pair_heads(const float* cost, const int32_t* q_lens, int q_group, const int32_t* c_lens, int B, int Sq,
                              int Sc, float temp, float* top2, float* att, float* att_probs, asp_stream_t stream) {
    ASP_REQUIRE(cost && q_lens && c_lens, "asp_pair_heads: NULL input pointer");
    ASP_REQUIRE(B >= 0 && Sq >= 1 && Sc >= 1 && Sq * Sc <= 1024, "asp_pair_heads: bad shape B=%d Sq=%d Sc=%d", B, Sq, Sc);
    ASP_REQUIRE(q_group >= 1 && temp > 0.f, "asp_pair_heads: q_group >= 1 and temp > 0 required");
    ASP_REQUIRE(top2 || att || att_probs, "asp_pair_heads: no output requested");
    if (B == 0) return ASP_OK;
    asp::pair_heads_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(cost, q_lens, q_group, c_lens, B, Sq, Sc,
                                                                         1.0f / temp, top2, att, att_probs);
    ASP_LAUNCH_CHECK("pair_heads_kernel");
    return ASP_OK;
}
