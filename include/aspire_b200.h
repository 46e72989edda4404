/*
 * aspire_b200 -- C ABI of the B200-native (sm_100a) implementation of Aspire's fine-grained
 * document-similarity scoring path.
 *
 * The reference (allenai/aspire) is pure Python and has NO native interface; each entry point below
 * names the reference Python code it replaces (paths relative to the reference root).  A maintainer binds
 * these with `ctypes`; the in-tree binding is aspire_b200/_abi.py.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (ASP_ERR_*); asp_last_error() gives the
 *     thread-local message of the last failing call on the calling thread;
 *   - all data pointers are DEVICE pointers to contiguous row-major fp32 / int32 arrays owned by the
 *     caller, 16-byte aligned, unless a parameter is documented as host memory;
 *   - nothing allocates, nothing synchronises: work is enqueued on `stream` (a cudaStream_t, may be 0);
 *   - optional outputs may be NULL;
 *   - "lens" arrays are int32 on the device; sentences beyond a document's length are padding and
 *     never influence a result.
 */
#ifndef ASPIRE_B200_H_
#define ASPIRE_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASP_OK 0
#define ASP_ERR_INVALID (-1)     /* bad argument (shape, alignment, NULL) */
#define ASP_ERR_CUDA (-2)        /* a CUDA runtime call failed */
#define ASP_ERR_UNSUPPORTED (-3) /* valid request outside the built kernels' limits */

#define ASP_MAX_SENTS 128   /* max sentences per document handled by the scoring kernels */
#define ASP_MAX_EPS 512     /* max entries of an epsilon schedule */

typedef void* asp_stream_t; /* cudaStream_t */

/* Library version: major*10000 + minor*100 + patch. */
int asp_version(void);
/* Message of the last error raised on this thread ("" if none). Never NULL. */
const char* asp_last_error(void);
/* Number of SMs of the current device (used by callers to size persistent grids); <0 on error. */
int asp_sm_count(void);
/* Kernels launched by this library since load (all threads); used by bench.py's gpu_launches. */
long long asp_launch_count(void);
/* Tuning/testing knobs.  "ot_kernel": 0 auto, 1 force warp-per-pair, 2 force thread-per-pair. */
int asp_set_option(const char* key, int value);

/* ---- K1: per-sentence token-span mean pooling -------------------------------------------------
 * Replaces the Python pooling loop of AspireConSent.consent_reps_bert
 *   examples/ex_aspire_consent.py:75-100  (= src/learning/facetid_models/disent_models.py:509-534).
 * hidden [B,L,D] fp32 (BERT last_hidden_state); spans [B,Smax,2] int32 half-open token ranges
 * [start,end) (start<0 or end<=start => missing sentence => zero row).
 * sent_reps [B,Smax,D] = sum(hidden[b,start:end]) / max(end-start,1);  cls_reps [B,D] = hidden[b,0] (may be NULL).
 */
int asp_span_mean_pool(const float* hidden, const int32_t* spans, int B, int L, int D, int Smax,
                       float* sent_reps, float* cls_reps, asp_stream_t stream);

/* ---- K2: pairwise sentence-sentence L2 cost matrix --------------------------------------------
 * Replaces torch.cdist at src/learning/facetid_models/pair_distances.py:49-50,167 and geomloss's
 * `distances()` (sqrt(clamp_min(|x|^2 - 2x.y + |y|^2, 1e-8))).
 * q [Bq,Sq,D], c [B,Sc,D]; Bq == B (paired) or q_broadcast != 0 (one query [1,Sq,D] against all B).
 * cost [B,Sq,Sc]: L2 distance for i<q_len, j<c_len; 0 elsewhere.
 */
int asp_pair_cost(const float* q, const int32_t* q_lens, int q_broadcast, const float* c,
                  const int32_t* c_lens, int B, int Sq, int Sc, int D, float* cost, asp_stream_t stream);

/* ---- K2+K3: tsAspire single best match ---------------------------------------------------------
 * Replaces allpair_masked_dist_l2max, src/learning/facetid_models/pair_distances.py:138-186
 * (numpy twin: src/pre_process/pp_gen_nearest.py:942-961).
 * best [B] = max_{i<ql,j<cl} -||q_i-c_j||;  flat_idx [B] = i*Sc+j of that entry, first occurrence
 * (pair_distances.py:176);  pair_sims [B,Sq,Sc] (optional) = -dist inside the valid block, -1e9 outside.
 * A pair with an empty side yields best = -1e9, flat_idx = 0 (what torch.max gives on an all-masked row).
 */
int asp_l2max(const float* q, const int32_t* q_lens, int q_broadcast, const float* c, const int32_t* c_lens,
              int B, int Sq, int Sc, int D, float* best, int32_t* flat_idx, float* pair_sims,
              asp_stream_t stream);

/* ---- K2+K4: otAspire masked Sinkhorn optimal transport -----------------------------------------
 * Replaces AllPairMaskedWasserstein.compute_distance, src/learning/facetid_models/pair_distances.py:21-92
 * (release copy examples/ex_aspire_consent_multimatch.py:118-189) INCLUDING the third-party solver it
 * calls, geomloss==0.2.4 SamplesLoss("sinkhorn", p=1, debias=False) (pair_distances.py:68-72,88-91):
 * marginals = softmax over sentences of the best match / temp; log-domain epsilon-scaling Sinkhorn with
 * symmetric averaged updates over the schedule eps[0..n_eps) and one final un-averaged extrapolation at
 * eps[n_eps-1].
 * eps_host: HOST array of n_eps floats (the caller derives it in float64 from the diameter exactly as
 * geomloss does -- see aspire_b200.distances.epsilon_schedule).
 * Outputs (all optional, device):
 *   dual [B]     = <alpha,f>+<beta,g>                       (return_pair_sims=False branch, :87-92)
 *   primal [B]   = sum_ij P_ij * (-C_ij)                    (return_pair_sims=True branch, :61-86)
 *   f [B,Sq], g [B,Sc], alpha [B,Sq], beta [B,Sc]          (0 on padding)
 *   neg_cost [B,Sq,Sc] = -C inside the valid block, 0 outside (:66)
 *   plan [B,Sq,Sc], weighted [B,Sq,Sc] = plan * neg_cost   (0 on padding)
 * cost_workspace: caller-owned scratch of B*Sq*Sc floats (device); holds the cost tensor on return.
 */
typedef struct asp_ot_outputs {
    float* dual;
    float* primal;
    float* f;
    float* g;
    float* alpha;
    float* beta;
    float* neg_cost;
    float* plan;
    float* weighted;
} asp_ot_outputs;

int asp_ot_sinkhorn(const float* q, const int32_t* q_lens, int q_broadcast, const float* c,
                    const int32_t* c_lens, int B, int Sq, int Sc, int D, const float* eps_host, int n_eps,
                    float temp, float* cost_workspace, const asp_ot_outputs* out, asp_stream_t stream);

/* Same solver on a precomputed cost tensor [B,Sq,Sc] (as written by asp_pair_cost). */
int asp_ot_sinkhorn_from_cost(const float* cost, const int32_t* q_lens, int q_broadcast, const int32_t* c_lens,
                              int B, int Sq, int Sc, const float* eps_host, int n_eps, float temp,
                              const asp_ot_outputs* out, asp_stream_t stream);

/* ---- geomloss `max_diameter`: bounding-box diagonal of all rows of x and y ----------------------
 * Replaces geomloss.sinkhorn_divergence.max_diameter as reached from pair_distances.py:68-72 (the
 * schedule's first epsilon).  x [nx,D], y [ny,D] (pad rows included, as geomloss sees them);
 * workspace: >= 2*D floats (device); diameter_out: 1 float (device).
 */
int asp_bbox_diameter(const float* x, long long nx, const float* y, long long ny, int D, float* workspace,
                      float* diameter_out, asp_stream_t stream);

/* ---- K5: per-query top-k over a score matrix ----------------------------------------------------
 * Replaces the Python `sorted(..., reverse=True)` at src/evaluation/evaluate.py:76 and
 * src/pre_process/pp_gen_nearest.py:339 for the head of the ranking.
 * scores [Q,N] fp32 (higher = better); ids returned are base_id + column.  Order: score descending, ties
 * by ascending id (stable like Python's sort; shard-count invariant).
 * out_scores [Q,k], out_ids [Q,k] int64; if k > N the tail is filled with -inf / -1.
 */
int asp_topk(const float* scores, int Q, long long N, int k, long long base_id, float* out_scores,
             long long* out_ids, asp_stream_t stream);

/* Merge R per-shard top-k lists (as gathered over NCCL) into one: in_scores/in_ids [Q,R*k] -> out [Q,k]. */
int asp_topk_merge(const float* in_scores, const long long* in_ids, int Q, int R, int k, float* out_scores,
                   long long* out_ids, asp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ASPIRE_B200_H_ */
