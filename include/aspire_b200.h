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
/* Tuning/testing knobs (process-wide; the defaults are the product path).
 *   "ot_kernel"     0 auto, 1 force warp-per-pair (never fuse), 2 force thread-per-pair
 *   "ot_varlen"     1 (default) one-kernel path for 11..32-sentence documents, 0 cost tensor + Sinkhorn kernels;
 *                   "vl_flags" developer bits of that kernel (1 every pair on the 32x32 variant, 2 no shape sort)
 *   "ot_fused_tc"   0 (default) the FFMA2 1 x N kernel; 1 pools of >= 128 candidates per query on the tcgen05 prototype
 *                   (ot_fused_tc.cu: same results, measured slower)
 *   "oa_warps"      Sinkhorn warps per CTA of the Q x C all-pairs otAspire kernel: 12 (default) or 8
 *   "gemm_kernel"   encoder GEMM: 3 persistent kernel, tile width picked per shape (default); 1 / 4 / 2 force 128- /
 *                   192- / 256-wide tiles; 0 one tile per CTA
 *   "gemm_cluster"  1 (default), 2 or 4 CTAs per cluster sharing W tiles by TMA multicast
 *   "gemm_pair"     CTA pairs (tcgen05 cta_group::2): -1 (default) 256-wide pair tiles for QKV, FFN1 and FFN2 at >= 16384
 *                   rows, 0 never, 1 / 2 always with 128- / 256-wide pair tiles
 *   "pdl"           1 (default) encoder kernels use programmatic dependent launch, 0 plain stream order
 *   "attn_tc"       plain-bf16 attention, L <= 256: 5 (default) pipelined tcgen05 kernel (producer + two worker groups per
 *                   SM, P in tensor memory); 3 / 4 / 1 / 2 earlier tcgen05 variants (one tile per CTA with V MN-major, + P in
 *                   tensor memory, V transposed in shared memory, first persistent kernel); 0 mma.sync.  1-5 agree bit for bit
 *   "ln_on_read"    1 (default) inner LayerNorms write the bf16 GEMM operand + per-row (mean, rstd) and the next
 *                   residual epilogue normalises the pre-LayerNorm rows as it reads them; 0 every LayerNorm writes the
 *                   fp32 residual stream (bit-identical outputs)
 *   "span_tma"      1 (default) span pooling staged by cp.async.bulk, 0 plain loads
 * Every setting computes bit-identical GEMM results (tests/test_gemm_gpu.py). */
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

/* ---- K0: tensor-core GEMM of the encoder (tcgen05 / TMEM / TMA) --------------------------------------------
 * Replaces the nn.Linear layers of HF BertModel as called from AspireConSent.consent_reps_bert
 * (examples/ex_aspire_consent.py:72).  out[M,N] = epilogue(A[M,K] . W[N,K]^T + bias[N]); A and W are bf16,
 * row-major with K contiguous (W = the PyTorch Linear weight as stored), fp32 accumulation in TMEM.
 * a_lo / w_lo: NULL -> plain bf16 operands.  Both given -> "bf16x3": operands are (hi, lo) bf16 pairs of fp32
 *   values and the kernel accumulates hi.hi + hi.lo + lo.hi (fp32-equivalent products).
 * epilogue: 0 = bf16 out (out_hi, and out_lo = bf16 of the rounding residual when non-NULL)
 *           1 = exact-erf GELU, then as 0        2 = + residual[M,N] (fp32) -> out_f32        3 = fp32 -> out_f32
 * Requirements: N % 128 == 0, K % 64 == 0, 16-byte aligned pointers.
 */
int asp_gemm_bf16_tn(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                     const float* residual, int M, int N, int K, int epilogue, void* out_hi, void* out_lo,
                     float* out_f32, asp_stream_t stream);

/* ---- K0: BERT-base encoder forward ---------------------------------------------------------------------
 * Replaces `self.bert_encoder(tokid_tt, token_type_ids=seg_tt, attention_mask=attnmask_tt)` of
 * AspireConSent.consent_reps_bert (examples/ex_aspire_consent.py:72; = disent_models.py:505) and yields its
 * `last_hidden_state`.  Architecture = HF BertModel (post-LN, exact-erf GELU, LayerNorm eps from the config).
 * All weight pointers are DEVICE pointers; the structs themselves live in HOST memory.
 *   *_hi : bf16 [out,in] = the nn.Linear weight rounded to bf16;  *_lo : bf16 of (weight - hi) or NULL.
 *   wqkv = rows of query | key | value weights stacked ([3*hidden, hidden]); bqkv likewise.
 */
typedef struct asp_bert_layer {
    const void *wqkv_hi, *wqkv_lo; const float* bqkv;
    const void *wo_hi, *wo_lo;     const float* bo;
    const float *ln1_g, *ln1_b;
    const void *w1_hi, *w1_lo;     const float* b1;
    const void *w2_hi, *w2_lo;     const float* b2;
    const float *ln2_g, *ln2_b;
} asp_bert_layer;

typedef struct asp_bert_weights {
    int hidden, heads, layers, intermediate, vocab, max_pos;
    float ln_eps;
    const float *word_emb, *pos_emb, *type_emb;  /* fp32 [vocab,hidden], [max_pos,hidden], [2,hidden] */
    const float *emb_ln_g, *emb_ln_b;
    const asp_bert_layer* layer;                 /* host array of `layers` entries */
} asp_bert_weights;

/* Bytes of device scratch asp_bert_forward needs for a [B,L] batch. */
size_t asp_bert_workspace_bytes(const asp_bert_weights* w, int B, int L, int precise);
/* ids / type_ids (may be NULL = all 0): int32 [B,L] right-padded; seq_lens int32 [B] = number of real tokens
 * (attention mask = 1 on the first seq_lens[b] positions, as prepare_bert_sentences builds it, :169-173).
 * precise = 0: bf16 tensor-core operands; 1: bf16x3 split operands (fp32-equivalent, needs the *_lo weights).
 * hidden_out: fp32 [B,L,hidden] (every position, pads included, like HF). */
int asp_bert_forward(const asp_bert_weights* w, const int32_t* ids, const int32_t* type_ids, const int32_t* seq_lens,
                     int B, int L, int precise, float* hidden_out, void* workspace, size_t workspace_bytes,
                     asp_stream_t stream);

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
/* Same with caller-owned device scratch of asp_l2max_workspace_bytes(B,Sq,Sc,D) bytes (may be 0 / NULL): long documents
 * (11..32 sentences, one-kernel path) use it to claim the pairs sorted by shape -- same results, ~2x faster on large,
 * ragged batches.  asp_l2max is this call without scratch. */
size_t asp_l2max_workspace_bytes(int B, int Sq, int Sc, int D);
int asp_l2max_ws(const float* q, const int32_t* q_lens, int q_broadcast, const float* c, const int32_t* c_lens,
                 int B, int Sq, int Sc, int D, float* best, int32_t* flat_idx, float* pair_sims, void* workspace,
                 size_t workspace_bytes, asp_stream_t stream);

/* ---- K2+K3 ALL PAIRS: tsAspire scores of every query document against every candidate document -------------
 * Replaces the numpy ranking path src/pre_process/pp_gen_nearest.py:939-961 (`-cdist(query_sents, pool_sents)`
 * then a per-candidate column-slice np.max; all-queries x all-corpus variant :788-816) and
 * allpair_masked_dist_l2max (pair_distances.py:138-186) applied to all NQ x NC pairs: one tcgen05 contraction
 * [NQ*S, D] x [D, NC*S] with fp32-equivalent (bf16x3) products and a segmented max epilogue.
 * q [NQ,S,D], c [NC,S,D] zero padded fp32, lens int32; S in [1,64], D % 64 == 0 (at most 16 candidate documents share a tile, so S < 10 uses part of it).
 * scores [NQ,NC] = max_{i<ql,j<cl} -||q_i-c_j|| (-1e9 if a side is empty); flat_idx [NQ,NC] (optional) = i*S+j
 * of the first maximum.  workspace: asp_l2max_allpairs_workspace_bytes() of device scratch (bf16 hi/lo copies).
 */
size_t asp_l2max_allpairs_workspace_bytes(int NQ, int NC, int S, int D);
int asp_l2max_allpairs(const float* q, const int32_t* q_lens, int NQ, const float* c, const int32_t* c_lens, int NC,
                       int S, int D, float* scores, int32_t* flat_idx, void* workspace, size_t workspace_bytes,
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

/* ---- K2+K4 FUSED: otAspire scores straight from the sentence representations (headline entry point) ----
 * Same arithmetic and outputs as asp_ot_sinkhorn, replacing pair_distances.py:21-92 (+ geomloss) in one
 * launch; the cost tensor never touches HBM for documents of up to 32 sentences (<= 10: ot_fused.cu, no workspace;
 * 11..32: ot_varlen.cu, whose workspace -- B int32 + a few counters, only for B >= 4096 -- holds the order in which
 * pairs are claimed, sorted by shape; without it the call still works, unsorted).  Longer documents (<= 128) go through
 * a cost tensor of B*Sq*Sc floats in the workspace.
 * q_group: number of CONSECUTIVE candidates that share one query: pair b uses query b / q_group
 *   (q [ceil(B/q_group),Sq,D], q_lens [ceil(B/q_group)]).  q_group = 1 is the reference's paired call
 *   (compute_distance, pair_distances.py:46); q_group = B is caching_score's "one query replicated B times"
 *   (disent_models.py:274-281) without the replication; anything between scores several query pools at once.
 * workspace: caller-owned device scratch of at least asp_ot_score_workspace_bytes(B,Sq,Sc,D) bytes (may be
 *   NULL when that is 0).
 */
size_t asp_ot_score_workspace_bytes(int B, int Sq, int Sc, int D);
int asp_ot_score(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens,
                 int B, int Sq, int Sc, int D, const float* eps_host, int n_eps, float temp,
                 const asp_ot_outputs* out, void* workspace, size_t workspace_bytes, asp_stream_t stream);

/*
 * Other aggregation heads on a distance tensor [B,Sq,Sc] written by asp_pair_cost (valid block only is read):
 *   top2      [B]        sum of the two largest (-dist) -- allpair_masked_dist_l2topk,
 *                        src/learning/facetid_models/pair_distances.py:295-345 (a missing runner-up is the reference's
 *                        mask constant -1e9);
 *   att       [B]        sum softmax(-dist/temp) * (-dist) over the valid block -- AllPairMaskedAttention,
 *                        pair_distances.py:95-135 + models_common/activations.py:35-61;
 *   att_probs [B,Sq,Sc]  that softmax (0 on padding).
 * Any of the three outputs may be NULL.  q_group = consecutive candidates sharing one entry of q_lens (1 = paired).
 */
int asp_pair_heads(const float* cost, const int32_t* q_lens, int q_group, const int32_t* c_lens, int B, int Sq, int Sc,
                   float temp, float* top2, float* att, float* att_probs, asp_stream_t stream);

/*
 * Score mixing of WordSentAlignBiEnc.caching_score (src/learning/facetid_models/disent_models.py:298-307), in place:
 *   scores[b] = sent_prop * scores[b] + abs_prop * ( -|| q_cls[b / q_group] - c_cls[b] + 1e-6 ||_2 )
 * (the second term is -torch.nn.functional.pairwise_distance(query_cls, cand_cls, p=2)).  scores [B] are the sentence-level
 * scores of asp_ot_score / asp_l2max / asp_pair_heads; q_cls [ceil(B/q_group), D], c_cls [B, D] the documents' CLS vectors.
 */
int asp_mix_cls_scores(float* scores, const float* q_cls, int q_group, const float* c_cls, int B, int D,
                       float sent_prop, float abs_prop, asp_stream_t stream);

/*
 * Q x C all-pairs mode of asp_ot_score (dual values only): scores[i*NC + j] = OT_eps(query i, candidate j) for every
 * query document i < NQ (q [NQ,Sq,D], q_lens [NQ]) and candidate document j < NC (c [NC,Sc,D], c_lens [NC]) -- the
 * all-queries x whole-corpus use of compute_distance (pair_distances.py:21-92 reached from pp_gen_nearest.py:131-204
 * for every query of a pool file).  Documents of <= 10 sentences (D % 64 == 0, NQ >= 2, NQ*NC < 2^31) run ONE launch of
 * the tcgen05 all-pairs kernel: the Gram matrices of 12 query x 16 candidate documents per tile on the tensor cores
 * (fp32-equivalent bf16 hi/lo split operands, fp32 accumulators in TMEM), the Sinkhorn solves on the MUFU pipe; it needs
 * asp_ot_score_allpairs_workspace_bytes() of scratch (bf16 hi/lo copies of q and c + their squared norms).  Other shapes,
 * or a smaller workspace (>= asp_ot_score_workspace_bytes with B = NC), run one 1 x N launch per query.
 */
size_t asp_ot_score_allpairs_workspace_bytes(int NQ, int NC, int Sq, int Sc, int D);
int asp_ot_score_allpairs(const float* q, const int32_t* q_lens, int NQ, const float* c, const int32_t* c_lens, int NC,
                          int Sq, int Sc, int D, const float* eps_host, int n_eps, float temp, float* scores,
                          void* workspace, size_t workspace_bytes, asp_stream_t stream);

/*
 * asp_ot_score with the candidates given as an index list into a corpus that stays in device memory: pair b scores
 * query b / q_group against corpus document c_index[b]; c [N,Sc,D] and c_lens [N] describe the WHOLE corpus, c_index
 * is int32 [B] on the device.  Nothing is gathered: the kernel walks the list.  This is how src/evaluation/evaluate.py:
 * 62-74 and pp_gen_nearest.py:154-202 present their pools (candidate id lists into one encodings cache).  Fused shapes
 * only (Sq, Sc <= 10, D % 128 == 0, D <= 768; ASP_ERR_UNSUPPORTED otherwise -- gather and call asp_ot_score).
 */
int asp_ot_score_indexed(const float* q, const int32_t* q_lens, int q_group, const float* c, const int32_t* c_lens,
                         const int32_t* c_index, int B, int Sq, int Sc, int D, const float* eps_host, int n_eps, float temp,
                         const asp_ot_outputs* out, asp_stream_t stream);

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

/* Single-pass top-k (k <= 128, ids < 2^32): the score matrix is read from HBM once (one CTA per 8192-score chunk keeps
 * its scores in registers, selects its own top-k and a second kernel merges the chunk lists of a query).  Same order
 * as asp_topk.  negate != 0 ranks by -scores (the OT distances of asp_ot_score: smaller = better) and returns the
 * negated values, so no separate negation pass is needed (pp_gen_nearest.py:194-202 sorts -distance).
 * out_packed [Q,k] (optional, uint64): the entries as sortable keys (order-preserving score bits << 32 | ~id; 0 =
 * filler) -- the ONE tensor a candidate-sharded job all-gathers; asp_topk_merge_packed merges the gathered
 * [R][Q][k] lists of R ranks in that layout (no concatenation) into out_scores / out_ids [Q,k].
 * workspace: asp_topk_workspace_bytes(Q,N,k) bytes of device scratch; when that is 0 (k > 128 or more than 8192/k
 * chunks) or no workspace is given the call falls back to asp_topk (negate / out_packed are then rejected). */
size_t asp_topk_workspace_bytes(int Q, long long N, int k);
int asp_topk_ws(const float* scores, int Q, long long N, int k, long long base_id, int negate, float* out_scores,
                long long* out_ids, unsigned long long* out_packed, void* workspace, size_t workspace_bytes,
                asp_stream_t stream);
int asp_topk_merge_packed(const unsigned long long* gathered, int R, int Q, int k, float* out_scores, long long* out_ids,
                          asp_stream_t stream);

/* ---- host front end of the encoder: BERT word pieces + sequence assembly (no GPU involved) -------------------
 * Replaces tokenizer.tokenize + convert_tokens_to_ids per sentence (examples/ex_aspire_consent.py:131-133 =
 * src/learning/batchers.py:579-581) and the per-document concatenation / 500-piece truncation / [CLS]..[SEP] /
 * padding / span bookkeeping of examples/ex_aspire_consent.py:120-173, multi-threaded.
 *
 * asp_wordpiece_create: vocabulary = n_vocab UTF-8 tokens concatenated in vocab_blob, token id = its index,
 *   vocab_offsets[n_vocab+1]; "##x" entries are continuation pieces; special_ids name the tokens ([SEP], [MASK] ...)
 *   that are matched verbatim in the raw text before normalisation.  Returns NULL on error (asp_last_error()).
 * asp_wordpiece_encode: n_sent sentences concatenated in `text` (offsets[n_sent+1], bytes).  out_ids needs room for
 *   offsets[n_sent]-offsets[0] entries; on return sentence i owns out_ids[out_offsets[i] .. out_offsets[i+1]).
 *   Text is cleaned, (optionally) lower-cased, split on whitespace and punctuation and cut into greedy longest-match
 *   word pieces; words longer than max_chars_per_word become the unknown token.  Without asp_wordpiece_set_unicode,
 *   sentences holding a byte >= 0x80 are NOT tokenised (needs_fallback[i] = 1, zero ids): the caller runs them
 *   through the original tokenizer.
 */
typedef struct asp_wordpiece asp_wordpiece;
asp_wordpiece* asp_wordpiece_create(const char* vocab_blob, const int64_t* vocab_offsets, int n_vocab, int lower_case,
                                    int unk_id, const int32_t* special_ids, int n_special);
void asp_wordpiece_destroy(asp_wordpiece* wp);
/* Optional: per-code-point tables of the Basic Multilingual Plane so that non-ASCII sentences are tokenised here too.
 *   norm_offsets[65537] / norm_blob: UTF-8 text each character is normalised to (may be empty, or longer than one
 *   character); out_class[65536]: 0 word character, 1 separates words, 2 punctuation (a word of its own), applied to
 *   the normalised text; fallback[65536]: 1 = a sentence holding this character is reported in needs_fallback.
 *   Characters beyond U+FFFF and malformed UTF-8 always are. */
int asp_wordpiece_set_unicode(asp_wordpiece* wp, const uint32_t* norm_offsets, const char* norm_blob,
                              const uint8_t* out_class, const uint8_t* fallback);
int asp_wordpiece_encode(const asp_wordpiece* wp, const char* text, const int64_t* offsets, int n_sent,
                         int max_chars_per_word, int threads, int32_t* out_ids, int64_t* out_offsets,
                         uint8_t* needs_fallback);
/* Document d owns doc_sents[d] consecutive sentences (element 0 = the title); sentence k's ids are
 * ids[sent_offsets[k] .. sent_offsets[k+1]).  plan: seq_lens[d] (with [CLS]/[SEP]) and abs_lens[d] (kept abstract
 * sentences) under the `budget`-piece truncation rule.  fill: tokid/seg/attn int64 [n_docs,width] (pad_id outside the
 * sequence, like the reference) and spans int32 [n_docs,max_sents,2] half-open token ranges, (-1,-1) = no sentence. */
int asp_abstracts_plan(const int64_t* sent_offsets, const int32_t* doc_sents, int n_docs, int budget, int32_t* seq_lens,
                       int32_t* abs_lens);
int asp_abstracts_fill(const int32_t* ids, const int64_t* sent_offsets, const int32_t* doc_sents, int n_docs, int budget,
                       int cls_id, int sep_id, int64_t pad_id, int width, int max_sents, int64_t* tokid, int64_t* seg,
                       int64_t* attn, int32_t* spans);

/* ---- host side of a pool: pad n per-paper encodings into one staging buffer (no GPU involved) ---------------
 * Replaces the per-candidate padding loop of caching_score (src/learning/facetid_models/disent_models.py:274-290).
 * srcs[j]: lens[j] rows of D floats, contiguous, host memory; dst: [n, smax, D] floats, rows lens[j]..smax zeroed. */
int asp_pack_pool(const float* const* srcs, const int32_t* lens, int n, int smax, int D, float* dst, int threads);

#ifdef __cplusplus
}
#endif
#endif /* ASPIRE_B200_H_ */
