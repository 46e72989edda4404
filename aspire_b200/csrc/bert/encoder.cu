// K0: BERT-base encoder forward (the `self.bert_encoder(...)` call of AspireConSent.consent_reps_bert,
// examples/ex_aspire_consent.py:72; architecture per HF BertModel: post-LN, exact-erf GELU, LayerNorm eps 1e-12).
// Orchestrates the kernels of this directory on one stream; no allocation, no host synchronisation.
//
//   x = LN(word[id] + type[tt] + pos[p])                                         embed_ln_kernel
//   for each layer:  qkv = x Wqkv^T + b          (fused Q|K|V projection)        gemm_tn_kernel  (tcgen05)
//                    ctx = softmax(q k^T / 8 + mask) v                           attention_kernel
//                    x   = LN(ctx Wo^T + b + x)                                  gemm_tn_kernel (+residual) , ln_kernel
//                    h   = gelu(x W1^T + b)                                      gemm_tn_kernel (+GELU)
//                    x   = LN(h W2^T + b + x)                                    gemm_tn_kernel (+residual) , ln_kernel
// The residual stream x stays fp32; GEMM operands are its bf16 copies written by the LN kernels (hi only, or hi+lo in
// the fp32-equivalent bf16x3 mode).
#include "../common.cuh"

namespace asp {

int gemm_bf16_tn(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                 const float* residual, int M, int N, int K, int epilogue, void* out_hi, void* out_lo, float* out_f32,
                 cudaStream_t stream, const LnOnRead* ln = nullptr);
int embed_ln_launch(const int32_t* ids, const int32_t* type_ids, int T, int L, int H, int vocab, int max_pos,
                    const float* word_emb, const float* pos_emb, const float* type_emb, const float* gamma, const float* beta,
                    float eps, float* out_f32, void* out_hi, void* out_lo, cudaStream_t stream);
int ln_launch(const float* in, int T, int H, const float* gamma, const float* beta, float eps, float* out_f32, void* out_hi,
              void* out_lo, cudaStream_t stream, void* stats = nullptr);
int g_ln_on_read = 1;
int attention_launch(const void* qkv_hi, const void* qkv_lo, const int32_t* seq_lens, int B, int L, int H, int heads,
                     void* ctx_hi, void* ctx_lo, cudaStream_t stream);

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct EncoderWs {
    char *x_hi, *x_lo, *qkv_hi, *qkv_lo, *ctx_hi, *ctx_lo, *h_hi, *h_lo;
    float* tmp;
    float *stats_a, *stats_b;  // per-row (mean, rstd) of the two LayerNorms of a layer (LayerNorm on read)
    size_t total;
};

static EncoderWs carve(char* base, size_t T, int H, int I, bool precise) {
    EncoderWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align256(bytes); return p; };
    w.x_hi = take(T * H * 2);
    w.qkv_hi = take(T * 3 * H * 2);
    w.ctx_hi = take(T * H * 2);
    w.h_hi = take(T * I * 2);
    w.tmp = reinterpret_cast<float*>(take(T * H * 4));
    w.stats_a = reinterpret_cast<float*>(take(T * 8));
    w.stats_b = reinterpret_cast<float*>(take(T * 8));
    if (precise) {
        w.x_lo = take(T * H * 2);
        w.qkv_lo = take(T * 3 * H * 2);
        w.ctx_lo = take(T * H * 2);
        w.h_lo = take(T * I * 2);
    }
    w.total = off;
    return w;
}

}  // namespace asp

extern "C" size_t asp_bert_workspace_bytes(const asp_bert_weights* w, int B, int L, int precise) {
    if (!w || B < 0 || L < 0) return 0;
    return asp::carve(nullptr, (size_t)B * L, w->hidden, w->intermediate, precise != 0).total;
}

extern "C" int asp_bert_forward(const asp_bert_weights* w, const int32_t* ids, const int32_t* type_ids,
                                const int32_t* seq_lens, int B, int L, int precise, float* hidden_out, void* workspace,
                                size_t workspace_bytes, asp_stream_t stream_) {
    using namespace asp;
    ASP_REQUIRE(w && ids && seq_lens && hidden_out, "asp_bert_forward: NULL argument");
    ASP_REQUIRE(w->layer && w->layers >= 1, "asp_bert_forward: weights hold no layers");
    ASP_REQUIRE(B >= 0 && L >= 1 && L <= w->max_pos, "asp_bert_forward: sequence length %d outside [1, %d]", L, w->max_pos);
    if (B == 0) return ASP_OK;
    const int H = w->hidden, I = w->intermediate;
    const size_t T = (size_t)B * L;
    const bool px = precise != 0;
    ASP_REQUIRE(workspace && workspace_bytes >= carve(nullptr, T, H, I, px).total, "asp_bert_forward: workspace too small");
    ASP_REQUIRE(T < (size_t)1 << 31, "asp_bert_forward: too many tokens");
    const EncoderWs ws = carve(static_cast<char*>(workspace), T, H, I, px);
    cudaStream_t stream = (cudaStream_t)stream_;
    float* x = hidden_out;  // fp32 residual stream lives in the output buffer
    int rc = embed_ln_launch(ids, type_ids, (int)T, L, H, w->vocab, w->max_pos, w->word_emb, w->pos_emb, w->type_emb,
                             w->emb_ln_g, w->emb_ln_b, w->ln_eps, x, ws.x_hi, ws.x_lo, stream);
    if (rc) return rc;
    // Residual stream.  With LayerNorm on read (g_ln_on_read) only two fp32 tensors exist per layer -- the pre-LayerNorm sums
    // A = attn-out + residual (ws.tmp) and B = FFN2 + residual (x) -- and each LayerNorm writes just the bf16 GEMM operand
    // and its row statistics; the residual epilogue that follows reads A (or B) and normalises on the fly.  The last
    // LayerNorm writes the fp32 output in place.  Without it: x is the normalised stream, ws.tmp the pre-LayerNorm sum.
    const bool lnr = g_ln_on_read != 0;
    for (int l = 0; l < w->layers; ++l) {
        const asp_bert_layer& y = w->layer[l];
        const bool last = l + 1 == w->layers;
        if (px) ASP_REQUIRE(y.wqkv_lo && y.wo_lo && y.w1_lo && y.w2_lo, "asp_bert_forward: precise mode needs the lo weight halves");
        if ((rc = gemm_bf16_tn(ws.x_hi, ws.x_lo, y.wqkv_hi, px ? y.wqkv_lo : nullptr, y.bqkv, nullptr, (int)T, 3 * H, H, 0,
                               ws.qkv_hi, ws.qkv_lo, nullptr, stream)))
            return rc;
        if ((rc = attention_launch(ws.qkv_hi, ws.qkv_lo, seq_lens, B, L, H, w->heads, ws.ctx_hi, ws.ctx_lo, stream))) return rc;
        // x holds the embedding LayerNorm's output (l = 0) or, with lnr, the previous layer's pre-LayerNorm sum B
        const LnOnRead prev{ws.stats_b, l ? w->layer[l - 1].ln2_g : nullptr, l ? w->layer[l - 1].ln2_b : nullptr};
        if ((rc = gemm_bf16_tn(ws.ctx_hi, ws.ctx_lo, y.wo_hi, px ? y.wo_lo : nullptr, y.bo, x, (int)T, H, H, 2, nullptr, nullptr,
                               ws.tmp, stream, (lnr && l) ? &prev : nullptr)))
            return rc;
        if ((rc = ln_launch(ws.tmp, (int)T, H, y.ln1_g, y.ln1_b, w->ln_eps, lnr ? nullptr : x, ws.x_hi, ws.x_lo, stream,
                            lnr ? ws.stats_a : nullptr)))
            return rc;
        if ((rc = gemm_bf16_tn(ws.x_hi, ws.x_lo, y.w1_hi, px ? y.w1_lo : nullptr, y.b1, nullptr, (int)T, I, H, 1, ws.h_hi,
                               ws.h_lo, nullptr, stream)))
            return rc;
        const LnOnRead mid{ws.stats_a, y.ln1_g, y.ln1_b};
        if ((rc = gemm_bf16_tn(ws.h_hi, ws.h_lo, y.w2_hi, px ? y.w2_lo : nullptr, y.b2, lnr ? ws.tmp : x, (int)T, H, I, 2, nullptr,
                               nullptr, lnr ? x : ws.tmp, stream, lnr ? &mid : nullptr)))
            return rc;
        if (lnr)
            rc = ln_launch(x, (int)T, H, y.ln2_g, y.ln2_b, w->ln_eps, last ? x : nullptr, ws.x_hi, ws.x_lo, stream,
                           last ? nullptr : ws.stats_b);
        else
            rc = ln_launch(ws.tmp, (int)T, H, y.ln2_g, y.ln2_b, w->ln_eps, x, ws.x_hi, ws.x_lo, stream);
        if (rc) return rc;
    }
    return ASP_OK;
}
