"""GPU parity of the sm_100a BERT encoder (asp_bert_forward) and of AspireConSent.forward built on it.

Reference = the plain PyTorch fp32 forward of the same HF ``BertModel`` (what the reference calls at
examples/ex_aspire_consent.py:72), weights seeded (no checkpoints exist offline).
Tolerances on last_hidden_state (post-LayerNorm values, O(1)):
  bf16x3 (default, fp32-equivalent split operands): max |err| <= 2e-4
  bf16   (plain bf16 tensor-core operands)        : relative L2 error <= 3e-2, cosine >= 0.999
and on the end-to-end README example against the golden output of the UNMODIFIED reference
(tests/golden/readme_encoder_2layer.npz): sentence reps <= 2e-4, tsAspire score <= 1e-3 absolute.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_shims

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _hf_reference(model, ids, lens):
    torch.backends.cuda.matmul.allow_tf32 = False
    model = model.cuda().float()
    B, L = ids.shape
    mask = (torch.arange(L)[None, :] < torch.as_tensor(lens)[:, None]).long().cuda()
    with torch.no_grad():
        out = model(ids.cuda(), token_type_ids=torch.zeros_like(ids).cuda(), attention_mask=mask).last_hidden_state
    return out.float()


@pytest.mark.parametrize("layers,B,L,lens", [(2, 3, 70, [70, 33, 5]), (12, 2, 300, [300, 121]), (12, 5, 129, [129, 64, 65, 2, 128])])
def test_encoder_matches_hf_fp32(layers, B, L, lens):
    from aspire_b200.encoder import B200BertEncoder
    model = ref_shims.seeded_bert(seed=3, num_hidden_layers=layers)
    g = torch.Generator().manual_seed(layers * 100 + L)
    ids = torch.randint(1000, 31000, (B, L), generator=g)
    for b, n in enumerate(lens):
        ids[b, n:] = 0
    ref = _hf_reference(model, ids, lens)
    enc = B200BertEncoder(model)
    valid = (torch.arange(L)[None, :] < torch.tensor(lens)[:, None]).cuda()
    got = enc.forward(ids, lens, precision="bf16x3")
    torch.cuda.synchronize()
    err = (got - ref).abs()
    assert torch.isfinite(got).all()
    assert err[valid].max().item() <= 2e-4, f"bf16x3 max err (valid tokens) {err[valid].max().item():.3e}"
    assert err.max().item() <= 2e-4, f"bf16x3 max err (pad positions too) {err.max().item():.3e}"
    fast = enc.forward(ids, lens, precision="bf16")
    torch.cuda.synchronize()
    rel = ((fast - ref)[valid].norm() / ref[valid].norm()).item()
    cos = torch.nn.functional.cosine_similarity(fast[valid], ref[valid], dim=-1).min().item()
    assert rel <= 3e-2 and cos >= 0.999, f"bf16 rel L2 {rel:.3e}, min cosine {cos:.5f}"


def test_readme_example_end_to_end_vs_reference_golden():
    """BASELINE config 1 on the GPU: prepare_abstracts -> AspireConSent.forward -> tsAspire score."""
    import transformers
    from oracle.make_golden import README_ABSTRACTS
    from aspire_b200 import allpair_masked_dist_l2max, rep_len_tup
    from aspire_b200.consent import AspireConSent, prepare_abstracts
    z = np.load(os.path.join(GOLDEN, "readme_encoder_2layer.npz"))
    orig = transformers.AutoModel.from_pretrained
    transformers.AutoModel.from_pretrained = staticmethod(lambda name, *a, **k: ref_shims.seeded_bert(0, num_hidden_layers=2))
    try:
        model = AspireConSent("allenai/aspire-contextualsentence-singlem-compsci")
    finally:
        transformers.AutoModel.from_pretrained = orig
    bb, al, sti = prepare_abstracts(batch_abs=README_ABSTRACTS, pt_lm_tokenizer=ref_shims.ToyTokenizer())
    with torch.no_grad():
        cls, reps = model.forward(bert_batch=bb, abs_lens=al, sent_tok_idxs=sti)
    assert not reps.is_cuda and tuple(reps.shape) == tuple(z["reps"].shape) and tuple(cls.shape) == tuple(z["cls"].shape)
    assert al == z["abs_lens"].tolist()
    assert np.abs(reps.numpy() - z["reps"]).max() <= 2e-4
    assert np.abs(cls.numpy() - z["cls"]).max() <= 2e-4
    assert (reps[0, al[0]:] == 0).all() and (reps[1, al[1]:] == 0).all()
    qt = rep_len_tup(embed=reps[0:1].permute(0, 2, 1), abs_lens=[al[0]])
    ct = rep_len_tup(embed=reps[1:2].permute(0, 2, 1), abs_lens=[al[1]])
    score, _ = allpair_masked_dist_l2max(query=qt, cand=ct, return_pair_sims=True)
    assert abs(score.item() - float(z["ts_score"].reshape(-1)[0])) <= 1e-3


def test_entity_augmented_models_encode_vs_reference_golden():
    """AspireNER (entities appended as sentences) and AspireContextNER (entity spans pooled by the span kernel over the
    same encoder forward) against the unmodified reference's encodings (utils/models.py:211-233, 607-734)."""
    import json
    from test_host_api import _ner_model
    with open(os.path.join(GOLDEN, "ner.json")) as fh:
        papers = json.load(fh)["papers"]
    z = np.load(os.path.join(GOLDEN, "ner.npz"))
    for name, key in (("aspire_context_ner_compsci", "ctx"), ("aspire_ner_compsci", "ner")):
        m = _ner_model(name)
        reps = m.encode(papers)
        assert len(reps) == len(papers)
        for i, r in enumerate(reps):
            want = z[f"{key}_{i}"]
            assert tuple(r.shape) == tuple(want.shape) and not r.is_cuda
            assert np.abs(r.numpy() - want).max() <= 2e-4, (name, i)
    # scoring an entity-augmented pair goes through the usual otAspire path
    s = m.get_similarity(reps[0], reps[1])
    assert np.isfinite(s) and s < 0


def test_encode_stream_matches_batchwise_encode():
    """Overlapped host preparation (span-table front end on worker threads) gives the same encodings as the
    reference-shaped prepare_abstracts -> forward path, in input order."""
    import transformers
    from oracle.make_golden import README_ABSTRACTS
    from aspire_b200.consent import AspireConSent, prepare_abstracts
    from aspire_b200.evaluation import encode_stream
    orig = transformers.AutoModel.from_pretrained
    transformers.AutoModel.from_pretrained = staticmethod(lambda name, *a, **k: ref_shims.seeded_bert(0, num_hidden_layers=2))
    try:
        model = AspireConSent("allenai/aspire-contextualsentence-singlem-compsci")
    finally:
        transformers.AutoModel.from_pretrained = orig
    tok = ref_shims.ToyTokenizer()
    papers = [README_ABSTRACTS[i % 2] for i in range(7)]
    got = list(encode_stream(model, tok, papers, batch_size=3, workers=2))
    assert len(got) == 7
    for s in range(0, 7, 3):
        bb, al, sti = prepare_abstracts(batch_abs=papers[s:s + 3], pt_lm_tokenizer=tok)
        with torch.no_grad():
            _, reps = model.forward(bert_batch=bb, abs_lens=al, sent_tok_idxs=sti)
        for i, n in enumerate(al):
            assert torch.equal(got[s + i], reps[i, :n])


@pytest.mark.parametrize("B,L,lens", [(3, 256, [256, 130, 7]), (4, 129, [129, 64, 65, 2]), (2, 70, [70, 33]), (2, 200, [1, 200])])
def test_tcgen05_attention_equals_mma_sync_attention(B, L, lens):
    """attention_tc.cu (plain bf16, L <= 256: Q K^T and P V on tcgen05, V transposed in shared memory, softmax straight
    from TMEM) against attention.cu (mma.sync fragments, online softmax) inside the same 2-layer encoder: both round Q, K,
    V and P to bf16 and accumulate in fp32, so the hidden states agree to bf16 round-off -- on pad positions too -- and
    both stay within the bf16 tolerance of the fp32 Hugging Face forward."""
    from aspire_b200 import _abi
    from aspire_b200.encoder import B200BertEncoder
    model = ref_shims.seeded_bert(seed=5, num_hidden_layers=2)
    g = torch.Generator().manual_seed(L)
    ids = torch.randint(1000, 31000, (B, L), generator=g)
    for b, n in enumerate(lens):
        ids[b, n:] = 0
    enc = B200BertEncoder(model)
    out = {}
    # 5: pipelined kernel (one CTA per SM, producer warp + two worker groups, operands one tile ahead),
    # 4: as 3 with P in tensor memory instead of shared memory,
    # 3: one tile per CTA with V as an MN-major operand (no V^T pass), 2: persistent warp-specialised kernel,
    # 1: one tile per CTA, V transposed in shared memory, 0: mma.sync
    for mode in (5, 4, 3, 2, 1, 0):
        _abi.set_option("attn_tc", mode)
        try:
            out[mode] = enc.forward(ids, lens, precision="bf16").clone()
            torch.cuda.synchronize()
        finally:
            _abi.set_option("attn_tc", 5)
    ref = _hf_reference(model, ids, lens)
    valid = (torch.arange(L)[None, :] < torch.tensor(lens)[:, None]).cuda()
    for mode in (5, 4, 3, 2, 1):
        assert torch.isfinite(out[mode]).all()
        rel = ((out[mode] - out[0]).norm() / out[0].norm()).item()
        assert rel <= 6e-3, f"tcgen05 (mode {mode}) vs mma.sync attention: relative L2 {rel:.3e}"
        rel_ref = ((out[mode] - ref)[valid].norm() / ref[valid].norm()).item()
        assert rel_ref <= 3e-2, f"tcgen05 attention (mode {mode}) vs HF fp32: relative L2 {rel_ref:.3e}"
    assert torch.equal(out[2], out[1])   # same arithmetic in the same order: the tcgen05 kernels agree bit for bit
    assert torch.equal(out[3], out[1])
    assert torch.equal(out[4], out[1])   # P kept in tensor memory (A operand of P V read from TMEM)
    assert torch.equal(out[5], out[1])   # pipelined kernel: producer warp + two worker groups


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_layernorm_on_read_is_bit_identical(precision):
    """asp_set_option("ln_on_read"): with 1 (default) the inner LayerNorms write only the bf16 GEMM operand and per-row
    (mean, rstd) and the next residual epilogue normalises the pre-LayerNorm rows as it reads them; with 0 every LayerNorm
    writes the fp32 residual stream.  Same expression, same operands: the hidden states agree bit for bit (3 layers, so
    both the layer-0 path -- residual = embedding LayerNorm output -- and the chained path are exercised; ragged lengths,
    a row count that is not a multiple of the 128-row tile)."""
    from aspire_b200 import _abi
    from aspire_b200.encoder import B200BertEncoder
    model = ref_shims.seeded_bert(seed=11, num_hidden_layers=3)
    B, L, lens = 5, 77, [77, 50, 1, 33, 64]
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(1000, 31000, (B, L), generator=g)
    for b, n in enumerate(lens):
        ids[b, n:] = 0
    enc = B200BertEncoder(model)
    out = {}
    for mode in (0, 1):
        _abi.set_option("ln_on_read", mode)
        try:
            out[mode] = enc.forward(ids, lens, precision=precision).clone()
            torch.cuda.synchronize()
        finally:
            _abi.set_option("ln_on_read", 1)
    assert torch.isfinite(out[1]).all()
    assert torch.equal(out[0], out[1])
    # the other GEMM kernels carry the same epilogue: one tile per CTA (row-per-lane stores) and CTA pairs
    for key, val in (("gemm_kernel", 0), ("gemm_pair", 1)):
        _abi.set_option(key, val)
        try:
            alt = enc.forward(ids, lens, precision=precision).clone()
            torch.cuda.synchronize()
        finally:
            _abi.set_option(key, 3 if key == "gemm_kernel" else -1)
        assert torch.equal(alt, out[1]), f"{key}={val} with LayerNorm on read differs"


@pytest.mark.parametrize("B", [3, 30])
def test_tcgen05_attention_exact_softmax_path(B):
    """The tcgen05 attention kernels read S from tensor memory once and shift by the maximum over a SUBSET of the keys; a tile
    with an exponent more than 100 above that shift redoes the softmax in the exact two-pass form (the kernels that keep P in
    tensor memory have to run Q K^T again for that).  Query / key projections scaled by 16 make the scores span hundreds of
    powers of two, so nearly every tile takes that path: all tcgen05 variants still agree bit for bit, stay finite, and stay
    close to the mma.sync kernel's online softmax."""
    from aspire_b200 import _abi
    from aspire_b200.encoder import B200BertEncoder
    model = ref_shims.seeded_bert(seed=7, num_hidden_layers=2)
    with torch.no_grad():
        for layer in model.encoder.layer:
            layer.attention.self.query.weight.mul_(16.0)
            layer.attention.self.key.weight.mul_(16.0)
    L = 256
    lens = ([256, 200, 90] * 10)[:B]   # B = 30: ~5 tiles per CTA of the pipelined kernel, most of them redone
    g = torch.Generator().manual_seed(17)
    ids = torch.randint(1000, 31000, (B, L), generator=g)
    for b, n in enumerate(lens):
        ids[b, n:] = 0
    enc = B200BertEncoder(model)
    out = {}
    for mode in (5, 4, 3, 2, 1, 0):
        _abi.set_option("attn_tc", mode)
        try:
            out[mode] = enc.forward(ids, lens, precision="bf16").clone()
            torch.cuda.synchronize()
        finally:
            _abi.set_option("attn_tc", 5)
    valid = (torch.arange(L)[None, :] < torch.tensor(lens)[:, None]).cuda()
    for mode in (5, 4, 3, 2):
        assert torch.isfinite(out[mode]).all()
        assert torch.equal(out[mode], out[1]), f"attn_tc={mode} differs from attn_tc=1 on the exact-softmax path"
    rel = ((out[1] - out[0])[valid].norm() / out[0][valid].norm()).item()
    assert rel <= 5e-2, f"tcgen05 vs mma.sync attention on peaked softmax rows: relative L2 {rel:.3e}"


@pytest.mark.parametrize("B,L", [(40, 256), (30, 100), (26, 131)])
def test_pipelined_attention_many_tiles_per_cta(B, L):
    """The pipelined attention kernel walks several tiles per CTA (operand slots, TMEM buffers and barrier phases are reused):
    960 / 360 / 624 tiles on 148 SMs, ragged lengths incl. 1 and L, one and two query blocks per document.  Bit-identical
    to the one-tile-per-CTA kernels, within bf16 round-off of the mma.sync kernel."""
    from aspire_b200 import _abi
    from aspire_b200.encoder import B200BertEncoder
    model = ref_shims.seeded_bert(seed=23, num_hidden_layers=2)
    g = torch.Generator().manual_seed(B * 1000 + L)
    lens = torch.randint(1, L + 1, (B,), generator=g).tolist()
    lens[0], lens[1] = L, 1
    ids = torch.randint(1000, 31000, (B, L), generator=g)
    for b, n in enumerate(lens):
        ids[b, n:] = 0
    enc = B200BertEncoder(model)
    out = {}
    for mode in (5, 3, 1, 0):
        _abi.set_option("attn_tc", mode)
        try:
            out[mode] = enc.forward(ids, lens, precision="bf16").clone()
            torch.cuda.synchronize()
        finally:
            _abi.set_option("attn_tc", 5)
    assert torch.isfinite(out[5]).all()
    assert torch.equal(out[5], out[1]) and torch.equal(out[3], out[1])
    rel = ((out[5] - out[0]).norm() / out[0].norm()).item()
    assert rel <= 6e-3, f"pipelined tcgen05 vs mma.sync attention: relative L2 {rel:.3e}"
