"""Developer benchmark of the sm_100a BERT encoder (not the judged bench): docs/s and TFLOP/s per precision mode,
next to the HF module run by PyTorch (cuBLAS / SDPA library kernels) on the same GPU."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from aspire_b200.encoder import B200BertEncoder


def seeded_bert(seed=0, vocab_size=31116, num_hidden_layers=12):
    """Random-init BERT-base in eval mode (no HF weights offline)."""
    from transformers import BertConfig, BertModel
    torch.manual_seed(seed)
    model = BertModel(BertConfig(vocab_size=vocab_size, num_hidden_layers=num_hidden_layers))
    model.eval()
    return model


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    quick = "--quick" in sys.argv
    for arg in sys.argv[1:]:  # developer switches: --opt key=value -> asp_set_option
        if arg.startswith("--opt="):
            from aspire_b200 import _abi
            k, v = arg[6:].split("=")
            _abi.set_option(k, int(v))
    shapes = [tuple(int(v) for v in a[8:].split(",")) for a in sys.argv[1:] if a.startswith("--shape=")]  # --shape=B,L
    precs = [a[7:] for a in sys.argv[1:] if a.startswith("--prec=")] or ["bf16", "bf16x3"]
    model = seeded_bert(seed=0, num_hidden_layers=12)
    enc = B200BertEncoder(model)
    hf32 = model.cuda().float()
    quick = quick or bool(shapes)
    for B, L in (shapes or ([(32, 256)] if quick else [(8, 256), (32, 256), (32, 512), (128, 256)])):
        ids = torch.randint(1000, 31000, (B, L)).cuda()
        lens = torch.full((B,), L, dtype=torch.int32).cuda()
        mask = torch.ones((B, L), dtype=torch.long).cuda()
        flops = B * L * (2 * 85.05e6 + 12 * 4 * L * 768)
        out = {}
        for prec in precs:
            out[prec] = timeit(lambda: enc.forward(ids, lens, precision=prec))
        if not quick:
            with torch.no_grad():
                out["hf_fp32"] = timeit(lambda: hf32(ids, attention_mask=mask).last_hidden_state, iters=3, warm=1)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    out["hf_autocast_bf16"] = timeit(lambda: hf32(ids, attention_mask=mask).last_hidden_state, iters=3, warm=1)
        print(f"B={B} L={L} ({flops / 1e12:.2f} TFLOP): " + "  ".join(
            f"{k} {v:.3f} ms ({B / v * 1e3:.0f} docs/s, {flops / v / 1e9:.0f} TFLOP/s)" for k, v in out.items()), flush=True)


if __name__ == "__main__":
    main()
