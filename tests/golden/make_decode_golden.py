"""Generates tests/golden/decode_golden.npz by running the reference's OWN gpt-fast model
(/root/reference/inference/model.py: Transformer) on CPU in float32 with dense nn.Linear layers, one token at a time.
Build-container only.  Shims (test harness only): ROPE_INIT_FUNCTIONS["default"] (missing in
transformers 5.5; the default init is inv_freq = 1/base^(arange(0,dim,2)/dim), scaling 1) and an `ap_gemv` module so
that `from plugin import *` in model.py imports.  Usage: python tests/golden/make_decode_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("REF_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import guidedquant_b200  # noqa: E402

guidedquant_b200.install_as_ap_gemv()
from transformers import modeling_rope_utils as mru  # noqa: E402

if "default" not in mru.ROPE_INIT_FUNCTIONS:
    def _default_rope(config=None, device=None, seq_len=None, **kw):
        dim, base = kw["dim"], kw["base"]
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
        return inv_freq, 1.0
    mru.ROPE_INIT_FUNCTIONS["default"] = _default_rope
sys.path.insert(0, os.path.join(REF, "inference"))
import model as refmodel  # noqa: E402  (the reference's inference/model.py)

torch.manual_seed(0)
cfg = refmodel.ModelArgs(block_size=64, vocab_size=256, n_layer=2, n_head=2, dim=256, intermediate_size=512,
                         n_local_heads=1, rope_base=500000, model_name="llama-tiny")
m = refmodel.Transformer(torch.float32, cfg, linear_class=torch.nn.Linear)
with torch.no_grad():
    for name, p in m.named_parameters():
        if "norm" in name:
            p.copy_((1.0 + 0.1 * torch.randn_like(p)).half().float())
        elif "tok_embeddings" in name:
            p.copy_(torch.randn_like(p).half().float())
        else:
            p.copy_((torch.randn_like(p) / (p.shape[1] ** 0.5)).half().float())
m.eval()
with torch.device("cpu"):
    m.setup_caches(max_batch_size=1, max_seq_length=32)
tokens = [1, 17, 200, 3, 99, 42]
logits = []
with torch.no_grad():
    for pos, tok in enumerate(tokens):
        out = m(torch.tensor([[tok]]), torch.tensor([pos]))
        logits.append(out[0, -1].float().numpy().copy())
out = {"tokens": np.array(tokens), "logits": np.stack(logits),
       "meta": np.array([cfg.n_layer, cfg.n_head, cfg.n_local_heads, cfg.dim, cfg.intermediate_size, cfg.vocab_size, 32])}
for k, v in m.state_dict().items():
    if v.dtype.is_floating_point and "cache" not in k and "mask" not in k:
        out["w:" + k] = v.half().numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "decode_golden.npz"), **out)
print("wrote decode_golden.npz", {k: v.shape for k, v in out.items() if k.startswith("w:")})
