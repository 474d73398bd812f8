"""Seeded synthetic HF-named packed checkpoint shared by the golden generator and the CPU test."""
import hashlib

import torch


def make_hf_checkpoint(n_layer=32, dim=64, kv=32, inter=96, vocab=50, planes=4):
    g = torch.Generator().manual_seed(2024)
    sd = {}

    def q(n, k):
        return torch.randint(-2**31, 2**31 - 1, (planes, n, k // 32), dtype=torch.int32, generator=g)

    sd["model.embed_tokens.weight"] = torch.randn((vocab, dim), generator=g).bfloat16()
    for i in range(n_layer):
        p = f"model.layers.{i}."
        for name, (n, k) in {"self_attn.q_proj": (dim, dim), "self_attn.k_proj": (kv, dim), "self_attn.v_proj": (kv, dim),
                             "self_attn.o_proj": (dim, dim), "mlp.gate_proj": (inter, dim), "mlp.up_proj": (inter, dim),
                             "mlp.down_proj": (dim, inter)}.items():
            sd[p + name + ".qweight"] = q(n, k)
            for b in (2, 3, 4):
                sd[p + name + f".lut{b}"] = torch.randn((n, 2 ** b), generator=g).half()  # pack.py:172 writes np.float16
        sd[p + "input_layernorm.weight"] = torch.randn(dim, generator=g).bfloat16()
        sd[p + "post_attention_layernorm.weight"] = torch.randn(dim, generator=g).bfloat16()
    sd["model.norm.weight"] = torch.randn(dim, generator=g).bfloat16()
    sd["lm_head.weight"] = torch.randn((vocab, dim), generator=g).bfloat16()
    return sd


def digest(t: torch.Tensor):
    raw = t.contiguous().view(torch.uint8).numpy().tobytes() if t.dtype != torch.bfloat16 else t.float().contiguous().view(torch.uint8).numpy().tobytes()
    return [list(t.shape), str(t.dtype), hashlib.sha1(raw).hexdigest()]
