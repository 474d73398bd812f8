"""Synthesise a tiny Any-Precision HF checkpoint directory in the format the reference packer writes
(any_precision/quantization/pack.py:133-203): config.json with an `anyprec` section + pytorch_model.bin holding
`<linear>.qweight [parent_bits, N, K/32]`, `<linear>.lut{b}` for b = seed..parent, and fp16 everything else.
Test infrastructure only (uses the oracle's packer)."""
import os

import numpy as np
import torch

LLAMA_MODULES = ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
                 "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"]


def tiny_llama_config(seed_precision=2, parent_precision=4, hidden=128, inter=256, layers=2, heads=4, kv_heads=2, vocab=320):
    from transformers import LlamaConfig

    cfg = LlamaConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers, num_attention_heads=heads,
                      num_key_value_heads=kv_heads, vocab_size=vocab, max_position_embeddings=256, rms_norm_eps=1e-5,
                      tie_word_embeddings=False, attention_bias=False, mlp_bias=False)
    cfg.architectures = ["LlamaForCausalLM"]
    cfg.anyprec = {"seed_precision": seed_precision, "parent_precision": parent_precision, "group_count": 1,
                   "arch_config": {"module_names": LLAMA_MODULES, "model_name": "model", "layers_name": "layers"}}
    return cfg


def write_checkpoint(path, oracle, cfg=None, seed=0):
    """returns (cfg, dense) where dense[b] is the fp16 state dict of the equivalent dense model at precision b."""
    cfg = cfg or tiny_llama_config()
    rng = np.random.default_rng(seed)
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    hd = H // cfg.num_attention_heads
    shapes = {"self_attn.q_proj": (H, H), "self_attn.k_proj": (cfg.num_key_value_heads * hd, H),
              "self_attn.v_proj": (cfg.num_key_value_heads * hd, H), "self_attn.o_proj": (H, H),
              "mlp.gate_proj": (I, H), "mlp.up_proj": (I, H), "mlp.down_proj": (H, I)}
    lo, hi = cfg.anyprec["seed_precision"], cfg.anyprec["parent_precision"]
    sd, dense = {}, {b: {} for b in range(lo, hi + 1)}

    def both(k, t):
        sd[k] = t
        for b in dense:
            dense[b][k] = t

    both("model.embed_tokens.weight", torch.from_numpy((rng.standard_normal((V, H)) * 0.5).astype(np.float16)))
    both("lm_head.weight", torch.from_numpy((rng.standard_normal((V, H)) * 0.05).astype(np.float16)))
    both("model.norm.weight", torch.from_numpy((1 + 0.1 * rng.standard_normal(H)).astype(np.float16)))
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        both(p + "input_layernorm.weight", torch.from_numpy((1 + 0.1 * rng.standard_normal(H)).astype(np.float16)))
        both(p + "post_attention_layernorm.weight", torch.from_numpy((1 + 0.1 * rng.standard_normal(H)).astype(np.float16)))
        for name, (N, K) in shapes.items():
            idx = rng.integers(0, 1 << hi, size=(N, K), dtype=np.uint8)
            sd[p + name + ".qweight"] = torch.from_numpy(oracle.pack(idx, hi))
            for b in range(lo, hi + 1):
                lut = (rng.standard_normal((N, 1 << b)) * (1.0 / np.sqrt(K))).astype(np.float16)
                sd[p + name + f".lut{b}"] = torch.from_numpy(lut)
                W = lut[np.arange(N)[:, None], idx >> (hi - b)]  # first b planes == the b-bit model
                dense[b][p + name + ".weight"] = torch.from_numpy(W)
    os.makedirs(path, exist_ok=True)
    torch.save(sd, os.path.join(path, "pytorch_model.bin"))
    cfg.save_pretrained(path)
    return cfg, dense
