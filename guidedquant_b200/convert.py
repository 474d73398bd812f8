"""Checkpoint conversion: HF-named packed Any-Precision checkpoint (what any_precision/quantization/pack.py:133-203
writes: `model.layers.{i}.self_attn.q_proj.qweight`, `...lut{b}`, ...) -> the fused gpt-fast names the decode runtime
loads.  Mirrors the reference's inference/sqllm_llama_convert_fuse.py:12-118 (same renames, `lut{bitwidth}` selection,
`qweight[:bitwidth]` slicing, q|k|v and gate|up concatenation along the output-row axis, bf16 -> fp16), with two
differences: the layer count is read from the keys instead of being guessed from the directory name (the reference
rejects everything but Llama-2 directory names, :62-69), and nothing is written unless asked.

    python -m guidedquant_b200.convert --ckpt_dir <dir> --bitwidth 2      # writes <dir>/converted_pytorch_model.bin
"""
from __future__ import annotations

import argparse
import os
import re

import torch

REPLACEMENTS = {  # sqllm_llama_convert_fuse.py:12-24
    "embed_tokens": "tok_embeddings",
    "self_attn": "attention",
    "o_proj": "wo",
    "mlp": "feed_forward",
    "down_proj": "w2",
    "lm_head": "output",
    "lookup_table": "lut",
}


def convert_state_dict(ckpt: dict, bitwidth: int) -> dict:
    new = {}
    for key, value in ckpt.items():
        k = key.replace("model.", "")
        for old, rep in REPLACEMENTS.items():
            k = k.replace(old, rep)
        if "lut" in k:  # keep only lut{bitwidth}, renamed to lut (:44-49)
            if f"lut{bitwidth}" in k:
                k = re.sub(r"(?<=lut)[2-8]", "", k)
            else:
                continue
        new[k] = value
    for k in list(new.keys()):
        v = new[k]
        if v.dtype == torch.bfloat16:
            v = v.half()
        if k.endswith(".lut"):
            v = v.half()
        if "qweight" in k:
            v = v.contiguous()[:bitwidth, :, :]  # any-precision: the first `bitwidth` planes are the model (:59-60)
        new[k] = v
    layers = sorted({int(m.group(1)) for m in (re.match(r"layers\.(\d+)\.", k) for k in new) if m})
    for i in layers:
        a, f = f"layers.{i}.attention.", f"layers.{i}.feed_forward."
        for suffix, dim in (("qweight", 1), ("lut", 0)):
            if a + "q_proj." + suffix in new:  # q | k | v along the output rows (:71-97)
                new[a + "wqkv." + suffix] = torch.cat([new.pop(a + p + "_proj." + suffix) for p in ("q", "k", "v")], dim=dim)
            if f + "gate_proj." + suffix in new:  # gate | up (:99-116)
                new[f + "w1w3." + suffix] = torch.cat([new.pop(f + p + "_proj." + suffix) for p in ("gate", "up")], dim=dim)
    return new


def read_checkpoint(model_path: str) -> dict:
    """pytorch_model.bin (what the reference packer writes, pack.py:199), or safetensors / sharded variants of either."""
    import json

    cands = ["pytorch_model.bin", "model.safetensors", "pytorch_model.bin.index.json", "model.safetensors.index.json"]
    for name in cands:
        p = os.path.join(model_path, name)
        if not os.path.exists(p):
            continue
        files = sorted(set(json.load(open(p))["weight_map"].values())) if name.endswith(".index.json") else [name]
        sd = {}
        for f in files:
            fp = os.path.join(model_path, f)
            if f.endswith(".safetensors"):
                from safetensors.torch import load_file

                sd.update(load_file(fp))
            else:
                sd.update(torch.load(fp, map_location="cpu", mmap=True, weights_only=True))
        return sd
    raise FileNotFoundError(f"no checkpoint file ({', '.join(cands)}) under {model_path}")


def rope_inv_freq(theta: float, scaling: dict | None = None, head_dim: int = 128) -> torch.Tensor:
    """fp32 inv_freq[head_dim / 2] of the rotary embedding (inference/model.py:353 -> transformers' ROPE_INIT_FUNCTIONS):
    "default" and the Llama-3.1 frequency rescaling ("llama3": low frequencies divided by `factor`, a smooth ramp between
    the two wavelength thresholds; attention_scaling stays 1).  Host-side: it only changes the 64-entry table."""
    import math

    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    kind = (scaling or {}).get("rope_type", (scaling or {}).get("type", "default"))
    if kind in (None, "default"):
        return inv
    if kind != "llama3":
        raise NotImplementedError(f"rope scaling '{kind}' is not implemented (default and llama3 only)")
    factor, lo, hi = scaling["factor"], scaling["low_freq_factor"], scaling["high_freq_factor"]
    old_len = scaling["original_max_position_embeddings"]
    low_wl, high_wl = old_len / lo, old_len / hi
    wavelen = 2 * math.pi / inv
    out = torch.where(wavelen > low_wl, inv / factor, inv)
    smooth = (old_len / wavelen - lo) / (hi - lo)
    smoothed = (1 - smooth) * out / factor + smooth * out
    medium = ~(wavelen < high_wl) & ~(wavelen > low_wl)
    return torch.where(medium, smoothed, out)


def arch_from_hf_config(hf: dict) -> tuple[dict, float, float]:
    """(model config, rope base, rms-norm eps) of a Llama-family config.json; refuses what the decode kernels do not
    implement (head_dim != 128, RoPE scalings other than llama3) instead of producing silently wrong logits.  A llama3
    scaling dict is returned inside the model config under "rope_scaling"."""
    n_head = hf["num_attention_heads"]
    head_dim = hf.get("head_dim") or hf["hidden_size"] // n_head
    if head_dim != 128:
        raise NotImplementedError(f"head_dim {head_dim}: the attention kernel is specialised for 128")
    rp = hf.get("rope_parameters") or {}                      # transformers >= 5 nests the RoPE settings
    scaling = hf.get("rope_scaling") or ({k: v for k, v in rp.items() if k != "rope_theta"} if rp else None)
    kind = (scaling or {}).get("rope_type", (scaling or {}).get("type", "default"))
    if kind not in (None, "default", "llama3"):
        raise NotImplementedError(f"rope scaling '{kind}' is not implemented (default and llama3 only, model.py:337-353)")
    theta = hf.get("rope_theta", rp.get("rope_theta", 10000.0))
    cfg = dict(dim=hf["hidden_size"], n_layer=hf["num_hidden_layers"], n_head=n_head,
               n_kv=hf.get("num_key_value_heads") or n_head, inter=hf["intermediate_size"], vocab=hf["vocab_size"])
    if kind == "llama3":
        cfg["rope_scaling"] = dict(scaling)
    return cfg, float(theta), float(hf.get("rms_norm_eps", 1e-5))


def convert_checkpoint(ckpt_dir: str, bitwidth: int, out_name: str = "converted_pytorch_model.bin") -> str:
    ckpt = read_checkpoint(ckpt_dir)
    out = os.path.join(ckpt_dir, out_name)
    torch.save(convert_state_dict(ckpt, bitwidth), out)
    return out


def load_converted(path: str, device="cuda") -> dict:
    """the runtime-side load (generate.py:237-238: mmap + weights_only)."""
    return torch.load(path, map_location=device, mmap=True, weights_only=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ckpt_dir", type=str, required=True)
    ap.add_argument("--bitwidth", type=int, required=True)
    a = ap.parse_args()
    print(convert_checkpoint(a.ckpt_dir, a.bitwidth))
