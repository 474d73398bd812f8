"""HF-side surface: AnyPrecisionForCausalLM over the reference packer's checkpoint format.
CPU part: construction / loading / precision bookkeeping.  GPU part: logits against the dense fp16 HF model whose
Linear weights are the oracle's dequantised matrices, for every supported precision, prefill and KV-cache decode."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
import hf_fixture  # noqa: E402


@pytest.fixture(scope="module")
def ckpt(tmp_path_factory, oracle):
    d = str(tmp_path_factory.mktemp("anyprec_ckpt"))
    cfg, dense = hf_fixture.write_checkpoint(d, oracle, seed=3)
    return d, cfg, dense


def test_load_structure_and_precisions_cpu(ckpt):
    from guidedquant_b200.AnyPrecisionForCausalLM import AnyPrecisionForCausalLM
    from guidedquant_b200.AnyPrecisionLinear import AnyPrecisionLinear

    d, cfg, _ = ckpt
    m = AnyPrecisionForCausalLM.from_quantized(d, device="cpu")
    assert m.supported_bits == [2, 3, 4] and m.precisions == [2, 3, 4] and m.precision == 4
    assert len(m.ap_linears) == 7 * cfg.num_hidden_layers and m.layer_type == "LlamaDecoderLayer"
    sd = torch.load(os.path.join(d, "pytorch_model.bin"), weights_only=True)
    l0 = m.model.model.layers[0]
    assert isinstance(l0.self_attn.q_proj, AnyPrecisionLinear) and isinstance(l0.mlp.down_proj, AnyPrecisionLinear)
    assert torch.equal(l0.mlp.down_proj.qweight, sd["model.layers.0.mlp.down_proj.qweight"])
    assert l0.mlp.down_proj.lut3.dtype == torch.float16
    assert torch.equal(l0.mlp.down_proj.lut3, sd["model.layers.0.mlp.down_proj.lut3"])
    assert torch.equal(m.model.lm_head.weight, sd["lm_head.weight"])
    # rotary inv_freq is a non-persistent buffer: must have been rebuilt, not left uninitialised
    hd = cfg.hidden_size // cfg.num_attention_heads
    expect = 1.0 / (cfg.rope_parameters["rope_theta"] if hasattr(cfg, "rope_parameters") else cfg.rope_theta) ** (
        torch.arange(0, hd, 2).float() / hd)
    assert torch.allclose(m.model.model.rotary_emb.inv_freq.float().cpu(), expect)
    m.set_precision(2)
    assert m.precision == 2 and all(q.precision == 2 for q in m.ap_linears)
    with pytest.raises(RuntimeError):
        m.set_precision(5)
    # no CPU fallback: a forward off the GPU must fail loudly, never compute on the host
    with pytest.raises(RuntimeError, match="GPU"):
        m(torch.zeros((1, 1), dtype=torch.long))


def test_prune_and_errors_cpu(ckpt, tmp_path):
    from guidedquant_b200.AnyPrecisionForCausalLM import AnyPrecisionForCausalLM

    d, cfg, _ = ckpt
    m = AnyPrecisionForCausalLM.from_quantized(d, device="cpu", precisions=[2, 3])
    q = m.ap_linears[0]
    assert m.precision == 3 and q.qweight.shape[0] == 3 and not hasattr(q, "lut4") and hasattr(q, "lut2")
    with pytest.raises(AssertionError):
        AnyPrecisionForCausalLM.from_quantized(d, device="cpu", precisions=[2, 5])
    with pytest.raises(FileNotFoundError):
        AnyPrecisionForCausalLM.from_quantized(d, device="cpu").__class__(str(tmp_path / "nope"), cfg, device="cpu")
    # a checkpoint lacking a tensor is refused
    sd = torch.load(os.path.join(d, "pytorch_model.bin"), weights_only=True)
    sd.pop("model.layers.1.mlp.up_proj.lut3")
    bad = tmp_path / "bad"
    bad.mkdir()
    torch.save(sd, bad / "pytorch_model.bin")
    cfg.save_pretrained(str(bad))
    with pytest.raises(RuntimeError, match="lacks"):
        AnyPrecisionForCausalLM.from_quantized(str(bad), device="cpu")


@pytest.mark.gpu
def test_logits_match_dense_model_gpu(ckpt):
    from transformers import LlamaForCausalLM

    from guidedquant_b200.AnyPrecisionForCausalLM import AnyPrecisionForCausalLM

    d, cfg, dense = ckpt
    m = AnyPrecisionForCausalLM.from_quantized(d)
    assert m.device.type == "cuda"
    ids = torch.tensor([[1, 17, 250, 33, 99, 7]], device="cuda")
    for bits in (4, 3, 2):
        ref = LlamaForCausalLM(cfg).half().cuda().eval()
        ref.load_state_dict(dense[bits], strict=True)
        with torch.inference_mode():
            want = ref(ids).logits.float()
            got = m(ids, precision=bits).logits.float()  # seq > 1: dequant + matmul path
            assert m.precision == 4  # restored
            err = float((got - want).abs().max() / want.abs().max())
            assert err <= 5e-3, (bits, err)
            # KV-cache decode: prefill 5 tokens, then one token through the GEMV kernels
            o = m(ids[:, :5], use_cache=True, precision=bits)
            step = m(ids[:, 5:], past_key_values=o.past_key_values, use_cache=True, precision=bits).logits.float()
            err = float((step[0, -1] - want[0, -1]).abs().max() / want[0, -1].abs().max())
            assert err <= 1e-2, (bits, err)
        del ref
    out = m.generate(ids[:, :2], max_new_tokens=4, do_sample=False, precision=3)
    assert out.shape == (1, 6) and m.precision == 4


def test_checkpoint_container_variants_cpu(ckpt, tmp_path):
    """the loader accepts what HF tooling may have re-saved the packer's output as: safetensors, and sharded files with
    an index json (any_precision checkpoints on the hub ship pytorch_model.bin; pack.py:199)."""
    import json

    from safetensors.torch import save_file

    from guidedquant_b200.AnyPrecisionForCausalLM import AnyPrecisionForCausalLM

    d, cfg, _ = ckpt
    sd = torch.load(os.path.join(d, "pytorch_model.bin"), weights_only=True)
    ref = AnyPrecisionForCausalLM.from_quantized(d, device="cpu").model.state_dict()

    st = tmp_path / "st"
    st.mkdir()
    save_file({k: v.contiguous() for k, v in sd.items()}, str(st / "model.safetensors"))
    cfg.save_pretrained(str(st))
    got = AnyPrecisionForCausalLM.from_quantized(str(st), device="cpu").model.state_dict()
    assert got.keys() == ref.keys() and all(torch.equal(got[k], ref[k]) for k in ref)

    sh = tmp_path / "sharded"
    sh.mkdir()
    keys = sorted(sd)
    parts = {"pytorch_model-00001-of-00002.bin": keys[: len(keys) // 2], "pytorch_model-00002-of-00002.bin": keys[len(keys) // 2:]}
    for fn, ks in parts.items():
        torch.save({k: sd[k] for k in ks}, sh / fn)
    json.dump({"metadata": {}, "weight_map": {k: fn for fn, ks in parts.items() for k in ks}},
              open(sh / "pytorch_model.bin.index.json", "w"))
    cfg.save_pretrained(str(sh))
    got = AnyPrecisionForCausalLM.from_quantized(str(sh), device="cpu").model.state_dict()
    assert got.keys() == ref.keys() and all(torch.equal(got[k], ref[k]) for k in ref)

    empty = tmp_path / "empty"
    empty.mkdir()
    cfg.save_pretrained(str(empty))
    with pytest.raises(FileNotFoundError, match="no checkpoint file"):
        AnyPrecisionForCausalLM.from_quantized(str(empty), device="cpu")
