"""AnyPrecisionForCausalLM — the HF-side model wrapper of the reference (SURVEY.md §8 f-3;
any_precision/modules/AnyPrecisionForCausalLM.py:28-208), running on this repo's ap_gemv drop-in.

Same user surface: `AnyPrecisionForCausalLM.from_quantized(path, precisions=...)`, `.forward(..., precision=b)`,
`.generate(..., precision=b)`, `.set_precision(b)`, `.prune_precisions()`, `.precision`, `.precisions`,
`.supported_bits`, `.ap_linears`, `.device`, and the checkpoint format written by the reference packer
(any_precision/quantization/pack.py:133-203): a directory holding `config.json` with an `anyprec` dict
{seed_precision, parent_precision, group_count, arch_config{model_name, layers_name, module_names}} and
`pytorch_model.bin` whose quantized Linears are stored as `<...>.qweight [parent_bits, N, K/32]` + `<...>.lut{b}`.

What is different, and why:
  * no `accelerate` (not in this image, and device_map="auto" sharding is not how this repo scales — tensor
    parallelism lives in guidedquant_b200.model): the skeleton is built on the meta device, the Linears named by
    `arch_config` are swapped for AnyPrecisionLinear, then the whole model is materialised on ONE device and the
    state dict is loaded directly;
  * no hub download (no network): `model_path` must be a local directory;
  * the set of Linears to replace comes straight from `config.anyprec['arch_config']` — the reference routes the
    same three fields through its ModelAnalyzer (any_precision/analyzer/analyzer.py:14-15, 90-105).
"""
from __future__ import annotations

import gc
import os

import torch
import torch.nn as nn

from .AnyPrecisionLinear import AnyPrecisionLinear
from .convert import read_checkpoint


def _resolve(root, dotted: str):
    m = root
    for part in dotted.split("."):
        m = m[int(part)] if part.isdigit() else getattr(m, part)
    return m


def _swap(root, dotted: str, new: nn.Module) -> None:
    head, _, leaf = dotted.rpartition(".")
    setattr(_resolve(root, head) if head else root, leaf, new)


class AnyPrecisionForCausalLM(nn.Module):
    def __init__(self, model_path, config, precisions=None, torch_dtype=torch.float16, fuse_layers=False,
                 trust_remote_code=True, local_dir=None, device=None):
        super().__init__()
        from transformers import AutoModelForCausalLM

        if not hasattr(config, "anyprec"):
            raise RuntimeError("config has no `anyprec` section: not an Any-Precision checkpoint")
        self.config = config
        ap = config.anyprec
        self.supported_bits = list(range(ap["seed_precision"], ap["parent_precision"] + 1))
        if precisions is None:
            self.precisions = self.supported_bits
        else:
            assert len(precisions) == len(set(precisions)), "Precisions must be unique"
            assert all(bit in self.supported_bits for bit in precisions), \
                f"Supported bits {precisions} must be a subset of model supported bits {self.supported_bits}"
            self.precisions = list(precisions)
        self.precision = max(self.precisions)
        if not os.path.isdir(model_path):
            raise FileNotFoundError(f"{model_path}: not a local directory (hub download is not available here)")
        if device is None:
            device = "cuda"
        device = torch.device(device)

        with torch.device("meta"):
            try:
                self.model = AutoModelForCausalLM.from_config(config, dtype=torch_dtype, trust_remote_code=trust_remote_code)
            except TypeError:  # older transformers spell it torch_dtype
                self.model = AutoModelForCausalLM.from_config(config, torch_dtype=torch_dtype,
                                                              trust_remote_code=trust_remote_code)
        self.ap_linears: list[AnyPrecisionLinear] = []
        self._load_quantized_modules()
        self.tie_weights()
        self._materialize(model_path, device, torch_dtype)
        if fuse_layers:
            self.fuse_layers()
        self.prune_precisions()
        self.model.eval()

    # -- construction ---------------------------------------------------------------------------------------------------
    def _load_quantized_modules(self):
        names = self.config.anyprec["arch_config"]["module_names"]
        for layer in self.get_model_layers():
            for name in names:
                lin = _resolve(layer, name)
                q = AnyPrecisionLinear(lin.in_features, lin.out_features, self.supported_bits,
                                       bias=lin.bias is not None, precisions=self.precisions, device="meta")
                self.ap_linears.append(q)
                _swap(layer, name, q)

    def _materialize(self, model_path, device, dtype):
        sd = read_checkpoint(model_path)
        # modules that own non-persistent buffers (rotary inv_freq) cannot be restored from a state dict: rebuild them
        rebuilt = []
        for name, m in list(self.model.named_modules()):
            if m._non_persistent_buffers_set and not isinstance(m, AnyPrecisionLinear):
                try:
                    fresh = type(m)(config=self.config, device=device)
                except Exception as e:  # loud: a silently uninitialised buffer would corrupt every logit
                    raise RuntimeError(f"cannot re-create {type(m).__name__} ({name}) holding non-persistent buffers") from e
                rebuilt.append((name, fresh))
        self.model.to_empty(device=device)
        for name, fresh in rebuilt:
            _swap(self.model, name, fresh)
        # lut{b} buffers are created without a dtype in the reference module (fp32 default) and cast per call; keep the
        # checkpoint's fp16 instead
        for q in self.ap_linears:
            for b in q.supported_bits:
                q._buffers[f"lut{b}"] = q._buffers[f"lut{b}"].to(torch.float16)
            if q.bias is not None:
                q._buffers["bias"] = q._buffers["bias"].to(dtype)
        res = self.model.load_state_dict(sd, strict=False, assign=False)
        missing = [k for k in res.missing_keys if "rotary_emb" not in k]
        if self.model.config.tie_word_embeddings:
            missing = [k for k in missing if not k.startswith("lm_head.")]
        if missing:
            raise RuntimeError(f"checkpoint {model_path} lacks {len(missing)} tensors, e.g. {missing[:4]}")
        self.tie_weights()
        del sd
        gc.collect()

    # -- reference surface ---------------------------------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        prev = self.precision
        if "precision" in kwargs:
            self.set_precision(kwargs.pop("precision"))
        try:
            return self.model.forward(*args, **kwargs)
        finally:
            self.set_precision(prev)

    def generate(self, *args, **kwargs):
        prev = self.precision
        if "precision" in kwargs:
            self.set_precision(kwargs.pop("precision"))
        try:
            with torch.inference_mode():
                return self.model.generate(*args, **kwargs)
        finally:
            self.set_precision(prev)

    @staticmethod
    def _load_config(model_path, trust_remote_code=True):
        from transformers import AutoConfig

        return AutoConfig.from_pretrained(model_path, trust_remote_code=trust_remote_code)

    @classmethod
    def from_quantized(cls, quant_model_path, trust_remote_code=True, fuse_layers=False, precisions=None, local_dir=None,
                       torch_dtype=torch.float16, device=None):
        config = cls._load_config(quant_model_path, trust_remote_code)
        return cls(model_path=quant_model_path, precisions=precisions, config=config, fuse_layers=fuse_layers,
                   trust_remote_code=trust_remote_code, local_dir=local_dir, torch_dtype=torch_dtype, device=device)

    def prune_precisions(self):
        for q in self.ap_linears:
            q.prune_precisions()
        gc.collect()

    def set_precision(self, precision):
        for q in self.ap_linears:
            q.set_precision(precision)
        self.precision = precision

    def tie_weights(self):
        if hasattr(self.model, "tie_weights"):
            self.model.tie_weights()

    def get_model_layers(self):
        arch = self.config.anyprec["arch_config"]
        return _resolve(_resolve(self.model, arch["model_name"]), arch["layers_name"])

    def fuse_layers(self):
        # the reference leaves this unimplemented as well (AnyPrecisionForCausalLM.py:190-194); the fused decode path of
        # this repo is guidedquant_b200.model.APTransformer, fed through guidedquant_b200.convert
        raise NotImplementedError("layer fusion lives in guidedquant_b200.model.APTransformer (see convert.py)")

    @property
    def layer_type(self):
        for layer in self.get_model_layers():
            if type(layer).__name__.endswith("DecoderLayer"):
                return type(layer).__name__
        return None

    @property
    def device(self):
        return self.model.device
