"""guidedquant_b200 — B200-native (sm_100a) Any-Precision LUT GEMV decode path of snu-mllab/GuidedQuant.

Public surface (mirrors the reference's inference/ files):
    guidedquant_b200.ap_gemv      drop-in for the native module `ap_gemv`
    guidedquant_b200.plugin       torch custom op `plugin::anyprec_gemv` + `anyprec_dequant`
    guidedquant_b200.APLinear     gpt-fast side Linear module
    guidedquant_b200.AnyPrecisionLinear  HF side Linear module (multi-precision lut{b})
    guidedquant_b200.AnyPrecisionForCausalLM  HF side model wrapper (from_quantized / set_precision / generate)
    guidedquant_b200.pack         packed bit-plane layout (pack / unpack / K-shard re-pack)
    guidedquant_b200.convert      HF-named packed checkpoint -> fused gpt-fast names (sqllm_llama_convert_fuse.py)
    guidedquant_b200.model        APTransformer: the whole decode step as one CUDA graph (single GPU or tensor parallel)
    guidedquant_b200.runtime      ApGemvChain: the per-token chain of APLinear GEMVs (the hot path alone)
    guidedquant_b200.tp           peer-memory plumbing of the fused one-shot all-reduce
"""
import sys as _sys

__version__ = "0.1.0"


def install_as_ap_gemv() -> None:
    """Make `import ap_gemv` resolve to this package's drop-in module (for the reference's own
    inference/plugin.py, APLinear.py and any_precision/modules/AnyPrecisionLinear.py)."""
    from . import ap_gemv as _m

    _sys.modules["ap_gemv"] = _m
