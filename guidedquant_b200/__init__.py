"""guidedquant_b200 — B200-native (sm_100a) Any-Precision LUT GEMV decode path of snu-mllab/GuidedQuant.

Public surface (mirrors the reference's inference/ files):
    guidedquant_b200.ap_gemv      drop-in for the native module `ap_gemv`
    guidedquant_b200.plugin       torch custom op `plugin::anyprec_gemv` + `anyprec_dequant`
    guidedquant_b200.APLinear     gpt-fast side Linear module
    guidedquant_b200.AnyPrecisionLinear  HF side Linear module (multi-precision lut{b})
    guidedquant_b200.pack         packed bit-plane layout (pack / unpack / K-shard re-pack)
    guidedquant_b200.sharding     K-sharded (row-parallel) APLinear with one all-reduce
"""
import sys as _sys

__version__ = "0.1.0"


def install_as_ap_gemv() -> None:
    """Make `import ap_gemv` resolve to this package's drop-in module (for the reference's own
    inference/plugin.py, APLinear.py and any_precision/modules/AnyPrecisionLinear.py)."""
    from . import ap_gemv as _m

    _sys.modules["ap_gemv"] = _m
