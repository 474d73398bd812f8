"""Operator boundary, mirroring the reference's inference/plugin.py:1-18.

    torch.ops.plugin.anyprec_gemv(x, q_weight, lut, output, bitwidth) -> None   (mutates `output`)
    anyprec_dequant(q_weight, lut, bitwidth) -> Tensor                           (plain function)

Same op name, schema, mutates_args and fake impl, so graphs traced against the reference's op (e.g.
generate.py's torch.compile(decode_one_token, fullgraph=True), generate.py:330-336) resolve to this
implementation.  Note the argument order differs from the extension's (plugin.py:9).
"""
import torch

from . import ap_gemv

_OP = "plugin::anyprec_gemv"


def _already_registered() -> bool:
    try:
        return hasattr(torch.ops.plugin, "anyprec_gemv") and torch.ops.plugin.anyprec_gemv is not None
    except (AttributeError, RuntimeError):
        return False


if not _already_registered():  # the reference registers this name twice (plugin.py:7, AnyPrecisionLinear.py:9)

    @torch.library.custom_op(_OP, mutates_args={"output"})
    def anyprec_gemv(x: torch.Tensor, q_weight: torch.Tensor, lut: torch.Tensor, output: torch.Tensor,
                     bitwidth: int) -> None:
        ap_gemv.anyprec_gemv(x, output, q_weight, lut, bitwidth)

    @anyprec_gemv.register_fake
    def _(x, q_weight, lut, output, bitwidth):
        return None

else:  # pragma: no cover - only when another module registered the op first
    anyprec_gemv = torch.ops.plugin.anyprec_gemv


def anyprec_dequant(q_weight: torch.Tensor, lut: torch.Tensor, bitwidth: int) -> torch.Tensor:
    return ap_gemv.anyprec_dequant(q_weight, lut, bitwidth)
