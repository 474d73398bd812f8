"""AnyPrecisionLinear — mirror of the reference's HF-side module any_precision/modules/AnyPrecisionLinear.py:17-89
(same constructor, buffers `qweight [max_bits, N, K/32]` + one `lut{b}` per supported bit-width, `set_precision`,
`prune_precisions`, the in-place clamp to +-(1 - 5e-3) * fp16 max, and the aliased persistent `self.output`).

Difference: the reference's extension insists on `qweight.size(0) == bitwidth` (gemv.cu:76), so running a pruned-down
precision on the shared bit-plane tensor raises there; here any `bitwidth <= qweight.size(0)` runs on the first
`bitwidth` planes (the any-precision property).
"""
import torch
import torch.nn as nn

from . import ap_gemv
from .plugin import anyprec_gemv


class AnyPrecisionLinear(nn.Module):
    def __init__(self, in_features, out_features, supported_bits, bias=True, precisions=None, device=None, dtype=None):
        super().__init__()
        if precisions is None:
            precisions = supported_bits
        if not isinstance(precisions, list):
            raise RuntimeError("supported_bits must be a list of integers.")
        if dtype is not None and dtype != torch.float16:
            raise RuntimeError("Only float16 is supported for now.")
        self.in_features = in_features
        self.out_features = out_features
        self.precisions = precisions
        self.precision = max(self.precisions)
        self.supported_bits = supported_bits
        self.register_buffer(
            "qweight", torch.empty((max(supported_bits), out_features, in_features // 32), dtype=torch.int32, device=device))
        for bit in supported_bits:
            self.register_buffer(f"lut{bit}", torch.empty((out_features, 2 ** bit), dtype=dtype, device=device))
        if bias:
            self.register_buffer("bias", torch.empty((out_features,), dtype=dtype, device=device))
        else:
            self.bias = None
        # the reference allocates this on 'cuda' at construction (AnyPrecisionLinear.py:55); here it is created on the
        # first GEMV call on qweight's device, so that a skeleton can be built on the meta device — still ONE
        # persistent tensor per module, returned by reference on every call
        self.output = None

    def prune_precisions(self):
        self.qweight = self.qweight[:max(self.precisions)]
        for bit in self.supported_bits:
            if bit not in self.precisions:
                delattr(self, f"lut{bit}")

    def forward(self, x, **kwargs):
        w_bits = kwargs["precision"] if "precision" in kwargs else self.precision
        lut = self._buffers[f"lut{w_bits}"].to(torch.float16)
        T = x.numel() // x.shape[-1]
        if 1 < T <= 8 and x.is_cuda:
            # short sequences: the batched LUT GEMV (kernel M dimension, gemv.cu:41) instead of dequant -> matmul
            out = torch.empty((T, 1, self.out_features), dtype=torch.float16, device=x.device)
            anyprec_gemv(x.to(torch.float16).reshape(T, 1, -1).contiguous(), self.qweight, lut, out, w_bits)
            x = out.to(x.dtype).reshape(*x.shape[:-1], self.out_features)
        elif T > 1 and x.is_cuda and ap_gemv.prefill_prefers_fused(self.qweight, w_bits, T):
            # fused dequant + tensor-core GEMM: the fp16 weight matrix is never written to HBM (csrc/prefill_tc.cuh)
            x = ap_gemv.anyprec_prefill_gemm(x.to(torch.float16), self.qweight, lut, w_bits).to(x.dtype)
        elif T > 1:
            weight = ap_gemv.anyprec_dequant(self.qweight, lut, w_bits).to(x.dtype)
            x = torch.matmul(x, weight.T)
        else:
            if self.output is None or self.output.device != self.qweight.device:
                self.output = torch.zeros((1, 1, self.out_features), dtype=torch.float16, device=self.qweight.device)
            anyprec_gemv(x.to(torch.float16).reshape(1, 1, -1), self.qweight, lut, self.output, w_bits)
            x = self.output.to(x.dtype).reshape(*x.shape[:-1], self.out_features)
        if self.bias is not None:
            x += self.bias
        return x.clamp_(torch.finfo(x.dtype).min * (1.0 - 5e-3), torch.finfo(x.dtype).max * (1.0 - 5e-3))

    def set_precision(self, precision):
        if precision not in self.precisions:
            raise RuntimeError(f"{self.precisions}-bit precisions are supported but {precision}-bit was specified.")
        self.precision = precision

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}"
