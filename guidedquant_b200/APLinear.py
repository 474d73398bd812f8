"""APLinear — mirror of the reference's inference/APLinear.py:6-60 (same constructor, buffers, aliasing
and forward semantics) on top of the B200 kernels.

Buffers: qweight int32 [bits, N, K/32], lut fp16 [N, 2^bits], optional bias; `self.output` is a
persistent NON-buffer [1,1,N] that is overwritten and RETURNED BY REFERENCE on every decode call
(APLinear.py:33, :60) — callers that keep the result across calls must clone it, as with the reference.
"""
import torch
import torch.nn as nn

from . import ap_gemv
from .plugin import anyprec_dequant, anyprec_gemv


class APLinear(nn.Module):
    def __init__(self, in_features, out_features, bitwidth, bias=False, dtype=torch.half, device="cuda"):
        super().__init__()
        assert in_features % 32 == 0, "in_features must be a multiple of 32 (packed layout)"
        self.in_features = in_features
        self.out_features = out_features
        self.bitwidth = bitwidth
        self.dtype = dtype
        self.register_buffer(
            "qweight", torch.empty((bitwidth, out_features, in_features // 32), dtype=torch.int32, device=device))
        self.register_buffer("lut", torch.empty((out_features, 2 ** bitwidth), dtype=self.dtype, device=device))
        if bias:
            self.register_buffer("bias", torch.empty((out_features,), dtype=self.dtype, device=device))
        else:
            self.bias = None
        self.output = torch.zeros((1, 1, self.out_features), dtype=self.dtype, device=device)

    def _apply(self, fn, *args, **kwargs):  # keep the non-buffer output tensor on the module's device
        super()._apply(fn, *args, **kwargs)
        self.output = fn(self.output)
        return self

    def gemm(self, x):
        # prefill / seq > 1 (APLinear.py:35-38 does dequant -> fp16 matmul).  Up to 8 tokens go through the batched LUT
        # GEMV instead (the kernel's M dimension, gemv.cu:41): the packed weights are read once and no fp16 copy of the
        # matrix is written to HBM.  Longer sequences: the fused dequant + tcgen05 GEMM kernel (csrc/prefill_tc.cuh), which
        # also never materialises the fp16 matrix, up to the measured cross-over token count; beyond it, and for shapes
        # the kernel does not take (bits > 4, K % 256 != 0): dequant -> cuBLAS like the reference.
        T = x.shape[1]
        if T <= 8 and x.dtype == torch.float16 and x.is_cuda:
            out = torch.empty((T, 1, self.out_features), dtype=torch.float16, device=x.device)
            anyprec_gemv(x.reshape(T, 1, self.in_features).contiguous(), self.qweight, self.lut, out, self.bitwidth)
            return out.reshape(1, T, self.out_features)
        if x.dtype == torch.float16 and x.is_cuda and ap_gemv.prefill_prefers_fused(self.qweight, self.bitwidth, T):
            return ap_gemv.anyprec_prefill_gemm(x, self.qweight, self.lut, self.bitwidth)
        weight = anyprec_dequant(self.qweight, self.lut, self.bitwidth)
        return torch.matmul(x, weight.T)

    def forward(self, x, **kwargs):
        assert x.shape[0] == 1
        if x.shape[1] > 1:
            output = self.gemm(x)
            if self.bias is not None:
                output += self.bias
            return output
        # the reference zeroes self.output first (APLinear.py:53); the kernel overwrites every element, so
        # the memset is skipped here (one launch per Linear instead of two)
        anyprec_gemv(x, self.qweight, self.lut, self.output, self.bitwidth)
        if self.bias is not None:
            self.output += self.bias
        return self.output
