"""Drop-in replacement of the reference's native extension module `ap_gemv`
(inference/ap_gemv/bindings.cpp:12-17, gemv.h:17-39): same function names, argument order, tensor
contract and error behaviour (RuntimeError on bad shape/dtype, like TORCH_CHECK in gemv.cu:64-90),
implemented as a thin torch -> C-ABI shim over libapgemv_b200.so (hand-written sm_100a kernels).

    anyprec_gemv(input, output, qweight, lut, bitwidth) -> None      # output [M,1,N] overwritten
    anyprec_dequant(qweight, lut, bitwidth) -> fp16 Tensor [N, K]

`import guidedquant_b200; guidedquant_b200.install_as_ap_gemv()` registers this module as
sys.modules["ap_gemv"], so the reference's inference/plugin.py, APLinear.py and
any_precision/modules/AnyPrecisionLinear.py run unmodified on top of it.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

__all__ = ["anyprec_gemv", "anyprec_dequant", "anyprec_gemv_ex", "anyprec_prefill_gemm", "prefill_supported", "prefill_prefers_fused"]


def _req(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _check_gemv_args(input, output, qweight, lut, bitwidth):
    # mirrors anyprec_gemv_stream, inference/ap_gemv/gemv.cu:64-90 (same messages where it has them)
    _req(2 <= bitwidth <= 8, "Bitwidth must be between 2 and 8.")
    _req(input.dtype == lut.dtype == output.dtype, "Mismatched data types between input, lut, and output tensors.")
    # the reference reinterprets any matching dtype as half (gemv.cu:46-49); we refuse instead
    _req(input.dtype == torch.float16, "input, lut and output tensors must be float16.")
    _req(qweight.dtype == torch.int32, "qweight tensor must be of type int.")
    _req(input.dim() == 3, "input tensor must be of shape (batch_size, seq_len, hidden_size).")
    _req(output.dim() == 3, "output tensor must be of shape (batch_size, seq_len, hidden_size).")
    N, K = output.size(2), input.size(2)
    _req(lut.dim() == 2 and lut.size(1) == (1 << bitwidth) and lut.size(0) == N,
         f"lut tensor must be of shape (output_feat, 2 ** bitwidth). Expected ({N}, {1 << bitwidth}), got {tuple(lut.shape)}.")
    # the reference requires qweight.size(0) == bitwidth (gemv.cu:76); a multi-precision tensor with MORE planes is
    # accepted here and its first `bitwidth` planes are used (plane-major layout, any-precision property)
    _req(qweight.dim() == 3 and qweight.size(0) >= bitwidth and qweight.size(2) == K // 32 and qweight.size(1) == N,
         f"qweight tensor must be of shape (bitwidth, output_feat, input_feat / 32). Expected ({bitwidth}, {N}, {K // 32}), got {tuple(qweight.shape)}.")
    _req(input.size(1) == 1, "Only sequence length of 1 is supported.")
    _req(output.size(1) == 1, "Only sequence length of 1 is supported.")
    _req(input.is_cuda and output.is_cuda, "input and output tensors must be on GPU.")
    _req(qweight.is_cuda and lut.is_cuda and qweight.device == input.device == output.device == lut.device,
         "all tensors must be on the same GPU.")
    _req(input.is_contiguous(), "input tensor must be contiguous.")
    _req(output.is_contiguous(), "output tensor must be contiguous.")
    _req(qweight.is_contiguous(), "qweight tensor must be contiguous.")
    _req(lut.is_contiguous(), "lut tensor must be contiguous.")
    _req(output.size(0) == input.size(0), "input and output batch sizes differ.")
    _req(1 <= input.size(0) <= 8, "batch size must be between 1 and 8.")
    _req(K % 32 == 0 and K > 0, "input_feat must be a positive multiple of 32.")
    return input.size(0), N, K


def anyprec_gemv_ex(input, output, qweight, lut, bitwidth, flags: int = 0, partial=None, ctas_per_sm: int = 0,
                    prefetch_next: torch.Tensor | None = None) -> None:
    """anyprec_gemv with the C-ABI's options; `prefetch_next` = the packed weights the NEXT launch on this stream will
    read (L2 prefetch hint, performance only)."""
    M, N, K = _check_gemv_args(input, output, qweight, lut, bitwidth)
    if prefetch_next is not None:
        _lib.lib().apg_prefetch_hint(prefetch_next.data_ptr(), prefetch_next.numel() * prefetch_next.element_size())
    if partial is not None:
        _req(partial.dtype == torch.float32 and partial.is_contiguous() and partial.numel() == M * N
             and partial.device == input.device, "partial must be a contiguous float32 [M, N] tensor on the same GPU.")
    with torch.cuda.device(qweight.device):  # the reference does cudaSetDevice(qweight.device) (gemv.cu:103)
        st = _lib.lib().apg_gemv_ex(
            input.data_ptr(), output.data_ptr(), partial.data_ptr() if partial is not None else None,
            qweight.data_ptr(), lut.data_ptr(), M, N, K, bitwidth, flags, ctas_per_sm, _stream(qweight.device))
    _lib.check(st, "anyprec_gemv")


def anyprec_gemv(input: torch.Tensor, output: torch.Tensor, qweight: torch.Tensor, lut: torch.Tensor,
                 bitwidth: int) -> None:
    """ap_gemv.anyprec_gemv (inference/ap_gemv/gemv.cu:96-107).  Launches on the current stream."""
    anyprec_gemv_ex(input, output, qweight, lut, bitwidth, 0)


def anyprec_dequant(qweight: torch.Tensor, lut: torch.Tensor, bitwidth: int) -> torch.Tensor:
    """ap_gemv.anyprec_dequant (inference/ap_gemv/gemv.cu:109-134): allocates and returns fp16 [N, K]."""
    _req(2 <= bitwidth <= 8, "Bitwidth must be between 2 and 8.")
    _req(qweight.is_cuda and lut.is_cuda and qweight.device == lut.device, "qweight and lut must be on the same GPU.")
    _req(qweight.dtype == torch.int32 and qweight.dim() == 3 and qweight.is_contiguous(),
         "qweight tensor must be a contiguous int tensor of shape (bitwidth, output_feat, input_feat / 32).")
    _req(qweight.size(0) >= bitwidth, "qweight has fewer bit-planes than bitwidth.")
    N, K = qweight.size(1), qweight.size(2) * 32
    _req(lut.dtype == torch.float16 and lut.is_contiguous() and tuple(lut.shape) == (N, 1 << bitwidth),
         f"lut tensor must be a contiguous float16 tensor of shape ({N}, {1 << bitwidth}).")
    weight = torch.empty((N, K), dtype=torch.float16, device=qweight.device)
    with torch.cuda.device(qweight.device):
        st = _lib.lib().apg_dequant(qweight.data_ptr(), lut.data_ptr(), weight.data_ptr(), N, K, bitwidth,
                                    _stream(qweight.device))
    _lib.check(st, "anyprec_dequant")
    return weight


def prefill_supported(qweight: torch.Tensor, bitwidth: int) -> bool:
    """shapes the fused tensor-core prefill kernel takes (include/apgemv_b200.h, apg_prefill_gemm)"""
    return 2 <= bitwidth <= 4 and (qweight.size(2) * 32) % 256 == 0


# Measured cross-over (B200, profiles/r2_prefill_fused_vs_dequant_matmul.jsonl): the fused kernel dequantises every weight
# once per 256-token tile, the reference route once per call but pays an fp16 [N, K] write + read.  Fused wins while that
# HBM round trip dominates, i.e. up to these token counts: (N*K < 40M, N*K >= 40M) per bit-width.
_FUSED_MAX_TOKENS = {2: (128, 1024), 3: (64, 512), 4: (32, 128)}


def prefill_prefers_fused(qweight: torch.Tensor, bitwidth: int, tokens: int) -> bool:
    """True when `anyprec_prefill_gemm` is the faster route for `tokens` rows of input (else: anyprec_dequant + matmul)"""
    if not prefill_supported(qweight, bitwidth):
        return False
    N, K = qweight.size(1), qweight.size(2) * 32
    return tokens <= _FUSED_MAX_TOKENS[bitwidth][1 if N * K >= 40_000_000 else 0]


_workspaces: dict = {}  # (device index, stream) -> fp32 scratch of the split-K partial sums, grown on demand


def anyprec_prefill_gemm(input: torch.Tensor, qweight: torch.Tensor, lut: torch.Tensor, bitwidth: int) -> torch.Tensor:
    """x [..., K] fp16 -> x @ dequant(qweight, lut).T [..., N] fp16 in ONE kernel (csrc/prefill_tc.cuh): the fused form of
    the reference's prefill branch `anyprec_dequant` + `torch.matmul` (inference/ap_gemv/APLinear.py:35-38)."""
    _req(input.is_cuda and qweight.is_cuda and lut.is_cuda and input.device == qweight.device == lut.device,
         "input, qweight and lut must be on the same GPU.")
    _req(input.dtype == torch.float16, "input must be float16.")
    _req(qweight.dtype == torch.int32 and qweight.dim() == 3 and qweight.is_contiguous(),
         "qweight tensor must be a contiguous int tensor of shape (bitwidth, output_feat, input_feat / 32).")
    _req(qweight.size(0) >= bitwidth, "qweight has fewer bit-planes than bitwidth.")
    N, K = qweight.size(1), qweight.size(2) * 32
    _req(input.shape[-1] == K, f"input feature size {input.shape[-1]} does not match the weight ({K}).")
    _req(lut.dtype == torch.float16 and lut.is_contiguous() and tuple(lut.shape) == (N, 1 << bitwidth),
         f"lut tensor must be a contiguous float16 tensor of shape ({N}, {1 << bitwidth}).")
    _req(prefill_supported(qweight, bitwidth), "the fused prefill kernel needs bits in 2..4 and K % 256 == 0.")
    x2 = input.reshape(-1, K)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    T = x2.shape[0]
    out = torch.empty((T, N), dtype=torch.float16, device=input.device)
    L = _lib.lib()
    with torch.cuda.device(input.device):
        stream = _stream(input.device)
        plan = (ctypes.c_uint32 * 8)()
        need = ctypes.c_uint64(0)
        sms = torch.cuda.get_device_properties(input.device).multi_processor_count
        _lib.check(L.apg_prefill_plan(T, N, K, bitwidth, sms, plan, ctypes.byref(need)), "apg_prefill_plan")
        ws = None
        if need.value:
            key = (input.device.index, stream)
            ws = _workspaces.get(key)
            if ws is None or ws.numel() * 4 < need.value:
                ws = _workspaces[key] = torch.empty((need.value + 3) // 4, dtype=torch.float32, device=input.device)
        st = L.apg_prefill_gemm(x2.data_ptr(), out.data_ptr(), qweight.data_ptr(), lut.data_ptr(), T, N, K, bitwidth,
                                ws.data_ptr() if ws is not None else None, need.value, stream)
    _lib.check(st, "anyprec_prefill_gemm")
    return out.reshape(*input.shape[:-1], N)


def lutgemm_gemv(*args, **kwargs):
    """Present only so that `from plugin import *` of the reference finds the symbol; the LUT-GEMM (BCQ)
    format is outside this path's scope."""
    raise NotImplementedError("lutgemm_gemv (BCQ format) is out of scope of the Any-Precision path")
