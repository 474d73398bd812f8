"""Test helper: ctypes access to the UNMODIFIED reference kernels compiled into oracle/_ref/
(oracle/build_ref.sh).  Test infrastructure only."""
import ctypes
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libapgemv_ref.so")
_ref = None


def available() -> bool:
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(REF_SO)
        vp, u32, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int
        _ref.ref_anyprec_gemv.restype = i32
        _ref.ref_anyprec_gemv.argtypes = [vp, vp, vp, vp, u32, u32, u32, i32, vp]
        _ref.ref_anyprec_dequant.restype = i32
        _ref.ref_anyprec_dequant.argtypes = [vp, vp, vp, u32, u32, i32, vp]
    return _ref


def ref_gemv(x, qweight, lut, bits):
    """reference anyprec_matmul on [M,1,K] fp16 x -> [M,1,N] fp16 (needs N % 4 == 0; N % 16 == 0 for M > 1)."""
    M, _, K = x.shape
    N = qweight.shape[1]
    out = torch.zeros((M, 1, N), dtype=torch.float16, device=x.device)
    rc = ref().ref_anyprec_gemv(x.data_ptr(), out.data_ptr(), qweight.data_ptr(), lut.data_ptr(), M, N, K, bits,
                                torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    return out


def ref_dequant(qweight, lut, bits):
    N, K = qweight.shape[1], qweight.shape[2] * 32
    w = torch.empty((N, K), dtype=torch.float16, device=qweight.device)
    rc = ref().ref_anyprec_dequant(qweight.data_ptr(), lut.data_ptr(), w.data_ptr(), N, K, bits,
                                   torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    return w
