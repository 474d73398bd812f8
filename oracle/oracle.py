"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

numpy/ctypes front-end of the CPU oracle (oracle/apgemv_oracle.c) that restates the reference's
Any-Precision LUT GEMV path (inference/ap_gemv/anyprec.cu, any_precision/quantization/pack.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; guidedquant_b200/ never does.  Parity status: PINNED (see the header of
apgemv_oracle.c and DESIGN.md §3).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "apgemv_oracle.c")
_SO = os.path.join(_HERE, "libapgemv_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc the C restatement into oracle/libapgemv_oracle.so (a second or two)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-o", _SO, _SRC, "-lm"]
        )
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.apo_f64_to_f16.restype = ctypes.c_uint16
        _lib.apo_f64_to_f16.argtypes = [ctypes.c_double]
        _lib.apo_f16_to_f64.restype = ctypes.c_double
        _lib.apo_f16_to_f64.argtypes = [ctypes.c_uint16]
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _u16(a) -> np.ndarray:
    """fp16 array (numpy float16 or uint16 bit patterns) -> contiguous uint16 bit patterns."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.float16:
        return a.view(np.uint16)
    assert a.dtype == np.uint16, a.dtype
    return a


# ---------------------------------------------------------------------------------- layout
def pack(idx: np.ndarray, bits: int) -> np.ndarray:
    """uint8 idx [N,K] -> int32 qweight [bits,N,K/32]  (pack.py:304-321)."""
    idx = np.ascontiguousarray(idx, dtype=np.uint8)
    N, K = idx.shape
    assert K % 32 == 0
    q = np.zeros((bits, N, K // 32), dtype=np.uint32)
    rc = lib().apo_pack(_p(idx), ctypes.c_uint32(N), ctypes.c_uint32(K), ctypes.c_int(bits), _p(q))
    assert rc == 0
    return q.view(np.int32)


def unpack(qweight: np.ndarray, bits: int | None = None) -> np.ndarray:
    """int32 qweight [bits,N,K/32] -> uint8 idx [N,K]  (pack.py:324-347)."""
    q = np.ascontiguousarray(qweight).view(np.uint32)
    b, N, words = q.shape
    bits = b if bits is None else bits
    q = np.ascontiguousarray(q[:bits])
    K = words * 32
    idx = np.zeros((N, K), dtype=np.uint8)
    rc = lib().apo_unpack(_p(q), ctypes.c_uint32(N), ctypes.c_uint32(K), ctypes.c_int(bits), _p(idx))
    assert rc == 0
    return idx


# --------------------------------------------------------------------------------- dequant
def dequant(qweight: np.ndarray, lut: np.ndarray, bits: int) -> np.ndarray:
    """W[n,k] = lut[n, idx[n,k]] as float16 [N,K]  (anyprec.cu:294-359, eval.py:102-108)."""
    q = np.ascontiguousarray(np.ascontiguousarray(qweight).view(np.uint32)[:bits])
    _, N, words = q.shape
    K = words * 32
    l = _u16(lut)
    assert l.shape == (N, 1 << bits), (l.shape, N, bits)
    scratch = np.zeros((N, K), dtype=np.uint8)
    W = np.zeros((N, K), dtype=np.uint16)
    rc = lib().apo_dequant(_p(q), _p(l), ctypes.c_uint32(N), ctypes.c_uint32(K), ctypes.c_int(bits),
                           _p(scratch), _p(W))
    assert rc == 0
    return W.view(np.float16)


# ------------------------------------------------------------------------------------ GEMV
def _wx(W, x):
    W = _u16(W)
    x = _u16(x)
    N, K = W.shape
    x = x.reshape(-1, K)
    return W, x, x.shape[0], N, K


def gemv_f64(W: np.ndarray, x: np.ndarray) -> np.ndarray:
    """fp64 truth, [M,N] float64."""
    W, x, M, N, K = _wx(W, x)
    y = np.zeros((M, N), dtype=np.float64)
    lib().apo_gemv_f64(_p(W), _p(x), ctypes.c_uint32(M), ctypes.c_uint32(N), ctypes.c_uint32(K), _p(y))
    return y


def gemv_ref_order_f16(W: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Bit-exact emulation of matmul_kbit_32's fp16 arithmetic (anyprec.cu:424-541), [M,N] float16."""
    W, x, M, N, K = _wx(W, x)
    y = np.zeros((M, N), dtype=np.uint16)
    lib().apo_gemv_ref_order_f16(_p(W), _p(x), ctypes.c_uint32(M), ctypes.c_uint32(N), ctypes.c_uint32(K), _p(y))
    return y.view(np.float16)


def gemv_dequant_matmul_f16(W: np.ndarray, x: np.ndarray) -> np.ndarray:
    """BASELINE config 0 (APLinear.gemm, APLinear.py:35-38): fp32-accumulated fp16 matmul, [M,N] float16."""
    W, x, M, N, K = _wx(W, x)
    y = np.zeros((M, N), dtype=np.uint16)
    lib().apo_gemv_dequant_matmul_f16(_p(W), _p(x), ctypes.c_uint32(M), ctypes.c_uint32(N), ctypes.c_uint32(K), _p(y))
    return y.view(np.float16)


def ap_linear_forward(qweight: np.ndarray, lut: np.ndarray, x: np.ndarray, bits: int, mode: str = "f64") -> np.ndarray:
    """Whole path: unpack -> gather -> GEMV.  mode in {"f64", "ref_f16", "matmul_f16"}."""
    W = dequant(qweight, lut, bits)
    return {"f64": gemv_f64, "ref_f16": gemv_ref_order_f16, "matmul_f16": gemv_dequant_matmul_f16}[mode](W, x)


# ----------------------------------------------------------------- synthetic inputs
def synth_layer(N: int, K: int, bits: int, seed: int = 0, M: int = 1, sorted_lut: bool = False):
    """Seeded synthetic (idx, qweight, lut, x): idx uniform, lut ~ fp16 N(0,0.02) unsorted (LNQ-like),
    x ~ fp16 N(0,1)."""
    rng = np.random.default_rng(1000 * bits + seed)
    idx = rng.integers(0, 1 << bits, size=(N, K), dtype=np.uint8)
    lut = (rng.standard_normal((N, 1 << bits)) * 0.02).astype(np.float16)
    if sorted_lut:
        lut = np.sort(lut, axis=1)
    x = np.random.default_rng(7 + seed).standard_normal((M, 1, K)).astype(np.float16)
    return idx, pack(idx, bits), lut, x


# ------------------------------------------------- CPU baseline of the reference path (timed leg)
def cpu_reference_linear(qweight: np.ndarray, lut: np.ndarray, x: np.ndarray, bits: int, threads: int | None = None,
                         dequant_each_call: bool = True, repeats: int = 1, matmul_dtype: str = "f16"):
    """The reference's own CPU-runnable arm (BASELINE.json configs[0]): dequant -> fp16 torch.matmul
    on the host cores, mirroring APLinear.gemm (APLinear.py:35-38: anyprec_dequant then
    torch.matmul(x, W.T)).  Returns (y float16 [M,N], seconds per call, threads used).
    The dequant is the C oracle's unpack + gather, row blocks dealt to `threads` host threads."""
    import torch

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    q = np.ascontiguousarray(qweight)
    K = q.shape[2] * 32
    xt = torch.from_numpy(np.ascontiguousarray(x).reshape(-1, K))
    lut = np.ascontiguousarray(lut)

    def deq():
        # the C oracle's dequant (unpack + gather) is single-threaded; deal row blocks to `threads` host threads (ctypes
        # releases the GIL), which is what a threaded CPU implementation of anyprec_dequant would do
        N = q.shape[1]
        nblk = max(1, min(threads, N // 64))
        bounds = [N * i // nblk for i in range(nblk + 1)]
        out = np.empty((N, K), dtype=np.float16)

        def one(i):
            r0, r1 = bounds[i], bounds[i + 1]
            out[r0:r1] = dequant(np.ascontiguousarray(q[:bits, r0:r1]), lut[r0:r1], bits)

        if nblk == 1:
            one(0)
        else:
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(nblk) as ex:
                list(ex.map(one, range(nblk)))
        return torch.from_numpy(out)

    W = None if dequant_each_call else deq()
    if matmul_dtype == "f32":  # resident fp32 copy of the dense weight ("fp16 and fp32")
        xt = xt.float()
        if W is not None:
            W = W.float()
    t0 = time.perf_counter()
    for _ in range(repeats):
        Wc = deq() if dequant_each_call else W
        if matmul_dtype == "f32" and dequant_each_call:
            Wc = Wc.float()
        try:
            y = torch.matmul(xt, Wc.T)
        except RuntimeError:  # no fp16 CPU matmul in this torch build
            y = torch.matmul(xt.float(), Wc.float().T).half()
    y = y.half()
    dt = (time.perf_counter() - t0) / repeats
    return y.numpy(), dt, threads
