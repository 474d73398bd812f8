"""Loader/builder of the C-ABI library libapgemv_b200.so (include/apgemv_b200.h).

The library is built IN-TREE with nvcc for sm_100a only (guidedquant_b200/lib/), so it travels with
the working tree.  There is no fallback: if the library cannot be built or loaded every entry point
of this package raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
_CSRC = os.path.join(_PKG, "csrc")
_LIBDIR = os.path.join(_PKG, "lib")
LIB_PATH = os.path.join(_LIBDIR, "libapgemv_b200.so")
_SOURCES = ["apgemv_capi.cu", "decode_capi.cu", "persist_capi.cu", "prefill_capi.cu"]
_HEADERS = ["apgemv_common.cuh", "apgemv_fast.cuh", "apgemv_generic.cuh", "apgemv_wide.cuh", "decode_kernels.cuh",
            "apgemv_persist.cuh", "prefill_tc.cuh"]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC", "-t", "4",
    "-gencode", "arch=compute_100a,code=sm_100a",
]

EXPORTS = (
    "apg_version", "apg_status_string", "apg_last_cuda_error", "apg_gemv", "apg_gemv_ex", "apg_dequant",
    "apg_round_f32_to_f16", "apg_prefetch_hint", "apg_gemv_fused", "apg_gemv_fused_push", "apg_allreduce_finish",
    "apd_embed", "apd_attn_decode", "apd_lm_head", "apd_argmax_advance", "apd_argmax_advance_tp",
    "apd_sample_topk_advance", "apd_sample_topk_advance_tp", "apg_plan_fast",
    "apg_persist_job_bytes", "apg_persist_smem", "apg_persist_job_gemv", "apg_persist_job_attn", "apg_persist_job_pack",
    "apg_persist_job_reduce", "apg_persist_launch", "apg_prefill_plan", "apg_prefill_gemm",
)


def _source_hash() -> str:
    import hashlib

    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for d in [os.path.join(_CSRC, f) for f in sorted(_SOURCES + _HEADERS)] + [os.path.join(_ROOT, "include", "apgemv_b200.h"),
                                                                            os.path.join(_ROOT, "include", "apdecode_b200.h")]:
        if os.path.exists(d):
            h.update(os.path.basename(d).encode())
            h.update(open(d, "rb").read())
    return h.hexdigest()


_HASH_PATH = LIB_PATH + ".srchash"


def _stale() -> bool:
    """the library is stale when it was built from other sources than the ones in the tree (content hash, not mtimes:
    a snapshot copy of the tree does not preserve them)"""
    if not os.path.exists(LIB_PATH) or not os.path.exists(_HASH_PATH):
        return True
    return open(_HASH_PATH).read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> guidedquant_b200/lib/libapgemv_b200.so
    Built into a temporary file and renamed into place under a file lock, so concurrent ranks (torchrun on a fresh tree)
    neither write the same output nor load a partially written library."""
    import fcntl

    if not force and not _stale():
        return LIB_PATH
    os.makedirs(_LIBDIR, exist_ok=True)
    with open(os.path.join(_LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():  # another process built it while we waited
                return LIB_PATH
            srcs = [os.path.join(_CSRC, f) for f in _SOURCES if os.path.exists(os.path.join(_CSRC, f))]
            tmp = LIB_PATH + f".tmp{os.getpid()}"
            cmd = ["nvcc", *NVCC_FLAGS, "-I" + os.path.join(_ROOT, "include"), "-I" + _CSRC, "-o", tmp, *srcs]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed building libapgemv_b200.so:\n" + res.stdout + res.stderr)
            if verbose:
                print(res.stderr)
            os.replace(tmp, LIB_PATH)
            with open(_HASH_PATH + ".tmp", "w") as f:
                f.write(_source_hash())
            os.replace(_HASH_PATH + ".tmp", _HASH_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    """ctypes handle with typed signatures.  Raises (never falls back) if the library is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if _stale():  # missing, or built from other sources than the tree holds (an ABI mismatch waiting to happen)
        try:
            build()
        except Exception as e:  # no nvcc on this box and no matching prebuilt library
            raise RuntimeError(
                f"libapgemv_b200.so is missing or stale ({LIB_PATH}) and could not be built: {e}. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` where nvcc is available."
            ) from e
    L = ctypes.CDLL(LIB_PATH)
    if L.apg_version() != APG_VERSION:
        raise RuntimeError(f"libapgemv_b200.so reports version {L.apg_version()}, this package expects {APG_VERSION}")
    vp, u32, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int
    L.apg_version.restype = i32
    L.apg_status_string.restype = ctypes.c_char_p
    L.apg_status_string.argtypes = [i32]
    L.apg_last_cuda_error.restype = i32
    L.apg_gemv.restype = i32
    L.apg_gemv.argtypes = [vp, vp, vp, vp, u32, u32, u32, i32, vp]
    L.apg_gemv_ex.restype = i32
    L.apg_gemv_ex.argtypes = [vp, vp, vp, vp, vp, u32, u32, u32, i32, u32, i32, vp]
    L.apg_dequant.restype = i32
    L.apg_dequant.argtypes = [vp, vp, vp, u32, u32, i32, vp]
    L.apg_prefill_plan.restype = i32
    L.apg_prefill_plan.argtypes = [u32, u32, u32, i32, i32, ctypes.POINTER(u32), ctypes.POINTER(ctypes.c_uint64)]
    L.apg_prefill_gemm.restype = i32
    L.apg_prefill_gemm.argtypes = [vp, vp, vp, vp, u32, u32, u32, i32, vp, ctypes.c_uint64, vp]
    L.apg_gemv_fused.restype = i32
    L.apg_gemv_fused.argtypes = [vp, vp, vp, vp, vp, u32, u32, i32, vp, ctypes.c_float, i32, vp, u32, vp]
    f32 = ctypes.c_float
    L.apd_embed.restype = i32
    L.apd_embed.argtypes = [vp, vp, vp, u32, u32, u32, vp]
    L.apd_attn_decode.restype = i32
    L.apd_attn_decode.argtypes = [vp, vp, vp, vp, vp, vp, vp, u32, u32, u32, u32, f32, u32, vp]
    L.apd_lm_head.restype = i32
    L.apd_lm_head.argtypes = [vp, vp, f32, vp, vp, u32, u32, vp, vp, ctypes.POINTER(u32), u32, u32, vp]
    L.apd_argmax_advance_tp.restype = i32
    L.apd_argmax_advance_tp.argtypes = [vp, vp, u32, u32, u32, ctypes.POINTER(vp), vp, vp, vp, vp, u32, u32, vp]
    L.apd_argmax_advance.restype = i32
    L.apd_argmax_advance.argtypes = [vp, vp, u32, vp, vp, vp, u32, u32, vp]
    L.apg_plan_fast.restype = i32
    L.apg_plan_fast.argtypes = [u32, u32, i32, i32, i32, ctypes.POINTER(u32 * 16)]
    L.apd_sample_topk_advance.restype = i32
    L.apd_sample_topk_advance.argtypes = [vp, u32, f32, u32, vp, vp, vp, vp, u32, u32, vp]
    L.apd_sample_topk_advance_tp.restype = i32
    L.apd_sample_topk_advance_tp.argtypes = [vp, u32, f32, u32, vp, u32, u32, ctypes.POINTER(vp), u32, vp, vp, vp, vp, u32, u32, vp]
    L.apg_gemv_fused_push.restype = i32
    L.apg_gemv_fused_push.argtypes = [vp, vp, vp, u32, u32, i32, vp, ctypes.c_float, i32, u32, u32,
                                      ctypes.POINTER(vp), vp, vp, u32, vp]
    L.apg_allreduce_finish.restype = i32
    L.apg_allreduce_finish.argtypes = [vp, vp, vp, vp, vp, u32, u32, u32, vp]
    L.apg_prefetch_hint.restype = i32
    L.apg_prefetch_hint.argtypes = [vp, ctypes.c_uint64]
    L.apg_round_f32_to_f16.restype = i32
    L.apg_round_f32_to_f16.argtypes = [vp, vp, u32, vp]
    L.apg_persist_job_bytes.restype = u32
    L.apg_persist_job_bytes.argtypes = []
    L.apg_persist_smem.restype = i32
    L.apg_persist_smem.argtypes = [i32, u32, ctypes.POINTER(u32), ctypes.POINTER(u32)]
    L.apg_persist_job_gemv.restype = i32
    L.apg_persist_job_gemv.argtypes = [vp, u32, u32, i32, i32, u32, vp, vp, vp, vp, vp, vp, f32, vp, u32, u32, vp, u32, u32, u32]
    L.apg_persist_job_attn.restype = i32
    L.apg_persist_job_attn.argtypes = [vp, vp, vp, vp, vp, vp, vp, u32, u32, u32, f32, u32, u32]
    L.apg_persist_job_pack.restype = i32
    L.apg_persist_job_pack.argtypes = [vp, vp, vp, u32, u32, vp, vp, u32]
    L.apg_persist_job_reduce.restype = i32
    L.apg_persist_job_reduce.argtypes = [vp, vp, u32, u32, vp, vp, vp, u32, u32, u32]
    L.apg_persist_launch.restype = i32
    L.apg_persist_launch.argtypes = [vp, u32, i32, u32, vp, vp, vp, vp, i32, u32, vp, vp]
    _lib = L
    return L


def check(status: int, what: str) -> None:
    if status != 0:
        L = lib()
        msg = L.apg_status_string(status).decode()
        extra = f" (cudaError {L.apg_last_cuda_error()})" if status == 6 else ""
        raise RuntimeError(f"{what}: {msg}{extra}")


APG_VERSION = 202  # include/apgemv_b200.h: major*100 + minor
APG_FLAG_REF_ORDER = 0x1
APG_FLAG_GENERIC = 0x2
APG_FLAG_PDL = 0x4
