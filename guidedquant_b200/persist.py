"""Host side of the persistent token kernel (csrc/apgemv_persist.cuh, C-ABI apg_persist_* in include/apgemv_b200.h).

A `PersistentProgram` is a list of dependent jobs — fused Any-Precision GEMVs, attention, embedding-row packing, all-reduce
finishers — that ONE cooperative kernel launch executes per token.  Activation vectors between jobs are "LL buffers":
8-byte packets (half2, tag) that consumers spin on, so there is no grid barrier and no kernel boundary between the ~160
dependent steps of a token (the reference runs one launch per op under torch.compile CUDA graphs, generate.py:330-336).

    prog = PersistentProgram(bits, device)
    x = prog.buffer(dim); qkv = prog.buffer(n_qkv) ...
    prog.pack(src_rows, x, row_index=token)                    # x := emb[token]
    prog.gemv(x, qweight, lut, qkv, norm_w=w, eps=1e-5)        # qkv := W . rmsnorm(x)
    prog.attn(qkv, rope_cs, k_cache, v_cache, att, H, Hkv, S, scale)
    ...
    prog.finalize(); prog.launch(pos)                          # per token (CUDA-graph capturable)
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

PF_NORM, PF_RESIDUAL, PF_GLU, PF_PUSH = 1, 4, 8, 16
_NO_TAG = 0xFFFFFFFF


class LLBuf:
    """activation vector of n halfs as n/2 (half2, tag) packets; `tag` = index of the job that wrote it last"""

    def __init__(self, n: int, device):
        assert n % 2 == 0
        self.n = n
        self.t = torch.zeros((n // 2, 2), dtype=torch.int32, device=device)
        self.tag: int | None = None

    def ptr(self) -> int:
        return self.t.data_ptr()


class PersistentProgram:
    def __init__(self, bits: int, device=None):
        self.L = _lib.lib()
        self.bits = bits
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.job_bytes = int(self.L.apg_persist_job_bytes())
        self.host = bytearray()
        self.n_jobs = 0
        self.n_gemv = 0
        self.max_k = 128                # largest GEMV input length: sizes the kernel's shared-memory x staging
        self.keep: list = []            # tensors the job table points to
        self.jobs_dev: torch.Tensor | None = None
        self.epoch = torch.zeros(2, dtype=torch.int32, device=self.device)   # token counter (packets carry epoch*n_jobs + job)
        # watchdog word in pinned HOST memory (device-mapped): still readable after a device-side trap killed the context
        self.err = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.done = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.prof: torch.Tensor | None = None   # enable_profile(): int64 [SMs, n_jobs, 4] clock stamps of the last launch

    # ------------------------------------------------------------------ buffers
    def buffer(self, n_halfs: int) -> LLBuf:
        b = LLBuf(n_halfs, self.device)
        self.keep.append(b.t)
        return b

    def _append(self, raw: ctypes.Array) -> int:
        self.host += bytes(raw)
        self.n_jobs += 1
        return self.n_jobs - 1

    @staticmethod
    def _p(t):
        return None if t is None else (t.ptr() if isinstance(t, LLBuf) else t.data_ptr())

    # ------------------------------------------------------------------ jobs
    def gemv(self, x: LLBuf, qweight: torch.Tensor, lut: torch.Tensor, out: LLBuf | None, norm_w: torch.Tensor | None = None,
             eps: float = 1e-5, residual: LLBuf | None = None, glu: bool = False, out_plain: torch.Tensor | None = None,
             push=None) -> int:
        """out := W . f(x) (+ residual)   W given by (qweight [bits, N, K/32], lut [N, 2^bits]); f = RMSNorm * norm_w if given.
        glu: rows of W are interleaved (gate_i, up_i) and out (N/2 halfs) receives silu(y[2i]) * y[2i+1].
        push = (world, rank, peer_ptrs): K-sharded Linear, fp32 partial sums pushed to every rank's receive buffer."""
        bits, N, K = self.bits, qweight.shape[1], qweight.shape[2] * 32
        assert qweight.dtype == torch.int32 and qweight.shape[0] >= bits and qweight.is_contiguous()
        assert lut.dtype == torch.float16 and tuple(lut.shape) == (N, 1 << bits) and lut.is_contiguous()
        assert x.tag is not None, "the input vector has no producer job"
        assert x.n >= K, (x.n, K)
        flags = (PF_NORM if norm_w is not None else 0) | (PF_RESIDUAL if residual is not None else 0) | (PF_GLU if glu else 0)
        world = rank = 0
        peers = None
        if push is not None:
            world, rank, ptrs = push
            peers = (ctypes.c_void_p * 8)(*ptrs)
            flags |= PF_PUSH
        if out is not None:
            assert out.n >= (N // 2 if glu else N)
        if residual is not None:
            assert residual.tag is not None and residual.n >= N
        raw = (ctypes.c_uint8 * self.job_bytes)()
        idx = self.n_jobs
        st = self.L.apg_persist_job_gemv(raw, N, K, bits, self.sms, flags, x.ptr(), qweight.data_ptr(), lut.data_ptr(),
                                         self._p(out), self._p(out_plain), self._p(norm_w), float(eps), self._p(residual),
                                         world, rank, peers, x.tag, residual.tag if residual is not None else _NO_TAG, idx)
        _lib.check(st, f"apg_persist_job_gemv N={N} K={K} bits={bits}")
        self.keep += [qweight, lut, norm_w, out_plain]
        self.n_gemv += 1
        self.max_k = max(self.max_k, K)
        if out is not None:
            out.tag = idx
        return self._append(raw)

    def attn(self, qkv: LLBuf, rope_cs: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, out: LLBuf, H: int, Hkv: int,
             S: int, scale: float, out_plain: torch.Tensor | None = None) -> int:
        assert qkv.tag is not None and H <= self.sms, "one CTA per head"
        assert rope_cs.dtype == torch.float16 and tuple(rope_cs.shape) == (S, 128) and rope_cs.is_contiguous()
        raw = (ctypes.c_uint8 * self.job_bytes)()
        idx = self.n_jobs
        st = self.L.apg_persist_job_attn(raw, qkv.ptr(), rope_cs.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), out.ptr(),
                                         self._p(out_plain), H, Hkv, S, float(scale), qkv.tag, idx)
        _lib.check(st, "apg_persist_job_attn")
        self.keep += [rope_cs, k_cache, v_cache, out_plain]
        out.tag = idx
        return self._append(raw)

    def pack(self, src_rows: torch.Tensor, out: LLBuf, row_index: torch.Tensor | None = None, out_plain: torch.Tensor | None = None) -> int:
        """out := src_rows[row_index[0]] (or row 0) as packets: the embedding row / an externally provided input vector"""
        assert src_rows.dtype == torch.float16 and src_rows.is_contiguous()
        n = src_rows.shape[-1]
        n_rows = src_rows.numel() // n
        raw = (ctypes.c_uint8 * self.job_bytes)()
        idx = self.n_jobs
        st = self.L.apg_persist_job_pack(raw, src_rows.data_ptr(), self._p(row_index), n, n_rows, out.ptr(), self._p(out_plain), idx)
        _lib.check(st, "apg_persist_job_pack")
        self.keep += [src_rows, row_index, out_plain]
        out.tag = idx
        return self._append(raw)

    def reduce(self, recv_ptr: int, tag_push: int, N: int, world: int, out: LLBuf, residual: LLBuf | None = None,
               out_plain: torch.Tensor | None = None) -> int:
        """out := sum over ranks of the fp32 packets pushed by job `tag_push` (+ residual), rounded once"""
        raw = (ctypes.c_uint8 * self.job_bytes)()
        idx = self.n_jobs
        st = self.L.apg_persist_job_reduce(raw, recv_ptr, N, world, self._p(residual), out.ptr(), self._p(out_plain), tag_push,
                                           residual.tag if residual is not None else _NO_TAG, idx)
        _lib.check(st, "apg_persist_job_reduce")
        self.keep += [out_plain]
        out.tag = idx
        return self._append(raw)

    # ------------------------------------------------------------------ run
    def finalize(self):
        assert self.n_jobs > 0
        self.jobs_dev = torch.frombuffer(bytearray(self.host), dtype=torch.uint8).clone().to(self.device)
        return self

    def launch(self, pos: torch.Tensor | None = None, bump_epoch: bool = True, cooperative: bool = True):
        """one launch = one pass over the job list; asynchronous on the current stream, CUDA-graph capturable"""
        if self.jobs_dev is None:
            self.finalize()
        st = self.L.apg_persist_launch(self.jobs_dev.data_ptr(), self.n_jobs, self.bits, self.max_k, self.epoch.data_ptr(),
                                       pos.data_ptr() if pos is not None else None, self.err.data_ptr(), self.done.data_ptr(),
                                       1 if bump_epoch else 0, 0 if cooperative else 1,
                                       self.prof.data_ptr() if self.prof is not None else None, torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "apg_persist_launch")

    def enable_profile(self):
        """debug aid: per-CTA, per-job clock64 stamps (job start, x in registers, stages done, job end) of every launch"""
        self.prof = torch.zeros((self.sms, self.n_jobs, 4), dtype=torch.int64, device=self.device)
        return self

    def check(self):
        """synchronising: raise if a device-side watchdog fired (a wait that never completed)"""
        e = int(self.err[0]) & 0xFFFFFFFF
        if e:
            raise RuntimeError(f"persistent kernel watchdog: code {e & 0xff} (1/2 packet wait, 3 weight stage, 4 ring slot, "
                               f"5 bad job, 6 smem alignment), cta {(e >> 8) & 0xfff}, thread {e >> 20}")
