"""Peer-memory plumbing for the fused one-shot all-reduce of the K-sharded Linears.

One symmetric allocation per process group (torch.distributed._symmetric_memory: every rank maps every peer's
buffer), carved into `sites`: one receive buffer fp32 [world][n_max] + one arrival counter per all-reduce site of the
decode graph (2 per transformer block), so that a site is only reused one token later and no double buffering is
needed.  The GEMV kernel pushes into the peers' buffers (apg_gemv_fused_push); apg_allreduce_finish consumes.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib


class PushAllReduce:
    def __init__(self, n_sites: int, n_max: int, group=None, device=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        assert 2 <= self.world <= 8
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.n_sites, self.n_max = n_sites, n_max
        self.site_bytes = self.world * n_max * 8          # uint2 (value, epoch) packets [world][n_max]
        self.buf = symm_mem.empty(n_sites * self.site_bytes, dtype=torch.uint8, device=self.device)
        self.buf.zero_()                                   # epoch 0 everywhere: the first use pushes epoch 1
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.peer_base = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(self.peer_base) == self.world and self.peer_base[self.rank] == self.buf.data_ptr()
        self.epoch = torch.zeros(n_sites, dtype=torch.int32, device=self.device)     # per-site epoch counters
        self.scratch = torch.zeros(n_max, dtype=torch.float32, device=self.device)
        self.done = torch.zeros(n_sites, dtype=torch.int32, device=self.device)      # CTA tickets of the finisher kernels
        torch.cuda.synchronize()
        dist.barrier(self.group)
        vp = ctypes.c_void_p
        self._recv_arr = [(vp * self.world)(*[b + s * self.site_bytes for b in self.peer_base]) for s in range(n_sites)]

    def site_ptrs(self, site: int):
        """host array of the peers' base pointers of a site (for exchanges other than the GEMV all-reduce)."""
        return self._recv_arr[site]

    def epoch_ptr(self, site: int) -> int:
        return self.epoch.data_ptr() + 4 * site

    def gemv_push(self, site: int, x, qweight, lut, N: int, K: int, bits: int, norm=None, eps: float = 1e-5,
                  silu_mul: int = 0, flags: int = 0):
        """K-shard GEMV whose epilogue pushes (fp32 partial sum, epoch) packets into every peer's receive slot."""
        assert N == self.n_max, "sites are laid out [world][n_max]"
        st = _lib.lib().apg_gemv_fused_push(
            x.data_ptr(), qweight.data_ptr(), lut.data_ptr(), N, K, bits, norm.data_ptr() if norm is not None else None,
            eps, silu_mul, self.world, self.rank, self._recv_arr[site], self.epoch.data_ptr() + 4 * site,
            self.scratch.data_ptr(), flags, torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "apg_gemv_fused_push")

    def finish(self, site: int, out, N: int, residual=None, flags: int = 0):
        """poll all ranks' packets of this site, sum in rank order (+ residual), round to fp16 into `out`."""
        st = _lib.lib().apg_allreduce_finish(self.peer_base[self.rank] + site * self.site_bytes,
                                             self.epoch.data_ptr() + 4 * site, self.done.data_ptr() + 4 * site,
                                             residual.data_ptr() if residual is not None else None, out.data_ptr(), N,
                                             self.world, flags, torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "apg_allreduce_finish")
