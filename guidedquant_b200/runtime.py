"""Decode-step runtime around the GEMV: the per-token chain of APLinear GEMVs of a Llama model
(wqkv -> wo -> w1w3 -> w2 per block, the fused shapes of the reference's inference/model.py:176-183,
248-253), launched through the C-ABI with programmatic dependent launch and replayed as ONE CUDA graph
(the reference gets its graphs from torch.compile(mode="max-autotune"), generate.py:330-336).

`ApGemvChain` is the public API bench.py measures:
    chain = ApGemvChain("llama3-8b", bits=2)           # synthetic packed weights, resident in HBM
    chain.step()                                       # one token's 4*L GEMVs, inputs resident on device
    y = chain.step_host(x_pinned)                      # H2D(x) -> graph -> D2H(y): the end-to-end call
With torch.distributed initialised (world_size W > 1) the chain is Megatron-sharded:
wqkv / w1w3 split by output rows (no collective), wo / w2 split along K (re-packed shards) with ONE
all-reduce (NCCL, fp32 partial sums, rounded once) after each.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, ap_gemv
from . import pack as packmod

# (dim, n_layers, n_heads, n_kv_heads, intermediate, vocab) — inference/model.py:53-61 (+ Llama-3-70B, which the
# reference's table lacks)
MODEL_CONFIGS = {
    "llama3-8b": dict(dim=4096, n_layer=32, n_head=32, n_kv=8, inter=14336, vocab=128256),
    "llama2-7b": dict(dim=4096, n_layer=32, n_head=32, n_kv=32, inter=11008, vocab=32000),
    "llama2-70b": dict(dim=8192, n_layer=80, n_head=64, n_kv=8, inter=28672, vocab=32000),
    "llama3-70b": dict(dim=8192, n_layer=80, n_head=64, n_kv=8, inter=28672, vocab=128256),
    "tiny": dict(dim=1024, n_layer=2, n_head=8, n_kv=2, inter=2048, vocab=1024),
}


def linear_shapes(cfg: dict) -> dict:
    """fused Linear shapes (N, K) of one block (model.py:176-183, 248-253)."""
    hd = cfg["dim"] // cfg["n_head"]
    return {
        "wqkv": ((cfg["n_head"] + 2 * cfg["n_kv"]) * hd, cfg["dim"]),
        "wo": (cfg["dim"], cfg["dim"]),
        "w1w3": (2 * cfg["inter"], cfg["dim"]),
        "w2": (cfg["dim"], cfg["inter"]),
    }


# Default engine by world size.  Measured on B200 (profiles/r2_engines_by_world.txt), Llama-3-8B 2-bit tok/s, launches vs
# persistent: 1 GPU 800 vs 619; 2 GPUs 997 vs 856; 4 GPUs ~1100 vs 990; 8 GPUs 1174 vs 1017 (Llama-3-70B: 292 vs 223 at 4,
# 341 vs 283 at 8).  The persistent engine led at 4-8 GPUs only while the all-reduce finisher of the launches engine was a
# single CTA; since that kernel runs on several CTAs the per-launch engine wins everywhere and the persistent engine is
# opt-in (engine="persistent").
PERSISTENT_MIN_WORLD = 1 << 30


def persistent_supported(cfg: dict, bits: int, world: int = 1) -> bool:
    """the persistent token kernel takes K <= 16384 per GEMV job (one 1024-chunk per consumer warp) with 4 rows of all planes
    within a 32 KB stage; Llama-70B's w2 (K = 28672) on ONE GPU does not fit and runs on the per-launch engine"""
    if bits < 2 or bits > 4:
        return False
    for name, (N, K) in linear_shapes(cfg).items():
        k_local = K // world if name in ("wo", "w2") else K
        if k_local > 16384 or k_local % 128 or 4 * (k_local // 8) * bits > 32768:
            return False
    return True


def gemv_algo_bytes(N: int, K: int, bits: int, M: int = 1) -> int:
    """algorithmic bytes of one GEMV call: planes + LUT + x + y."""
    return bits * N * K // 8 + N * (1 << bits) * 2 + M * K * 2 + M * N * 2


@dataclass
class _Lin:
    name: str
    N: int          # local rows
    K: int          # local input features
    qweight: torch.Tensor
    lut: torch.Tensor
    k_shard: bool   # K-split: produces fp32 partial sums that must be all-reduced


class ApGemvChain:
    def __init__(self, model: str = "llama3-8b", bits: int = 2, device=None, seed: int = 0, n_layer: int | None = None,
                 pdl: bool = True, world_size: int = 1, rank: int = 0, process_group=None, ctas_per_sm: int = 0,
                 l2_prefetch: bool = False, collective: str = "push", engine: str | None = None):
        self.cfg = dict(MODEL_CONFIGS[model])
        # "persistent": the whole chain is ONE cooperative launch of the persistent token kernel (persist.py);
        # "launches": one PDL launch per Linear under a CUDA graph (the round-1 path)
        # default: per-launch kernels on one GPU (measured faster there: the hand-over between two Linears costs about the
        # same either way and the per-launch kernel keeps 18 warps per SM busy); the persistent kernel under
        # tensor parallelism, where the per-GPU Linears are small and the launch count is what bounds a token
        ok = world_size >= PERSISTENT_MIN_WORLD and collective == "push" and persistent_supported(self.cfg, bits, world_size)
        self.engine = engine or ("persistent" if ok else "launches")
        assert self.engine in ("launches", "persistent"), f"unknown engine {engine!r}"
        self.prog = None
        if n_layer is not None:
            self.cfg["n_layer"] = n_layer
        self.model, self.bits, self.pdl = model, bits, pdl
        self.world, self.rank, self.pg = world_size, rank, process_group
        self.ctas = ctas_per_sm
        self.l2_prefetch = l2_prefetch
        self.collective = collective if world_size > 1 else "none"   # "push": fused one-shot all-reduce; "nccl"
        self.push = None
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.shapes = linear_shapes(self.cfg)
        self.layers: list[list[_Lin]] = []
        g = torch.Generator(device=self.device).manual_seed(1234 + seed)  # same full-model weights on every rank
        W = self.world
        for li in range(self.cfg["n_layer"]):
            lins = []
            for name, (N, K) in self.shapes.items():
                k_shard = W > 1 and name in ("wo", "w2")
                n_shard = W > 1 and not k_shard
                # synthetic packed weights: uniform random bit-planes == uniform random indices; codebooks
                # N(0, 1/K) so activations keep unit scale through the chain
                if W == 1:
                    q = torch.randint(-2**31, 2**31 - 1, (bits, N, K // 32), dtype=torch.int32, device=self.device, generator=g)
                    lut = (torch.randn((N, 1 << bits), device=self.device, generator=g) / math.sqrt(K)).half()
                    lins.append(_Lin(name, N, K, q, lut, False))
                    continue
                # sharded: generate the FULL tensor deterministically, keep this rank's shard
                qf = torch.randint(-2**31, 2**31 - 1, (bits, N, K // 32), dtype=torch.int32, device=self.device, generator=g)
                lf = (torch.randn((N, 1 << bits), device=self.device, generator=g) / math.sqrt(K)).half()
                if n_shard:
                    n0, n1 = N * rank // W, N * (rank + 1) // W
                    lins.append(_Lin(name, n1 - n0, K, qf[:, n0:n1].contiguous(), lf[n0:n1].contiguous(), False))
                else:
                    k0, k1 = packmod.shard_bounds(K, W)[rank]
                    if k0 % 1024 == 0 and (k1 % 1024 == 0 or k1 == K):
                        qs = qf[:, :, k0 // 32:k1 // 32].contiguous()
                    else:  # the cut falls inside a 1024-chunk: unpack -> slice -> re-pack
                        qs = packmod.shard_k_torch(qf, k0, k1)
                    lins.append(_Lin(name, N, k1 - k0, qs, lf, True))
                del qf, lf
            self.layers.append(lins)
        d = self.cfg["dim"]
        hd = d // self.cfg["n_head"]
        self.x_in = torch.zeros((1, 1, d), dtype=torch.float16, device=self.device)
        # activation buffers (full-size on every rank; sharded Linears write / read their slice)
        self.buf = {
            "h": torch.zeros((1, 1, d), dtype=torch.float16, device=self.device),
            "qkv": torch.zeros((1, 1, self.shapes["wqkv"][0]), dtype=torch.float16, device=self.device),
            "o": torch.zeros((1, 1, d), dtype=torch.float16, device=self.device),
            "gu": torch.zeros((1, 1, self.shapes["w1w3"][0]), dtype=torch.float16, device=self.device),
            "h2": torch.zeros((1, 1, d), dtype=torch.float16, device=self.device),
        }
        self.part = torch.zeros((1, d), dtype=torch.float32, device=self.device) if W > 1 else None
        if self.collective == "push":
            from .tp import PushAllReduce

            # every K-sharded Linear of the chain outputs `dim` rows: 2 sites per block
            self.push = PushAllReduce(2 * self.cfg["n_layer"], d, group=process_group, device=self.device)
        self._site = 0
        self.hd = hd
        self.graph = None
        self.stream = torch.cuda.Stream(device=self.device)
        self.y_host = torch.zeros((1, 1, d), dtype=torch.float16).pin_memory()
        self.launches_per_step = 0

    # ------------------------------------------------------------------ accounting
    def algo_bytes_per_step(self) -> int:
        """algorithmic bytes this RANK moves per token (planes + LUT + x + y of every local GEMV)."""
        return sum(gemv_algo_bytes(l.N, l.K, self.bits) for lins in self.layers for l in lins)

    def weight_bytes(self) -> int:
        return sum(l.qweight.numel() * 4 + l.lut.numel() * 2 for lins in self.layers for l in lins)

    # ------------------------------------------------------------------ one token
    def _gemv(self, lin: _Lin, x: torch.Tensor, out: torch.Tensor, nxt: _Lin | None = None):
        flags = _lib.APG_FLAG_PDL if self.pdl else 0
        self.launches_per_step += 1
        pf = nxt.qweight if (nxt is not None and self.l2_prefetch) else None
        if lin.k_shard and self.push is not None:
            site = self._site
            self._site += 1
            self.push.gemv_push(site, x, lin.qweight, lin.lut, lin.N, lin.K, self.bits, flags=flags)
            self.push.finish(site, out, lin.N, flags=flags)
            self.launches_per_step += 1
        elif lin.k_shard:
            ap_gemv.anyprec_gemv_ex(x, out, lin.qweight, lin.lut, self.bits, flags=flags, partial=self.part,
                                    ctas_per_sm=self.ctas, prefetch_next=pf)
            torch.distributed.all_reduce(self.part, group=self.pg)
            st = _lib.lib().apg_round_f32_to_f16(self.part.data_ptr(), out.data_ptr(), out.numel(),
                                                 torch.cuda.current_stream().cuda_stream)
            _lib.check(st, "apg_round_f32_to_f16")
            self.launches_per_step += 1
        else:
            ap_gemv.anyprec_gemv_ex(x, out, lin.qweight, lin.lut, self.bits, flags=flags, ctas_per_sm=self.ctas,
                                    prefetch_next=pf)

    def _token(self):
        """the 4*L dependent GEMVs of one token.  Between Linears the (out-of-scope) attention / SwiGLU are
        replaced by slicing, in the Megatron data flow: every rank feeds its K-sharded wo (w2) with the first
        wo.K (w2.K) entries of ITS OWN wqkv (w1w3) output shard — what a head-split attention / a split SwiGLU
        would hand it.  With world_size 1 that is: attention-out := qkv[:dim], mlp-act := gu[:inter]."""
        b = self.buf
        x = self.x_in
        self.launches_per_step = 0
        self._site = 0
        for li, lins in enumerate(self.layers):
            wqkv, wo, w1w3, w2 = lins
            h_out = b["h"] if li % 2 == 0 else b["h2"]
            nxt_wqkv = self.layers[li + 1][0] if li + 1 < len(self.layers) else self.layers[0][0]
            self._gemv(wqkv, x, self._view(b["qkv"], 0, wqkv.N), wo)
            self._gemv(wo, self._view(b["qkv"], 0, wo.K), b["o"], w1w3)
            self._gemv(w1w3, b["o"], self._view(b["gu"], 0, w1w3.N), w2)
            self._gemv(w2, self._view(b["gu"], 0, w2.K), h_out, nxt_wqkv)
            x = h_out
        self.y_dev = x
        return x

    @staticmethod
    def _view(t: torch.Tensor, a: int, n: int) -> torch.Tensor:
        return t.reshape(-1)[a:a + n].reshape(1, 1, n)

    # ------------------------------------------------------------------ persistent engine
    def _build_program(self):
        """the same chain as _token() as one job list: x_in -> packets, 4*L GEMV jobs (K-sharded ones push their fp32
        partial sums to every rank and are followed by a reduce job), last output also written as plain fp16"""
        from .persist import PersistentProgram

        prog = PersistentProgram(self.bits, self.device)
        d, W = self.cfg["dim"], self.world
        bx, qkv, o = prog.buffer(d), prog.buffer(self.shapes["wqkv"][0]), prog.buffer(d)
        gu, hh = prog.buffer(self.shapes["w1w3"][0]), [prog.buffer(d), prog.buffer(d)]
        self.y_dev = torch.zeros((1, 1, d), dtype=torch.float16, device=self.device)
        prog.pack(self.x_in.reshape(1, d), bx)
        x = bx
        site = 0
        nl = len(self.layers)

        def lin(l, xin, out, plain=None):
            nonlocal site
            if not l.k_shard:
                prog.gemv(xin, l.qweight, l.lut, out, out_plain=plain)
                return
            ptrs = [b + site * self.push.site_bytes for b in self.push.peer_base]
            j = prog.gemv(xin, l.qweight, l.lut, None, push=(W, self.rank, ptrs))
            prog.reduce(ptrs[self.rank], j, l.N, W, out, out_plain=plain)
            site += 1

        for li, (wqkv, wo, w1w3, w2) in enumerate(self.layers):
            lin(wqkv, x, qkv)
            lin(wo, qkv, o)
            lin(w1w3, o, gu)
            lin(w2, gu, hh[li % 2], plain=self.y_dev if li == nl - 1 else None)
            x = hh[li % 2]
        self.prog = prog.finalize()
        self.launches_per_step = 1

    # ------------------------------------------------------------------ graph
    def capture(self):
        if self.engine == "persistent":
            with torch.cuda.device(self.device):
                if self.prog is None:
                    self._build_program()
                if self.world > 1:  # ranks finish building at different times: line them up before the first exchange
                    torch.cuda.synchronize()
                    torch.distributed.barrier(self.pg)
                with torch.cuda.stream(self.stream):
                    self.prog.launch()  # warm-up (function attributes)
                self.stream.synchronize()
                self.prog.check()
            self.graph = False  # a single launch per step: no graph needed
            return self
        with torch.cuda.device(self.device):
            s = self.stream
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._token()  # warm-up (lazy attribute setup, NCCL communicator)
                s.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=s):
                    self._token()
            torch.cuda.current_stream().wait_stream(s)
        return self

    def step(self):
        """one token, inputs resident in HBM; asynchronous on self.stream."""
        if self.graph is None:
            self.capture()
        with torch.cuda.stream(self.stream):
            self._replay()

    def _replay(self):
        if self.engine == "persistent":
            self.prog.launch()
        else:
            self.graph.replay()

    def step_host(self, x_host: torch.Tensor) -> torch.Tensor:
        """end-to-end call: pinned-host activations in, pinned-host result out (synchronises)."""
        if self.graph is None:
            self.capture()
        with torch.cuda.stream(self.stream):
            self.x_in.copy_(x_host, non_blocking=True)
            self._replay()
            self.y_host.copy_(self.y_dev, non_blocking=True)
        self.stream.synchronize()
        return self.y_host

    def eager_token(self, x: torch.Tensor) -> torch.Tensor:
        """un-graphed reference execution of the same chain (for tests)."""
        self.x_in.copy_(x)
        if self.engine == "persistent":
            if self.prog is None:
                self._build_program()
            self.prog.launch()
            return self.y_dev.clone()
        return self._token().clone()
