"""Decode runtime: the reference's gpt-fast Transformer (inference/model.py) at batch 1 / sequence 1 on top of the
native kernels, one CUDA graph per token.

Per block 5 launches (the reference's eager graph has 11 ops + what Inductor fuses):
    wqkv  = apg_gemv_fused(x,  norm=input_layernorm)                       model.py:152-153, 211
    att   = apd_attn_decode(wqkv)          RoPE + KV append + attention     model.py:206-236
    h     = apg_gemv_fused(att, residual=x)                       wo        model.py:152, 236
    gu    = apg_gemv_fused(h,  norm=post_attention_layernorm)     w1w3      model.py:159, 261
    x'    = apg_gemv_fused(gu, silu_mul, residual=h)              w2        model.py:165, 266
plus embedding, lm_head (final RMSNorm fused) and greedy sampling (generate.py:55-73 at temperature 0) per token.
The token id and the position live in device memory and are advanced by the sampling kernel, so `generate()`
replays the same graph without any host round trip.

State-dict names follow the reference's converted checkpoints (inference/sqllm_llama_convert_fuse.py:71-116):
    tok_embeddings.weight, layers.{i}.attention.{wqkv,wo}.{qweight,lut}, layers.{i}.feed_forward.{w1w3,w2}.{qweight,lut},
    layers.{i}.{input_layernorm,post_attention_layernorm}.weight, norm.weight, output.weight
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .runtime import MODEL_CONFIGS, PERSISTENT_MIN_WORLD, gemv_algo_bytes, linear_shapes, persistent_supported

ROPE_BASE = {"llama3-8b": 500000.0, "llama3-70b": 500000.0, "llama2-7b": 10000.0, "llama2-70b": 10000.0, "tiny": 500000.0,
             "golden-tiny": 500000.0, "tiny128": 500000.0, "tiny128kv4": 500000.0}
MODEL_CONFIGS.setdefault("golden-tiny", dict(dim=256, n_layer=2, n_head=2, n_kv=1, inter=512, vocab=256))
MODEL_CONFIGS.setdefault("tiny128", dict(dim=1024, n_layer=2, n_head=8, n_kv=2, inter=2048, vocab=1024))
MODEL_CONFIGS.setdefault("tiny128kv4", dict(dim=1024, n_layer=2, n_head=8, n_kv=4, inter=2048, vocab=1024))


class APTransformer:
    def __init__(self, model: str = "llama3-8b", bits: int = 2, max_seq_len: int = 256, device=None, pdl: bool = True,
                 norm_eps: float = 1e-5, n_layer: int | None = None, attn_splits: int | None = None,
                 world_size: int = 1, rank: int = 0, process_group=None, glu_epilogue: bool | None = None,
                 engine: str | None = None):
        self.cfg = dict(MODEL_CONFIGS[model])
        # "launches" (default, measured fastest at every world size): one PDL launch per op;
        # "persistent" (opt-in): embedding + all blocks of a token are ONE cooperative launch of the persistent token kernel
        # (persist.py / csrc/apgemv_persist.cuh), then lm_head + sampling
        self.engine = engine or ("persistent" if world_size >= PERSISTENT_MIN_WORLD and persistent_supported(self.cfg, bits, world_size)
                                 else "launches")
        assert self.engine in ("launches", "persistent"), f"unknown decode engine {engine!r}"
        self.prog = None
        if n_layer is not None:
            self.cfg["n_layer"] = n_layer
        c = self.cfg
        assert c["dim"] // c["n_head"] == 128, "the attention kernel is specialised for head_dim 128"
        self.model, self.bits, self.S, self.eps = model, bits, max_seq_len, norm_eps
        # w1w3's rows are kept interleaved (gate_i, up_i) so that its epilogue writes silu(gate)*up once per element instead
        # of every CTA of the w2 launch recomputing the activation in its prologue: same roundings -> identical logits,
        # measured 5 % faster per token on B200 (profiles/r2_ab_glu.log).  Always on in the persistent engine, the default
        # of the launches engine (under tensor parallelism every rank interleaves its own gate / up columns).
        if self.engine == "persistent":
            assert glu_epilogue in (None, True), "the persistent engine always uses the SwiGLU epilogue"
            self.glu_epilogue = True
        else:
            if glu_epilogue is None:   # the epilogue variant exists for one chunk per warp and 8-row stages (launch_fast)
                import ctypes

                plan = (ctypes.c_uint32 * 16)()
                W_ = max(1, world_size)
                ok = _lib.lib().apg_plan_fast(2 * (c["inter"] // W_), c["dim"], bits, 0, 148, plan) == 0
                glu_epilogue = ok and plan[0] == 1 and plan[3] == 8 and (2 * (c["inter"] // W_)) % 4 == 0
            self.glu_epilogue = bool(glu_epilogue)
        self.flags = _lib.APG_FLAG_PDL if pdl else 0
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.rope_base = ROPE_BASE.get(model, 10000.0)
        self.shapes = linear_shapes(c)
        # tensor parallelism (Megatron pairing): heads / MLP columns split over `world` ranks; wo and w2
        # are K-sharded and their all-reduce is fused into the GEMV epilogue (PushAllReduce); lm_head is replicated
        self.world, self.rank, self.pg = world_size, rank, process_group
        W = world_size
        assert c["n_head"] % W == 0 and c["n_kv"] % W == 0 and c["inter"] % (128 * W) == 0 and c["dim"] % (128 * W) == 0
        self.H_l, self.Hkv_l, self.inter_l, self.dk_l = c["n_head"] // W, c["n_kv"] // W, c["inter"] // W, c["dim"] // W
        self.lshapes = {"wqkv": ((self.H_l + 2 * self.Hkv_l) * 128, c["dim"]), "wo": (c["dim"], self.dk_l),
                        "w1w3": (2 * self.inter_l, c["dim"]), "w2": (c["dim"], self.inter_l)}
        self.push = None
        if W > 1:
            from .tp import PushAllReduce

            # 2 all-reduce sites per block + 1 site for the arg-max exchange of the vocab-sharded lm_head
            self.push = PushAllReduce(2 * c["n_layer"] + 1, c["dim"], group=process_group, device=self.device)
            assert c["vocab"] % W == 0
        d, dev, f16 = c["dim"], self.device, torch.float16
        self.nsplit = attn_splits if attn_splits is not None else max(1, min(32, max_seq_len // 512))
        self.sd: dict[str, torch.Tensor] = {}
        # activations / state
        self.x = torch.zeros(d, dtype=f16, device=dev)
        self.h = torch.zeros(d, dtype=f16, device=dev)
        self.qkv = torch.zeros(self.lshapes["wqkv"][0], dtype=f16, device=dev)
        self.att = torch.zeros(self.dk_l, dtype=f16, device=dev)
        self.gu = torch.zeros(self.lshapes["w1w3"][0], dtype=f16, device=dev)
        self.V_l = c["vocab"] // W                      # lm_head rows of this rank (vocab-sharded under TP)
        self.logits = torch.zeros(self.V_l, dtype=f16, device=dev)
        self.best_val = torch.zeros(8192, dtype=torch.float32, device=dev)   # per-CTA arg-max partials of lm_head
        self.best_idx = torch.zeros(8192, dtype=torch.int32, device=dev)
        self.token = torch.zeros(1, dtype=torch.int32, device=dev)
        self.pos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.history = torch.zeros(max_seq_len + 1, dtype=torch.int32, device=dev)
        hd = 128
        from .convert import rope_inv_freq

        self.inv_freq = rope_inv_freq(self.rope_base, c.get("rope_scaling"), hd).to(dev)   # default or llama3-scaled table
        self.k_cache = [torch.zeros((self.Hkv_l, max_seq_len, hd), dtype=f16, device=dev) for _ in range(c["n_layer"])]
        self.v_cache = [torch.zeros((self.Hkv_l, max_seq_len, hd), dtype=f16, device=dev) for _ in range(c["n_layer"])]
        self.part = torch.zeros(self.H_l * self.nsplit * 132, dtype=torch.float32, device=dev) if self.nsplit > 1 else None
        self.graph = None
        self.use_graph = True   # False: launch the token's kernels eagerly instead of replaying a captured CUDA graph
        self.stream = torch.cuda.Stream(device=dev)
        self.tok_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.launches_per_token = 0
        self._pos_host = 0   # host mirror of the device-side position (bound check in step / step_host)
        # measured on B200: launching the 1 GB lm_head stream programmatically (early, beside the last w2) costs
        # ~250 us/token; it is launched as a plain stream-ordered kernel instead
        self.lm_head_no_pdl = _lib.APG_FLAG_PDL
        self.debug_skip: set[str] = set()  # profiling aid only: {"attn", "lm_head", "sample", "embed", "fusion"}
        # sampling (generate.py:55-73): temperature 0 -> greedy arg-max kernel; > 0 -> apd_sample_topk_advance.  The seed
        # lives in device memory so that set_sampling(seed=...) does not invalidate a captured graph.
        self.temperature, self.top_k = 0.0, None
        self.seed = torch.full((1,), 1234, dtype=torch.int64, device=dev)

    # ------------------------------------------------------------------ weights
    def random_init(self, seed: int = 0):
        """synthetic packed weights of the right shapes (the reference's --random_init, generate.py:235,412)."""
        c, dev, bits = self.cfg, self.device, self.bits
        g = torch.Generator(device=dev).manual_seed(4321 + seed)
        sd = self.sd
        sd["tok_embeddings.weight"] = torch.randn((c["vocab"], c["dim"]), device=dev, generator=g).half()
        for i in range(c["n_layer"]):
            for mod, names in (("attention", ("wqkv", "wo")), ("feed_forward", ("w1w3", "w2"))):
                for nm in names:
                    N, K = self.shapes[nm]
                    sd[f"layers.{i}.{mod}.{nm}.qweight"] = torch.randint(-2**31, 2**31 - 1, (bits, N, K // 32), dtype=torch.int32, device=dev, generator=g)
                    sd[f"layers.{i}.{mod}.{nm}.lut"] = (torch.randn((N, 1 << bits), device=dev, generator=g) * (1.6 / math.sqrt(K))).half()
            sd[f"layers.{i}.input_layernorm.weight"] = (1 + 0.1 * torch.randn(c["dim"], device=dev, generator=g)).half()
            sd[f"layers.{i}.post_attention_layernorm.weight"] = (1 + 0.1 * torch.randn(c["dim"], device=dev, generator=g)).half()
        sd["norm.weight"] = (1 + 0.1 * torch.randn(c["dim"], device=dev, generator=g)).half()
        sd["output.weight"] = (torch.randn((c["vocab"], c["dim"]), device=dev, generator=g) / math.sqrt(c["dim"])).half()
        if self.world > 1 or self.glu_epilogue:  # shard / re-order through the common path
            self.sd = {}
            self.load_state_dict(sd)
        self.prog = self.graph = None
        return self

    def expected_keys(self) -> list[str]:
        c = self.cfg
        keys = ["tok_embeddings.weight", "norm.weight", "output.weight"]
        for i in range(c["n_layer"]):
            keys += [f"layers.{i}.input_layernorm.weight", f"layers.{i}.post_attention_layernorm.weight"]
            for mod, names in (("attention", ("wqkv", "wo")), ("feed_forward", ("w1w3", "w2"))):
                for nm in names:
                    keys += [f"layers.{i}.{mod}.{nm}.qweight", f"layers.{i}.{mod}.{nm}.lut"]
        return keys

    def _validate(self, name: str, t: torch.Tensor) -> torch.Tensor:
        """shape / dtype check of one FULL (unsharded) tensor against (bits, N, K) — what load_state_dict(strict=True) of the
        reference's nn.Module does (generate.py:239); a qweight with more planes than `bits` (any-precision parent) is
        sliced to its first `bits` planes like sqllm_llama_convert_fuse.py:59-60."""
        c = self.cfg
        parts = name.split(".")
        if name.endswith(".qweight") or name.endswith(".lut"):
            N, K = self.shapes[parts[-2]]
            if name.endswith(".qweight"):
                if t.dtype != torch.int32 or t.dim() != 3 or t.shape[0] < self.bits or tuple(t.shape[1:]) != (N, K // 32):
                    raise ValueError(f"{name}: expected int32 [>={self.bits}, {N}, {K // 32}], got {t.dtype} {tuple(t.shape)}")
                return t[:self.bits]
            if t.dtype != torch.float16 or tuple(t.shape) != (N, 1 << self.bits):
                raise ValueError(f"{name}: expected float16 [{N}, {1 << self.bits}] (a {self.bits}-bit codebook), got {t.dtype} "
                                 f"{tuple(t.shape)} — was the checkpoint converted for another bitwidth?")
            return t
        want = {"tok_embeddings.weight": (c["vocab"], c["dim"]), "output.weight": (c["vocab"], c["dim"])}.get(name, (c["dim"],))
        if t.dtype != torch.float16 or tuple(t.shape) != want:
            raise ValueError(f"{name}: expected float16 {want}, got {t.dtype} {tuple(t.shape)}")
        return t

    def load_state_dict(self, sd: dict):
        """full (unsharded) state dict in; with world_size > 1 the Linears are sharded for this rank on the way.  Strict:
        every expected key must be present with the shape / dtype (bits, N, K) imply, unexpected keys are refused."""
        exp = self.expected_keys()
        missing = [k for k in exp if k not in sd]
        unexpected = [k for k in sd if k not in set(exp)]
        if missing or unexpected:
            raise KeyError(f"state dict mismatch: missing {missing[:6]}{'...' if len(missing) > 6 else ''}, "
                           f"unexpected {unexpected[:6]}{'...' if len(unexpected) > 6 else ''}")
        for k in exp:
            t = self._shard(k, self._validate(k, sd[k]).to(self.device))
            if self.glu_epilogue and ".w1w3." in k:  # rows (gate | up) -> (gate_0, up_0, gate_1, up_1, ...)
                inter = self.inter_l  # after sharding the rows are this rank's gate columns then its up columns
                idx = torch.stack([torch.arange(inter), inter + torch.arange(inter)], dim=1).reshape(-1).to(t.device)
                t = t.index_select(1 if k.endswith(".qweight") else 0, idx)
            self.sd[k] = t.contiguous()
        self.prog = self.graph = None  # the job table / graph point into the old tensors
        return self

    def _shard(self, name: str, t: torch.Tensor) -> torch.Tensor:
        W, r, c = self.world, self.rank, self.cfg
        if W > 1 and name == "output.weight":   # lm_head: vocab-sharded rows
            return t[r * (c["vocab"] // W):(r + 1) * (c["vocab"] // W)]
        if W == 1 or ".attention." not in name and ".feed_forward." not in name:
            return t
        from . import pack as packmod

        H, Hkv, inter = c["n_head"], c["n_kv"], c["inter"]
        rows_dim = 1 if name.endswith(".qweight") else 0
        if ".wqkv." in name:   # rows: this rank's q heads, k heads, v heads (model.py:211 split order q | k | v)
            q0, k0, v0 = 0, H * 128, (H + Hkv) * 128
            idx = torch.cat([torch.arange(q0 + r * self.H_l * 128, q0 + (r + 1) * self.H_l * 128),
                             torch.arange(k0 + r * self.Hkv_l * 128, k0 + (r + 1) * self.Hkv_l * 128),
                             torch.arange(v0 + r * self.Hkv_l * 128, v0 + (r + 1) * self.Hkv_l * 128)]).to(t.device)
            return t.index_select(rows_dim, idx)
        if ".w1w3." in name:   # rows: this rank's gate columns then its up columns (model.py:261 split order w1 | w3)
            idx = torch.cat([torch.arange(r * self.inter_l, (r + 1) * self.inter_l),
                             torch.arange(inter + r * self.inter_l, inter + (r + 1) * self.inter_l)]).to(t.device)
            return t.index_select(rows_dim, idx)
        # wo / w2: K-sharded.  lut is per output row -> replicated; qweight is cut along K (re-packed if the cut falls
        # inside a 1024-weight chunk)
        if name.endswith(".lut"):
            return t
        K = t.shape[2] * 32
        k0, k1 = r * (K // W), (r + 1) * (K // W)
        if k0 % 1024 == 0 and (k1 % 1024 == 0 or k1 == K):
            return t[:, :, k0 // 32:k1 // 32]
        return packmod.shard_k_torch(t, k0, k1)

    @classmethod
    def from_checkpoint(cls, ckpt_dir: str, bitwidth: int, max_seq_len: int = 2048, **kw) -> "APTransformer":
        """Load a packed Any-Precision checkpoint directory as written by the reference's packer
        (any_precision/quantization/pack.py:133-203: `pytorch_model.bin` with HF names + `config.json`), converting it
        on the fly like inference/sqllm_llama_convert_fuse.py, or its already converted `converted_pytorch_model.bin`
        (loaded mmap'd like generate.py:237-238).  The architecture comes from config.json, not from a name table."""
        import json
        import os

        from .convert import arch_from_hf_config, convert_state_dict, read_checkpoint

        hf = json.load(open(os.path.join(ckpt_dir, "config.json")))
        name = "ckpt:" + os.path.basename(os.path.normpath(ckpt_dir))
        MODEL_CONFIGS[name], ROPE_BASE[name], eps = arch_from_hf_config(hf)
        m = cls(name, bits=bitwidth, max_seq_len=max_seq_len, norm_eps=eps, **kw)
        conv = os.path.join(ckpt_dir, "converted_pytorch_model.bin")
        if os.path.exists(conv):
            sd = torch.load(conv, map_location="cpu", mmap=True, weights_only=True)
        else:
            sd = convert_state_dict(read_checkpoint(ckpt_dir), bitwidth)
        sd = {k: (v.half() if v.is_floating_point() else v) for k, v in sd.items()}
        if "output.weight" not in sd and hf.get("tie_word_embeddings"):
            sd["output.weight"] = sd["tok_embeddings.weight"]
        sd = {k: v for k, v in sd.items() if not k.endswith("rotary_emb.inv_freq")}
        return m.load_state_dict(sd)  # validates every tensor against (bitwidth, N, K): a file converted for another bitwidth is refused

    # ------------------------------------------------------------------ accounting
    def algo_bytes_per_token(self, pos: int = 0) -> dict:
        c = self.cfg
        gemv = c["n_layer"] * sum(gemv_algo_bytes(N, K, self.bits) for (N, K) in self.shapes.values())
        lm_head = 2 * c["vocab"] * c["dim"]
        kv = c["n_layer"] * 2 * c["n_kv"] * 128 * 2 * (pos + 1)
        misc = c["n_layer"] * 2 * c["dim"] * 2 + 2 * c["dim"] + 2 * c["vocab"] * 2
        return {"gemv": gemv, "lm_head": lm_head, "kv": kv, "misc": misc, "total": gemv + lm_head + kv + misc}

    # ------------------------------------------------------------------ one token
    def _fused(self, x, out, name, N, K, norm=None, silu_mul=0, residual=None):
        L = _lib.lib()
        q, lut = self.sd[name + ".qweight"], self.sd[name + ".lut"]
        if "fusion" in self.debug_skip:
            norm, silu_mul, residual = None, 0, None
        if "norm" in self.debug_skip:
            norm = None
        if "silu" in self.debug_skip:
            silu_mul = 0
        if "residual" in self.debug_skip:
            residual = None
        st = L.apg_gemv_fused(x.data_ptr(), out.data_ptr(), None, q.data_ptr(), lut.data_ptr(), N, K, self.bits,
                              norm.data_ptr() if norm is not None else None, self.eps, silu_mul,
                              residual.data_ptr() if residual is not None else None, self.flags,
                              torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "apg_gemv_fused " + name)
        self.launches_per_token += 1

    def _build_program(self):
        """the embedding row + every block of a token as ONE job list for the persistent token kernel (module docstring's
        5 ops per block; K-sharded wo / w2 push their fp32 partial sums to every rank and are finished by a reduce job)"""
        from .persist import PersistentProgram

        c, sd, W = self.cfg, self.sd, self.world
        prog = PersistentProgram(self.bits, self.device)
        d = c["dim"]
        X = [prog.buffer(d), prog.buffer(d)]
        QKV, ATT = prog.buffer(self.lshapes["wqkv"][0]), prog.buffer(self.dk_l)
        Hb, GU = prog.buffer(d), prog.buffer(self.inter_l)
        scale = 1.0 / math.sqrt(128.0)
        # cos / sin of every position in fp32, rounded to fp16 like the reference's rotary embedding (model.py:396-405)
        ang = torch.arange(self.S, dtype=torch.float32, device=self.device)[:, None] * self.inv_freq[None, :].float()
        rope_cs = torch.cat([torch.cos(ang), torch.sin(ang)], dim=1).half().contiguous()
        prog.pack(sd["tok_embeddings.weight"], X[0], row_index=self.token)
        x = X[0]
        nl = c["n_layer"]

        def k_sharded(site, xin, name, out, residual, plain):
            ptrs = [b + site * self.push.site_bytes for b in self.push.peer_base]
            j = prog.gemv(xin, sd[name + ".qweight"], sd[name + ".lut"], None, push=(W, self.rank, ptrs))
            prog.reduce(ptrs[self.rank], j, d, W, out, residual=residual, out_plain=plain)

        for i in range(nl):
            p = f"layers.{i}."
            prog.gemv(x, sd[p + "attention.wqkv.qweight"], sd[p + "attention.wqkv.lut"], QKV, norm_w=sd[p + "input_layernorm.weight"],
                      eps=self.eps)
            prog.attn(QKV, rope_cs, self.k_cache[i], self.v_cache[i], ATT, self.H_l, self.Hkv_l, self.S, scale)
            if W == 1:
                prog.gemv(ATT, sd[p + "attention.wo.qweight"], sd[p + "attention.wo.lut"], Hb, residual=x)
            else:
                k_sharded(2 * i, ATT, p + "attention.wo", Hb, x, None)
            prog.gemv(Hb, sd[p + "feed_forward.w1w3.qweight"], sd[p + "feed_forward.w1w3.lut"], GU,
                      norm_w=sd[p + "post_attention_layernorm.weight"], eps=self.eps, glu=True)
            xn = X[(i + 1) % 2]
            plain = self.x if i == nl - 1 else None  # the fp16 hidden state the lm_head kernel reads
            if W == 1:
                prog.gemv(GU, sd[p + "feed_forward.w2.qweight"], sd[p + "feed_forward.w2.lut"], xn, residual=Hb, out_plain=plain)
            else:
                k_sharded(2 * i + 1, GU, p + "feed_forward.w2", xn, Hb, plain)
            x = xn
        self.prog = prog.finalize()

    def _blocks_launches(self):
        """launches engine: embedding + L blocks, 5 PDL launches per block"""
        L, c, sd, fl = _lib.lib(), self.cfg, self.sd, self.flags
        st = torch.cuda.current_stream().cuda_stream
        if "embed" not in self.debug_skip:
            _lib.check(L.apd_embed(sd["tok_embeddings.weight"].data_ptr(), self.token.data_ptr(), self.x.data_ptr(), c["dim"], c["vocab"], fl, st), "apd_embed")
            self.launches_per_token += 1
        scale = 1.0 / math.sqrt(128.0)
        for i in range(c["n_layer"]):
            p = f"layers.{i}."
            (nq, kq), (no, ko), (ng, kg), (n2, k2) = (self.lshapes[n] for n in ("wqkv", "wo", "w1w3", "w2"))
            self._fused(self.x, self.qkv, p + "attention.wqkv", nq, kq, norm=sd[p + "input_layernorm.weight"])
            if "attn" not in self.debug_skip:
              _lib.check(L.apd_attn_decode(self.qkv.data_ptr(), self.inv_freq.data_ptr(), self.k_cache[i].data_ptr(),
                                         self.v_cache[i].data_ptr(), self.pos.data_ptr(), self.att.data_ptr(),
                                         self.part.data_ptr() if self.part is not None else None, self.H_l, self.Hkv_l,
                                         self.S, self.nsplit, scale, fl, st), "apd_attn_decode")
              self.launches_per_token += 1 + (1 if self.nsplit > 1 else 0)
            if self.push is None:
                self._fused(self.att, self.h, p + "attention.wo", no, ko, residual=self.x)
                glu = self.glu_epilogue
                self._fused(self.h, self.gu, p + "feed_forward.w1w3", ng, kg, norm=sd[p + "post_attention_layernorm.weight"],
                            silu_mul=2 if glu else 0)
                self._fused(self.gu, self.x, p + "feed_forward.w2", n2, k2, silu_mul=0 if glu else 1, residual=self.h)
            else:  # K-sharded wo / w2: partial sums pushed to every peer from the GEMV epilogue, residual added by the finisher
                self.push.gemv_push(2 * i, self.att, sd[p + "attention.wo.qweight"], sd[p + "attention.wo.lut"], no, ko,
                                    self.bits, flags=fl)
                self.push.finish(2 * i, self.h, no, residual=self.x, flags=fl)
                glu = self.glu_epilogue
                self._fused(self.h, self.gu, p + "feed_forward.w1w3", ng, kg, norm=sd[p + "post_attention_layernorm.weight"],
                            silu_mul=2 if glu else 0)
                self.push.gemv_push(2 * i + 1, self.gu, sd[p + "feed_forward.w2.qweight"], sd[p + "feed_forward.w2.lut"],
                                    n2, k2, self.bits, silu_mul=0 if glu else 1, eps=self.eps, flags=fl)
                self.push.finish(2 * i + 1, self.x, n2, residual=self.h, flags=fl)
                self.launches_per_token += 4

    def _head_and_sample(self):
        """final RMSNorm + lm_head on self.x, then the sampler: writes self.token, history[pos + 1], pos += 1 on the device"""
        L, c, sd, fl = _lib.lib(), self.cfg, self.sd, self.flags
        st = torch.cuda.current_stream().cuda_stream
        if "lm_head" not in self.debug_skip:
          import ctypes
          npart = ctypes.c_uint32(0)
          _lib.check(L.apd_lm_head(self.x.data_ptr(), sd["norm.weight"].data_ptr(), self.eps, sd["output.weight"].data_ptr(),
                                 self.logits.data_ptr(), self.V_l, c["dim"], self.best_val.data_ptr(),
                                 self.best_idx.data_ptr(), ctypes.byref(npart), self.rank * self.V_l,
                                 fl & ~self.lm_head_no_pdl, st), "apd_lm_head")
          self._npart = npart.value
        if "sample" not in self.debug_skip and self.push is not None and self.temperature > 0.0:
          site = 2 * c["n_layer"]   # vocab-sharded sampling: top-k pivot + winner exchanged through the peers' buffers
          _lib.check(L.apd_sample_topk_advance_tp(self.logits.data_ptr(), self.V_l, float(self.temperature), int(self.top_k or 0),
                                                self.seed.data_ptr(), self.world, self.rank, self.push.site_ptrs(site),
                                                self.push.n_max, self.push.epoch_ptr(site), self.token.data_ptr(),
                                                self.pos.data_ptr(), self.history.data_ptr(), self.history.numel(), fl, st),
                     "apd_sample_topk_advance_tp")
        elif "sample" not in self.debug_skip and self.push is not None:
          site = 2 * c["n_layer"]
          _lib.check(L.apd_argmax_advance_tp(self.best_val.data_ptr(), self.best_idx.data_ptr(), self._npart, self.world,
                                           self.rank, self.push.site_ptrs(site), self.push.epoch_ptr(site),
                                           self.token.data_ptr(), self.pos.data_ptr(), self.history.data_ptr(),
                                           self.history.numel(), fl, st), "apd_argmax_advance_tp")
        elif "sample" not in self.debug_skip and self.temperature > 0.0:
          _lib.check(L.apd_sample_topk_advance(self.logits.data_ptr(), self.V_l, float(self.temperature), int(self.top_k or 0),
                                             self.seed.data_ptr(), self.token.data_ptr(), self.pos.data_ptr(),
                                             self.history.data_ptr(), self.history.numel(), fl, st), "apd_sample_topk_advance")
        elif "sample" not in self.debug_skip:
          _lib.check(L.apd_argmax_advance(self.best_val.data_ptr(), self.best_idx.data_ptr(), self._npart,
                                        self.token.data_ptr(), self.pos.data_ptr(),
                                        self.history.data_ptr(), self.history.numel(), fl, st), "apd_argmax_advance")

    def decode_step(self):
        """embedding -> L blocks -> lm_head -> greedy sample; reads self.token / self.pos on the device and advances them."""
        L, c, sd, fl = _lib.lib(), self.cfg, self.sd, self.flags
        st = torch.cuda.current_stream().cuda_stream
        self.launches_per_token = 0
        if self.engine == "persistent":
            if self.prog is None:
                self._build_program()
            self.prog.launch(self.pos)
            self.launches_per_token = 1
        else:
            self._blocks_launches()
        self._head_and_sample()
        self.launches_per_token += 2

    # ------------------------------------------------------------------ prompt prefill
    def _prefill_linear(self, x: torch.Tensor, name: str) -> torch.Tensor:
        """[T, K] -> [T, N] through the same routes as APLinear.gemm: batched LUT GEMV up to 8 rows, the fused dequant + tcgen05
        GEMM kernel up to the measured cross-over, dequant -> library matmul beyond"""
        from . import ap_gemv

        q, lut = self.sd[name + ".qweight"], self.sd[name + ".lut"]
        T = x.shape[0]
        if T <= 8:
            out = torch.empty((T, 1, q.shape[1]), dtype=torch.float16, device=x.device)
            ap_gemv.anyprec_gemv(x.reshape(T, 1, -1).contiguous(), out, q, lut, self.bits)
            return out.reshape(T, -1)
        if ap_gemv.prefill_prefers_fused(q, self.bits, T):
            return ap_gemv.anyprec_prefill_gemm(x.contiguous(), q, lut, self.bits)
        return torch.matmul(x, ap_gemv.anyprec_dequant(q, lut, self.bits).T)

    def _rmsnorm(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:  # model.py:280-285: fp32 normalise, cast, * weight
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)).to(x.dtype) * w

    @torch.no_grad()
    def prefill(self, tokens: list[int]) -> int:
        """Process a whole prompt at once — the reference's prefill (generate.py:146-166: one model forward over the prompt, SDPA
        with a causal mask, APLinear.gemm for every Linear) — instead of one decode step per prompt token.  The Linears run on
        the fused tensor-core kernel; the glue (RMSNorm, RoPE, attention, SwiGLU, residuals) is torch with the reference's fp16
        roundings.  Fills the KV caches, samples the first new token with the decode sampler and leaves the model at position
        len(tokens); returns that token.  Tensor parallel: K-sharded Linears are summed with one NCCL all-reduce each."""
        T, c, sd = len(tokens), self.cfg, self.sd
        assert 1 <= T < self.S, (T, self.S)
        if self.graph is None:
            self.capture()   # also lines the ranks up and builds the persistent program
        import torch.nn.functional as F

        H, Hkv, hd = self.H_l, self.Hkv_l, 128

        def allreduce(y):
            if self.world == 1:
                return y
            yf = y.float()
            torch.distributed.all_reduce(yf, group=self.pg)
            return yf.half()

        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            if self.world > 1:
                self.stream.synchronize()  # no NCCL collective while launches of the persistent engine are in flight
            tok = torch.tensor(tokens, dtype=torch.long, device=self.device)
            x = sd["tok_embeddings.weight"][tok]                                   # [T, dim] fp16
            ang = torch.arange(T, dtype=torch.float32, device=self.device)[:, None] * self.inv_freq[None, :].float()
            emb = torch.cat((ang, ang), dim=-1)
            cos, sin = emb.cos().half()[:, None, :], emb.sin().half()[:, None, :]  # fp32 -> fp16 (model.py:396-405)

            def rope(t):  # HF rotate-half in fp16 (model.py:268-272, 309-314)
                rot = torch.cat((-t[..., hd // 2:], t[..., :hd // 2]), dim=-1)
                return t * cos + rot * sin

            for i in range(c["n_layer"]):
                p = f"layers.{i}."
                qkv = self._prefill_linear(self._rmsnorm(x, sd[p + "input_layernorm.weight"]), p + "attention.wqkv")
                q = rope(qkv[:, : H * hd].reshape(T, H, hd))
                k = rope(qkv[:, H * hd: (H + Hkv) * hd].reshape(T, Hkv, hd))
                v = qkv[:, (H + Hkv) * hd:].reshape(T, Hkv, hd)
                self.k_cache[i][:, :T] = k.transpose(0, 1)
                self.v_cache[i][:, :T] = v.transpose(0, 1)
                kk = k.transpose(0, 1).repeat_interleave(H // Hkv, dim=0)          # GQA (model.py:229-230)
                vv = v.transpose(0, 1).repeat_interleave(H // Hkv, dim=0)
                y = F.scaled_dot_product_attention(q.transpose(0, 1)[None], kk[None], vv[None], is_causal=True)[0]
                y = y.transpose(0, 1).reshape(T, H * hd)
                h = x + allreduce(self._prefill_linear(y, p + "attention.wo"))
                gu = self._prefill_linear(self._rmsnorm(h, sd[p + "post_attention_layernorm.weight"]), p + "feed_forward.w1w3")
                if self.glu_epilogue:   # rows interleaved (gate_i, up_i)
                    gate, up = gu[:, 0::2], gu[:, 1::2]
                else:
                    gate, up = gu[:, : self.inter_l], gu[:, self.inter_l:]
                x = h + allreduce(self._prefill_linear(F.silu(gate) * up, p + "feed_forward.w2"))
            # hand over to the decode kernels: last token's hidden state -> lm_head -> sampler (advances pos to T)
            self.x.copy_(x[T - 1])
            self.history[:T] = tok.to(torch.int32)
            self.pos.fill_(T - 1)
            self._head_and_sample()
            self._pos_host = T
        self.stream.synchronize()
        return int(self.token.cpu()[0])

    # ------------------------------------------------------------------ graph + generation
    def set_sampling(self, temperature: float = 0.0, top_k: int | None = None, seed: int | None = None):
        """temperature / top_k of the reference's sample() (generate.py:55-73).  temperature 0 (the reference CLI default,
        generate.py:400) keeps the greedy kernel.  Changing temperature / top_k drops the captured graph; a new seed does not."""
        if temperature > 0.0 and self.world > 1 and top_k is not None and top_k < self.cfg["vocab"]:
            # vocab-sharded sampler (apd_sample_topk_advance_tp): every rank contributes its k largest logits to the pivot
            if not (0 < top_k <= 256 and top_k < self.V_l and top_k + 2 <= self.cfg["dim"]):
                raise ValueError(f"tensor-parallel top-k sampling needs top_k <= min(256, vocab/world - 1, dim - 2), got {top_k}")
        if (float(temperature), top_k) != (self.temperature, self.top_k):
            self.temperature, self.top_k, self.graph = float(temperature), top_k, None
        if seed is not None:
            with torch.cuda.stream(self.stream):
                self.seed.fill_(int(seed))
            self.stream.synchronize()
        return self

    def capture(self):
        with torch.cuda.device(self.device):
            s = self.stream
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                saved = (self.token.clone(), self.pos.clone())
                if self.world > 1:  # ranks finish loading at different times: line them up before the first exchange
                    s.synchronize()
                    torch.distributed.barrier(self.pg)
                self.decode_step()  # warm-up (function attributes, lazy init); its side effects are undone below
                s.synchronize()
                if self.prog is not None:
                    self.prog.check()  # a device-side watchdog would have fired on the very first token
                if self.use_graph:
                    self.graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph, stream=s):
                        self.decode_step()
                else:
                    self.graph = False  # eager: decode_step() per token (3 launches with the persistent engine)
                self.token.copy_(saved[0])
                self.pos.copy_(saved[1])
                s.synchronize()
            torch.cuda.current_stream().wait_stream(s)
        return self

    def reset(self, first_token: int = 1):
        self._pos_host = 0
        with torch.cuda.stream(self.stream):
            self.token.fill_(first_token)
            self.pos.zero_()
            self.history.zero_()
            self.history[0] = first_token
        self.stream.synchronize()

    def _advance_host_pos(self):
        # the attention kernel refuses positions >= max_seq_len (it leaves `out` untouched) while the sampler would still
        # advance: track the position on the host and raise instead of producing tokens from stale attention output
        if self._pos_host >= self.S:
            raise RuntimeError(f"KV cache full: position {self._pos_host} >= max_seq_len {self.S} (reset() or build with a larger max_seq_len)")
        self._pos_host += 1

    def step(self):
        if self.graph is None:
            self.capture()
        self._advance_host_pos()
        with torch.cuda.stream(self.stream):
            self._replay()

    def _replay(self):
        if self.graph is False:
            self.decode_step()
        else:
            self.graph.replay()

    def step_host(self, token_host: torch.Tensor) -> int:
        """end-to-end call for one token: pinned-host token id -> H2D -> graph -> D2H of the sampled token."""
        if self.graph is None:
            self.capture()
        self._advance_host_pos()
        with torch.cuda.stream(self.stream):
            self.token.copy_(token_host, non_blocking=True)
            self._replay()
            self.tok_host.copy_(self.token, non_blocking=True)
        self.stream.synchronize()
        return int(self.tok_host[0])

    @torch.no_grad()
    def generate(self, prompt: list[int], max_new_tokens: int, temperature: float | None = None, top_k: int | None = None,
                 seed: int | None = None, prefill: bool = True) -> list[int]:
        """the reference's generate() (generate.py:146-186): prompts longer than one token go through prefill() (one batched pass,
        Linears on the fused tensor-core kernel) unless prefill=False (one decode step per prompt token); a BOS-only prompt — the
        reference's benchmark protocol, generate.py:310-313 — is pure decode.  Greedy unless a temperature > 0 is given / set."""
        assert len(prompt) >= 1 and len(prompt) + max_new_tokens <= self.S
        if temperature is not None:
            self.set_sampling(temperature, top_k, seed)
        elif seed is not None:
            self.set_sampling(self.temperature, self.top_k, seed)
        self.reset(prompt[0])
        if prefill and len(prompt) > 1 and max_new_tokens >= 1:
            self.prefill(prompt)
            for _ in range(max_new_tokens - 1):
                self.step()
            self.stream.synchronize()
            return self.history[: len(prompt) + max_new_tokens].cpu().tolist()
        for t in prompt[1:]:  # teacher-forced, one decode step per prompt token
            self.step()
            with torch.cuda.stream(self.stream):
                self.token.fill_(t)
                self.history[int(self.pos_host())] = t
        for _ in range(max_new_tokens):
            self.step()
        self.stream.synchronize()
        n = len(prompt) + max_new_tokens
        return self.history[:n].cpu().tolist()

    def pos_host(self) -> int:
        self.stream.synchronize()
        return int(self.pos.cpu()[0])
