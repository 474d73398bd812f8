"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU restatement (torch, float32 arithmetic with explicit fp16 roundings) of ONE decode step of the reference's
gpt-fast Transformer at batch 1 (inference/model.py: Transformer.forward :121-131, TransformerBlock.forward :151-167,
Attention.forward :206-236, FeedForward.forward :259-266, RMSNorm :274-285, rotate_half / apply_rotary_pos_emb
:268-272, :309-314, LlamaRotaryEmbedding.forward :381-405) and of greedy sampling (inference/generate.py:55-73 at
temperature 0).  Linear layers take DENSE weights (for APLinear: the oracle's dequantised W, i.e. the
"dequant -> fp16 matmul" semantics of APLinear.gemm, APLinear.py:35-38).

Parity status: PINNED to the reference's own model.py run on CPU in float32 (tests/golden/make_decode_golden.py ->
tests/golden/decode_golden.npz): with half_rounding=False this module reproduces those logits to float32 accuracy.
half_rounding=True additionally rounds to fp16 wherever the reference holds fp16 tensors when the model is .half().
"""
from __future__ import annotations

import math

import torch


class DecodeOracle:
    def __init__(self, weights: dict, n_layer: int, n_head: int, n_kv: int, dim: int, max_seq: int, rope_base: float,
                 eps: float = 1e-5, half_rounding: bool = True):
        """weights: name -> float tensor, names as in the reference's state dict:
        tok_embeddings.weight, layers.{i}.attention.wqkv.weight / .wo.weight, layers.{i}.feed_forward.w1w3.weight /
        .w2.weight, layers.{i}.input_layernorm.weight / .post_attention_layernorm.weight, norm.weight, output.weight"""
        self.w = {k: v.float() for k, v in weights.items()}
        self.L, self.H, self.Hkv, self.dim, self.S = n_layer, n_head, n_kv, dim, max_seq
        self.hd = dim // n_head
        self.eps, self.hr = eps, half_rounding
        # default rope init: inv_freq = 1 / base^(arange(0, dim, 2) / dim), attention_scaling = 1
        self.inv_freq = 1.0 / (rope_base ** (torch.arange(0, self.hd, 2, dtype=torch.int64).float() / self.hd))
        self.k_cache = [torch.zeros(n_kv, max_seq, self.hd) for _ in range(n_layer)]
        self.v_cache = [torch.zeros(n_kv, max_seq, self.hd) for _ in range(n_layer)]

    def r(self, t: torch.Tensor) -> torch.Tensor:  # "this tensor is fp16 in the reference"
        return t.half().float() if self.hr else t

    def rmsnorm(self, x, w):  # model.py:280-285
        n = x * torch.rsqrt(torch.mean(x * x, dim=-1, keepdim=True) + self.eps)
        return self.r(self.r(n) * w)

    def linear(self, x, W):  # fp16 GEMV with fp32 accumulation, fp16 output
        return self.r(W @ x)

    @staticmethod
    def rotate_half(x):  # model.py:268-272
        h = x.shape[-1] // 2
        return torch.cat((-x[..., h:], x[..., :h]), dim=-1)

    def step(self, token: int, pos: int) -> torch.Tensor:
        """logits (float) for `token` at position `pos`; updates the KV caches."""
        w, H, Hkv, hd = self.w, self.H, self.Hkv, self.hd
        x = self.r(w["tok_embeddings.weight"][token])
        freqs = self.inv_freq * float(pos)
        emb = torch.cat((freqs, freqs))
        cos, sin = self.r(emb.cos()), self.r(emb.sin())  # computed in fp32, cast to x.dtype (model.py:396-405)
        for i in range(self.L):
            p = f"layers.{i}."
            xn = self.rmsnorm(x, w[p + "input_layernorm.weight"])
            qkv = self.linear(xn, w[p + "attention.wqkv.weight"])
            q = qkv[: H * hd].view(H, hd)
            k = qkv[H * hd: (H + Hkv) * hd].view(Hkv, hd)
            v = qkv[(H + Hkv) * hd:].view(Hkv, hd)
            q = self.r(self.r(q * cos) + self.r(self.rotate_half(q) * sin))  # model.py:309-314
            k = self.r(self.r(k * cos) + self.r(self.rotate_half(k) * sin))
            self.k_cache[i][:, pos] = k
            self.v_cache[i][:, pos] = v
            kk = self.k_cache[i][:, : pos + 1].repeat_interleave(H // Hkv, dim=0)  # [H, T, hd] (model.py:229-230)
            vv = self.v_cache[i][:, : pos + 1].repeat_interleave(H // Hkv, dim=0)
            att = torch.softmax(torch.einsum("hd,htd->ht", q, kk) / math.sqrt(hd), dim=-1)  # SDPA, causal row
            y = self.r(torch.einsum("ht,htd->hd", att, vv)).reshape(-1)
            h = self.r(x + self.linear(y, w[p + "attention.wo.weight"]))  # model.py:152-155
            hn = self.rmsnorm(h, w[p + "post_attention_layernorm.weight"])
            gu = self.linear(hn, w[p + "feed_forward.w1w3.weight"])
            inter = gu.numel() // 2
            act = self.r(self.r(torch.nn.functional.silu(gu[:inter])) * gu[inter:])  # model.py:261-266
            x = self.r(h + self.linear(act, w[p + "feed_forward.w2.weight"]))  # model.py:165
        xn = self.rmsnorm(x, w["norm.weight"])
        return self.linear(xn, w["output.weight"])

    @staticmethod
    def greedy(logits: torch.Tensor) -> int:  # generate.py:55-73 with temperature 0 == argmax
        return int(torch.argmax(logits))


# ---------------------------------------------------------------------------------------------------------------------
# temperature / top-k sampling (generate.py:55-73) — oracle of apd_sample_topk_advance
# ---------------------------------------------------------------------------------------------------------------------
def sample_uniform(seed: int, pos: int, n: int):
    """u_i in (0,1), i < n: the documented counter hash of the kernel (decode_kernels.cuh sample_uniform), in numpy uint64."""
    import numpy as np

    with np.errstate(over="ignore"):
        ctr = (np.uint64(pos) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * ctr
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(41)).astype(np.float64) + 0.5) / 8388608.0


def sample_topk_reference(logits_f16, temperature: float, top_k, q):
    """The reference's own formulation, line by line (logits_to_probs + multinomial_sample_one_no_sync, generate.py:55-73),
    with the Exp(1) noise `q` passed in instead of drawn from torch's generator."""
    logits = torch.as_tensor(logits_f16).float()[None, :]
    logits = logits / max(temperature, 1e-5)
    if top_k is not None:
        v, _ = torch.topk(logits, min(top_k, logits.size(-1)))
        pivot = v.select(-1, -1).unsqueeze(-1)
        logits = torch.where(logits < pivot, -float("Inf"), logits)
    probs = torch.nn.functional.softmax(logits, dim=-1)
    return int(torch.argmax(probs / torch.as_tensor(q, dtype=torch.float32)[None, :], dim=-1)), probs[0]


def sample_topk_scores(logits_f16, temperature: float, top_k, seed: int, pos: int):
    """float64 scores l_i/T - log q_i (-inf outside the top-k, ties with the k-th kept); argmax == the sampled token.
    Equal to sample_topk_reference in exact arithmetic: softmax's normaliser is common to every i."""
    import numpy as np

    l = np.asarray(logits_f16, dtype=np.float16).astype(np.float64)
    n = l.shape[0]
    s = l / max(temperature, 1e-5)
    if top_k is not None and 0 < top_k < n:
        pivot = np.sort(l[~np.isnan(l)])[::-1][min(top_k, n) - 1] if np.count_nonzero(~np.isnan(l)) >= top_k else -np.inf
        s = np.where(l < pivot, -np.inf, s)
    s = np.where(np.isnan(l), -np.inf, s)
    q = -np.log(sample_uniform(seed, pos, n))
    return s - np.log(q), q


def sample_topk_sharded(logits_f16, temperature: float, top_k, seed: int, pos: int, world: int):
    """The vocab-sharded sampler's algorithm (apd_sample_topk_advance_tp, csrc/decode_kernels.cuh) restated on the CPU:
    every rank contributes its k largest logits, the k-th largest of that union is the global pivot, every rank then
    proposes its best score among its logits >= pivot, and the best proposal (ties: smaller global index) wins.
    Must equal argmax(sample_topk_scores(...)) on the unsharded logits — checked in tests/test_decode_oracle_cpu.py."""
    import numpy as np

    l = np.asarray(logits_f16, dtype=np.float16).astype(np.float64)
    n = l.shape[0]
    assert n % world == 0
    vl = n // world
    t = max(temperature, 1e-5)
    pivot = -np.inf
    if top_k is not None and 0 < top_k < n:
        assert top_k < vl
        union = np.concatenate([np.sort(l[r * vl:(r + 1) * vl])[::-1][:top_k] for r in range(world)])
        pivot = np.sort(union)[::-1][top_k - 1]
    q = -np.log(sample_uniform(seed, pos, n))  # hashed from the GLOBAL index
    best = (-np.inf, n)
    for r in range(world):
        sl = l[r * vl:(r + 1) * vl]
        s = np.where(sl < pivot, -np.inf, sl / t) - np.log(q[r * vl:(r + 1) * vl])
        i = int(np.argmax(s))  # first maximum = smallest index
        if s[i] > best[0] or (s[i] == best[0] and r * vl + i < best[1]):
            best = (s[i], r * vl + i)
    return best[1]
