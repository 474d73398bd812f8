"""Packed bit-plane layout of Any-Precision weights — host-side (numpy) pack / unpack / K-shard re-pack.

Layout contract (reference: any_precision/quantization/pack.py:12-83, 304-347):
    qweight int32 [bits, N, K/32]; plane j = 0 is the MSB.  Word w = i*32 + t of a (plane, row);
    bit 31-(8c+e) of that word is the plane's bit of   k = i*1024 + c*8*eff + 8t + e,
    eff = 32 for full 1024-weight chunks and (K % 1024)/32 (t < eff) for the tail chunk.

The reference packs with np.packbits + a numba byte permutation; this module uses the closed form
directly (vectorised gather/scatter over a precomputed k-permutation), and is checked against vectors
produced by the reference packer (tests/golden/pack_golden.npz).
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np


@lru_cache(maxsize=64)
def k_of_bit(K: int) -> np.ndarray:
    """perm[w*32 + p] = k stored at word w, bit position 31-p (p = 8c+e), for one (plane,row)."""
    assert K % 32 == 0 and K > 0, "K must be a positive multiple of 32"
    words = K // 32
    w = np.arange(words)
    i, t = w // 32, w % 32
    eff = np.where(i < K // 1024, 32, (K % 1024) // 32)
    p = np.arange(32)
    c, e = p // 8, p % 8
    k = (i * 1024 + 8 * t)[:, None] + c[None, :] * (8 * eff)[:, None] + e[None, :]
    k = k.reshape(-1)
    assert np.array_equal(np.sort(k), np.arange(K))
    return k


def pack_indices(idx: np.ndarray, bits: int) -> np.ndarray:
    """uint8 idx [N, K] (values < 2^bits) -> int32 qweight [bits, N, K/32]  (pack_single_weight, pack.py:304-321)."""
    idx = np.asarray(idx)
    if idx.ndim == 3:  # reference shape [N, group_count, group_size]
        idx = idx.reshape(idx.shape[0], -1)
    N, K = idx.shape
    perm = k_of_bit(K)
    g = np.ascontiguousarray(idx[:, perm]).astype(np.uint8).reshape(N, K // 32, 32)  # [N, word, p]
    out = np.empty((bits, N, K // 32), dtype=np.uint32)
    for j in range(bits):
        plane = (g >> (bits - 1 - j)) & 1
        by = np.packbits(plane, axis=2, bitorder="big")  # [N, word, 4]; byte 0 = p 0..7 = most significant byte
        out[j] = (by[..., 0].astype(np.uint32) << 24) | (by[..., 1].astype(np.uint32) << 16) | \
                 (by[..., 2].astype(np.uint32) << 8) | by[..., 3].astype(np.uint32)
    return out.view(np.int32)


def unpack_indices(qweight: np.ndarray, bits: int | None = None) -> np.ndarray:
    """int32 qweight [>=bits, N, K/32] -> uint8 idx [N, K] using the first `bits` planes
    (unpack_single_weight, pack.py:324-347; the any-precision property: the first b planes are the b-bit model)."""
    q = np.ascontiguousarray(qweight).view(np.uint32)
    P, N, words = q.shape
    bits = P if bits is None else bits
    K = words * 32
    perm = k_of_bit(K)
    acc = np.zeros((N, words, 32), dtype=np.uint8)
    for j in range(bits):
        w = q[j]
        by = np.stack([(w >> 24) & 0xFF, (w >> 16) & 0xFF, (w >> 8) & 0xFF, w & 0xFF], axis=-1).astype(np.uint8)
        plane = np.unpackbits(by, axis=2, bitorder="big")  # [N, word, 32] in p order
        acc |= (plane << (bits - 1 - j)).astype(np.uint8)
    idx = np.empty((N, K), dtype=np.uint8)
    idx[:, perm] = acc.reshape(N, K)
    return idx


def shard_k(qweight: np.ndarray, k0: int, k1: int) -> np.ndarray:
    """Re-pack the input-feature range [k0, k1) of a packed weight as a self-contained packed tensor
    (K-sharding for row-parallel multi-GPU).  Cuts at multiples of 1024 that end on a
    chunk boundary (or at K) are pure slices of the word axis; anything else goes through
    unpack -> slice -> pack because k is interleaved inside a chunk."""
    q = np.ascontiguousarray(qweight)
    bits, N, words = q.shape
    K = words * 32
    assert 0 <= k0 < k1 <= K and k0 % 32 == 0 and k1 % 32 == 0
    if k0 % 1024 == 0 and (k1 % 1024 == 0 or k1 == K):
        return np.ascontiguousarray(q[:, :, k0 // 32:k1 // 32])
    idx = unpack_indices(q, bits)
    return pack_indices(idx[:, k0:k1], bits)


def shard_bounds(K: int, world: int, align: int = 128) -> list[tuple[int, int]]:
    """Even split of K over `world` ranks with every cut a multiple of `align` (128 keeps the 128-bit
    vector path, K % 128 == 0 per shard)."""
    assert K % align == 0, f"K={K} not a multiple of {align}"
    units = K // align
    cuts = [(units * r) // world * align for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_k_torch(qweight, k0: int, k1: int):
    """shard_k for torch tensors on any device (used at model-shard time, where the numpy path costs seconds per
    Linear): same contract and bit-identical results."""
    import torch

    bits, N, words = qweight.shape
    K = words * 32
    assert 0 <= k0 < k1 <= K and k0 % 32 == 0 and k1 % 32 == 0
    if k0 % 1024 == 0 and (k1 % 1024 == 0 or k1 == K):
        return qweight[:, :, k0 // 32:k1 // 32].contiguous()
    dev = qweight.device
    sh = (31 - torch.arange(32, device=dev, dtype=torch.int32)).view(1, 1, 1, 32)
    # plane bits in stored (word, p) order, then scattered to k order
    b = ((qweight.unsqueeze(-1) >> sh) & 1).to(torch.uint8).reshape(bits, N, K)
    perm = torch.from_numpy(k_of_bit(K)).to(dev)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(K, device=dev)
    bk = b.index_select(2, inv)[:, :, k0:k1]           # [bits, N, K'] in k order
    Kn = k1 - k0
    permn = torch.from_numpy(k_of_bit(Kn)).to(dev)
    bs = bk.index_select(2, permn).reshape(bits, N, Kn // 32, 32).to(torch.int64)
    w = (bs << (31 - torch.arange(32, device=dev, dtype=torch.int64)).view(1, 1, 1, 32)).sum(dim=-1)
    w = torch.where(w >= 2**31, w - 2**32, w)
    return w.to(torch.int32).contiguous()
