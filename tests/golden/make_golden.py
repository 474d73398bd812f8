"""Generates tests/golden/pack_golden.npz by IMPORTING the reference's own packer
(/root/reference/any_precision/quantization/pack.py: pack_single_weight / unpack_single_weight).
Runs only in the build container (the reference tree is absent on the GPU box); the vectors it
writes are committed.  Usage:  NUMBA_CACHE_DIR=/tmp/numba python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")  # never write into /root/reference
sys.dont_write_bytecode = True
import numpy as np
import torch

REF = os.environ.get("REF_ROOT", "/root/reference")
spec = importlib.util.spec_from_file_location("ref_pack", os.path.join(REF, "any_precision/quantization/pack.py"))
ref_pack = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_pack)

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
cases = []
# (N, K, bits): full chunks, tail chunks (eff=24: 11008, eff=16: 13824/3584), tail-only (96, 32), odd sizes
for bits in (2, 3, 4, 5, 8):
    for (N, K) in ((4, 1024), (8, 2048), (4, 11008), (4, 13824), (8, 96), (4, 32), (12, 3584), (4, 1056)):
        if bits in (5, 8) and K > 2048:
            continue
        cases.append((N, K, bits))
for ci, (N, K, bits) in enumerate(cases):
    rng = np.random.default_rng(100 + ci)
    idx = rng.integers(0, 1 << bits, size=(N, 1, K), dtype=np.uint8)
    q = ref_pack.pack_single_weight(torch.from_numpy(idx), bits)          # int32 [bits,N,K/32]
    back = ref_pack.unpack_single_weight(torch.from_numpy(np.ascontiguousarray(q)), bits).numpy()
    assert np.array_equal(back, idx), "reference pack/unpack is not a round trip?!"
    out[f"c{ci}_idx"] = idx.reshape(N, K)
    out[f"c{ci}_q"] = np.ascontiguousarray(q).astype(np.int32)
    out[f"c{ci}_meta"] = np.array([N, K, bits], dtype=np.int64)
# all-zero / all-max edge rows
for bits in (2, 3, 4):
    for val, tag in ((0, "zero"), ((1 << bits) - 1, "max")):
        idx = np.full((4, 1, 2048 + 512), val, dtype=np.uint8)
        q = ref_pack.pack_single_weight(torch.from_numpy(idx), bits)
        out[f"e{bits}{tag}_idx"] = idx.reshape(4, -1)
        out[f"e{bits}{tag}_q"] = np.ascontiguousarray(q).astype(np.int32)
        out[f"e{bits}{tag}_meta"] = np.array([4, 2560, bits], dtype=np.int64)
np.savez_compressed(os.path.join(HERE, "pack_golden.npz"), **out)
print("wrote", len(out) // 3, "cases")
