#!/usr/bin/env python
"""Records outputs of the UNMODIFIED reference kernels (oracle/_ref/libapgemv_ref.so = anyprec.cu compiled for
sm_100a) on a B200 as golden fixtures: gpurun_out/ref_gpu_golden.npz -> commit as tests/golden/ref_gpu_b200.npz.
The CPU suite (tests/test_oracle_cpu.py::test_oracle_against_reference_gpu_golden) then pins the oracle's
dequant and reference-order GEMV emulation to the real reference, bit for bit, without a GPU."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import refgpu  # noqa: E402

CASES = [  # (N, K, bits, M): small enough to commit, covering full/tail chunks, bits 2..8, M 1..8
    (16, 1024, 2, 1), (16, 2048, 3, 1), (16, 4096, 4, 1), (8, 11008, 2, 1), (8, 13824, 3, 1), (16, 96, 4, 1),
    (16, 1056, 2, 1), (16, 1024, 5, 1), (16, 2048, 6, 1), (16, 1024, 7, 1), (16, 1024, 8, 1),
    (16, 2048, 2, 2), (16, 1024, 3, 4), (16, 2048, 4, 8), (8, 14336, 2, 1), (4, 28672, 2, 1),
]
out = {}
for ci, (N, K, bits, M) in enumerate(CASES):
    idx, q, lut, x = O.synth_layer(N, K, bits, seed=500 + ci, M=M)
    qq, ll, xx = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (q, lut, x))
    y = refgpu.ref_gemv(xx, qq, ll, bits)
    w = refgpu.ref_dequant(qq, ll, bits)
    torch.cuda.synchronize()
    y = y.cpu().numpy().reshape(M, N)
    w = w.cpu().numpy()
    nm = f"c{ci}"
    out[nm + "_meta"] = np.array([N, K, bits, M], dtype=np.int64)
    out[nm + "_q"], out[nm + "_lut"], out[nm + "_x"], out[nm + "_y"] = q, lut, x, y
    out[nm + "_w0"], out[nm + "_w1"] = w[0].copy(), w[-1].copy()
    out[nm + "_wsum"] = np.array(int(w.view(np.uint16).astype(np.uint64).sum()), dtype=np.uint64)
    # report agreement with the oracle right here as well
    W = O.dequant(q, lut, bits)
    ok_w = np.array_equal(W.view(np.uint16), w.view(np.uint16))
    ok_y = np.array_equal(O.gemv_ref_order_f16(W, x).view(np.uint16), y.view(np.uint16))
    print(f"case {ci} N={N} K={K} bits={bits} M={M}: dequant==oracle {ok_w}, gemv==oracle-emulation {ok_y}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_gpu_golden.npz"), **out)
print("wrote gpurun_out/ref_gpu_golden.npz")
