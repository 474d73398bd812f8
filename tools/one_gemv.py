#!/usr/bin/env python
"""Launch a few GEMVs of one shape over rotating weights (for ncu).  Usage: one_gemv.py SHAPE BITS CTAS [N_LAUNCH]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guidedquant_b200 import ap_gemv
from tools.microbench import SHAPES
name, bits, ctas = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
nl = int(sys.argv[4]) if len(sys.argv) > 4 else 8
N, K = SHAPES[name]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
qs = [torch.randint(-2**31, 2**31 - 1, (bits, N, K // 32), dtype=torch.int32, device=dev, generator=g) for _ in range(nl)]
lut = (torch.randn((N, 1 << bits), device=dev, generator=g) * 0.02).half()
x = torch.randn((1, 1, K), device=dev, generator=g).half()
out = torch.zeros((1, 1, N), dtype=torch.float16, device=dev)
for q in qs:
    ap_gemv.anyprec_gemv_ex(x, out, q, lut, bits, ctas_per_sm=ctas)
torch.cuda.synchronize()
print("done", float(out.float().abs().sum()))
