"""one fused prefill Linear (for ncu): python tools/one_prefill.py N K T bits [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from guidedquant_b200 import ap_gemv  # noqa: E402

N, K, T, bits = (int(v) for v in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randint(-2**31, 2**31 - 1, (bits, N, K // 32), dtype=torch.int32, device="cuda", generator=g)
lut = (torch.randn((N, 1 << bits), device="cuda", generator=g) * 0.02).half()
x = torch.randn((T, K), device="cuda", generator=g).half()
for _ in range(reps):
    y = ap_gemv.anyprec_prefill_gemm(x, q, lut, bits)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
