#!/usr/bin/env python
"""Stand-alone timing of apd_lm_head and apd_attn_decode (graph of repeated launches)."""
import ctypes, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guidedquant_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
V, D = 128256, 4096
W = (torch.randn((V, D), device=dev) / 64).half()
x = torch.randn(D, device=dev).half(); nw = torch.ones(D, device=dev).half()
logits = torch.zeros(V, device=dev).half(); bv = torch.zeros(4096, device=dev); bi = torch.zeros(4096, dtype=torch.int32, device=dev)
def timeit(fn, n=20, reps=5):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n): fn()
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps): g.replay()
        e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3
for flags in (0, 4):
    t = timeit(lambda: L.apd_lm_head(x.data_ptr(), nw.data_ptr(), 1e-5, W.data_ptr(), logits.data_ptr(), V, D, bv.data_ptr(), bi.data_ptr(), None, 0, flags, torch.cuda.current_stream().cuda_stream))
    print(f"lm_head flags={flags}: {t:.1f} us  -> {2*V*D/t/1e3:.0f} GB/s")
H, Hkv, S = 32, 8, 512
qkv = torch.randn((H + 2 * Hkv) * 128, device=dev).half()
inv = (1.0 / (500000.0 ** (torch.arange(0, 128, 2).float() / 128))).to(dev)
kc = torch.randn((Hkv, S, 128), device=dev).half(); vc = torch.randn((Hkv, S, 128), device=dev).half()
out = torch.zeros(H * 128, device=dev).half(); pos = torch.zeros(1, dtype=torch.int32, device=dev)
part = torch.zeros(H * 8 * 132, device=dev)
for p_ in (0, 16, 64, 128, 256, 500):
    pos.fill_(p_)
    for ns in (1, 4):
        t = timeit(lambda: L.apd_attn_decode(qkv.data_ptr(), inv.data_ptr(), kc.data_ptr(), vc.data_ptr(), pos.data_ptr(), out.data_ptr(), part.data_ptr(), H, Hkv, S, ns, 1 / math.sqrt(128), 4, torch.cuda.current_stream().cuda_stream))
        print(f"attn pos={p_} nsplit={ns}: {t:.2f} us")
