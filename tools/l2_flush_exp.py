#!/usr/bin/env python
"""Does an L2-thrashing stream between tokens slow the GEMV chain?  chain alone vs chain + 1 GB read (torch sum) per token."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guidedquant_b200.runtime import ApGemvChain
ch = ApGemvChain("llama3-8b", bits=2).capture()
big = torch.zeros(512 * 1024 * 1024, dtype=torch.float16, device="cuda")  # 1 GB
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ch.stream):
        e0.record()
        for _ in range(n): fn()
        e1.record()
    ch.stream.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def chain_only(): ch.step()
def stream_only():
    with torch.cuda.stream(ch.stream): big.sum()
def both():
    ch.step()
    with torch.cuda.stream(ch.stream): big.sum()
a, b, c = t(chain_only), t(stream_only), t(both)
print(f"chain {a:.1f} us, 1GB read {b:.1f} us, both {c:.1f} us, excess {c - a - b:.1f} us")
