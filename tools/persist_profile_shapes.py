#!/usr/bin/env python
"""Per-job timing of the persistent engine on an arbitrary chain of Linear shapes (one GPU, no exchange), e.g. the per-rank
shard shapes of Llama-3-70B at 8 GPUs:   python tools/persist_profile_shapes.py 2 6 1280x8192 8192x1024 7168x8192 8192x3584
(args: bits, layers, then the N x K of the Linears of one block; every Linear feeds the next, widths are matched by slicing)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guidedquant_b200.persist import PersistentProgram  # noqa: E402

bits, layers = int(sys.argv[1]), int(sys.argv[2])
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[3:]]
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
prog = PersistentProgram(bits, dev)
width = max(max(s) for s in shapes)
src = torch.randn((1, width), device=dev, generator=g).half()
x = prog.buffer(width)
prog.pack(src, x)
bufs = [prog.buffer(width), prog.buffer(width)]
names = ["pack"]
wbytes = 0
for li in range(layers):
    for si, (N, K) in enumerate(shapes):
        q = torch.randint(-2**31, 2**31 - 1, (bits, N, K // 32), dtype=torch.int32, device=dev, generator=g)
        lut = (torch.randn((N, 1 << bits), device=dev, generator=g) * (1.0 / K ** 0.5)).half()
        out = bufs[(li * len(shapes) + si) % 2]
        prog.gemv(x, q, lut, out)
        wbytes += q.numel() * 4
        x = out
        names.append(f"{N}x{K}")
prog.finalize()
prog.launch()
torch.cuda.synchronize()
prog.check()
prog.enable_profile()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    prog.launch()
torch.cuda.synchronize()
ev0.record()
for _ in range(20):
    prog.launch()
ev1.record()
torch.cuda.synchronize()
prog.check()
tok_us = ev0.elapsed_time(ev1) * 1e3 / 20
P = prog.prof.cpu().numpy().astype(np.float64)
mhz = 1965.0
by = {}
for j in range(1, P.shape[1]):
    st, xr, sd, en = P[:, j, 0], P[:, j, 1], P[:, j, 2], P[:, j, 3]
    has = xr > 0
    if not has.any():
        continue
    by.setdefault(names[j], []).append(((xr - st)[has], (sd - xr)[has], (en - sd)[has], int(has.sum())))
print(json.dumps({"launch_us": round(tok_us, 1), "per_layer_us": round(tok_us / layers, 2), "GBs": round(wbytes / tok_us / 1e3, 1)}))
for n, rs in by.items():
    xw = np.concatenate([r[0] for r in rs]) / mhz
    stg = np.concatenate([r[1] for r in rs]) / mhz
    ep = np.concatenate([r[2] for r in rs]) / mhz
    print(n, json.dumps({"ctas": rs[0][3], "x_wait_load_us": [round(float(np.median(xw)), 2), round(float(xw.max()), 2)],
                         "stages_us": [round(float(np.median(stg)), 2), round(float(stg.max()), 2)],
                         "epilogue_us": [round(float(np.median(ep)), 2), round(float(ep.max()), 2)]}))
