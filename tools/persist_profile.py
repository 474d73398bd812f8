#!/usr/bin/env python
"""Where does a persistent-engine token go?  Per-job clock stamps of every CTA (PersistentProgram.enable_profile):
prints, per job type / Linear, the median over CTAs of: wait for x + load, stages, epilogue, and the job's span over all CTAs."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guidedquant_b200.runtime import ApGemvChain  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "llama3-8b"
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 2
layers = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ch = ApGemvChain(model, bits=bits, n_layer=layers, engine="persistent")
ch.capture()
ch.x_in.copy_(torch.randn((1, 1, ch.cfg["dim"]), device="cuda").half())
ch.prog.enable_profile()
for _ in range(3):
    ch.step()
ch.stream.synchronize()
ch.prog.check()
P = ch.prog.prof.cpu().numpy().astype(np.float64)        # [sms, jobs, 4]
mhz = 1965.0
names = ["pack"] + [n for _ in range(layers) for n in ("wqkv", "wo", "w1w3", "w2")]
t0 = P[:, 0, 0].min()
rows = []
for j in range(P.shape[1]):
    st, xr, sd, en = P[:, j, 0], P[:, j, 1], P[:, j, 2], P[:, j, 3]
    has = xr > 0
    rec = {"job": j, "name": names[j] if j < len(names) else "?",
           "start_us_min": (st.min() - t0) / mhz, "end_us_max": (en.max() - t0) / mhz,
           "span_us": (en.max() - st.min()) / mhz,
           "x_wait_load_us_med": float(np.median((xr - st)[has])) / mhz if has.any() else None,
           "stages_us_med": float(np.median((sd - xr)[has])) / mhz if has.any() else None,
           "stages_us_max": float(np.max((sd - xr)[has])) / mhz if has.any() else None,
           "epilogue_us_med": float(np.median((en - sd)[has])) / mhz if has.any() else None}
    rows.append(rec)
for r in rows[: 1 + 4 * min(layers, 3)]:
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()}))
tot = (P[:, -1, 3].max() - t0) / mhz
print(json.dumps({"total_us": round(tot, 1), "per_layer_us": round(tot / layers, 2)}))
by = {}
for r in rows[1:]:
    by.setdefault(r["name"], []).append(r)
for n, rs in by.items():
    print(n, {k: round(float(np.mean([r[k] for r in rs])), 2) for k in ("span_us", "x_wait_load_us_med", "stages_us_med", "stages_us_max", "epilogue_us_med")})
