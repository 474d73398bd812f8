#!/usr/bin/env python
"""In-graph time of the decode step with components removed (profiling aid): which part of the token costs what."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guidedquant_b200.model import APTransformer
model = sys.argv[1] if len(sys.argv) > 1 else "llama3-8b"
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tf = APTransformer(model, bits=bits, max_seq_len=512, pdl=os.environ.get("NOPDL") is None).random_init()
def run(skip, n=100):
    tf.debug_skip = set(skip); tf.graph = None; tf.capture(); tf.reset(1)
    for _ in range(10): tf.step()
    tf.stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(tf.stream):
        e0.record()
        for _ in range(n): tf.graph.replay()
        e1.record()
    tf.stream.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
import itertools
cfgs = [[], ["sample", "attn", "fusion"], ["sample", "attn", "lm_head", "fusion"], ["sample", "attn", "lm_head", "norm"],
        ["sample", "attn", "lm_head", "residual"], ["sample", "attn", "lm_head", "silu"], ["sample", "attn", "lm_head", "norm", "residual"],
        ["sample", "attn", "lm_head"]]
if len(sys.argv) > 3:
    cfgs = [c.split(",") if c else [] for c in sys.argv[3].split(";")]
for skip in cfgs:
    print(json.dumps({"skip": skip, "us_per_token": round(run(skip), 1)}), flush=True)
