#!/usr/bin/env python
"""A/B of the experimental GLU epilogue (APTransformer(glu_epilogue=True)) against the default decode
step on one GPU: first checks that both variants generate the same tokens on a small model, then times the captured
token graph of each.  Usage: ab_glu.py [model] [bits] [n_tokens]      (writes JSON lines to stdout)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guidedquant_b200.model import APTransformer  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "llama3-8b"
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 200

a = APTransformer("tiny128", bits=bits, max_seq_len=32).random_init(5)
b = APTransformer("tiny128", bits=bits, max_seq_len=32, glu_epilogue=True).random_init(5)
same = a.generate([1, 7, 3], 12) == b.generate([1, 7, 3], 12) and bool(torch.equal(a.logits, b.logits))
print(json.dumps({"check": "tiny128 tokens and logits identical", "ok": same}), flush=True)
del a, b


def time_tokens(glu):
    tf = APTransformer(model, bits=bits, max_seq_len=512, glu_epilogue=glu).random_init()
    tf.capture()
    tf.reset(1)
    for _ in range(20):
        tf.step()
    tf.stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(tf.stream):
        e0.record()
        for _ in range(n):
            tf.graph.replay()
        e1.record()
    tf.stream.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    del tf
    torch.cuda.empty_cache()
    return us


for glu in (False, True, False, True):
    us = time_tokens(glu)
    print(json.dumps({"model": model, "bits": bits, "glu_epilogue": glu, "us_per_token": round(us, 1), "tok_s": round(1e6 / us, 1)}), flush=True)
