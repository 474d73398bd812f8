#!/usr/bin/env python
"""Run a few decode tokens eagerly (no graph) for ncu launch lists.  Usage: one_decode.py [model] [bits] [ntok]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guidedquant_b200.model import APTransformer
model = sys.argv[1] if len(sys.argv) > 1 else "llama3-8b"
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ntok = int(sys.argv[3]) if len(sys.argv) > 3 else 3
tf = APTransformer(model, bits=bits, max_seq_len=512, engine=os.environ.get("ENGINE") or None).random_init()
tf.reset(1)
with torch.cuda.stream(tf.stream):
    for _ in range(ntok):
        tf.decode_step()
tf.stream.synchronize()
print("tokens", tf.history[: ntok + 1].cpu().tolist())
