"""A/B: GEMV chain with / without the L2 prefetch of the next Linear's weights (ApGemvChain(l2_prefetch=...))"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from guidedquant_b200.runtime import ApGemvChain
model = sys.argv[1] if len(sys.argv) > 1 else "llama3-8b"
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 2
for rep in range(2):
    for pf in (False, True):
        ch = ApGemvChain(model, bits=bits, engine="launches", l2_prefetch=pf)
        ch.capture()
        for _ in range(5):
            ch.step()
        ch.stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ch.stream):
            e0.record()
            for _ in range(50):
                ch.step()
            e1.record()
        ch.stream.synchronize()
        print(json.dumps({"l2_prefetch": pf, "chain_us": round(e0.elapsed_time(e1) / 50 * 1e3, 1)}), flush=True)
        ch.graph = None
        del ch
        torch.cuda.empty_cache()
