#!/usr/bin/env python
"""anyprec_dequant (ours) vs the UNMODIFIED reference dequant kernel (oracle/_ref) on one GPU: CUDA-graph replay over a
rotating set of packed weights, CUDA events.  Output bytes = 2*N*K (HBM-write bound).  Writes JSON lines."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guidedquant_b200 import _lib  # noqa: E402
from tests import refgpu  # noqa: E402
from tools.microbench import SHAPES, time_graph  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="l3-8b.wo,l3-8b.w1w3")
    ap.add_argument("--bits", default="2,3,4,8")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "dequant_bench.jsonl"))
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    L = _lib.lib()
    fout = open(a.out, "a")
    for name in a.shapes.split(","):
        N, K = SHAPES[name]
        for bits in map(int, a.bits.split(",")):
            g = torch.Generator(device=dev).manual_seed(bits)
            nrot = 4
            qs = [torch.randint(-2**31, 2**31 - 1, (bits, N, K // 32), dtype=torch.int32, device=dev, generator=g) for _ in range(nrot)]
            lut = (torch.randn((N, 1 << bits), device=dev, generator=g) * 0.02).half()
            outs = [torch.empty((N, K), dtype=torch.float16, device=dev) for _ in range(nrot)]
            st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
            rec = {"shape": name, "N": N, "K": K, "bits": bits, "out_bytes": 2 * N * K}
            fns = [(lambda q=q, o=o: _lib.check(L.apg_dequant(q.data_ptr(), lut.data_ptr(), o.data_ptr(), N, K, bits, st()), "dq"))
                   for q, o in zip(qs, outs)]
            t = time_graph(fns)
            rec["ours_us"], rec["ours_GBs"] = round(t * 1e6, 2), round((2 * N * K + bits * N * K // 8) / t / 1e9, 1)
            if refgpu.available():
                fns = [(lambda q=q, o=o: refgpu.ref().ref_anyprec_dequant(q.data_ptr(), lut.data_ptr(), o.data_ptr(), N, K, bits, st()))
                       for q, o in zip(qs, outs)]
                t = time_graph(fns)
                rec["ref_us"], rec["ref_GBs"] = round(t * 1e6, 2), round((2 * N * K + bits * N * K // 8) / t / 1e9, 1)
                w = torch.empty((N, K), dtype=torch.float16, device=dev)
                refgpu.ref().ref_anyprec_dequant(qs[0].data_ptr(), lut.data_ptr(), w.data_ptr(), N, K, bits, st())
                torch.cuda.synchronize()
                rec["bit_identical"] = bool(torch.equal(w.view(torch.int16), outs[0].view(torch.int16)))
            print(json.dumps(rec), flush=True)
            fout.write(json.dumps(rec) + "\n")
            fout.flush()


if __name__ == "__main__":
    main()
