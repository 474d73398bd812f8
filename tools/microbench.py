#!/usr/bin/env python
"""Per-Linear microbenchmark of the GEMV kernels on one GPU (SURVEY.md §8d protocol): a CUDA graph that
walks a rotating set of >= 512 MB of distinct packed weights (so nothing is served from the 126 MB L2),
timed with CUDA events over many replays.  Times our kernel (several grid settings) and, when
oracle/_ref is present, the UNMODIFIED reference kernel on the same tensors.  Writes JSON lines."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guidedquant_b200 import _lib, ap_gemv  # noqa: E402
from tests import refgpu  # noqa: E402

SHAPES = {  # name: (N, K)
    "l3-8b.wqkv": (6144, 4096), "l3-8b.wo": (4096, 4096), "l3-8b.w1w3": (28672, 4096), "l3-8b.w2": (4096, 14336),
    "70b.wqkv": (10240, 8192), "70b.wo": (8192, 8192), "70b.w1w3": (57344, 8192), "70b.w2": (8192, 28672),
}


def algo_bytes(N, K, bits, M=1):
    return bits * N * K // 8 + N * (1 << bits) * 2 + M * K * 2 + M * N * 2


def time_graph(fn_list, replays=20, warm=3):
    """fn_list: callables launching one kernel each on the current stream; captured into one graph."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for f in fn_list[: min(4, len(fn_list))]:
            f()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for f in fn_list:
                f()
        for _ in range(warm):
            g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(replays):
            g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / (replays * len(fn_list))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="l3-8b.wqkv,l3-8b.wo,l3-8b.w1w3,l3-8b.w2")
    ap.add_argument("--bits", default="2,3,4")
    ap.add_argument("--ctas", default="0,1,2,3,4")
    ap.add_argument("--rot-mb", type=int, default=512)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "microbench.jsonl"))
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--M", type=int, default=1, help="batch rows (1..8); M > 1 and bits > 4 run the generic kernel")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    dev = torch.device("cuda:0")
    fout = open(a.out, "a")
    for name in a.shapes.split(","):
        N, K = SHAPES[name]
        for bits in map(int, a.bits.split(",")):
            wbytes = bits * N * K // 8
            nrot = max(4, min(256, (a.rot_mb << 20) // wbytes + 1))
            g = torch.Generator(device=dev).manual_seed(bits)
            qs = [torch.randint(-2**31, 2**31 - 1, (bits, N, K // 32), dtype=torch.int32, device=dev, generator=g)
                  for _ in range(nrot)]
            lut = (torch.randn((N, 1 << bits), device=dev, generator=g) * 0.02).half()
            x = torch.randn((a.M, 1, K), device=dev, generator=g).half()
            out = torch.zeros((a.M, 1, N), dtype=torch.float16, device=dev)
            ab = algo_bytes(N, K, bits, a.M)
            rec = {"shape": name, "N": N, "K": K, "bits": bits, "M": a.M, "algo_bytes": ab, "nrot": nrot}
            for c in map(int, a.ctas.split(",")):
                for pdl in (0, 1):
                    flags = _lib.APG_FLAG_PDL if pdl else 0
                    fns = [(lambda q=q: ap_gemv.anyprec_gemv_ex(x, out, q, lut, bits, flags=flags, ctas_per_sm=c)) for q in qs]
                    t = time_graph(fns)
                    rec[f"ours_c{c}_pdl{pdl}_us"] = round(t * 1e6, 3)
                    rec[f"ours_c{c}_pdl{pdl}_GBs"] = round(ab / t / 1e9, 1)
            if refgpu.available() and not a.no_ref and N % (4 if a.M == 1 else 16) == 0:
                fns = [(lambda q=q: refgpu.ref().ref_anyprec_gemv(x.data_ptr(), out.data_ptr(), q.data_ptr(), lut.data_ptr(),
                                                                   a.M, N, K, bits, torch.cuda.current_stream().cuda_stream)) for q in qs]
                t = time_graph(fns)
                rec["ref_us"] = round(t * 1e6, 3)
                rec["ref_GBs"] = round(ab / t / 1e9, 1)
            print(json.dumps(rec), flush=True)
            fout.write(json.dumps(rec) + "\n")
            fout.flush()
            del qs
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
