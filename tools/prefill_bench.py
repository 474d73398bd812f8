"""Prefill Linear: fused dequant + tcgen05 GEMM (apg_prefill_gemm) vs the reference's route (anyprec_dequant + library
fp16 matmul, inference/ap_gemv/APLinear.py:35-38), CUDA-event timed, inputs > L2 rotated between iterations.
    python tools/prefill_bench.py [--bits 2 3 4] [--seqs 16 64 256 1024 2048] [--shapes 4096x4096 ...] > out.jsonl"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from guidedquant_b200 import ap_gemv  # noqa: E402


def timed(fn, iters, calls, warm=2):
    """median us per call of `fn`; `calls` consecutive calls (rotating weight copies) are captured into ONE CUDA graph so
    that the host-side cost of a call (allocation, plan, tensor-map encode: tens of us) is not what is measured"""
    for _ in range(warm * calls):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(calls):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        g.replay()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 / calls for i in range(iters))
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", type=int, nargs="+", default=[2, 3, 4])
    ap.add_argument("--seqs", type=int, nargs="+", default=[16, 32, 64, 128, 256, 512, 1024, 2048])
    ap.add_argument("--shapes", nargs="+", default=["6144x4096", "4096x4096", "28672x4096", "4096x14336"])
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(1)
    for shape in a.shapes:
        N, K = map(int, shape.split("x"))
        for bits in a.bits:
            # several copies of the weights so that consecutive iterations do not find them in the 126 MB L2
            ncopy = max(2, min(8, (192 << 20) // (N * K * bits // 8) + 1))
            qs = [torch.randint(-2**31, 2**31 - 1, (bits, N, K // 32), dtype=torch.int32, device="cuda", generator=g) for _ in range(ncopy)]
            lut = (torch.randn((N, 1 << bits), device="cuda", generator=g) * 0.02).half()
            for T in a.seqs:
                x = torch.randn((T, K), device="cuda", generator=g).half()
                it = [0]

                def fused():
                    it[0] += 1
                    return ap_gemv.anyprec_prefill_gemm(x, qs[it[0] % ncopy], lut, bits)

                def ref():
                    it[0] += 1
                    return torch.matmul(x, ap_gemv.anyprec_dequant(qs[it[0] % ncopy], lut, bits).T)

                W = ap_gemv.anyprec_dequant(qs[0], lut, bits)

                def gemm_only():
                    return torch.matmul(x, W.T)

                y = ap_gemv.anyprec_prefill_gemm(x, qs[0], lut, bits)
                truth = x.double() @ W.double().T
                err = float((y.double() - truth).abs().max() / truth.abs().max())
                t_f, t_r, t_g = timed(fused, a.iters, ncopy), timed(ref, a.iters, ncopy), timed(gemm_only, a.iters, ncopy)
                flops = 2.0 * T * N * K
                print(json.dumps({"N": N, "K": K, "bits": bits, "T": T, "fused_us": round(t_f, 2), "dequant_matmul_us": round(t_r, 2),
                                  "matmul_only_us": round(t_g, 2), "speedup": round(t_r / t_f, 3), "fused_tflops": round(flops / t_f / 1e6, 1),
                                  "err_vs_f64": err}), flush=True)
                del W, truth, y


if __name__ == "__main__":
    main()
