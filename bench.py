#!/usr/bin/env python
"""bench.py — decode tok/s of the Any-Precision LUT GEMV hot path (BASELINE.json metric) on N B200s.

A step = ONE decoded token's pass over the hot path: the 4*L dependent APLinear GEMVs of the model
(wqkv, wo, w1w3, w2 per block; fused shapes of the reference's inference/model.py), launched through the
C-ABI library with programmatic dependent launch and replayed as one CUDA graph.

  value      tok/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e        tok/s through the public API ApGemvChain.step_host(): pinned-host x -> H2D -> graph -> D2H y
  roofline   achieved HBM GB/s of the dominant kernel (gemv_fast_kernel): algorithmic bytes of all GEMV
             launches of a step / measured step time, against the measured copy bandwidth
  cpu_baseline / --impl reference
             the reference's own CPU-runnable arm (BASELINE.json configs[0]): dequant -> fp16 torch.matmul of
             the same Linears on the host cores (oracle port), on a bounded sample (one block = 4 Linears)

N > 1 (torchrun): Megatron sharding of the same model (strong scaling): wqkv/w1w3 by rows, wo/w2 along K
with one NCCL all-reduce each.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default=None, help="default llama3-8b (BASELINE configs[1]) at every N; llama3-70b / llama2-70b / llama2-7b selectable")
    ap.add_argument("--bits", type=int, default=2)
    ap.add_argument("--layers", type=int, default=None, help="override layer count (debug only; invalidates the number)")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--l2-prefetch", action="store_true", help="L2-prefetch the next Linear (measured slower on B200; off)")
    return ap.parse_args()


METRIC = "decode tok/s Llama-3-8B 2-bit bs=1; ap_gemv HBM GB/s vs 8 TB/s peak"
UNIT = "tok/s"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBs"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s, MEASURED_PEAKS.json absent)"


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons of one GPU while the timed region runs (pynvml, 50 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_arm(model: str, bits: int, steps: int, warmup: int):
    """The reference's CPU-runnable arm: dequant -> fp16 torch.matmul (APLinear.gemm, APLinear.py:35-38) of the
    4 Linears of ONE block per step (a bounded 1/L sample of a token), all host cores."""
    import numpy as np
    import torch

    from guidedquant_b200.runtime import MODEL_CONFIGS, linear_shapes
    from oracle import oracle as O

    cfg = MODEL_CONFIGS[model]
    shapes = linear_shapes(cfg)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(0)
    lins = []
    for name, (N, K) in shapes.items():
        q = rng.integers(-2**31, 2**31 - 1, size=(bits, N, K // 32), dtype=np.int64).astype(np.int32)
        lut = (rng.standard_normal((N, 1 << bits)) / np.sqrt(K)).astype(np.float16)
        x = rng.standard_normal((1, 1, K)).astype(np.float16)
        lins.append((q, lut, x))

    def block():
        t = 0.0
        for q, lut, x in lins:
            _, dt, _ = O.cpu_reference_linear(q, lut, x, bits, threads=cores, dequant_each_call=True)
            t += dt
        return t

    for _ in range(max(0, min(warmup, 1))):
        block()
    steps = max(1, min(steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        block()
    t_block = (time.perf_counter() - t0) / steps
    tok_s = 1.0 / (t_block * cfg["n_layer"])
    return tok_s, t_block, cores, steps


def main():
    a = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    model = a.model or "llama3-8b"  # same workload at every N so the scaling series is comparable (70B: --model llama3-70b)
    workload = f"{model} {a.bits}-bit bs=1 decode, ap_gemv hot path ({'4*L' if a.layers is None else a.layers} APLinear GEMVs/token chain)"

    if a.impl == "reference":
        if rank != 0:
            return
        tok_s, t_block, cores, steps = cpu_reference_arm(model, a.bits, a.steps, a.warmup)
        sample = f"1 block (wqkv, wo, w1w3, w2) of {model} per step, dequant->fp16 torch.matmul on CPU, value scaled by n_layer"
        line = {
            "impl": "reference", "metric": METRIC, "value": tok_s, "unit": UNIT, "n_gpus": 0, "steps": steps, "warmup": 1,
            "ms_per_step": t_block * 1e3, "higher_is_better": True, "scaling": "strong" if max(world, a.gpus) > 1 else "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": workload, "reference_arm": "CPU port of APLinear.gemm (oracle/oracle.py)"},
            "cpu_baseline": {"value": tok_s, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tok_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line), flush=True)
        return

    import torch

    from guidedquant_b200 import _lib
    from guidedquant_b200.runtime import ApGemvChain

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback exists)")
    _lib.lib()  # fail loudly if the native library is missing
    torch.cuda.set_device(local_rank)
    pg = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        pg = dist.group.WORLD

    chain = ApGemvChain(model, bits=a.bits, n_layer=a.layers, pdl=not a.no_pdl, world_size=world, rank=rank,
                        process_group=pg, ctas_per_sm=a.ctas, l2_prefetch=a.l2_prefetch)
    chain.capture()
    d = chain.cfg["dim"]
    x_host = torch.randn((1, 1, d)).half().pin_memory()
    chain.x_in.copy_(x_host)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing
    for _ in range(max(3, a.warmup)):
        chain.step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(chain.stream):
        e0.record()
    for _ in range(a.steps):
        chain.step()
    with torch.cuda.stream(chain.stream):
        e1.record()
    barrier()
    dt = e0.elapsed_time(e1) * 1e-3
    # ---------------- end-to-end timing (host buffers, copies inside the timed region)
    for _ in range(3):
        chain.step_host(x_host)
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(chain.stream):
        e2 = torch.cuda.Event(enable_timing=True)
        e2.record()
    for _ in range(a.steps):
        y = chain.step_host(x_host)
    with torch.cuda.stream(chain.stream):
        e3 = torch.cuda.Event(enable_timing=True)
        e3.record()
    barrier()
    dt_e2e = max(e2.elapsed_time(e3) * 1e-3, 0.0)
    dt_e2e_wall = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    if world > 1:
        t = torch.tensor([dt, dt_e2e], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        dt, dt_e2e = float(t[0]), float(t[1])
    def finish():
        # a live CUDA graph that captured NCCL kernels makes communicator teardown hang: drop the graph, sync,
        # and leave without tearing the communicator down
        if world > 1:
            chain.graph = None
            torch.cuda.synchronize()
            sys.stdout.flush()
            os._exit(0)

    if rank != 0:
        finish()
        return

    tok_s = a.steps / dt
    ms = dt / a.steps * 1e3
    peak, peak_src = measured_peak_gbs()
    ab = chain.algo_bytes_per_step()                 # per rank, all GEMV launches of one token
    n_gemv = 4 * chain.cfg["n_layer"]
    achieved = ab / (dt / a.steps) / 1e9             # GB/s per GPU over the whole step (includes launch gaps)
    line = {
        "metric": METRIC, "value": tok_s, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {
            "workload": workload, "bits": a.bits, "gemv_launches_per_token": n_gemv,
            "parallelism": "single GPU" if world == 1 else f"tp{world}: wqkv/w1w3 N-sharded, wo/w2 K-sharded + NCCL all-reduce",
            "l2_policy": "inputs larger than L2: every GEMV reads its own distinct weights (%.2f GB/token/GPU), streamed evict-first" % (chain.weight_bytes() / 1e9),
            "pdl": not a.no_pdl, "l2_prefetch_next_linear": a.l2_prefetch, "accumulate": "fp16 chains of 8 -> fp32",
        },
        "clocks": sampler.result(),
        "e2e": {"value": a.steps / dt_e2e if dt_e2e > 0 else None, "unit": UNIT, "h2d_bytes_per_step": d * 2,
                "d2h_bytes_per_step": d * 2, "wall_value": a.steps / dt_e2e_wall},
        "gpu_launches": chain.launches_per_step * a.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "kernel": "apg::gemv_fast_kernel",
                     "algorithmic_bytes_per_step": ab, "launches_per_step": n_gemv,
                     "frac_of_8TBs": achieved / 8000.0},
    }
    # CPU baseline (rank 0, N = 1 only): bounded sample
    if world == 1:
        try:
            v, t_block, cores, st = cpu_reference_arm(model, a.bits, 2, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{st} x 1 block (4 Linears) of {model}, dequant->fp16 torch.matmul, scaled by n_layer; {t_block:.2f} s/block"}
        except Exception as e:  # the GPU number must not be lost to a host-side failure
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    main()
