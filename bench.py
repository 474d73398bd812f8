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
    ap.add_argument("--workload", default="decode", choices=["decode", "gemv-chain"],
                    help="decode: the whole token step (tensor-parallel for N > 1); gemv-chain: only the 4*L APLinear GEMVs")
    ap.add_argument("--max-seq", type=int, default=512)
    ap.add_argument("--collective", default="push", choices=["push", "nccl"],
                    help="N > 1: fused one-shot all-reduce pushed from the GEMV epilogue over NVLink peer memory, or NCCL")
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
                    v = d[k]
                    if isinstance(v, dict):  # tolerate {"value": ..} / {"burst": ..} wrappers
                        v = v.get("value", v.get("burst", v.get("gbs")))
                    return float(v), "measured (MEASURED_PEAKS.json)"
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


def cpu_reference_arm(model: str, bits: int, steps: int, warmup: int, lm_head: bool = True):
    """The reference's CPU-runnable arm: dequant -> fp16 torch.matmul (APLinear.gemm, APLinear.py:35-38) of the
    4 Linears of ONE block per step (a bounded 1/L sample of a token), all host cores; with `lm_head` the fp16 output
    projection (model.py:94,129) is timed once and added per token.  Attention / norms / sampling at batch 1 are < 1 % of
    the CPU token and are left out (which only flatters the CPU arm)."""
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
    t_head = 0.0
    if lm_head:
        Wh = torch.empty((cfg["vocab"], cfg["dim"]), dtype=torch.float16).normal_(0, 0.02)
        xh = torch.randn((1, cfg["dim"])).half()
        torch.matmul(xh, Wh.T)
        t1 = time.perf_counter()
        torch.matmul(xh, Wh.T)
        t_head = time.perf_counter() - t1
    tok_s = 1.0 / (t_block * cfg["n_layer"] + t_head)
    return tok_s, t_block, cores, steps, t_head


def main():
    a = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    model = a.model or "llama3-8b"  # same workload at every N so the scaling series is comparable (70B: --model llama3-70b)
    full_decode = a.workload == "decode" and a.impl == "ours"
    workload = (f"{model} {a.bits}-bit bs=1 decode, full token step: embed + L x (wqkv|attn|wo|w1w3|w2) + lm_head + greedy sample"
                if full_decode else
                f"{model} {a.bits}-bit bs=1 decode, ap_gemv hot path only ({'4*L' if a.layers is None else a.layers} APLinear GEMVs/token chain)")

    if a.impl == "reference":
        if rank != 0:
            return
        whole = a.workload == "decode"
        tok_s, t_block, cores, steps, t_head = cpu_reference_arm(model, a.bits, a.steps, a.warmup, lm_head=whole)
        if whole:  # same workload as our arm's default
            workload = f"{model} {a.bits}-bit bs=1 decode, full token step: embed + L x (wqkv|attn|wo|w1w3|w2) + lm_head + greedy sample"
        sample = (f"1 block (wqkv, wo, w1w3, w2) of {model} per step, dequant->fp16 torch.matmul on CPU, value = 1 / (n_layer x block"
                  + (f" + fp16 lm_head {t_head:.2f} s)" if whole else ")"))
        line = {
            "impl": "reference", "metric": METRIC, "value": tok_s, "unit": UNIT, "n_gpus": max(world, a.gpus), "steps": steps, "warmup": 1,
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
    from guidedquant_b200.model import APTransformer
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

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, stream, steps, warm):
        for _ in range(warm):
            step_fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
        for _ in range(steps):
            step_fn()
        with torch.cuda.stream(stream):
            e1.record()
        barrier()
        dt = e0.elapsed_time(e1) * 1e-3
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            dt = float(t[0])
        return dt

    warm = max(3, a.warmup)
    full = a.workload == "decode"
    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---------------- the hot path alone: the per-token GEMV chain (always measured: it carries the roofline figure)
    chain = ApGemvChain(model, bits=a.bits, n_layer=a.layers, pdl=not a.no_pdl, world_size=world, rank=rank,
                        process_group=pg, ctas_per_sm=a.ctas, l2_prefetch=a.l2_prefetch, collective=a.collective)
    chain.capture()
    d = chain.cfg["dim"]
    x_host = torch.randn((1, 1, d)).half().pin_memory()
    chain.x_in.copy_(x_host)
    args_steps_chain = a.steps
    dt_chain = timed(chain.step, chain.stream, a.steps, warm)
    if not full:
        dt = dt_chain
        dt_e2e = timed(lambda: chain.step_host(x_host), chain.stream, a.steps, 3)
        launches, h2d, d2h = chain.launches_per_step, d * 2, d * 2
    ab_chain = chain.algo_bytes_per_step()
    n_gemv = 4 * chain.cfg["n_layer"]
    wbytes = chain.weight_bytes()

    # ---------------- the whole decode step (1 GPU): embed + L blocks (5 launches each) + lm_head + greedy sample
    bytes_tok = None
    tf = None
    if full:
        chain.graph = None
        chain = None
        torch.cuda.empty_cache()
        tf = APTransformer(model, bits=a.bits, max_seq_len=a.max_seq, pdl=not a.no_pdl, n_layer=a.layers,
                           world_size=world, rank=rank, process_group=pg).random_init()
        tf.capture()
        n_tok = min(a.steps, a.max_seq - 2)

        def run_tokens(step_fn, n, w):
            # every timed step is a NEW token at the next position (BOS-only prompt protocol, generate.py:310-313)
            tf.reset(1)
            return timed(step_fn, tf.stream, n, w)

        tok_pinned = torch.ones(1, dtype=torch.int32).pin_memory()
        warm_d = min(warm, max(1, a.max_seq - 2 - n_tok))
        dt = run_tokens(tf.step, n_tok, warm_d)
        dt_e2e = run_tokens(lambda: tf.step_host(tok_pinned), n_tok, min(3, warm_d))
        a.steps = n_tok
        launches, h2d, d2h = tf.launches_per_token, 4, 4
        bytes_tok = tf.algo_bytes_per_token(pos=n_tok // 2)

    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    def finish():
        # a live CUDA graph that captured NCCL kernels makes communicator teardown hang: drop the graph, sync,
        # and leave without tearing the communicator down
        if world > 1:
            if chain is not None:
                chain.graph = None
            if tf is not None:
                tf.graph = None
            torch.cuda.synchronize()
            sys.stdout.flush()
            os._exit(0)

    if rank != 0:
        finish()
        return

    tok_s = a.steps / dt
    ms = dt / a.steps * 1e3
    peak, peak_src = measured_peak_gbs()
    t_chain = dt_chain / args_steps_chain  # seconds per token of the GEMV chain alone
    achieved = ab_chain / t_chain / 1e9          # GB/s per GPU of the GEMV launches (launch gaps included)
    traffic = None                                # measured DRAM bytes per launch (ncu --set full), when recorded for this config
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_gemv_traffic.json")))
        if world == 1 and tj.get("model") == model and tj.get("bits") == a.bits and a.layers is None:
            traffic = tj["per_layer_bytes"] / 4.0
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": tok_s, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": warm,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {
            "workload": workload, "bits": a.bits, "gemv_launches_per_token": n_gemv,
            "parallelism": "single GPU" if world == 1 else f"tp{world}: wqkv/w1w3 N-sharded (heads / MLP columns), wo/w2 K-sharded + " + ("one-shot all-reduce fused into the GEMV epilogue (NVLink peer stores)" if (a.collective == "push" or full) else "NCCL all-reduce") + ("; lm_head vocab-sharded + arg-max exchange over peer memory" if full else ""),
            "l2_policy": "inputs larger than L2: every GEMV reads its own distinct weights (%.2f GB/token/GPU), streamed evict-first" % (wbytes / 1e9),
            "pdl": not a.no_pdl, "l2_prefetch_next_linear": a.l2_prefetch, "accumulate": "fp16 chains of 8 -> fp32",
        },
        "clocks": sampler.result(),
        "e2e": {"value": a.steps / dt_e2e if dt_e2e > 0 else None, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches * a.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_note": "avg DRAM bytes per GEMV launch from profiles/r1_gemv_traffic.json (ncu --set full); algorithmic avg %.0f" % (ab_chain / n_gemv),
                     "peak_source": peak_src, "kernel": "apg::gemv_fast_kernel",
                     "algorithmic_bytes_per_step": ab_chain, "launches_per_step": n_gemv,
                     "timed_as": "the %d GEMV launches of a token replayed as their own CUDA graph in this process: %.1f us/token" % (n_gemv, t_chain * 1e6),
                     "frac_of_8TBs": achieved / 8000.0},
    }
    if full:
        line["config"]["max_seq_len"] = a.max_seq
        line["config"]["launches_per_token"] = launches
        line["config"]["sampling"] = "greedy (temperature 0), token and position advanced on the device"
        line["decode_bytes"] = {"per_token": bytes_tok, "achieved_GBs_all_bytes": bytes_tok["total"] / (dt / a.steps) / 1e9,
                                "frac_of_peak_all_bytes": bytes_tok["total"] / (dt / a.steps) / 1e9 / peak}
    # CPU baseline (rank 0, N = 1 only): bounded sample
    if world == 1:
        try:
            v, t_block, cores, st, t_head = cpu_reference_arm(model, a.bits, 2, 1, lm_head=full_decode)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{st} x 1 block (4 Linears) of {model}, dequant->fp16 torch.matmul, scaled by n_layer; {t_block:.2f} s/block"
                                              + (f" + fp16 lm_head {t_head:.2f} s/token" if full_decode else "")}
        except Exception as e:  # the GPU number must not be lost to a host-side failure
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    main()
