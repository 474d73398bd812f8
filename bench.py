#!/usr/bin/env python
"""bench.py — decode tok/s of the Any-Precision LUT GEMV hot path (BASELINE.json metric) on N B200s.

A step = ONE decoded token of Llama-3-8B 2-bit at batch 1 (BASELINE configs[1]): embedding, 32 blocks (fused wqkv ->
RoPE/KV/attention -> wo -> w1w3 -> w2, every Linear an Any-Precision LUT GEMV), fp16 lm_head, greedy sampling.

  value        tok/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e          tok/s through the public API APTransformer.step_host(): pinned token id -> H2D -> token step -> D2H token
  roofline     achieved HBM GB/s of the hot path alone: the 4*L GEMVs of a token run back to back by the same engine,
               algorithmic bytes / measured time, against the measured copy bandwidth (MEASURED_PEAKS.json)
  parity       MEASURED max error of the benchmarked path vs the unmodified reference kernel and vs fp64, this process
  ref_kernel_baseline   the UNMODIFIED reference kernel (oracle/_ref, anyprec.cu for sm_100a) chained over the same
               tensors under a CUDA graph: the GPU kernel to beat
  plugin_path  the same chain through the drop-in boundary: APLinear.forward -> torch.ops.plugin.anyprec_gemv -> C-ABI,
               captured in a CUDA graph like the reference's generate.py runs it
  extra        BASELINE configs 3/4/5: 3-/4-bit Llama-3-8B, Llama-2-70B and Llama-3-70B 2-bit decode (bounded samples)
  cpu_baseline / --impl reference
               the reference's CPU-runnable arm: dequant once -> timed fp16 torch.matmul per Linear
               (primary) and dequant-each-call (APLinear.gemm), all host cores, bounded sample (one block per step)

N > 1 (torchrun): tensor-parallel decode of the same model (strong scaling): wqkv/w1w3 by rows, wo/w2 along K with the
all-reduce fused into the GEMV kernels over NVLink peer memory.
"""
import argparse
import ctypes
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
                    help="N > 1 gemv-chain: all-reduce fused into the GEMV kernels over NVLink peer memory, or NCCL")
    ap.add_argument("--engine", default=None, choices=[None, "persistent", "launches"],
                    help="decode engine: one persistent kernel per token (default where available) or one launch per op")
    ap.add_argument("--no-extra", action="store_true", help="skip the BASELINE config 3/4/5 sweep and the baselines (quick runs)")
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


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_arm(model: str, bits: int, steps: int, warmup: int, lm_head: bool = True):
    """The reference's CPU-runnable arm (BASELINE.md §3, configs[0] scaled to the model): the 4 Linears of ONE block per
    step (a bounded 1/L sample of a token) on all host cores, two variants:
      once  dequant once (untimed, like a loaded dense model), then time fp16 torch.matmul(x, W.T) per Linear   [primary]
      each  dequant (threaded C unpack + gather) + matmul every call, what APLinear.gemm does (APLinear.py:35-38)
    With `lm_head` the fp16 output projection (model.py:94,129) is timed once and added per token.  Attention / norms /
    sampling at batch 1 are < 1 % of the CPU token and are left out (which only flatters the CPU arm)."""
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
    steps = max(1, min(steps, 5))

    # variant "each": dequant every call
    def block_each():
        return sum(O.cpu_reference_linear(q, lut, x, bits, threads=cores, dequant_each_call=True)[1] for q, lut, x in lins)

    for _ in range(max(0, min(warmup, 1))):
        block_each()
    t_each = sum(block_each() for _ in range(max(1, steps // 2))) / max(1, steps // 2)

    # variant "once": W resident, only the matmul is timed — fp16 (what the reference's half model does) and fp32 (CPUs have
    # no fast fp16 GEMV path); the faster one is the primary CPU number
    t_once16 = t_once32 = 0.0
    for q, lut, x in lins:
        t_once16 += O.cpu_reference_linear(q, lut, x, bits, threads=cores, dequant_each_call=False, repeats=max(3, steps))[1]
        t_once32 += O.cpu_reference_linear(q, lut, x, bits, threads=cores, dequant_each_call=False, repeats=max(3, steps),
                                           matmul_dtype="f32")[1]
    t_once = min(t_once16, t_once32)
    t_head = 0.0
    if lm_head:
        Wh = torch.empty((cfg["vocab"], cfg["dim"]), dtype=torch.float16).normal_(0, 0.02)
        xh = torch.randn((1, cfg["dim"])).half()
        torch.matmul(xh, Wh.T)
        t1 = time.perf_counter()
        torch.matmul(xh, Wh.T)
        t_head = time.perf_counter() - t1
    L = cfg["n_layer"]
    return {"once": 1.0 / (t_once * L + t_head), "each": 1.0 / (t_each * L + t_head), "t_block_once": t_once,
            "t_block_once_f16": t_once16, "t_block_once_f32": t_once32,
            "t_block_each": t_each, "t_head": t_head, "cores": cores, "steps": steps}


def cpu_baseline_object(model, bits, r, whole):
    sample = (f"1 block (wqkv, wo, w1w3, w2) of {model} {bits}-bit per step on {r['cores']} host threads, scaled by n_layer"
              + (f" + fp16 lm_head {r['t_head']:.3f} s/token" if whole else "")
              + f"; primary = dequant once + timed torch.matmul, faster of fp16 {r['t_block_once_f16'] * 1e3:.1f} / fp32 {r['t_block_once_f32'] * 1e3:.1f} ms/block; "
              f"dequant-each-call variant (APLinear.gemm) {r['t_block_each'] * 1e3:.1f} ms/block")
    return {"value": r["once"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample,
            "variants": {"dequant_once_matmul": r["once"], "dequant_each_call_matmul": r["each"]}}


# ------------------------------------------------------------------------------------------------ GPU-side baselines
_REF_SO = os.path.join(ROOT, "oracle", "_ref", "libapgemv_ref.so")
_ref_lib = None


def ref_lib():
    """the UNMODIFIED reference kernels compiled for sm_100a (oracle/build_ref.sh); baseline / checker only"""
    global _ref_lib
    if _ref_lib is None and os.path.exists(_REF_SO):
        L = ctypes.CDLL(_REF_SO)
        vp, u32, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int
        L.ref_anyprec_gemv.restype = i32
        L.ref_anyprec_gemv.argtypes = [vp, vp, vp, vp, u32, u32, u32, i32, vp]
        _ref_lib = L
    return _ref_lib


def graph_time(torch, fn, stream, steps, warm):
    """capture fn() (launches on the current stream) into a CUDA graph on `stream`, replay, seconds per replay"""
    with torch.cuda.stream(stream):
        fn()
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            fn()
        for _ in range(warm):
            g.replay()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            g.replay()
        e1.record(stream)
        stream.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / steps


def prefill_probe(torch, model, bits):
    """the prompt-prefill branch of a Linear, outside the timed region: fused dequant + tcgen05 GEMM kernel
    (apg_prefill_gemm) vs the reference's route (anyprec_dequant + library fp16 matmul, APLinear.py:35-38) on the model's
    w1w3 shape, each captured in a CUDA graph over rotating weight copies (> L2), plus the measured error vs fp64"""
    from guidedquant_b200 import ap_gemv
    from guidedquant_b200.runtime import MODEL_CONFIGS, linear_shapes

    N, K = linear_shapes(MODEL_CONFIGS[model])["w1w3"]
    g = torch.Generator(device="cuda").manual_seed(7)
    ncopy = 4
    qs = [torch.randint(-2**31, 2**31 - 1, (bits, N, K // 32), dtype=torch.int32, device="cuda", generator=g) for _ in range(ncopy)]
    lut = (torch.randn((N, 1 << bits), device="cuda", generator=g) * 0.02).half()
    out = {"linear": f"w1w3 {N}x{K}", "bits": bits, "what": "us per Linear call, CUDA graph of 4 calls over distinct weight copies",
           "kernel": "apg::ptc::prefill_tc_kernel (tcgen05.mma, A operand dequantised into TMEM, token tiles by TMA)", "tokens": {}}
    st = torch.cuda.Stream()
    for T in (64, 512):
        x = torch.randn((T, K), device="cuda", generator=g).half()

        def fused():
            for q in qs:
                ap_gemv.anyprec_prefill_gemm(x, q, lut, bits)

        def ref():
            for q in qs:
                torch.matmul(x, ap_gemv.anyprec_dequant(q, lut, bits).T)

        W = ap_gemv.anyprec_dequant(qs[0], lut, bits)
        y = ap_gemv.anyprec_prefill_gemm(x, qs[0], lut, bits)
        truth = x.double() @ W.double().T
        err = float((y.double() - truth).abs().max() / truth.abs().max())
        del W, truth
        t_f = graph_time(torch, fused, st, 10, 2) / ncopy
        t_r = graph_time(torch, ref, st, 10, 2) / ncopy
        out["tokens"][str(T)] = {"fused_us": t_f * 1e6, "dequant_matmul_us": t_r * 1e6, "speedup": t_r / t_f,
                                 "fused_TFLOPs": 2.0 * T * N * K / t_f / 1e12, "err_vs_f64": err}
    return out


def parity_probe(torch, model, bits):
    """measured max error of the default GEMV path at the model's four Linear shapes vs the reference kernel and fp64"""
    from guidedquant_b200 import ap_gemv
    from guidedquant_b200.runtime import MODEL_CONFIGS, linear_shapes

    R = ref_lib()
    out = {"tolerance_vs_f64": 1.2e-3, "tolerance_vs_ref": 2.5e-3, "shapes": [], "max_err_vs_ref": None, "max_err_vs_f64": 0.0,
           "ref_err_vs_f64": None, "metric": "max|y - y_ref| / max|y_ref|"}
    g = torch.Generator(device="cuda").manual_seed(77)
    for name, (N, K) in linear_shapes(MODEL_CONFIGS[model]).items():
        q = torch.randint(-2**31, 2**31 - 1, (bits, N, K // 32), dtype=torch.int32, device="cuda", generator=g)
        lut = (torch.randn((N, 1 << bits), device="cuda", generator=g) * 0.02).half()
        x = torch.randn((1, 1, K), device="cuda", generator=g).half()
        y = torch.empty((1, 1, N), dtype=torch.float16, device="cuda")
        ap_gemv.anyprec_gemv(x, y, q, lut, bits)
        W = ap_gemv.anyprec_dequant(q, lut, bits)
        y64 = torch.empty(N, dtype=torch.float64, device="cuda")
        step = max(1, (1 << 26) // K)
        for r0 in range(0, N, step):
            y64[r0:r0 + step] = (W[r0:r0 + step].double() @ x.double().reshape(K, 1)).reshape(-1)
        e64 = float((y.double().reshape(-1) - y64).abs().max() / y64.abs().max())
        out["max_err_vs_f64"] = max(out["max_err_vs_f64"], e64)
        rec = {"linear": name, "N": N, "K": K, "err_vs_f64": e64}
        if R is not None and N % 4 == 0:
            yr = torch.zeros((1, 1, N), dtype=torch.float16, device="cuda")
            R.ref_anyprec_gemv(x.data_ptr(), yr.data_ptr(), q.data_ptr(), lut.data_ptr(), 1, N, K, bits,
                               torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            er = float((y.double() - yr.double()).abs().max() / yr.double().abs().max())
            rr = float((yr.double().reshape(-1) - y64).abs().max() / y64.abs().max())
            rec["err_vs_ref"], rec["ref_err_vs_f64"] = er, rr
            out["max_err_vs_ref"] = max(out["max_err_vs_ref"] or 0.0, er)
            out["ref_err_vs_f64"] = max(out["ref_err_vs_f64"] or 0.0, rr)
        out["shapes"].append(rec)
        del q, lut, W
    return out


def ref_kernel_chain(torch, chain, steps, warm):
    """the reference kernel chained over the chain's own tensors, same data flow, one CUDA graph (kernel only: the
    reference's APLinear.forward additionally zeroes the output first, APLinear.py:53)"""
    R = ref_lib()
    if R is None:
        return {"unavailable": "oracle/_ref/libapgemv_ref.so not built"}
    bits = chain.bits

    def token():
        st = torch.cuda.current_stream().cuda_stream
        b = chain.buf
        x = chain.x_in
        for li, (wqkv, wo, w1w3, w2) in enumerate(chain.layers):
            h_out = b["h"] if li % 2 == 0 else b["h2"]
            for lin, xi, yo in ((wqkv, x, b["qkv"]), (wo, b["qkv"], b["o"]), (w1w3, b["o"], b["gu"]), (w2, b["gu"], h_out)):
                R.ref_anyprec_gemv(xi.data_ptr(), yo.data_ptr(), lin.qweight.data_ptr(), lin.lut.data_ptr(), 1, lin.N, lin.K,
                                   bits, st)
            x = h_out

    t = graph_time(torch, token, chain.stream, steps, warm)
    ab = chain.algo_bytes_per_step()
    return {"tok_s_equiv": 1.0 / t, "us_per_token": t * 1e6, "GBs": ab / t / 1e9, "launches": 4 * len(chain.layers),
            "what": "unmodified reference anyprec_matmul (anyprec.cu:587-620, compiled for sm_100a) over the same packed tensors, CUDA graph"}


def plugin_chain(torch, chain, steps, warm):
    """the drop-in boundary: 4*L APLinear.forward calls -> torch.ops.plugin.anyprec_gemv -> C-ABI, CUDA-graph captured
    (how the reference's generate.py runs its APLinear modules, generate.py:330-336)"""
    from guidedquant_b200.APLinear import APLinear

    mods = []
    for lins in chain.layers:
        row = []
        for lin in lins:
            m = APLinear(lin.K, lin.N, chain.bits, device="meta")
            m.qweight, m.lut = lin.qweight, lin.lut          # alias the chain's resident tensors (no copy)
            m.output = torch.zeros((1, 1, lin.N), dtype=torch.float16, device=chain.device)
            row.append(m)
        mods.append(row)

    def view(t, n):
        return t.reshape(-1)[:n].reshape(1, 1, n)

    def token():
        x = chain.x_in
        for wqkv, wo, w1w3, w2 in mods:
            qkv = wqkv(x)
            o = wo(view(qkv, wo.in_features))
            gu = w1w3(o)
            x = w2(view(gu, w2.in_features))
        return x

    t = graph_time(torch, token, chain.stream, steps, warm)
    ab = chain.algo_bytes_per_step()
    return {"tok_s_equiv": 1.0 / t, "us_per_token": t * 1e6, "GBs": ab / t / 1e9, "launches": 4 * len(chain.layers),
            "what": "APLinear.forward -> torch.ops.plugin.anyprec_gemv -> apg_gemv (C-ABI) per Linear, CUDA graph of 4*L ops"}


def main():
    a = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    model = a.model or "llama3-8b"  # same workload at every N so the scaling series is comparable
    full_decode = a.workload == "decode"
    wl_decode = f"{model} {a.bits}-bit bs=1 decode, full token step: embed + L x (wqkv|attn|wo|w1w3|w2) + lm_head + greedy sample"
    wl_chain = f"{model} {a.bits}-bit bs=1 decode, ap_gemv hot path only ({'4*L' if a.layers is None else a.layers} APLinear GEMVs/token chain)"
    workload = wl_decode if full_decode else wl_chain

    if a.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_arm(model, a.bits, a.steps, a.warmup, lm_head=full_decode)
        cb = cpu_baseline_object(model, a.bits, r, full_decode)
        tok_s = cb["value"]
        line = {
            "impl": "reference", "metric": METRIC, "value": tok_s, "unit": UNIT, "n_gpus": max(world, a.gpus), "steps": r["steps"],
            "warmup": 1, "ms_per_step": r["t_block_once"] * 1e3, "higher_is_better": True,
            "scaling": "strong" if max(world, a.gpus) > 1 else "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": workload, "reference_arm": "CPU port of the reference's dequant -> fp16 torch.matmul path (oracle/oracle.py); the GPU kernel baseline is ref_kernel_baseline in the main line"},
            "cpu_baseline": cb,
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
        # drain the device BEFORE the NCCL barrier: a persistent token kernel occupies every SM and waits for its peers'
        # packets, so an NCCL kernel slipped between two queued token launches on one rank would dead-lock the group
        torch.cuda.synchronize()
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
    peak, peak_src = measured_peak_gbs()
    eng_kw = {} if a.engine is None else {"engine": a.engine}

    def run_chain(mdl, bits, steps, extras=False):
        """the hot path alone (4*L GEMVs of a token, same engine as the decode step): seconds per token + accounting"""
        ch = ApGemvChain(mdl, bits=bits, n_layer=a.layers, pdl=not a.no_pdl, world_size=world, rank=rank, process_group=pg,
                         ctas_per_sm=a.ctas, collective=a.collective, **eng_kw)
        ch.capture()
        d = ch.cfg["dim"]
        xh = torch.randn((1, 1, d)).half().pin_memory()
        ch.x_in.copy_(xh)
        dt = timed(ch.step, ch.stream, steps, warm) / steps
        res = {"t": dt, "algo_bytes": ch.algo_bytes_per_step(), "n_gemv": 4 * ch.cfg["n_layer"], "wbytes": ch.weight_bytes(),
               "launches": ch.launches_per_step, "engine": getattr(ch, "engine", "launches"), "d": d}
        if not full_decode and mdl == model and bits == a.bits:
            res["t_e2e"] = timed(lambda: ch.step_host(xh), ch.stream, steps, 3) / steps
        if extras and world == 1:
            try:
                res["ref_kernel_baseline"] = ref_kernel_chain(torch, ch, max(5, steps // 4), 3)
            except Exception as e:
                res["ref_kernel_baseline"] = {"failed": repr(e)}
            try:
                res["plugin_path"] = plugin_chain(torch, ch, max(5, steps // 4), 3)
            except Exception as e:
                res["plugin_path"] = {"failed": repr(e)}
        ch.graph = None
        del ch
        torch.cuda.empty_cache()
        return res

    def run_decode(mdl, bits, steps, max_seq, e2e=False):
        tf = APTransformer(mdl, bits=bits, max_seq_len=max_seq, pdl=not a.no_pdl, n_layer=a.layers, world_size=world, rank=rank,
                           process_group=pg, **eng_kw).random_init()
        tf.capture()
        n_tok = min(steps, max_seq - 2)
        warm_d = min(warm, max(1, max_seq - 2 - n_tok))

        def run_tokens(step_fn, n, w):
            # every timed step is a NEW token at the next position (BOS-only prompt protocol, generate.py:310-313)
            tf.reset(1)
            return timed(step_fn, tf.stream, n, w)

        res = {"t": run_tokens(tf.step, n_tok, warm_d) / n_tok, "n_tok": n_tok, "launches": tf.launches_per_token,
               "bytes_tok": tf.algo_bytes_per_token(pos=n_tok // 2), "engine": getattr(tf, "engine", "launches")}
        if e2e:
            tok_pinned = torch.ones(1, dtype=torch.int32).pin_memory()
            res["t_e2e"] = run_tokens(lambda: tf.step_host(tok_pinned), n_tok, min(3, warm_d)) / n_tok
        tf.graph = None
        del tf
        torch.cuda.empty_cache()
        return res

    sampler = ClockSampler(local_rank)
    sampler.start()
    want_extra = not a.no_extra and a.layers is None
    ch = run_chain(model, a.bits, a.steps, extras=want_extra)
    dec = run_decode(model, a.bits, a.steps, a.max_seq, e2e=True) if full_decode else None
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    # ---------------- BASELINE configs 3/4/5 (bounded: 20 tokens each) — every rank takes part under tensor parallelism
    extra = {}
    if want_extra and full_decode and model == "llama3-8b" and a.bits == 2:
        sweep = [("llama3-8b", 3), ("llama3-8b", 4), ("llama2-70b", 2), ("llama3-70b", 2)] if world == 1 else [("llama3-70b", 2)]
        for mdl, bits in sweep:
            key = f"{mdl}_{bits}bit"
            try:
                c2 = run_chain(mdl, bits, 10)
                d2 = run_decode(mdl, bits, 20, 64)
                extra[key] = {"tok_s": 1.0 / d2["t"], "ms_per_token": d2["t"] * 1e3, "chain_us": c2["t"] * 1e6,
                              "chain_GBs": c2["algo_bytes"] / c2["t"] / 1e9, "frac": c2["algo_bytes"] / c2["t"] / 1e9 / peak,
                              "frac_all_bytes": d2["bytes_tok"]["total"] / world / d2["t"] / 1e9 / peak, "tokens_timed": d2["n_tok"],
                              "n_gpus": world}
            except Exception as e:  # the headline number must not be lost to a sweep failure
                extra[key] = {"failed": repr(e)[:300]}
    parity = None
    if want_extra and world == 1:
        try:
            parity = parity_probe(torch, model, a.bits)
        except Exception as e:
            parity = {"failed": repr(e)[:300]}

    def finish():
        # a live CUDA graph that captured NCCL kernels makes communicator teardown hang: sync and leave without tearing
        # the communicator down
        if world > 1:
            torch.cuda.synchronize()
            sys.stdout.flush()
            os._exit(0)

    if rank != 0:
        finish()
        return

    if full_decode:
        t_step, t_e2e, steps, launches, h2d, d2h = dec["t"], dec["t_e2e"], dec["n_tok"], dec["launches"], 4, 4
    else:
        t_step, t_e2e, steps, launches, h2d, d2h = ch["t"], ch["t_e2e"], a.steps, ch["launches"], ch["d"] * 2, ch["d"] * 2
    achieved = ch["algo_bytes"] / ch["t"] / 1e9          # GB/s per GPU of the hot path (launch / sync gaps included)
    traffic = None                                        # measured DRAM bytes per GEMV (ncu --set full), when recorded for this config
    traffic_note = "no ncu capture recorded for this configuration"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "gemv_traffic.json")))
        if world == 1 and tj.get("model") == model and tj.get("bits") == a.bits and a.layers is None and tj.get("engine") == ch["engine"]:
            traffic = tj["dram_bytes_per_gemv"]
            traffic_note = tj.get("note", "")
    except Exception:
        pass
    engine = (dec or ch)["engine"]
    line = {
        "metric": METRIC, "value": 1.0 / t_step, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {
            "workload": workload, "bits": a.bits, "gemv_per_token": ch["n_gemv"], "engine": engine,
            "parallelism": "single GPU" if world == 1 else f"tp{world}: wqkv/w1w3 N-sharded (heads / MLP columns), wo/w2 K-sharded + one-shot all-reduce fused into the GEMV kernels (NVLink peer stores); lm_head vocab-sharded + arg-max exchange over peer memory",
            "l2_policy": "inputs larger than L2: every GEMV reads its own distinct weights (%.2f GB/token/GPU), streamed evict-first" % (ch["wbytes"] / 1e9),
            "pdl": not a.no_pdl, "accumulate": "fp16 chains of 8 -> fp32",
        },
        "clocks": sampler.result(),
        "e2e": {"value": 1.0 / t_e2e if t_e2e else None, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches * steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                     "kernel": "apg::decode_persistent_kernel (GEMV jobs)" if engine == "persistent" else "apg::gemv_fast_kernel",
                     "algorithmic_bytes_per_step": ch["algo_bytes"], "gemv_per_step": ch["n_gemv"],
                     "algorithmic_bytes_per_gemv_avg": ch["algo_bytes"] / ch["n_gemv"],
                     "timed_as": "the %d GEMVs of a token run back to back by the %s engine in this process (CUDA events): %.1f us/token"
                                 % (ch["n_gemv"], engine, ch["t"] * 1e6),
                     "frac_of_8TBs": achieved / 8000.0},
    }
    if full_decode:
        line["config"]["max_seq_len"] = a.max_seq
        line["config"]["launches_per_token"] = launches
        line["config"]["sampling"] = "greedy (temperature 0), token and position advanced on the device"
        bt = dec["bytes_tok"]
        line["decode_bytes"] = {"per_token": bt, "achieved_GBs_all_bytes": bt["total"] / t_step / 1e9,
                                "frac_of_peak_all_bytes": bt["total"] / t_step / 1e9 / peak}
    if parity is not None:
        line["parity"] = parity
    if want_extra and world == 1:
        try:
            line["prefill_tc"] = prefill_probe(torch, model, a.bits)
        except Exception as e:
            line["prefill_tc"] = {"failed": repr(e)[:300]}
    for k in ("ref_kernel_baseline", "plugin_path"):
        if k in ch:
            line[k] = ch[k]
    if "ref_kernel_baseline" in ch and "tok_s_equiv" in ch["ref_kernel_baseline"]:
        line["ref_kernel_baseline"]["ours_chain_tok_s_equiv"] = 1.0 / ch["t"]
        line["ref_kernel_baseline"]["speedup_of_chain"] = ch["ref_kernel_baseline"]["us_per_token"] / (ch["t"] * 1e6)
    if extra:
        line["extra"] = extra
    # CPU baseline (rank 0, N = 1 only): bounded sample
    if world == 1:
        try:
            r = cpu_reference_arm(model, a.bits, 2, 1, lm_head=full_decode)
            line["cpu_baseline"] = cpu_baseline_object(model, a.bits, r, full_decode)
        except Exception as e:  # the GPU number must not be lost to a host-side failure
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    main()
