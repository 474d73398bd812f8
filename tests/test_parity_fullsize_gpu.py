"""Full-size GPU parity at the shapes bench.py runs (BASELINE configs 1/3/4/5): the DEFAULT fast path —
the launch geometry the benchmark uses, including ring-slot re-use — against

  (a) the UNMODIFIED reference kernel (oracle/_ref: anyprec.cu compiled for sm_100a) on the same tensors, tolerance TOL_REF;
  (b) fp64 dequant -> matmul on the GPU, where the dequant is bit-identical to the oracle (test_dequant_bit_exact), TOL_TRUTH.

Stated tolerances: max|y - y64| / max|y64| <= 1.2e-3 and max|y - y_ref| / max|y_ref| <= 2.5e-3 (the reference's own
all-fp16 accumulation sits 1.0-1.5e-3 from the fp64 truth, so its noise floor bounds the second figure).
The measured figures are printed (pytest -s) and repeated in the bench line's `parity` object.
"""
import ctypes

import pytest
import torch

from tests import refgpu

pytestmark = pytest.mark.gpu

TOL_TRUTH = 1.2e-3
TOL_REF = 2.5e-3

SHAPES = [  # (N, K): Llama-3-8B wqkv / wo / w1w3 / w2, 70B wqkv / wo(shard) / w1w3 / w2 / w2 shard at 8 GPUs
    (4096, 4096), (6144, 4096), (28672, 4096), (4096, 14336),
    (10240, 8192), (57344, 8192), (8192, 28672), (8192, 3584),
]


def _plan(N, K, bits, ctas, sms):
    from guidedquant_b200 import _lib

    arr = (ctypes.c_uint32 * 16)()
    st = _lib.lib().apg_plan_fast(N, K, bits, ctas, sms, ctypes.byref(arr))
    assert st == 0
    names = ("cpw", "nwk", "groups", "rs", "nslots", "stage_bytes", "grid", "rows_per_cta", "threads", "unit_rows",
             "units_q", "units_rem", "smem")
    return dict(zip(names, list(arr)))


def _f64_truth(q, lut, x, bits):
    """dequant (bit-exact index unpack, checked elsewhere) -> fp64 matmul on the GPU, row blocks to bound memory."""
    from guidedquant_b200 import ap_gemv

    W = ap_gemv.anyprec_dequant(q, lut, bits)
    xd = x.double().reshape(-1, 1)
    out = torch.empty(W.shape[0], dtype=torch.float64, device=W.device)
    step = max(1, (1 << 27) // W.shape[1])
    for r0 in range(0, W.shape[0], step):
        out[r0:r0 + step] = (W[r0:r0 + step].double() @ xd).reshape(-1)
    return out


def _nerr(y, ref):
    return float((y.double().reshape(-1) - ref.double().reshape(-1)).abs().max() / ref.double().abs().max())


def _synth(N, K, bits, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randint(-2**31, 2**31 - 1, (bits, N, K // 32), dtype=torch.int32, device="cuda", generator=g)
    lut = (torch.randn((N, 1 << bits), device="cuda", generator=g) * 0.02).half()
    x = torch.randn((1, 1, K), device="cuda", generator=g).half()
    return q, lut, x


@pytest.mark.skipif(not refgpu.available(), reason="oracle/_ref/libapgemv_ref.so not built")
@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("N,K", SHAPES)
def test_benchmark_shapes_vs_reference_kernel_and_f64(N, K, bits):
    from guidedquant_b200 import ap_gemv

    q, lut, x = _synth(N, K, bits, seed=N + K + bits)
    y_ref = refgpu.ref_gemv(x, q, lut, bits)
    y64 = _f64_truth(q, lut, x, bits)
    worst = (0.0, 0.0)
    for ctas in (0, 1, 3):
        out = torch.full((1, 1, N), float("nan"), dtype=torch.float16, device="cuda")
        ap_gemv.anyprec_gemv_ex(x, out, q, lut, bits, ctas_per_sm=ctas)
        torch.cuda.synchronize()
        assert not torch.isnan(out).any()
        e64, eref = _nerr(out, y64), _nerr(out, y_ref)
        worst = (max(worst[0], e64), max(worst[1], eref))
        assert e64 <= TOL_TRUTH, (N, K, bits, ctas, e64)
        assert eref <= TOL_REF, (N, K, bits, ctas, eref)
    print(f"\n[parity] {N}x{K} {bits}-bit: max err vs f64 {worst[0]:.2e}, vs reference kernel {worst[1]:.2e}, "
          f"reference vs f64 {_nerr(y_ref, y64):.2e}")


@pytest.mark.parametrize("bits", [2, 3, 4])
def test_ring_slot_reuse_is_exercised(bits):
    """at least one of the shapes above runs more stages per CTA than the ring has slots (the slot re-use path: parity
    flip of the full barrier, producer waiting on the empty barrier) for every bit-width, at the default geometry."""
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    wrapped = []
    for (N, K) in SHAPES:
        for ctas in (0, 1, 3):
            p = _plan(N, K, bits, ctas, sms)
            if p["rows_per_cta"] // p["rs"] > p["nslots"]:
                wrapped.append((N, K, ctas))
    assert wrapped, "no benchmark shape wraps the ring"
    print(f"\n[plan] {bits}-bit shapes with stages_per_cta > nslots: {wrapped}")


def _ref_fused(q, lut, x_in, bits, K, norm_w=None, eps=1e-5, silu_mul=False, residual=None):
    """torch restatement of the fused Linear with the reference's fp16 roundings (inference/model.py:151-167, 261-285):
    x := fp16(silu(g)) * u | fp16(fp16(x * rsqrt(mean(x^2) + eps)) * w);  y := fp16(W x) (+ residual in fp16)."""
    x = x_in.reshape(-1)
    if silu_mul:
        g, u = x[:K].float(), x[K:2 * K]
        x = (g / (1.0 + torch.exp(-g))).half() * u
    if norm_w is not None:
        xf = x.float()
        x = (xf * torch.rsqrt((xf * xf).mean() + eps)).half() * norm_w
    y = _f64_truth(q, lut, x.reshape(1, 1, K), bits).float().half()
    if residual is not None:
        y = y + residual.reshape(-1)
    return y


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("variant", ["norm", "silu_mul", "residual", "norm+residual", "silu_mul+residual"])
def test_fused_variants_full_size(variant, bits):
    """the fused prologues / epilogues of the decode step at a benchmark shape with ring wrap."""
    from guidedquant_b200 import _lib

    silu = "silu_mul" in variant
    N, K = (4096, 14336) if silu else (28672, 4096)
    q, lut, _ = _synth(N, K, bits, seed=5 * bits + len(variant))
    g = torch.Generator(device="cuda").manual_seed(99)
    xin = torch.randn((2 * K if silu else K,), device="cuda", generator=g).half()
    norm_w = (1 + 0.1 * torch.randn(K, device="cuda", generator=g)).half() if "norm" in variant else None
    res = torch.randn(N, device="cuda", generator=g).half() if "residual" in variant else None
    out = torch.full((N,), float("nan"), dtype=torch.float16, device="cuda")
    st = _lib.lib().apg_gemv_fused(xin.data_ptr(), out.data_ptr(), None, q.data_ptr(), lut.data_ptr(), N, K, bits,
                                   norm_w.data_ptr() if norm_w is not None else None, 1e-5, 1 if silu else 0,
                                   res.data_ptr() if res is not None else None, 0,
                                   torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "apg_gemv_fused")
    torch.cuda.synchronize()
    ref = _ref_fused(q, lut, xin, bits, K, norm_w=norm_w, silu_mul=silu, residual=res)
    assert not torch.isnan(out).any()
    # the residual is added after the fp16 rounding of y, so compare on the scale of the Linear's own output
    y_lin = ref.float() - (res.float() if res is not None else 0.0)
    denom = float(y_lin.abs().max())
    err = float((out.float() - ref.float()).abs().max()) / denom
    assert err <= 2.5e-3, (variant, bits, err)
    print(f"\n[parity] fused {variant} {N}x{K} {bits}-bit: max err {err:.2e} (of max|Wx|)")
