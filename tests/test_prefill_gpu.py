"""GPU parity of the fused tensor-core prefill kernel (apg_prefill_gemm, csrc/prefill_tc.cuh) against the path it replaces:
`anyprec_dequant` (bit-exact gather, tested elsewhere) followed by a matmul (inference/ap_gemv/APLinear.py:35-38).

Truth = the dequantised fp16 weights times x in fp64.  Stated tolerance: max|y - y64| / max|y64| <= 1e-3 (fp32 tensor-core
accumulation + ONE fp16 rounding of the output: 2^-11 = 4.9e-4 relative to the element, less relative to the max).
Covered: bits 2/3/4, token counts 9..1100 (partial token tiles, several token tiles), N not a multiple of the 128-row tile,
K with a partial last 1024-chunk (11008), split-K (small T*N) with and without a workspace, ring wrap (K/64 >> stages).
"""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-3


def _synth(N, K, bits, T, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randint(-2**31, 2**31 - 1, (bits, N, K // 32), dtype=torch.int32, device="cuda", generator=g)
    lut = (torch.randn((N, 1 << bits), device="cuda", generator=g) * 0.02).half()
    lut, _ = torch.sort(lut, dim=1)
    x = torch.randn((T, K), device="cuda", generator=g).half()
    return q, lut.contiguous(), x


def _truth(q, lut, x, bits):
    from guidedquant_b200 import ap_gemv

    W = ap_gemv.anyprec_dequant(q, lut, bits)
    return x.double() @ W.double().T


def _nerr(y, ref):
    return float((y.double() - ref).abs().max() / ref.abs().max())


CASES = [  # (N, K, T)
    (128, 256, 16), (256, 1024, 9), (200, 512, 33), (512, 4096, 64), (4096, 4096, 200), (1024, 4096, 256),
    (384, 11008, 77), (256, 14336, 300), (6144, 4096, 1100), (1000, 2048, 130),
    (28672, 4096, 520), (4096, 14336, 2048),   # the benchmarked w1w3 / w2 shapes: 12-waves grids, three token tiles / eight
]


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("N,K,T", CASES)
def test_prefill_gemm_vs_f64(bits, N, K, T):
    from guidedquant_b200 import ap_gemv

    q, lut, x = _synth(N, K, bits, T, seed=N + K + T + bits)
    y = ap_gemv.anyprec_prefill_gemm(x, q, lut, bits)
    torch.cuda.synchronize()
    assert y.shape == (T, N) and y.dtype == torch.float16
    ref = _truth(q, lut, x, bits)
    err = _nerr(y, ref)
    lib_err = _nerr(x @ ap_gemv.anyprec_dequant(q, lut, bits).T, ref)
    print(f"prefill bits={bits} N={N} K={K} T={T}: err {err:.2e} (dequant + library matmul {lib_err:.2e})")
    assert err <= TOL, (bits, N, K, T, err)


def test_prefill_without_workspace_walks_all_of_k():
    """split-K needs scratch; without it the same call must still give the right answer (one CTA per tile, whole K)"""
    from guidedquant_b200 import _lib

    L = _lib.lib()
    N, K, T, bits = 256, 4096, 40, 3
    q, lut, x = _synth(N, K, bits, T, seed=5)
    plan = (ctypes.c_uint32 * 8)()
    need = ctypes.c_uint64(0)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    assert L.apg_prefill_plan(T, N, K, bits, sms, plan, ctypes.byref(need)) == 0
    assert plan[3] > 1 and need.value == plan[3] * T * N * 4
    out = torch.zeros((T, N), dtype=torch.float16, device="cuda")
    st = L.apg_prefill_gemm(x.data_ptr(), out.data_ptr(), q.data_ptr(), lut.data_ptr(), T, N, K, bits, None, 0,
                            torch.cuda.current_stream().cuda_stream)
    assert st == 0
    torch.cuda.synchronize()
    assert _nerr(out, _truth(q, lut, x, bits)) <= TOL


def test_prefill_is_deterministic_and_matches_rowwise_gemv():
    from guidedquant_b200 import ap_gemv

    N, K, T, bits = 1024, 4096, 48, 2
    q, lut, x = _synth(N, K, bits, T, seed=9)
    y1 = ap_gemv.anyprec_prefill_gemm(x, q, lut, bits)
    y2 = ap_gemv.anyprec_prefill_gemm(x, q, lut, bits)
    assert torch.equal(y1, y2)
    # every token row against the decode GEMV of the same Linear
    out = torch.zeros((1, 1, N), dtype=torch.float16, device="cuda")
    for t in (0, 17, T - 1):
        ap_gemv.anyprec_gemv(x[t].reshape(1, 1, K), out, q, lut, bits)
        e = float((out.reshape(-1).double() - y1[t].double()).abs().max() / y1[t].double().abs().max())
        assert e <= 2.5e-3, (t, e)


def test_prefill_unsupported_shapes_are_refused_not_miscomputed():
    from guidedquant_b200 import _lib, ap_gemv

    q = torch.zeros((5, 128, 8), dtype=torch.int32, device="cuda")
    lut = torch.zeros((128, 32), dtype=torch.float16, device="cuda")
    x = torch.zeros((16, 256), dtype=torch.float16, device="cuda")
    assert not ap_gemv.prefill_supported(q, 5)
    with pytest.raises(RuntimeError):
        ap_gemv.anyprec_prefill_gemm(x, q, lut, 5)
    L = _lib.lib()
    out = torch.zeros((16, 128), dtype=torch.float16, device="cuda")
    st = L.apg_prefill_gemm(x.data_ptr(), out.data_ptr(), q.data_ptr(), lut.data_ptr(), 16, 128, 256, 5, None, 0, None)
    assert st == 8  # APG_ERR_UNSUPPORTED
    q2 = torch.zeros((2, 128, 4), dtype=torch.int32, device="cuda")  # K = 128: not a multiple of 256
    assert not ap_gemv.prefill_supported(q2, 2)


def test_aplinear_prefill_uses_the_fused_kernel_and_matches_dequant_matmul():
    from guidedquant_b200 import ap_gemv
    from guidedquant_b200.APLinear import APLinear

    N, K, T, bits = 512, 1024, 24, 4
    q, lut, x = _synth(N, K, bits, T, seed=21)
    assert ap_gemv.prefill_prefers_fused(q, bits, T)
    lin = APLinear(K, N, bits, device="cuda")
    lin.qweight.copy_(q)
    lin.lut.copy_(lut)
    y = lin(x.reshape(1, T, K))
    ref = (x.double() @ ap_gemv.anyprec_dequant(q, lut, bits).double().T).reshape(1, T, N)
    assert y.shape == (1, T, N)
    assert _nerr(y, ref) <= TOL
