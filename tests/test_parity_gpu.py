"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C-ABI library
(guidedquant_b200.ap_gemv -> libapgemv_b200.so); the CPU oracle (oracle/) and the compiled reference
kernels (oracle/_ref) are the checkers.

Bars:
  * index unpack: dequant output bit-identical to the oracle AND to the reference kernel;
  * APG_FLAG_REF_ORDER GEMV: bit-identical to the reference kernel and to the oracle's fp16 emulation;
  * fast GEMV (fp32 cross-chain accumulation): max|y - y_f64| / max|y_f64| <= 1.2e-3 (fp16 output rounding alone is up to 4.9e-4) and
    max|y - y_ref| / max|y_ref| <= 2.5e-3  (the reference's own all-fp16 accumulation sits 1.0-1.5e-3 from
    the fp64 truth, SURVEY.md §7.3-2, so its noise floor bounds the second figure).
"""
import numpy as np
import pytest
import torch

from tests import refgpu

pytestmark = pytest.mark.gpu

TOL_TRUTH = 1.2e-3
TOL_REF = 2.5e-3


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.cuda() if dtype is None else t.to(dtype).cuda()


def _nerr(y, ref):
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    return float(np.abs(y - ref).max() / max(np.abs(ref).max(), 1e-30))


def _run(x, q, lut, bits, flags=0, ctas=0):
    from guidedquant_b200 import ap_gemv

    out = torch.full((x.shape[0], 1, q.shape[1]), float("nan"), dtype=torch.float16, device="cuda")
    ap_gemv.anyprec_gemv_ex(x, out, q, lut, bits, flags=flags, ctas_per_sm=ctas)
    torch.cuda.synchronize()
    return out


CASES_FAST = [  # (N, K, bits)  M = 1
    (64, 4096, 2), (64, 4096, 3), (64, 4096, 4),
    (100, 4096, 2), (7, 4096, 3), (5, 4096, 4), (1, 4096, 2),          # ragged N (no N % 4 requirement)
    (48, 11008, 2), (48, 11008, 3), (48, 11008, 4),                   # tail chunk eff = 24 (Llama-2-7B w2)
    (32, 13824, 2), (32, 14336, 2), (32, 14336, 3), (32, 14336, 4),   # multi-slab, eff = 16 / half slab
    (32, 8192, 2), (16, 28672, 2), (16, 28672, 4), (24, 3584, 2),     # 70B shapes, 70B w2 shard 3584
    (32, 128, 2), (32, 1152, 3), (32, 2048, 4), (40, 5120, 2),
]


@pytest.mark.parametrize("N,K,bits", CASES_FAST)
def test_fast_gemv_vs_oracle(oracle, N, K, bits):
    from guidedquant_b200._lib import APG_FLAG_GENERIC

    idx, q, lut, x = oracle.synth_layer(N, K, bits, seed=N + K)
    W = oracle.dequant(q, lut, bits)
    y64 = oracle.gemv_f64(W, x)
    yref = oracle.gemv_ref_order_f16(W, x)
    xq, qq, ll = _t(x), _t(q), _t(lut)
    for ctas in (0, 1, 3):
        y = _run(xq, qq, ll, bits, ctas=ctas).cpu().numpy().reshape(1, N)
        assert not np.isnan(y).any()
        assert _nerr(y, y64) <= TOL_TRUTH, (N, K, bits, ctas, _nerr(y, y64))
        assert _nerr(y, yref) <= TOL_REF
    yg = _run(xq, qq, ll, bits, flags=APG_FLAG_GENERIC).cpu().numpy().reshape(1, N)
    assert _nerr(yg, y64) <= TOL_TRUTH


@pytest.mark.parametrize("N,K,bits,M", [(32, 4096, 2, 1), (32, 2048, 3, 2), (48, 11008, 4, 3), (16, 1024, 5, 1),
                                         (16, 2080, 6, 4), (16, 1024, 7, 8), (16, 4096, 8, 1), (32, 96, 2, 5),
                                         (16, 4096, 4, 8)])
def test_ref_order_bit_exact_vs_oracle_and_generic(oracle, N, K, bits, M):
    from guidedquant_b200._lib import APG_FLAG_GENERIC, APG_FLAG_REF_ORDER

    idx, q, lut, x = oracle.synth_layer(N, K, bits, seed=3 * N + K + M, M=M)
    W = oracle.dequant(q, lut, bits)
    y_emul = oracle.gemv_ref_order_f16(W, x)
    y64 = oracle.gemv_f64(W, x)
    xq, qq, ll = _t(x), _t(q), _t(lut)
    y = _run(xq, qq, ll, bits, flags=APG_FLAG_REF_ORDER).cpu().numpy().reshape(M, N)
    assert np.array_equal(y.view(np.uint16), y_emul.view(np.uint16))
    yg = _run(xq, qq, ll, bits, flags=APG_FLAG_GENERIC).cpu().numpy().reshape(M, N)
    assert _nerr(yg, y64) <= TOL_TRUTH
    y0 = _run(xq, qq, ll, bits).cpu().numpy().reshape(M, N)  # default dispatch (fast or generic)
    assert _nerr(y0, y64) <= TOL_TRUTH


@pytest.mark.parametrize("bits", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("N,K,M", [(33, 4096, 1), (7, 96, 2), (16, 1056, 3), (9, 11008, 4), (24, 2048, 5), (1, 1024, 8),
                                   (40, 4128, 8), (8203, 1056, 3), (8200, 2048, 1)])  # N > 8192: two rows per warp
def test_wide_gemv_vs_oracle(oracle, bits, N, K, M):
    """the 'wide' kernel (bits 5..8, M 2..8, K % 128 != 0; apgemv_wide.cuh): default dispatch vs the fp64 truth, and
    every batch row equal to the same row run alone (rows are independent)."""
    idx, q, lut, x = oracle.synth_layer(N, K, bits, seed=17 * N + K + M + bits, M=M)
    W = oracle.dequant(q, lut, bits)
    y64 = oracle.gemv_f64(W, x)
    xq, qq, ll = _t(x), _t(q), _t(lut)
    y = _run(xq, qq, ll, bits)
    yn = y.cpu().numpy().reshape(M, N)
    assert not np.isnan(yn).any()
    assert _nerr(yn, y64) <= TOL_TRUTH, (bits, N, K, M, _nerr(yn, y64))
    if M > 1 and (bits > 4 or K % 128):  # single rows take the same kernel family: bit-identical
        for m in (0, M - 1):
            y1 = _run(xq[m:m + 1].contiguous(), qq, ll, bits)
            assert torch.equal(y1.view(torch.int16).reshape(-1), y[m].view(torch.int16).reshape(-1)), (bits, m)


@pytest.mark.parametrize("N,K,bits", [(64, 4096, 2), (64, 4096, 3), (64, 4096, 4), (16, 11008, 2), (16, 13824, 3),
                                       (8, 96, 4), (12, 1024, 5), (8, 2048, 8), (7, 1056, 2), (9, 1120, 6), (5, 3072, 7),
                                       (3, 11008, 4), (6, 2080, 3)])
def test_dequant_bit_exact(oracle, N, K, bits):
    from guidedquant_b200 import ap_gemv

    idx, q, lut, _ = oracle.synth_layer(N, K, bits, seed=N)
    # hostile codebook values: +-65504, denormals, -0
    lut = lut.copy()
    lut[0, 0], lut[0, 1], lut[-1, -1], lut[-1, 0] = 65504.0, -65504.0, np.float16(6e-8), np.float16(-0.0)
    W = ap_gemv.anyprec_dequant(_t(q), _t(lut), bits).cpu().numpy()
    Wo = oracle.dequant(q, lut, bits)
    assert W.shape == (N, K)
    assert np.array_equal(W.view(np.uint16), Wo.view(np.uint16))
    assert np.array_equal(W.view(np.uint16), lut.view(np.uint16)[np.arange(N)[:, None], idx])


@pytest.mark.skipif(not refgpu.available(), reason="oracle/_ref/libapgemv_ref.so not built")
@pytest.mark.parametrize("N,K,bits,M", [(64, 4096, 2, 1), (64, 4096, 3, 1), (64, 4096, 4, 1), (32, 11008, 2, 1),
                                         (32, 14336, 2, 1), (32, 13824, 3, 1), (32, 28672, 2, 1), (32, 2048, 5, 1),
                                         (32, 4096, 3, 4), (32, 4096, 4, 8), (32, 4096, 2, 2)])
def test_against_compiled_reference_kernels(oracle, N, K, bits, M):
    """The UNMODIFIED reference kernels (anyprec.cu compiled for sm_100a) on the same tensors."""
    from guidedquant_b200 import ap_gemv
    from guidedquant_b200._lib import APG_FLAG_REF_ORDER

    idx, q, lut, x = oracle.synth_layer(N, K, bits, seed=11 * N + K, M=M)
    xq, qq, ll = _t(x), _t(q), _t(lut)
    y_ref = refgpu.ref_gemv(xq, qq, ll, bits)
    w_ref = refgpu.ref_dequant(qq, ll, bits)
    torch.cuda.synchronize()
    # (1) dequant: bit-identical
    w = ap_gemv.anyprec_dequant(qq, ll, bits)
    assert torch.equal(w.view(torch.int16), w_ref.view(torch.int16))
    # (2) reference-order mode: bit-identical to the reference kernel, and the oracle's emulation too
    y_exact = _run(xq, qq, ll, bits, flags=APG_FLAG_REF_ORDER)
    assert torch.equal(y_exact.view(torch.int16), y_ref.view(torch.int16))
    W = oracle.dequant(q, lut, bits)
    y_emul = oracle.gemv_ref_order_f16(W, x)
    assert np.array_equal(y_ref.cpu().numpy().reshape(M, N).view(np.uint16), y_emul.view(np.uint16))
    # (3) default path within the stated tolerance of the reference
    y = _run(xq, qq, ll, bits)
    assert _nerr(y.cpu().numpy(), y_ref.cpu().numpy()) <= TOL_REF


def test_full_size_properties(oracle):
    """BASELINE-size layer (4096x4096, 2-bit): size-independent properties instead of the slow oracle.
    linearity in x, permutation of output rows, all-equal codebook -> y = c * sum(x)."""
    from guidedquant_b200 import ap_gemv

    N = K = 4096
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randint(-2**31, 2**31 - 1, (2, N, K // 32), dtype=torch.int32, device="cuda", generator=g)
    lut = (torch.randn((N, 4), device="cuda", generator=g) * 0.02).half()
    x1 = torch.randn((1, 1, K), device="cuda", generator=g).half()
    y1 = _run(x1, q, lut, 2).float()
    # linearity: y(2x) == 2 y(x) up to fp16-denormal effects inside the short fp16 chains
    y2 = _run((x1 * 2).half(), q, lut, 2).float()
    assert float((y2 - 2 * y1).abs().max() / (2 * y1).abs().max()) <= 1e-3
    # dequant -> fp64 matmul (torch) agrees
    W = ap_gemv.anyprec_dequant(q, lut, 2).double()
    yt = (W @ x1.double().reshape(K, 1)).reshape(1, 1, N)
    assert float((y1.double() - yt).abs().max() / yt.abs().max()) <= TOL_TRUTH
    # row permutation of (qweight, lut) permutes y
    perm = torch.randperm(N, device="cuda", generator=g)
    yp = _run(x1, q[:, perm].contiguous(), lut[perm].contiguous(), 2).float()
    assert torch.equal(yp, y1[..., perm])
    # constant codebook: every weight equals c_n -> y_n = c_n * sum(x)
    lut_c = lut[:, :1].repeat(1, 4).contiguous()
    yc = _run(x1, q, lut_c, 2).float().reshape(N)
    expect = lut_c[:, 0].float() * x1.float().sum()
    # (rounding errors of the fp16 chains are fully correlated across k in this degenerate case)
    assert float((yc - expect).abs().max() / expect.abs().max()) <= 5e-3


def test_error_behaviour():
    from guidedquant_b200 import ap_gemv

    x = torch.zeros((1, 1, 128), dtype=torch.float16, device="cuda")
    q = torch.zeros((2, 8, 4), dtype=torch.int32, device="cuda")
    lut = torch.zeros((8, 4), dtype=torch.float16, device="cuda")
    out = torch.zeros((1, 1, 8), dtype=torch.float16, device="cuda")
    ap_gemv.anyprec_gemv(x, out, q, lut, 2)
    with pytest.raises(RuntimeError, match="Bitwidth"):
        ap_gemv.anyprec_gemv(x, out, q, lut, 9)
    with pytest.raises(RuntimeError, match="lut tensor must be of shape"):
        ap_gemv.anyprec_gemv(x, out, q, lut[:, :2].contiguous(), 2)
    with pytest.raises(RuntimeError, match="qweight tensor must be of shape"):
        ap_gemv.anyprec_gemv(x, out, q[:, :4].contiguous(), lut, 2)
    with pytest.raises(RuntimeError, match="sequence length"):
        ap_gemv.anyprec_gemv(x.reshape(1, 1, 128).expand(1, 2, 128).contiguous(), out, q, lut, 2)
    with pytest.raises(RuntimeError, match="contiguous"):
        ap_gemv.anyprec_gemv(x, out, q.transpose(1, 2).contiguous().transpose(1, 2), lut, 2)
    with pytest.raises(RuntimeError, match="float16"):
        ap_gemv.anyprec_gemv(x.bfloat16(), out.bfloat16(), q, lut.bfloat16(), 2)
    with pytest.raises(RuntimeError, match="type int"):
        ap_gemv.anyprec_gemv(x, out, q.long(), lut, 2)


def test_any_precision_linear_module(oracle):
    """HF-side module surface: one shared 4-plane qweight, lut3 / lut4, set_precision, prefill path == decode path."""
    from guidedquant_b200.AnyPrecisionLinear import AnyPrecisionLinear

    N, K = 96, 2048
    rng = np.random.default_rng(9)
    idx4 = rng.integers(0, 16, size=(N, K), dtype=np.uint8)
    q4 = oracle.pack(idx4, 4)
    lut4 = (rng.standard_normal((N, 16)) * 0.02).astype(np.float16)
    lut3 = (rng.standard_normal((N, 8)) * 0.02).astype(np.float16)
    m = AnyPrecisionLinear(K, N, [3, 4], bias=False, device="cuda", dtype=torch.float16)
    m.qweight.copy_(torch.from_numpy(q4))
    m.lut4.copy_(torch.from_numpy(lut4))
    m.lut3.copy_(torch.from_numpy(lut3))
    x = torch.from_numpy(rng.standard_normal((1, 1, K)).astype(np.float16)).cuda()
    for bits, lut, idx in ((4, lut4, idx4), (3, lut3, idx4 >> 1)):
        m.set_precision(bits)
        y = m(x).float().cpu().numpy().reshape(1, N)
        W = lut[np.arange(N)[:, None], idx]
        y64 = oracle.gemv_f64(W, x.cpu().numpy())
        assert _nerr(y, y64) <= TOL_TRUTH, bits
        yp = m(x.expand(1, 3, K).contiguous())  # 1 < seq <= 8 -> batched LUT GEMV
        assert yp.shape == (1, 3, N) and _nerr(yp[:, 2].float().cpu().numpy(), y64) <= 2e-3
        yl = m(x.expand(1, 11, K).contiguous())  # seq > 8 -> dequant + matmul (the reference's gemm path)
        assert yl.shape == (1, 11, N) and _nerr(yl[:, 10].float().cpu().numpy(), y64) <= 2e-3
    with pytest.raises(RuntimeError):
        m.set_precision(2)


def test_aplinear_module_and_custom_op(oracle):
    from guidedquant_b200.APLinear import APLinear

    N, K, bits = 128, 4096, 2
    idx, q, lut, x = oracle.synth_layer(N, K, bits, seed=77)
    lin = APLinear(K, N, bits)
    lin.qweight.copy_(torch.from_numpy(q))
    lin.lut.copy_(torch.from_numpy(lut))
    xq = torch.from_numpy(x).cuda()
    y = lin(xq)
    assert y.data_ptr() == lin.output.data_ptr()  # aliasing contract of the reference (APLinear.py:33,60)
    y64 = oracle.gemv_f64(oracle.dequant(q, lut, bits), x)
    assert _nerr(y.float().cpu().numpy(), y64) <= TOL_TRUTH
    y2 = torch.zeros_like(lin.output)
    torch.ops.plugin.anyprec_gemv(xq, lin.qweight, lin.lut, y2, bits)
    assert torch.equal(y2, lin.output)
    yp = lin(xq.expand(1, 4, K).contiguous())   # 1 < seq <= 8 -> batched LUT GEMV
    assert yp.shape == (1, 4, N) and _nerr(yp[:, 1].float().cpu().numpy(), y64) <= 2e-3
    yl = lin(xq.expand(1, 12, K).contiguous())  # seq > 8 -> dequant + matmul (APLinear.py:35-38)
    assert yl.shape == (1, 12, N) and _nerr(yl[:, 11].float().cpu().numpy(), y64) <= 2e-3
    # CUDA-graph capture of the op (how generate.py runs it)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        lin(xq)
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            lin(xq)
        lin.output.zero_()
        g.replay()
        s.synchronize()
    assert torch.equal(lin.output, y2)
