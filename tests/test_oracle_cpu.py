"""CPU tests of the oracle (test infrastructure) against the reference-derived golden vectors, and of
the product's host-side layout code against the same vectors."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _cases():
    g = np.load(os.path.join(GOLD, "pack_golden.npz"))
    names = sorted({k.rsplit("_", 1)[0] for k in g.files})
    return g, names


def test_oracle_pack_unpack_match_reference_packer(oracle):
    g, names = _cases()
    assert len(names) >= 30
    for nm in names:
        N, K, bits = map(int, g[nm + "_meta"])
        idx, q = g[nm + "_idx"], g[nm + "_q"]
        assert np.array_equal(oracle.pack(idx, bits), q), nm
        assert np.array_equal(oracle.unpack(q, bits), idx), nm


def test_product_pack_unpack_match_reference_packer():
    from guidedquant_b200 import pack as P

    g, names = _cases()
    for nm in names:
        N, K, bits = map(int, g[nm + "_meta"])
        idx, q = g[nm + "_idx"], g[nm + "_q"]
        assert np.array_equal(P.pack_indices(idx, bits), q), nm
        assert np.array_equal(P.unpack_indices(q, bits), idx), nm


def test_any_precision_prefix_property(oracle):
    """first b planes of a P-bit packing are the b-bit model's indices (idx >> (P-b))."""
    from guidedquant_b200 import pack as P

    rng = np.random.default_rng(3)
    idx = rng.integers(0, 16, size=(6, 2080), dtype=np.uint8)
    q = P.pack_indices(idx, 4)
    for b in (2, 3, 4):
        assert np.array_equal(P.unpack_indices(q, b), idx >> (4 - b))
        assert np.array_equal(oracle.unpack(q, b), idx >> (4 - b))


def test_half_conversion_matches_numpy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.standard_normal(3000) * s for s in (1e-8, 6e-8, 1e-5, 6.1e-5, 1e-3, 1, 100, 3e4, 6.55e4)])
    v = np.concatenate([v, [0.0, -0.0, 65504.0, 65519.99, 65520.0, 2**-24, 2**-25, 1.5 * 2**-24, 2.5 * 2**-24]])
    with np.errstate(over="ignore"):
        exp = v.astype(np.float16).view(np.uint16)
    got = np.array([L.apo_f64_to_f16(float(t)) for t in v], dtype=np.uint16)
    assert np.array_equal(got, exp)
    allh = np.arange(65536, dtype=np.uint16)
    back = np.array([L.apo_f16_to_f64(int(h)) for h in allh])
    ref = allh.view(np.float16).astype(np.float64)
    assert np.array_equal(np.isnan(back), np.isnan(ref))
    assert np.array_equal(back[~np.isnan(ref)], ref[~np.isnan(ref)])


def _py_ref_order(W, x):
    """independent numpy restatement of SURVEY.md Appendix A (fp16 via exact float64 ops + one rounding)."""
    f16 = np.float16
    N, K = W.shape
    y = np.zeros(N, dtype=f16)
    nchunk = (K + 1023) // 1024
    for n in range(N):
        part = np.zeros(32, dtype=f16)
        for t in range(32):
            p = f16(0)
            for i in range(nchunk):
                eff = 32 if i < K // 1024 else (K % 1024) // 32
                if t >= eff:
                    break
                sx = sy = f16(0)
                for c in (3, 2, 1, 0):
                    k0 = i * 1024 + c * 8 * eff + 8 * t
                    for m in range(4):
                        # binary16 products are exact in float64 and the float64 add is a single rounding that a
                        # second rounding to binary16 cannot double-round (oracle/apgemv_oracle.c hfma comment)
                        sx = f16(np.float64(W[n, k0 + 2 * m]) * np.float64(x[k0 + 2 * m]) + np.float64(sx))
                        sy = f16(np.float64(W[n, k0 + 2 * m + 1]) * np.float64(x[k0 + 2 * m + 1]) + np.float64(sy))
                p = f16(np.float64(p) + np.float64(f16(np.float64(sx) + np.float64(sy))))
            part[t] = p
        for off in (16, 8, 4, 2, 1):
            nxt = part.copy()
            for t in range(32):
                nxt[t] = f16(np.float64(part[t]) + np.float64(part[t + off] if t + off < 32 else part[t]))
            part = nxt
        y[n] = part[0]
    return y


@pytest.mark.parametrize("K,bits", [(1024, 2), (2048 + 512, 3), (96, 4)])
def test_ref_order_emulation_vs_independent_numpy(oracle, K, bits):
    idx, q, lut, x = oracle.synth_layer(3, K, bits, seed=K)
    W = oracle.dequant(q, lut, bits)
    assert np.array_equal(W, lut[np.arange(3)[:, None], idx])
    y = oracle.gemv_ref_order_f16(W, x)
    y2 = _py_ref_order(W, x.reshape(-1))
    assert np.array_equal(y.reshape(-1).view(np.uint16), y2.view(np.uint16))
    y64 = oracle.gemv_f64(W, x).reshape(-1)
    assert np.allclose(y64, W.astype(np.float64) @ x.reshape(-1).astype(np.float64), rtol=1e-12, atol=1e-12)
    ym = oracle.gemv_dequant_matmul_f16(W, x).reshape(-1)
    assert np.abs(ym.astype(np.float64) - y64).max() <= 2e-3 * np.abs(y64).max() + 1e-6


def test_oracle_against_reference_gpu_golden(oracle):
    """Outputs of the UNMODIFIED reference kernels recorded on a B200 (tests/golden/ref_gpu_*.npz, written by
    tools/make_ref_gpu_golden.py): the oracle must reproduce them bit for bit."""
    files = sorted(glob.glob(os.path.join(GOLD, "ref_gpu_*.npz")))
    if not files:
        pytest.skip("no reference GPU golden recorded yet")
    for f in files:
        g = np.load(f)
        names = sorted({k.rsplit("_", 1)[0] for k in g.files})
        for nm in names:
            N, K, bits, M = map(int, g[nm + "_meta"])
            q, lut, x = g[nm + "_q"], g[nm + "_lut"], g[nm + "_x"]
            W = oracle.dequant(q, lut, bits)
            if nm + "_wsum" in g.files:  # dequant checked through a checksum + a full first/last row
                assert np.array_equal(W[0].view(np.uint16), g[nm + "_w0"].view(np.uint16)), nm
                assert np.array_equal(W[-1].view(np.uint16), g[nm + "_w1"].view(np.uint16)), nm
                assert int(W.view(np.uint16).astype(np.uint64).sum()) == int(g[nm + "_wsum"]), nm
            y = oracle.gemv_ref_order_f16(W, x)
            assert np.array_equal(y.view(np.uint16), g[nm + "_y"].reshape(M, N).view(np.uint16)), nm
