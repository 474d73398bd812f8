"""CPU mirror of the fast kernel's integer path (guidedquant_b200/csrc/apgemv_fast.cuh: WordDot<2|3|4>): the same
bit-select / shift / mask / byte-extract steps in numpy, checked against the oracle's unpack of the same packed words.
Guards the (plane words) -> (table offset, x-pair slot) mapping without a GPU:
    word bit 31-o  <->  k offset o inside the lane's 32 weights,  o = 8c + e,  c = 3 - b (byte b of an offset word)."""
import numpy as np
import pytest

U = np.uint32


def bitsel(a, b, m):
    return (a & U(m)) | (b & ~U(m))


def byte(w, b):
    return (w >> U(8 * b)) & U(0xFF)


def lane_indices(oracle, planes):
    """ground truth: 2^bits index of each of the 32 weights (k offset o = 0..31) of one word per plane."""
    bits = len(planes)
    q = np.array(planes, dtype=np.uint32).reshape(bits, 1, 1)
    q = np.concatenate([q, np.zeros((bits, 1, 31), dtype=np.uint32)], axis=2)  # one 1024-chunk, lane t = 0
    idx = oracle.unpack(q.view(np.int32), bits)[0]          # [1024]; lane 0 owns k = c*256 + e
    return np.array([idx[(o // 8) * 256 + (o % 8)] for o in range(32)])


@pytest.mark.parametrize("seed", range(4))
def test_worddot2_mapping(oracle, seed):
    rng = np.random.default_rng(seed)
    H, L = U(rng.integers(0, 2**32)), U(rng.integers(0, 2**32))
    truth = lane_indices(oracle, [H, L])
    zh = bitsel(H, L >> U(2), 0xCCCCCCCC)
    zl = bitsel(H << U(2), L, 0xCCCCCCCC)
    a0, a1 = (zh << U(2)) & U(0x3C3C3C3C), (zh >> U(2)) & U(0x3C3C3C3C)
    a2, a3 = (zl << U(2)) & U(0x3C3C3C3C), (zl >> U(2)) & U(0x3C3C3C3C)
    for b in range(4):
        c = 3 - b
        for word, e2 in ((a1, 0), (a3, 1), (a0, 2), (a2, 3)):   # the kernel pairs w0..w3 with xr[4c + e2]
            off = int(byte(word, b))
            assert off % 4 == 0 and off < 64
            p = off // 4                                      # table entry (hA hB lA lB)
            idx_a = 2 * ((p >> 3) & 1) + ((p >> 1) & 1)
            idx_b = 2 * ((p >> 2) & 1) + (p & 1)
            o = 8 * c + 2 * e2
            assert (idx_a, idx_b) == (truth[o], truth[o + 1]), (b, e2)


@pytest.mark.parametrize("seed", range(4))
def test_worddot3_mapping(oracle, seed):
    rng = np.random.default_rng(10 + seed)
    P2, P1, P0 = (U(rng.integers(0, 2**32)) for _ in range(3))
    truth = lane_indices(oracle, [P2, P1, P0])
    t = [bitsel(P2 << U(6), bitsel(P1 << U(4), P0 << U(2), 0x30303030), 0xC0C0C0C0) & U(0xFCFCFCFC),
         bitsel(P2 << U(4), bitsel(P1 << U(2), P0, 0x30303030), 0xC0C0C0C0) & U(0xFCFCFCFC),
         bitsel(P2 << U(2), bitsel(P1, P0 >> U(2), 0x30303030), 0xC0C0C0C0) & U(0xFCFCFCFC),
         bitsel(P2, bitsel(P1 >> U(2), P0 >> U(4), 0x30303030), 0xC0C0C0C0) & U(0xFCFCFCFC)]
    for b in range(4):
        c = 3 - b
        for j in range(4):                                    # t[j] pairs with xr[4c + (3 - j)]
            off = int(byte(t[j], b))
            assert off % 4 == 0
            p = off // 4                                      # (a2 b2 a1 b1 a0 b0)
            idx_a = 4 * ((p >> 5) & 1) + 2 * ((p >> 3) & 1) + ((p >> 1) & 1)
            idx_b = 4 * ((p >> 4) & 1) + 2 * ((p >> 2) & 1) + (p & 1)
            o = 8 * c + 2 * (3 - j)
            assert (idx_a, idx_b) == (truth[o], truth[o + 1]), (b, j)


@pytest.mark.parametrize("seed", range(4))
def test_worddot4_mapping(oracle, seed):
    rng = np.random.default_rng(20 + seed)
    P3, P2, P1, P0 = (U(rng.integers(0, 2**32)) for _ in range(4))
    truth = lane_indices(oracle, [P3, P2, P1, P0])
    for sft in range(4):
        if sft == 3:
            y = bitsel(P3, bitsel(P2 >> U(1), bitsel(P1 >> U(2), P0 >> U(3), 0x22222222), 0x44444444), 0x88888888)
        elif sft == 2:
            y = bitsel(P3 << U(1), bitsel(P2, bitsel(P1 >> U(1), P0 >> U(2), 0x22222222), 0x44444444), 0x88888888)
        elif sft == 1:
            y = bitsel(P3 << U(2), bitsel(P2 << U(1), bitsel(P1, P0 >> U(1), 0x22222222), 0x44444444), 0x88888888)
        else:
            y = bitsel(P3 << U(3), bitsel(P2 << U(2), bitsel(P1 << U(1), P0, 0x22222222), 0x44444444), 0x88888888)
        ylo, yhi = (y << U(1)) & U(0x1E1E1E1E), (y >> U(3)) & U(0x1E1E1E1E)
        for b in range(4):
            c = 3 - b
            assert int(byte(ylo, b)) // 2 == truth[8 * c + 7 - sft]   # e_lo = 7 - sft
            assert int(byte(yhi, b)) // 2 == truth[8 * c + 3 - sft]   # e_hi = 3 - sft


@pytest.mark.parametrize("bits", [4, 5, 6, 7, 8])
@pytest.mark.parametrize("seed", range(3))
def test_wide_butterfly_mapping(oracle, bits, seed):
    """apgemv_wide.cuh WideDequant<4..8>: 8x8 bit-matrix butterfly transpose of the plane words; byte b of A[s] must be
    (index << SH) of the weight at bit position 8b + s, i.e. k offset o = 8(3 - b) + (7 - s)."""
    rng = np.random.default_rng(100 * bits + seed)
    pw = [U(rng.integers(0, 2**32)) for _ in range(bits)]
    truth = lane_indices(oracle, pw)
    SH = 1 if bits <= 7 else 0
    A = [U(0)] * 8
    for r in range(8):
        if r >= SH and r - SH < bits:
            A[r] = pw[bits - 1 - (r - SH)]
    for r in range(4):
        a, b = A[r], A[r + 4]
        A[r], A[r + 4] = bitsel(b << U(4), a, 0xF0F0F0F0), bitsel(b, a >> U(4), 0xF0F0F0F0)
    for r in (0, 1, 4, 5):
        a, b = A[r], A[r + 2]
        A[r], A[r + 2] = bitsel(b << U(2), a, 0xCCCCCCCC), bitsel(b, a >> U(2), 0xCCCCCCCC)
    for r in (0, 2, 4, 6):
        a, b = A[r], A[r + 1]
        A[r], A[r + 1] = bitsel(b << U(1), a, 0xAAAAAAAA), bitsel(b, a >> U(1), 0xAAAAAAAA)
    for s in range(8):
        for b in range(4):
            v = int(byte(A[s], b))
            assert v % (1 << SH) == 0 and (v >> SH) < (1 << bits)
            assert (v >> SH) == truth[8 * (3 - b) + (7 - s)], (bits, s, b)


def test_sampling_key_order_is_monotone():
    """decode_kernels.cuh f16_order_key: ascending with the fp16 value, NaN lowest, -0 == +0, and invertible."""
    h = np.arange(65536, dtype=np.uint32)
    v = h.astype(np.uint16).view(np.float16).astype(np.float64)
    nan = (h & 0x7FFF) > 0x7C00
    hh = np.where(h == 0x8000, 0, h)
    key = np.where(nan, 0, np.where(hh & 0x8000, ~hh & 0xFFFF, hh | 0x8000))
    ok = ~nan
    order = np.argsort(v[ok], kind="stable")
    ks = key[ok][order]
    assert np.all(np.diff(ks) >= 0)
    assert np.all((np.diff(v[ok][order]) > 0) == (np.diff(ks) > 0))     # equal values <-> equal keys (only +-0)
    back = np.where(key & 0x8000, key & 0x7FFF, ~key & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float64)
    assert np.array_equal(back[ok], np.where(v[ok] == 0, 0.0, v[ok]))
    assert key[ok].min() > 0
