/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, scalar) of the reference's Any-Precision LUT GEMV path, used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * checker.  Nothing under guidedquant_b200/ may import, link or call this file.
 *
 * Parity status: PINNED.  (1) pack/unpack are checked against the reference's own
 * any_precision/quantization/pack.py (imported in the build container by
 * tests/golden/make_golden.py; vectors committed under tests/golden/).  (2) dequant and the
 * reference-order fp16 GEMV emulation are checked bit-for-bit against the UNMODIFIED reference
 * kernels (inference/ap_gemv/anyprec.cu compiled for sm_100a into oracle/_ref/) run on a B200
 * (tests/test_parity_gpu.py; the outputs of that run are committed as tests/golden/ref_gpu_*.npz
 * so the CPU suite re-checks the oracle against them without a GPU).
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>

/* ------------------------------------------------------------------------------------------ */
/* IEEE binary16 helpers (round-to-nearest-even, denormals kept — the reference build has no   */
/* --use_fast_math, inference/ap_gemv/setup.py:12-27).                                          */
/* ------------------------------------------------------------------------------------------ */

static double h2d(uint16_t h) {
    uint32_t s = (h >> 15) & 1u, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    double v;
    if (e == 0) v = ldexp((double)m, -24);
    else if (e == 31) v = m ? NAN : INFINITY;
    else v = ldexp((double)(m | 0x400u), (int)e - 25);
    return s ? -v : v;
}

/* double -> binary16, one correctly rounded (RNE) step. */
static uint16_t d2h(double d) {
    uint16_t sign = 0;
    if (d != d) return 0x7fff;
    if (signbit(d)) { sign = 0x8000; d = -d; }
    if (d == 0.0) return sign;
    if (isinf(d)) return sign | 0x7c00;
    int ex;
    double fr = frexp(d, &ex); /* d = fr * 2^ex, fr in [0.5,1) */
    int e = ex - 1;            /* d = (2*fr) * 2^e, 2*fr in [1,2) */
    if (e < -14) {             /* subnormal range: quantum 2^-24 */
        double q = nearbyint(ldexp(d, 24)); /* default rounding mode = RNE */
        if (q >= 1024.0) return sign | 0x0400;
        return sign | (uint16_t)q;
    }
    double q = nearbyint(ldexp(fr, 11)); /* 11 significant bits: [1024, 2048] */
    if (q >= 2048.0) { q = 1024.0; e += 1; }
    if (e > 15) return sign | 0x7c00;
    return sign | (uint16_t)(((uint32_t)(e + 15) << 10) | ((uint32_t)q - 1024u));
}

/* fp16 fused multiply-add, single rounding (== __hfma2 per half).  a*b is exact in double
 * (22-bit product); fma() then rounds the exact sum once to 53 bits, and the 53->11-bit second
 * rounding cannot land on a false tie for binary16 operands (see DESIGN.md "oracle numerics"). */
static uint16_t hfma(uint16_t a, uint16_t b, uint16_t c) { return d2h(fma(h2d(a), h2d(b), h2d(c))); }
/* fp16 add (== half operator+): the double sum of two binary16 values is exact. */
static uint16_t hadd(uint16_t a, uint16_t b) { return d2h(h2d(a) + h2d(b)); }

uint16_t apo_f64_to_f16(double d) { return d2h(d); }
double apo_f16_to_f64(uint16_t h) { return h2d(h); }

/* ------------------------------------------------------------------------------------------ */
/* Packed layout (any_precision/quantization/pack.py:12-83, 304-347).                           */
/*                                                                                              */
/* qweight[j][n][w] (int32, plane j = 0 is the MSB, pack.py:312-316).  Bytes of a row-plane are  */
/* np.packbits (first k -> bit 7, pack.py:314) then permuted per 128-byte group so that lane t's */
/* little-endian word has byte (3-c) = original byte c*eff + t (pack.py:58-75; endianness flip   */
/* `^= 3` pack.py:72); eff = 32 for full 1024-weight chunks and (K%1024)/32 for the tail chunk   */
/* (pack.py:33-46).  Closed form: word w = i*32 + t, bit 31-(8c+e)  <->  k = i*1024 + c*8*eff +  */
/* 8t + e.                                                                                      */
/* ------------------------------------------------------------------------------------------ */

static inline uint32_t chunk_eff(uint32_t K, uint32_t i) {
    return (i < K / 1024u) ? 32u : (K % 1024u) / 32u;
}

/* idx: uint8 [N][K] with values < 2^bits  ->  qweight: uint32 [bits][N][K/32].
 * Restates pack_single_weight (pack.py:304-321). */
int apo_pack(const uint8_t *idx, uint32_t N, uint32_t K, int bits, uint32_t *qweight) {
    if (bits < 1 || bits > 8 || K % 32u) return 1;
    const uint32_t words = K / 32u, nchunk = (K + 1023u) / 1024u;
    for (int j = 0; j < bits; j++)
        for (uint32_t n = 0; n < N; n++) {
            uint32_t *row = qweight + ((size_t)j * N + n) * words;
            const uint8_t *irow = idx + (size_t)n * K;
            for (uint32_t i = 0; i < nchunk; i++) {
                const uint32_t eff = chunk_eff(K, i);
                for (uint32_t t = 0; t < eff; t++) {
                    uint32_t word = 0;
                    for (uint32_t c = 0; c < 4; c++)
                        for (uint32_t e = 0; e < 8; e++) {
                            const uint32_t k = i * 1024u + c * 8u * eff + 8u * t + e;
                            const uint32_t bit = (irow[k] >> (bits - 1 - j)) & 1u;
                            word |= bit << (31u - (8u * c + e));
                        }
                    row[i * 32u + t] = word;
                }
            }
        }
    return 0;
}

/* qweight -> idx uint8 [N][K].  Restates unpack_single_weight (pack.py:324-347). */
int apo_unpack(const uint32_t *qweight, uint32_t N, uint32_t K, int bits, uint8_t *idx) {
    if (bits < 1 || bits > 8 || K % 32u) return 1;
    const uint32_t words = K / 32u, nchunk = (K + 1023u) / 1024u;
    memset(idx, 0, (size_t)N * K);
    for (int j = 0; j < bits; j++)
        for (uint32_t n = 0; n < N; n++) {
            const uint32_t *row = qweight + ((size_t)j * N + n) * words;
            uint8_t *irow = idx + (size_t)n * K;
            for (uint32_t i = 0; i < nchunk; i++) {
                const uint32_t eff = chunk_eff(K, i);
                for (uint32_t t = 0; t < eff; t++) {
                    const uint32_t word = row[i * 32u + t];
                    for (uint32_t c = 0; c < 4; c++)
                        for (uint32_t e = 0; e < 8; e++) {
                            const uint32_t k = i * 1024u + c * 8u * eff + 8u * t + e;
                            irow[k] |= (uint8_t)(((word >> (31u - (8u * c + e))) & 1u) << (bits - 1 - j));
                        }
                }
            }
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Dequant: W[n][k] = lut[n][idx[n][k]]  (inference/ap_gemv/anyprec.cu:294-359 and the python    */
/* definition any_precision/evaluate/eval.py:102-108).  Pure gather on fp16 bit patterns.        */
/* ------------------------------------------------------------------------------------------ */
int apo_dequant(const uint32_t *qweight, const uint16_t *lut, uint32_t N, uint32_t K, int bits,
                uint8_t *idx_scratch, uint16_t *W) {
    int rc = apo_unpack(qweight, N, K, bits, idx_scratch);
    if (rc) return rc;
    const uint32_t nc = 1u << bits;
    for (uint32_t n = 0; n < N; n++)
        for (uint32_t k = 0; k < K; k++)
            W[(size_t)n * K + k] = lut[(size_t)n * nc + idx_scratch[(size_t)n * K + k]];
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* GEMV, three arithmetic models on the same dequantised weights (W = fp16 [N][K], x fp16 [M][K])*/
/* ------------------------------------------------------------------------------------------ */

/* (1) fp64 "truth": y[m][n] = sum_k W[n][k] * x[m][k] in double. */
void apo_gemv_f64(const uint16_t *W, const uint16_t *x, uint32_t M, uint32_t N, uint32_t K, double *y) {
    for (uint32_t m = 0; m < M; m++)
        for (uint32_t n = 0; n < N; n++) {
            double acc = 0.0;
            const uint16_t *w = W + (size_t)n * K, *xv = x + (size_t)m * K;
            for (uint32_t k = 0; k < K; k++) acc += h2d(w[k]) * h2d(xv[k]);
            y[(size_t)m * N + n] = acc;
        }
}

/* (2) Bit-exact emulation of the reference kernel's fp16 arithmetic and accumulation order
 * (matmul_kbit_32, inference/ap_gemv/anyprec.cu:424-541; warp_reduce_sum :362-370; SURVEY.md
 * Appendix A).  Per row, lane t in 0..31:
 *   partial_t = 0
 *   for chunk i: lanes t >= eff skip the tail chunk (:433-436)
 *       s = (0,0); for c = 3..0 (:497), m = 0..3 (:502-503):
 *           s = hfma2( (W[k0+2m], W[k0+2m+1]), (x[k0+2m], x[k0+2m+1]), s ),  k0 = i*1024 + c*8*eff + 8t
 *       partial_t += (s.x + s.y)                                   (:505, two half adds)
 *   shuffle tree offsets 16,8,4,2,1 (:367-368) in half; lane 0 stores (:538).
 * __shfl_down_sync with offset o: lane t reads lane t+o when t+o < 32, else its own value. */
void apo_gemv_ref_order_f16(const uint16_t *W, const uint16_t *x, uint32_t M, uint32_t N, uint32_t K,
                            uint16_t *y) {
    const uint32_t nchunk = (K + 1023u) / 1024u;
    for (uint32_t m = 0; m < M; m++)
        for (uint32_t n = 0; n < N; n++) {
            const uint16_t *w = W + (size_t)n * K, *xv = x + (size_t)m * K;
            uint16_t partial[32];
            for (uint32_t t = 0; t < 32; t++) {
                uint16_t p = 0;
                for (uint32_t i = 0; i < nchunk; i++) {
                    const uint32_t eff = chunk_eff(K, i);
                    if (t >= eff) break;
                    uint16_t sx = 0, sy = 0;
                    for (int c = 3; c >= 0; c--) {
                        const uint32_t k0 = i * 1024u + (uint32_t)c * 8u * eff + 8u * t;
                        for (uint32_t q = 0; q < 4; q++) {
                            sx = hfma(w[k0 + 2 * q], xv[k0 + 2 * q], sx);
                            sy = hfma(w[k0 + 2 * q + 1], xv[k0 + 2 * q + 1], sy);
                        }
                    }
                    p = hadd(p, hadd(sx, sy));
                }
                partial[t] = p;
            }
            for (uint32_t off = 16; off >= 1; off >>= 1) {
                uint16_t nxt[32];
                for (uint32_t t = 0; t < 32; t++)
                    nxt[t] = hadd(partial[t], (t + off < 32) ? partial[t + off] : partial[t]);
                memcpy(partial, nxt, sizeof(partial));
            }
            y[(size_t)m * N + n] = partial[0];
        }
}

/* (3) BASELINE config 0, "dequant -> fp16 matmul" as APLinear.gemm does it
 * (inference/APLinear.py:35-38): fp32 accumulation of fp16 products, one rounding to fp16 at the
 * end (what a CPU/cuBLAS fp16 matmul with fp32 accumulate returns).  Scalar; the multi-threaded
 * timing variant lives in oracle/oracle.py (torch.matmul on the host cores). */
void apo_gemv_dequant_matmul_f16(const uint16_t *W, const uint16_t *x, uint32_t M, uint32_t N, uint32_t K,
                                 uint16_t *y) {
    for (uint32_t m = 0; m < M; m++)
        for (uint32_t n = 0; n < N; n++) {
            float acc = 0.f;
            const uint16_t *w = W + (size_t)n * K, *xv = x + (size_t)m * K;
            for (uint32_t k = 0; k < K; k++) acc += (float)h2d(w[k]) * (float)h2d(xv[k]);
            y[(size_t)m * N + n] = d2h((double)acc);
        }
}
