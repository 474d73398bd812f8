// Fast Any-Precision LUT GEMV for M = 1, bits in {2,3,4}, K % 128 == 0, K <= 32768  (sm_100a).
//
// Replaces matmul_kbit_32<1,bits,false> (reference inference/ap_gemv/anyprec.cu:372-542).  Same math,
// same packed tensors, different machine mapping:
//
//   * WEIGHT STREAM: one producer thread per CTA streams the CTA's contiguous row range through a
//     shared-memory ring with 1-D bulk async copies (cp.async.bulk -> UBLKCP, completion on mbarriers):
//     a stage = RS rows x BITS planes x K/8 bytes.  No registers are spent on prefetch, the ring keeps
//     tens of KB per SM in flight, and under programmatic dependent launch the ring fills while the
//     previous kernel is still running (weights do not depend on it; only x does).
//   * CONSUMER WARPS own one 1024-wide K chunk each (CPW = 2: two chunks): lane t owns word t of the
//     chunk in every plane, i.e. k = i*1024 + c*8*eff + 8t + e (pack.py:58-75).  That map is
//     row-independent, so the lane's 32 activations per chunk live in 16 half2 REGISTERS for the whole
//     kernel (the reference re-reads x through L1 for every row); they are fetched with four fully
//     coalesced LDG.128.
//   * CODEBOOKS: each warp expands the stage's lut rows into conflict-free shared-memory tables
//     (2-bit: 16-entry half2 pair table = 64 B, one LDS yields two weights; 3-bit: 64-entry pair table;
//     4-bit: the 16 halfs as they are).  Table byte offsets are produced four at a time (masked words)
//     and every lookup address is formed by ONE PRMT (byte insert into a 256-B aligned table base).
//   * ARITHMETIC: fp16 HFMA2 chains of length 8 -> fp32 accumulators (the reference is fp16 end to
//     end); RS rows are reduced together with a transposing shuffle tree (9 SHFL for 8 rows); the
//     per-chunk partial sums meet in shared memory and are added in a fixed order (deterministic).
#pragma once
#include "apgemv_common.cuh"

namespace apg {

// ------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// producer-side wait: the single producer thread only needs to notice a free slot eventually, so it sleeps
// between polls instead of burning issue slots the consumer warps need
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(512);
    }
}
// global -> shared 1-D bulk copy, completion (bytes) signalled on an mbarrier; L2 evict-first (streamed once)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(pol)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// ------------------------------------------------------------------------------------------------
// per-bit-width configuration
// ------------------------------------------------------------------------------------------------
template <int BITS>
struct FastCfg;
template <>
struct FastCfg<2> {
    static constexpr int ROW_TBL_BYTES = 64;  // 16 x half2
};
template <>
struct FastCfg<3> {
    static constexpr int ROW_TBL_BYTES = 256;  // 64 x half2
};
template <>
struct FastCfg<4> {
    static constexpr int ROW_TBL_BYTES = 64;  // 16 x half in the first 32 bytes (64-byte stride: row 4 starts a 256-byte block)
};
template <int BITS, int RS>
struct FastWarpTbl {
    static constexpr int BYTES = (RS * FastCfg<BITS>::ROW_TBL_BYTES + 255) / 256 * 256;
};

// ------------------------------------------------------------------------------------------------
// Transposing warp reduction over RS accumulators: afterwards lane (r << SH) holds row r's total.
// ------------------------------------------------------------------------------------------------
template <int RS>
struct BatchReduce;
template <>
struct BatchReduce<8> {
    static constexpr int SH = 2;
    __device__ __forceinline__ static float run(const float (&s)[8], int lane) {
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float keep = b4 ? s[4 + i] : s[i], send = b4 ? s[i] : s[4 + i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
        float b[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const float keep = b3 ? a[2 + i] : a[i], send = b3 ? a[i] : a[2 + i];
            b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        const float keep = b2 ? b[1] : b[0], send = b2 ? b[0] : b[1];
        float v = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        return v;
    }
};
template <>
struct BatchReduce<4> {
    static constexpr int SH = 3;
    __device__ __forceinline__ static float run(const float (&s)[4], int lane) { return batch_reduce<4>(s, lane); }
};
template <>
struct BatchReduce<2> {
    static constexpr int SH = 4;
    __device__ __forceinline__ static float run(const float (&s)[2], int lane) { return batch_reduce<2>(s, lane); }
};

// ------------------------------------------------------------------------------------------------
// Codebook expansion.  Every lane first fetches the lut words it needs for a stage (prefetched one
// stage ahead, see the main loop), then writes its share of the warp's tables.
// ------------------------------------------------------------------------------------------------
template <int BITS, int RS>
struct Tables;

// 2-bit: entry p = (hA hB lA lB) -> half2( C[2hA+lA], C[2hB+lB] ), A = even k (low half).
template <int RS>
struct Tables<2, RS> {
    static constexpr int IT = (RS + 1) / 2;  // lane covers row it*2 + (lane >> 4), entry lane & 15
    struct Regs {
        uint2 c[IT];
    };
    __device__ __forceinline__ static void fetch(Regs &r, const __half *__restrict__ lut, uint32_t row0, uint32_t N,
                                                 int lane) {
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const uint32_t row = min(row0 + it * 2 + (lane >> 4), N - 1);
            r.c[it] = __ldg(reinterpret_cast<const uint2 *>(lut + (size_t)row * 4));
        }
    }
    __device__ __forceinline__ static void prefetch(const __half *__restrict__ lut, uint32_t row0, uint32_t N, int lane) {
#pragma unroll
        for (int it = 0; it < IT; it++)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(lut + (size_t)min(row0 + it * 2 + (lane >> 4), N - 1) * 4));
    }
    __device__ __forceinline__ static void store(const Regs &r, uint32_t tbl, int lane) {
        const int p = lane & 15;
        const uint32_t a = ((p >> 3) & 1) * 2 + ((p >> 1) & 1), b = ((p >> 2) & 1) * 2 + (p & 1);
        const uint32_t sel = (2 * a) | ((2 * a + 1) << 4) | ((2 * b) << 8) | ((2 * b + 1) << 12);
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int row = it * 2 + (lane >> 4);
            if (row < RS) {
                const uint32_t e = __byte_perm(r.c[it].x, r.c[it].y, sel);
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(tbl + row * 64 + p * 4), "r"(e) : "memory");
            }
        }
    }
};

// 3-bit: entry p = (a2 b2 a1 b1 a0 b0) -> half2( C[a], C[b] ), A = even k.
// lane covers row (lane >> 2) + 8*it and the 16 entries p = (lane & 3) * 16 + 0..15 of it.
template <int RS>
struct Tables<3, RS> {
    static constexpr int IT = (RS + 7) / 8;
    struct Regs {
        uint4 c[IT];
    };
    __device__ __forceinline__ static void fetch(Regs &r, const __half *__restrict__ lut, uint32_t row0, uint32_t N,
                                                 int lane) {
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const uint32_t row = min(row0 + it * 8 + (lane >> 2), N - 1);
            r.c[it] = __ldg(reinterpret_cast<const uint4 *>(lut + (size_t)row * 8));
        }
    }
    __device__ __forceinline__ static void prefetch(const __half *__restrict__ lut, uint32_t row0, uint32_t N, int lane) {
#pragma unroll
        for (int it = 0; it < IT; it++)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(lut + (size_t)min(row0 + it * 8 + (lane >> 2), N - 1) * 8));
    }
    __device__ __forceinline__ static uint32_t pick(const uint4 &c, uint32_t idx) {  // half idx (0..7) -> low 16 bits
        const uint32_t lo = (idx & 4) ? c.z : c.x, hi = (idx & 4) ? c.w : c.y;     // halfs 0..3 or 4..7
        const uint32_t w = (idx & 2) ? hi : lo;
        return (idx & 1) ? (w >> 16) : (w & 0xffffu);
    }
    __device__ __forceinline__ static void store(const Regs &r, uint32_t tbl, int lane) {
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int row = it * 8 + (lane >> 2);
            if (row < RS) {
                const uint32_t a_hi = (lane >> 1) & 1, b_hi = lane & 1;  // p bits 5,4 are fixed per lane
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                    uint32_t v[4];
#pragma unroll
                    for (int f = 0; f < 4; f++) {
                        const int lowp = e + f;  // p bits 3..0 = (a1 b1 a0 b0)
                        const uint32_t a = a_hi * 4 + ((lowp >> 3) & 1) * 2 + ((lowp >> 1) & 1);
                        const uint32_t b = b_hi * 4 + ((lowp >> 2) & 1) * 2 + (lowp & 1);
                        v[f] = pick(r.c[it], a) | (pick(r.c[it], b) << 16);
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tbl + row * 256 + (lane & 3) * 64 + e * 4),
                                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
                                 : "memory");
                }
            }
        }
    }
};

// 4-bit: the 16 halfs of each row, verbatim.  lane covers row (lane >> 2) + 8*it, 8-byte piece lane & 3.
template <int RS>
struct Tables<4, RS> {
    static constexpr int IT = (RS + 7) / 8;
    struct Regs {
        uint2 c[IT];
    };
    __device__ __forceinline__ static void fetch(Regs &r, const __half *__restrict__ lut, uint32_t row0, uint32_t N,
                                                 int lane) {
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const uint32_t row = min(row0 + it * 8 + (lane >> 2), N - 1);
            r.c[it] = __ldg(reinterpret_cast<const uint2 *>(lut + (size_t)row * 16) + (lane & 3));
        }
    }
    __device__ __forceinline__ static void prefetch(const __half *__restrict__ lut, uint32_t row0, uint32_t N, int lane) {
#pragma unroll
        for (int it = 0; it < IT; it++)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(lut + (size_t)min(row0 + it * 8 + (lane >> 2), N - 1) * 16 + (lane & 3) * 4));
    }
    __device__ __forceinline__ static void store(const Regs &r, uint32_t tbl, int lane) {
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int row = it * 8 + (lane >> 2);
            if (row < RS)
                asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(tbl + row * FastCfg<4>::ROW_TBL_BYTES + (lane & 3) * 8), "r"(r.c[it].x),
                             "r"(r.c[it].y)
                             : "memory");
        }
    }
};

// ------------------------------------------------------------------------------------------------
// One word per plane (32 weights of one row, this lane's share of a chunk) against 16 half2 of x.
// xr[4c + e2] = half2( x[k0(c) + 2 e2], x[k0(c) + 2 e2 + 1] ), k0(c) = i*1024 + c*8*eff + 8t.
// ROW_OFF = r * ROW_TBL_BYTES selects the row's table by an LDS immediate.
// ------------------------------------------------------------------------------------------------
template <int BITS, int ROW_OFF>
struct WordDot;

template <int ROW_OFF>
struct WordDot<2, ROW_OFF> {
    __device__ __forceinline__ static float run(float acc, const uint32_t (&pw)[2], const uint32_t *xr, uint32_t tbl) {
        const uint32_t H = pw[0], L = pw[1];
        // nibble m of zh = (H[4m+3] H[4m+2] L[4m+3] L[4m+2]); of zl = (H[4m+1] H[4m] L[4m+1] L[4m])
        const uint32_t zh = bitsel(H, L >> 2, 0xCCCCCCCCu);
        const uint32_t zl = bitsel(H << 2, L, 0xCCCCCCCCu);
        // byte b of each word = 4 * pair index = a complete table byte offset; byte b <-> c = 3 - b
        const uint32_t a0 = (zh << 2) & 0x3C3C3C3Cu;  // even nibbles of zh -> pair e2 = 2
        const uint32_t a1 = (zh >> 2) & 0x3C3C3C3Cu;  // odd  nibbles of zh -> pair e2 = 0
        const uint32_t a2 = (zl << 2) & 0x3C3C3C3Cu;  // even nibbles of zl -> pair e2 = 3
        const uint32_t a3 = (zl >> 2) & 0x3C3C3C3Cu;  // odd  nibbles of zl -> pair e2 = 1
        uint32_t s0 = 0u, s1 = 0u;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int xb = 4 * (3 - b);
            const uint32_t sel = 0x7650u | b;
            const uint32_t w0 = lds_b32_imm<ROW_OFF>(__byte_perm(a1, tbl, sel));
            const uint32_t w1 = lds_b32_imm<ROW_OFF>(__byte_perm(a3, tbl, sel));
            const uint32_t w2 = lds_b32_imm<ROW_OFF>(__byte_perm(a0, tbl, sel));
            const uint32_t w3 = lds_b32_imm<ROW_OFF>(__byte_perm(a2, tbl, sel));
            s0 = hfma2_u32(w0, xr[xb + 0], s0);
            s1 = hfma2_u32(w1, xr[xb + 1], s1);
            s0 = hfma2_u32(w2, xr[xb + 2], s0);
            s1 = hfma2_u32(w3, xr[xb + 3], s1);
        }
        return acc_add_h2(acc, hadd2_u32(s0, s1));
    }
};

template <int ROW_OFF>
struct WordDot<3, ROW_OFF> {
    __device__ __forceinline__ static float run(float acc, const uint32_t (&pw)[3], const uint32_t *xr, uint32_t tbl) {
        const uint32_t P2 = pw[0], P1 = pw[1], P0 = pw[2];
        // target j: byte b = (P2[p+1] P2[p] P1[p+1] P1[p] P0[p+1] P0[p] 0 0), p = 8b+2j  -> x pair e2 = 3-j
        uint32_t t[4];
        t[0] = bitsel(P2 << 6, bitsel(P1 << 4, P0 << 2, 0x30303030u), 0xC0C0C0C0u) & 0xFCFCFCFCu;
        t[1] = bitsel(P2 << 4, bitsel(P1 << 2, P0, 0x30303030u), 0xC0C0C0C0u) & 0xFCFCFCFCu;
        t[2] = bitsel(P2 << 2, bitsel(P1, P0 >> 2, 0x30303030u), 0xC0C0C0C0u) & 0xFCFCFCFCu;
        t[3] = bitsel(P2, bitsel(P1 >> 2, P0 >> 4, 0x30303030u), 0xC0C0C0C0u) & 0xFCFCFCFCu;
        uint32_t s0 = 0u, s1 = 0u;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int xb = 4 * (3 - b);
            const uint32_t sel = 0x7650u | b;
            const uint32_t w0 = lds_b32_imm<ROW_OFF>(__byte_perm(t[3], tbl, sel));
            const uint32_t w1 = lds_b32_imm<ROW_OFF>(__byte_perm(t[2], tbl, sel));
            const uint32_t w2 = lds_b32_imm<ROW_OFF>(__byte_perm(t[1], tbl, sel));
            const uint32_t w3 = lds_b32_imm<ROW_OFF>(__byte_perm(t[0], tbl, sel));
            s0 = hfma2_u32(w0, xr[xb + 0], s0);
            s1 = hfma2_u32(w1, xr[xb + 1], s1);
            s0 = hfma2_u32(w2, xr[xb + 2], s0);
            s1 = hfma2_u32(w3, xr[xb + 3], s1);
        }
        return acc_add_h2(acc, hadd2_u32(s0, s1));
    }
};

template <int ROW_OFF>
struct WordDot<4, ROW_OFF> {
    __device__ __forceinline__ static float run(float acc, const uint32_t (&pw)[4], const uint32_t *xr, uint32_t tbl) {
        const uint32_t P3 = pw[0], P2 = pw[1], P1 = pw[2], P0 = pw[3];
        __half s[4] = {__ushort_as_half(0), __ushort_as_half(0), __ushort_as_half(0), __ushort_as_half(0)};
#pragma unroll
        for (int sft = 0; sft < 4; sft++) {
            // nibble m of y = 4-bit index (P3 P2 P1 P0) of the weight at bit position 4m+sft
            uint32_t y;
            if (sft == 3)
                y = bitsel(P3, bitsel(P2 >> 1, bitsel(P1 >> 2, P0 >> 3, 0x22222222u), 0x44444444u), 0x88888888u);
            else if (sft == 2)
                y = bitsel(P3 << 1, bitsel(P2, bitsel(P1 >> 1, P0 >> 2, 0x22222222u), 0x44444444u), 0x88888888u);
            else if (sft == 1)
                y = bitsel(P3 << 2, bitsel(P2 << 1, bitsel(P1, P0 >> 1, 0x22222222u), 0x44444444u), 0x88888888u);
            else
                y = bitsel(P3 << 3, bitsel(P2 << 2, bitsel(P1 << 1, P0, 0x22222222u), 0x44444444u), 0x88888888u);
            // byte b = 2*index.  ylo: bit 8b+sft   -> k offset 31-8b-sft -> c = 3-b, e = 7-sft
            //                    yhi: bit 8b+4+sft -> k offset 27-8b-sft -> c = 3-b, e = 3-sft
            const uint32_t ylo = (y << 1) & 0x1E1E1E1Eu, yhi = (y >> 3) & 0x1E1E1E1Eu;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int xb = 4 * (3 - b);
                const uint32_t sel = 0x7650u | b;
                const int e_lo = 7 - sft, e_hi = 3 - sft;
                const __half wl = __ushort_as_half((unsigned short)lds_u16_imm<ROW_OFF>(__byte_perm(ylo, tbl, sel)));
                const __half wh = __ushort_as_half((unsigned short)lds_u16_imm<ROW_OFF>(__byte_perm(yhi, tbl, sel)));
                const __half2 xl = *reinterpret_cast<const __half2 *>(&xr[xb + e_lo / 2]);
                const __half2 xh = *reinterpret_cast<const __half2 *>(&xr[xb + e_hi / 2]);
                s[sft] = __hfma(wl, (e_lo & 1) ? __high2half(xl) : __low2half(xl), s[sft]);
                s[(sft + 2) & 3] = __hfma(wh, (e_hi & 1) ? __high2half(xh) : __low2half(xh), s[(sft + 2) & 3]);
            }
        }
        return acc + ((__half2float(s[0]) + __half2float(s[1])) + (__half2float(s[2]) + __half2float(s[3])));
    }
};

// compile-time unrolled loop over the RS rows of a stage (the row index must be a template constant so
// that the table offset becomes an LDS immediate).  All RS rows are always computed (no per-row branch,
// so the compiler can interleave rows); rows past the end of a partial stage read stale-but-mapped ring
// bytes and are discarded by the caller.
// REND = RS for a full stage, RS/2 for a stage that holds at most half the rows (the CTA's last one).
template <int BITS, int RS, int R, int REND>
struct RowLoop {
    __device__ __forceinline__ static void run(float (&acc)[RS], uint32_t pbase, uint32_t row_bytes, uint32_t plane_bytes,
                                               const uint32_t *xr, uint32_t tbl) {
        uint32_t pw[BITS];
#pragma unroll
        for (int j = 0; j < BITS; j++) pw[j] = lds_b32(pbase + j * plane_bytes);
        acc[R] = WordDot<BITS, R * FastCfg<BITS>::ROW_TBL_BYTES>::run(acc[R], pw, xr, tbl);
        RowLoop<BITS, RS, R + 1, REND>::run(acc, pbase + row_bytes, row_bytes, plane_bytes, xr, tbl);
    }
};
template <int BITS, int RS, int REND>
struct RowLoop<BITS, RS, REND, REND> {
    __device__ __forceinline__ static void run(float (&)[RS], uint32_t, uint32_t, uint32_t, const uint32_t *, uint32_t) {}
};

// ------------------------------------------------------------------------------------------------
// Kernel
// ------------------------------------------------------------------------------------------------
struct FastParams {
    const __half *x;         // [K]
    const uint8_t *W;        // [bits][N][K/8] bytes
    const __half *lut;       // [N][2^bits]
    __half *out;             // [N] or nullptr
    float *partial;          // [N] or nullptr
    uint32_t N, K;
    uint32_t nwk;            // consumer warps per row group = ceil(nchunk / CPW)
    uint32_t groups;         // row groups per CTA; consumer warps = groups * nwk; +1 producer warp
    uint32_t nslots;         // ring slots (multiple of groups)
    uint32_t stage_bytes;    // RS * BITS * K/8
    uint32_t unit_rows;      // rows are dealt to CTAs in units of RS or RS/2 rows (half stages: finer balance over the SMs)
    uint32_t units_q, units_rem;  // CTA b owns units_q + (b < units_rem) consecutive units
    uint32_t inv_nwk;        // ceil(65536 / nwk): warp / nwk == (warp * inv_nwk) >> 16 for warp < 64
    // ---- optional fusions of the ops that surround the Linear in the decode step (inference/model.py) ----
    const __half *norm_w;    // != nullptr: x := RMSNorm(x) * norm_w before the GEMV (model.py:280-285), eps below
    float norm_eps;
    uint32_t act_silu_mul;   // 1: x := silu(x[0:K]) * x[K:2K] (FeedForward, model.py:261-266); x holds 2K halfs
    const __half *residual;  // != nullptr: out := fp16(y) + residual in fp16 (TransformerBlock, model.py:151-167)
    // ---- optional fused one-shot all-reduce push (K-sharded multi-GPU path): every row's fp32 partial
    // sum is written, together with the current epoch, as ONE 8-byte store into slot `rank` of EVERY peer's receive
    // buffer over NVLink (peer_recv[p] = that peer's uint2 [world][N]).  Data and flag travel in the same store (the
    // "LL" idea): no fences, no atomics, one one-way NVLink latency; the receiver polls each slot for the epoch.
    uint32_t world, rank;
    uint2 *peer_recv[8];
    const uint32_t *epoch;   // device counter of this all-reduce site; the value pushed is *epoch + 1
    const uint8_t *prefetch; // optional: bytes the NEXT kernel on the stream will stream (its weights) ...
    uint64_t prefetch_bytes; // ... pulled into L2 by this kernel's producer threads while its warps compute
};

// smem: [barriers 2*nslots*8 | pad->256 | tables (per consumer warp) | ring nslots*stage_bytes | red floats]
// GLU = true (experimental, selected by apg_gemv_fused(silu_mul = 2)): the Linear's rows are interleaved (gate_i, up_i)
// and the epilogue writes out[i] = silu(y[2i]) * y[2i+1] — the FeedForward activation (model.py:261-266) computed once
// per element here instead of by every CTA of the w2 launch that follows.  The default instantiations are unchanged.
template <int BITS, int CPW, int RS, bool GLU = false>
__global__ void __launch_bounds__(544, 1) gemv_fast_kernel(const FastParams p) {
    constexpr int WTB = FastWarpTbl<BITS, RS>::BYTES;
    using Tb = Tables<BITS, RS>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nwk = p.nwk, G = p.groups, NS = p.nslots;
    const uint32_t ncons = G * nwk;
    const uint32_t K = p.K, N = p.N;
    const uint32_t row_bytes = K >> 3;  // bytes of one (row, plane)

    const uint32_t smem0 = smem_u32(smem_raw);
    const uint32_t bar_full = smem0, bar_empty = smem0 + NS * 8u;
    float *ssq = reinterpret_cast<float *>(smem_raw + 2u * NS * 8u);  // [ncons] sums of squares (fused RMSNorm)
    const uint32_t tbl0 = (smem0 + 2u * NS * 8u + 64u * 4u + 255u) & ~255u;
    const uint32_t ring0 = tbl0 + ncons * WTB;
    float *red = reinterpret_cast<float *>(smem_raw + (ring0 - smem0) + NS * p.stage_bytes);

    // rows of this CTA: consecutive units of unit_rows rows, walked in stages of RS rows; only the CTA's last stage can
    // be partial, and a stage holding at most RS/2 rows is computed with the half-length row loop
    const uint32_t u_begin = blockIdx.x * p.units_q + min(blockIdx.x, p.units_rem);
    const uint32_t nunits = p.units_q + (blockIdx.x < p.units_rem ? 1u : 0u);
    const uint32_t r_begin = min(u_begin * p.unit_rows, N);
    const uint32_t r_end = min((u_begin + nunits) * p.unit_rows, N);
    const uint32_t nrows = r_end - r_begin;
    const uint32_t nstages = (nrows + RS - 1) / RS;

    __half res_pref = __ushort_as_half(0);  // residual of row r_begin + threadIdx.x, prefetched by consumer threads
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < NS; s++) {
            mbar_init(bar_full + 8u * s, 1u);
            mbar_init(bar_empty + 8u * s, nwk);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if ((uint32_t)warp == ncons) {
        // ===================== producer: stream [r_begin, r_end) through the ring =====================
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            // L2 prefetch of what the next kernel on the stream will read (its packed weights): HBM is mostly idle
            // while this kernel's warps chew through their tables, and the successor's ring then fills at L2 latency.
            // Issued once, right after this CTA's own ring has been filled for the first time.
            auto prefetch_next = [&]() {
                if (!p.prefetch_bytes) return;
                const uint64_t per = ((p.prefetch_bytes / gridDim.x) + 15u) & ~15ull;
                const uint64_t off = per * blockIdx.x;
                if (off >= p.prefetch_bytes) return;
                uint64_t n = min(per, p.prefetch_bytes - off) & ~15ull;
                const uint8_t *src = p.prefetch + off;
                while (n) {
                    const uint32_t c = (uint32_t)min(n, (uint64_t)65536u);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(c) : "memory");
                    src += c, n -= c;
                }
            };
            const uint32_t pf_at = min(NS, nstages);
            uint32_t slot = 0, use = 0;
            if (pf_at == 0) prefetch_next();
            for (uint32_t s = 0; s < nstages; s++) {
                if (s == pf_at) prefetch_next();
                if (use > 0) mbar_wait_backoff(bar_empty + 8u * slot, (use - 1u) & 1u);
                const uint32_t row0 = r_begin + s * RS;
                const uint32_t rows = min((uint32_t)RS, r_end - row0);
                const uint32_t bytes = rows * row_bytes;
                mbar_arrive_expect_tx(bar_full + 8u * slot, bytes * BITS);
#pragma unroll
                for (int j = 0; j < BITS; j++)
                    bulk_g2s(ring0 + slot * p.stage_bytes + j * RS * row_bytes,
                             p.W + ((size_t)j * N + row0) * row_bytes, bytes, bar_full + 8u * slot, pol);
                if (++slot == NS) slot = 0, use++;
            }
            if (pf_at == nstages && nstages > 0) prefetch_next();
        }
    } else {
        // ===================== consumers =====================
        const uint32_t g = ((uint32_t)warp * p.inv_nwk) >> 16, wk = warp - g * nwk;
        const uint32_t tbl = tbl0 + warp * WTB;
        const uint32_t nchunk = (K + 1023u) >> 10;

        // codebook rows of this group's first stage (independent of the previous kernel)
        typename Tb::Regs lr;
        if (g < nstages) Tb::fetch(lr, p.lut, r_begin + g * RS, N, lane);

        // RMSNorm weights of this lane's k (static data: fetched before the dependency wait)
        uint32_t wn[CPW][16];
        if (p.norm_w) {
#pragma unroll
            for (int cc = 0; cc < CPW; cc++) {
                const uint32_t i = wk * CPW + cc;
                const uint32_t eff = (i < nchunk) ? chunk_eff(K, i) : 0u;
                if ((uint32_t)lane < eff) {
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const uint4 wv = __ldg(reinterpret_cast<const uint4 *>(p.norm_w + i * 1024u + c * 8u * eff + 8u * lane));
                        wn[cc][4 * c + 0] = wv.x, wn[cc][4 * c + 1] = wv.y, wn[cc][4 * c + 2] = wv.z, wn[cc][4 * c + 3] = wv.w;
                    }
                }
            }
        }

        // x -> registers (produced by the previous kernel on the stream: wait for it under PDL)
        pdl_wait_prior_grid();
        if (p.residual && threadIdx.x < nrows) res_pref = p.residual[r_begin + threadIdx.x];
        uint32_t xr[CPW][16];
        bool act[CPW];
        uint32_t woff[CPW];  // byte offset of this lane's word inside a (row, plane)
        float ss = 0.f;
#pragma unroll
        for (int cc = 0; cc < CPW; cc++) {
            const uint32_t i = wk * CPW + cc;
            const uint32_t eff = (i < nchunk) ? chunk_eff(K, i) : 0u;
            act[cc] = (uint32_t)lane < eff;
            woff[cc] = (i * 32u + lane) * 4u;
            if (act[cc]) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint32_t k0 = i * 1024u + c * 8u * eff + 8u * lane;
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p.x + k0));
                    xr[cc][4 * c + 0] = v.x, xr[cc][4 * c + 1] = v.y, xr[cc][4 * c + 2] = v.z, xr[cc][4 * c + 3] = v.w;
                    if (p.act_silu_mul) {  // x := silu(gate) * up, both rounded to fp16 like the reference's half tensors
                        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p.x + K + k0));
                        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float2 gf = __half22float2(*reinterpret_cast<const __half2 *>(&xr[cc][4 * c + e]));
                            // silu in fp32 with the fast reciprocal (2 ulp, invisible after the fp16 rounding); an IEEE
                            // division here costs ~400 serial instructions per warp on the critical path of every w2
                            const __half2 sg = __floats2half2_rn(__fdividef(gf.x, 1.f + __expf(-gf.x)),
                                                                 __fdividef(gf.y, 1.f + __expf(-gf.y)));
                            const __half2 r = __hmul2(sg, *reinterpret_cast<const __half2 *>(&uu[e]));
                            xr[cc][4 * c + e] = *reinterpret_cast<const uint32_t *>(&r);
                        }
                    }
                    if (p.norm_w) {
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&xr[cc][4 * c + e]));
                            ss = fmaf(f.x, f.x, ss);
                            ss = fmaf(f.y, f.y, ss);
                        }
                    }
                }
            }
        }
        if (p.norm_w) {
            // fused RMSNorm (model.py:280-285): every row group covers all K chunks once, so the group's warps
            // exchange their sums of squares through shared memory (fixed summation order)
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0) ssq[warp] = ss;
            asm volatile("bar.sync 1, %0;" ::"r"(ncons * 32u) : "memory");
            float tot = 0.f;
            for (uint32_t w = 0; w < nwk; w++) tot += ssq[g * nwk + w];
            const float rs = rsqrtf(tot / (float)K + p.norm_eps);
#pragma unroll
            for (int cc = 0; cc < CPW; cc++) {
                if (act[cc]) {
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&xr[cc][e]));
                        const __half2 n = __floats2half2_rn(f.x * rs, f.y * rs);  // .type_as(x)
                        const __half2 r = __hmul2(n, *reinterpret_cast<const __half2 *>(&wn[cc][e]));  // * weight
                        xr[cc][e] = *reinterpret_cast<const uint32_t *>(&r);
                    }
                }
            }
        }
        pdl_launch_dependents();

        uint32_t slot = g, use = 0;  // NS is a multiple of G: a slot always serves the same group
        for (uint32_t s = g; s < nstages; s += G) {
            const uint32_t row0 = r_begin + s * RS;
            const uint32_t rows = min((uint32_t)RS, r_end - row0);
            // tables for this stage from the prefetched codebook rows; prefetch the next stage's rows
            // (unconditional: the row index is clamped inside fetch, so no select/copy waits on the load)
            __syncwarp();
            Tb::store(lr, tbl, lane);
            Tb::fetch(lr, p.lut, row0 + G * RS, N, lane);
            __syncwarp();
            mbar_wait(bar_full + 8u * slot, use & 1u);

            const uint32_t stage = ring0 + slot * p.stage_bytes;
            // (a ROLLED loop over groups of four rows — a quarter of the code, friendlier to the instruction cache — was
            //  measured 8.6 % slower per token on B200 than this fully unrolled stage: profiles/r2_ab_rolled_rowloop.txt)
            float acc[RS];
#pragma unroll
            for (int r = 0; r < RS; r++) acc[r] = 0.f;
            if (RS >= 2 && rows <= (uint32_t)(RS / 2)) {
#pragma unroll
                for (int cc = 0; cc < CPW; cc++)
                    if (act[cc])
                        RowLoop<BITS, RS, 0, (RS >= 2 ? RS / 2 : RS)>::run(acc, stage + woff[cc], row_bytes, RS * row_bytes, xr[cc], tbl);
            } else {
#pragma unroll
                for (int cc = 0; cc < CPW; cc++)
                    if (act[cc]) RowLoop<BITS, RS, 0, RS>::run(acc, stage + woff[cc], row_bytes, RS * row_bytes, xr[cc], tbl);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8u * slot);

            const float v = BatchReduce<RS>::run(acc, lane);
            const int rl = lane >> BatchReduce<RS>::SH;
            if ((lane & ((1 << BatchReduce<RS>::SH) - 1)) == 0 && (uint32_t)rl < rows)
                red[(row0 - r_begin + rl) * nwk + wk] = v;
            slot += G;
            if (slot >= NS) slot -= NS, use++;
        }
    }

    __syncthreads();
    if constexpr (GLU) {
        // r_begin and nrows are even (units of RS/2 = 4 rows, N even): thread t combines the row pair (2t, 2t+1) with the
        // same roundings as the w2 prologue it replaces (y -> fp16, silu in fp32 -> fp16, fp16 product)
        for (uint32_t r = 2u * threadIdx.x; r + 1u < nrows; r += 2u * blockDim.x) {
            float vg = red[r * nwk], vu = red[(r + 1u) * nwk];
            for (uint32_t w = 1; w < nwk; w++) vg += red[r * nwk + w], vu += red[(r + 1u) * nwk + w];
            const float gf = __half2float(__float2half_rn(vg));
            const __half sg = __float2half_rn(__fdividef(gf, 1.f + __expf(-gf)));
            p.out[(r_begin + r) >> 1] = __hmul(sg, __float2half_rn(vu));
        }
        return;
    }
    const uint32_t ep = (p.world > 1) ? (*p.epoch + 1u) : 0u;
    // fixed-order combination of the per-chunk partial sums -> deterministic results
    for (uint32_t r = threadIdx.x; r < nrows; r += blockDim.x) {
        float v = red[r * nwk];
        for (uint32_t w = 1; w < nwk; w++) v += red[r * nwk + w];
        if (p.out) {
            __half h = __float2half_rn(v);
            if (p.residual) h = __hadd(h, (r == threadIdx.x && threadIdx.x < ncons * 32u) ? res_pref : p.residual[r_begin + r]);
            p.out[r_begin + r] = h;
        }
        if (p.partial) p.partial[r_begin + r] = v;
        if (p.world > 1) {
            const uint2 pkt = make_uint2(__float_as_uint(v), ep);
#pragma unroll 1
            for (uint32_t pr = 0; pr < p.world; pr++) {
                uint2 *dst = p.peer_recv[pr] + (size_t)p.rank * N + r_begin + r;
                asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(pkt.x), "r"(pkt.y) : "memory");
            }
        }
    }
}

}  // namespace apg
