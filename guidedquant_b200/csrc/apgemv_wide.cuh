// "Wide" Any-Precision LUT GEMV: every case the fast kernel does not take — bits 5..8, batch M = 2..8 (any bits),
// K not a multiple of 128, K > 32768 — at full generality (bits 2..8, M 1..8, any N, K % 32 == 0).
//
// Replaces matmul_kbit_32<maxm, bits, *> (reference inference/ap_gemv/anyprec.cu:372-542) for those cases.  Mapping:
//   * a warp owns R = 2 output rows at a time (R = 1 when N is too small to fill the GPU that way) and walks their 1024-wide K chunks with the reference's lane ownership
//     (lane t = word t of every plane, anyprec.cu:432-448); the next chunk's plane words are fetched (streaming,
//     L1 no-allocate) while the current chunk is computed;
//   * codebooks sit in per-warp shared-memory tables: the fast kernel's conflict-free PAIR tables for 2/3-bit (one
//     LDS = two weights), the plain 2^bits halfs for 4..8-bit; every lookup address is one PRMT byte insert into a
//     256-byte aligned base (bits <= 7; 8-bit needs base + 2*idx);
//   * the lane's 32 weights of each row are dequantised ONCE per chunk into 16 half2 registers and then used against
//     all M activation rows — the activations are read through L1 (same addresses for every warp of the SM) and each
//     128-bit load serves both rows;
//   * arithmetic like the fast kernel: fp16 HFMA2 chains of 8 -> fp32 accumulators (the reference is fp16 throughout).
#pragma once
#include "apgemv_common.cuh"
#include "apgemv_fast.cuh"

namespace apg {

__device__ __forceinline__ uint32_t ldg_stream_b32(const uint32_t *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
template <int IMM>
__device__ __forceinline__ uint32_t lds_u16_imm32(uint32_t addr) {  // zero-extended into a 32-bit register
    uint32_t v;
    asm volatile("{\n\t.reg .b16 t;\n\tld.shared.u16 t, [%1+%2];\n\tcvt.u32.u16 %0, t;\n\t}" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}

template <int BITS>
struct WideCfg {
    static constexpr int ROW_TBL_BYTES = BITS <= 4 ? FastCfg<(BITS <= 4 ? BITS : 2)>::ROW_TBL_BYTES : (2 << BITS);
};

// dq[4c + e2] = half2( w[k0(c) + 2 e2], w[k0(c) + 2 e2 + 1] ), k0(c) = i*1024 + c*8*eff + 8t — the order of the x registers
template <int BITS, int ROW_OFF>
struct WideDequant {  // plain table of 2^BITS halfs, BITS = 4..8
    __device__ __forceinline__ static void run(const uint32_t (&pw)[BITS], uint32_t tbl, uint32_t (&dq)[16]) {
        constexpr int SH = BITS <= 7 ? 1 : 0;  // bytes hold 2*index (a byte offset) when that fits
        // 8x8 bit-matrix transpose of every byte column (three butterfly stages, 4 instructions per row pair): row r
        // = the plane of index bit r - SH (plane j carries index bit BITS-1-j), zero rows elsewhere.  Afterwards
        // byte b of A[s] = (index << SH) of the weight at bit position 8b + s.
        uint32_t A[8];
#pragma unroll
        for (int r = 0; r < 8; r++) A[r] = (r >= SH && r - SH < BITS) ? pw[BITS - 1 - (r - SH) < 0 ? 0 : BITS - 1 - (r - SH)] : 0u;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t a = A[r], b = A[r + 4];
            A[r] = bitsel(b << 4, a, 0xF0F0F0F0u), A[r + 4] = bitsel(b, a >> 4, 0xF0F0F0F0u);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int r = (q & 1) + 4 * (q >> 1);  // 0, 1, 4, 5
            const uint32_t a = A[r], b = A[r + 2];
            A[r] = bitsel(b << 2, a, 0xCCCCCCCCu), A[r + 2] = bitsel(b, a >> 2, 0xCCCCCCCCu);
        }
#pragma unroll
        for (int r = 0; r < 8; r += 2) {
            const uint32_t a = A[r], b = A[r + 1];
            A[r] = bitsel(b << 1, a, 0xAAAAAAAAu), A[r + 1] = bitsel(b, a >> 1, 0xAAAAAAAAu);
        }
        uint32_t w[4][8];  // w[c][e], zero-extended halfs
#pragma unroll
        for (int sft = 0; sft < 8; sft++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {  // bit 8b + sft <-> k offset 31 - 8b - sft: c = 3 - b, e = 7 - sft
                if (BITS <= 7) {
                    w[3 - b][7 - sft] = lds_u16_imm32<ROW_OFF>(__byte_perm(A[sft], tbl, 0x7650u | b));
                } else {
                    const uint32_t idx = __byte_perm(A[sft], 0u, 0x4440u | b);
                    w[3 - b][7 - sft] = lds_u16_imm32<ROW_OFF>(tbl + 2u * idx);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int e2 = 0; e2 < 4; e2++) dq[4 * c + e2] = __byte_perm(w[c][2 * e2], w[c][2 * e2 + 1], 0x5410u);
    }
};

template <int ROW_OFF>
struct WideDequant<2, ROW_OFF> {  // 16-entry half2 pair table (Tables<2, RS>); index prep as in WordDot<2>
    __device__ __forceinline__ static void run(const uint32_t (&pw)[2], uint32_t tbl, uint32_t (&dq)[16]) {
        const uint32_t H = pw[0], L = pw[1];
        const uint32_t zh = bitsel(H, L >> 2, 0xCCCCCCCCu), zl = bitsel(H << 2, L, 0xCCCCCCCCu);
        const uint32_t a0 = (zh << 2) & 0x3C3C3C3Cu, a1 = (zh >> 2) & 0x3C3C3C3Cu;
        const uint32_t a2 = (zl << 2) & 0x3C3C3C3Cu, a3 = (zl >> 2) & 0x3C3C3C3Cu;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int xb = 4 * (3 - b);
            const uint32_t sel = 0x7650u | b;
            dq[xb + 0] = lds_b32_imm<ROW_OFF>(__byte_perm(a1, tbl, sel));
            dq[xb + 1] = lds_b32_imm<ROW_OFF>(__byte_perm(a3, tbl, sel));
            dq[xb + 2] = lds_b32_imm<ROW_OFF>(__byte_perm(a0, tbl, sel));
            dq[xb + 3] = lds_b32_imm<ROW_OFF>(__byte_perm(a2, tbl, sel));
        }
    }
};

template <int ROW_OFF>
struct WideDequant<3, ROW_OFF> {  // 64-entry half2 pair table (Tables<3, RS>); index prep as in WordDot<3>
    __device__ __forceinline__ static void run(const uint32_t (&pw)[3], uint32_t tbl, uint32_t (&dq)[16]) {
        const uint32_t P2 = pw[0], P1 = pw[1], P0 = pw[2];
        uint32_t t[4];
        t[0] = bitsel(P2 << 6, bitsel(P1 << 4, P0 << 2, 0x30303030u), 0xC0C0C0C0u) & 0xFCFCFCFCu;
        t[1] = bitsel(P2 << 4, bitsel(P1 << 2, P0, 0x30303030u), 0xC0C0C0C0u) & 0xFCFCFCFCu;
        t[2] = bitsel(P2 << 2, bitsel(P1, P0 >> 2, 0x30303030u), 0xC0C0C0C0u) & 0xFCFCFCFCu;
        t[3] = bitsel(P2, bitsel(P1 >> 2, P0 >> 4, 0x30303030u), 0xC0C0C0C0u) & 0xFCFCFCFCu;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int xb = 4 * (3 - b);
            const uint32_t sel = 0x7650u | b;
#pragma unroll
            for (int e2 = 0; e2 < 4; e2++) dq[xb + e2] = lds_b32_imm<ROW_OFF>(__byte_perm(t[3 - e2], tbl, sel));
        }
    }
};

template <int BITS, int MB, int R>
__global__ void __launch_bounds__(128) gemv_wide_kernel(const __half *__restrict__ x, const uint32_t *__restrict__ W,
                                                        const __half *__restrict__ lut, __half *__restrict__ out,
                                                        float *__restrict__ partial, uint32_t M, uint32_t N, uint32_t K) {
    constexpr int RTB = WideCfg<BITS>::ROW_TBL_BYTES;
    constexpr int WTB = (R * RTB + 255) / 256 * 256;
    __shared__ __align__(1024) uint8_t tb[4 * WTB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t row0 = (blockIdx.x * 4u + warp) * R;
    if (row0 >= N) return;  // whole warps leave; nothing below synchronises across warps
    uint32_t rowc[R];
#pragma unroll
    for (int r = 0; r < R; r++) rowc[r] = min(row0 + r, N - 1u);
    const uint32_t tbl = smem_u32(tb) + warp * WTB;

    if constexpr (BITS <= 3) {
        typename Tables<BITS, R>::Regs lr;
        Tables<BITS, R>::fetch(lr, lut, row0, N, lane);
        Tables<BITS, R>::store(lr, tbl, lane);
    } else {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint4 *src = reinterpret_cast<const uint4 *>(lut + ((size_t)rowc[r] << BITS));
            for (int e = lane; e < (2 << BITS) / 16; e += 32) {
                const uint4 v = __ldg(src + e);
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tbl + r * RTB + e * 16), "r"(v.x), "r"(v.y),
                             "r"(v.z), "r"(v.w)
                             : "memory");
            }
        }
    }
    __syncwarp();

    const uint32_t words = K >> 5, nchunk = (K + 1023u) >> 10;
    float acc[R][MB];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int m = 0; m < MB; m++) acc[r][m] = 0.f;

    uint32_t cur[R][BITS], nxt[R][BITS];
    auto load_planes = [&](uint32_t i, uint32_t (&dst)[R][BITS]) {
        if ((uint32_t)lane < chunk_eff(K, i)) {
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int j = 0; j < BITS; j++)
                    dst[r][j] = ldg_stream_b32(W + ((size_t)j * N + rowc[r]) * words + i * 32u + lane);
        }
    };
    load_planes(0, cur);
    for (uint32_t i = 0; i < nchunk; i++) {
        if (i + 1 < nchunk) load_planes(i + 1, nxt);
        const uint32_t eff = chunk_eff(K, i);
        if ((uint32_t)lane < eff) {  // tail chunk: lanes >= eff own nothing (anyprec.cu:433-436)
            uint32_t dq[R][16];
            WideDequant<BITS, 0>::run(cur[0], tbl, dq[0]);
            if constexpr (R == 2) WideDequant<BITS, RTB>::run(cur[R - 1], tbl, dq[R - 1]);
#pragma unroll
            for (int m = 0; m < MB; m++) {
                if ((uint32_t)m < M) {
                    const __half *xl = x + (size_t)m * K + i * 1024u + 8u * lane;
                    uint32_t s[R][2];
#pragma unroll
                    for (int r = 0; r < R; r++) s[r][0] = 0u, s[r][1] = 0u;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const uint4 xv = __ldg(reinterpret_cast<const uint4 *>(xl + c * 8u * eff));
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            s[r][0] = hfma2_u32(dq[r][4 * c + 0], xv.x, s[r][0]);
                            s[r][1] = hfma2_u32(dq[r][4 * c + 1], xv.y, s[r][1]);
                            s[r][0] = hfma2_u32(dq[r][4 * c + 2], xv.z, s[r][0]);
                            s[r][1] = hfma2_u32(dq[r][4 * c + 3], xv.w, s[r][1]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; r++) acc[r][m] = acc_add_h2(acc[r][m], hadd2_u32(s[r][0], s[r][1]));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int j = 0; j < BITS; j++) cur[r][j] = nxt[r][j];
    }

#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int m = 0; m < MB; m++) {
            float v = acc[r][m];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && row0 + r < N && (uint32_t)m < M) {
                if (out) out[(size_t)m * N + row0 + r] = __float2half_rn(v);
                if (partial) partial[(size_t)m * N + row0 + r] = v;
            }
        }
}

// W[n, k] = lut[n, idx[n, k]] (replaces dequant_kbit_store, anyprec.cu:294-359) on the wide kernel's machinery: one warp
// per row, shared-memory codebook table, butterfly / pair-table index extraction; the 16 half2 registers of a chunk
// are exactly four contiguous 16-byte pieces of the output row (k0(c) = i*1024 + c*8*eff + 8t), so every warp store is
// a fully coalesced 512 B.  A pure gather of the codebook's bit patterns: bit-identical to the reference by construction.
template <int BITS>
__global__ void __launch_bounds__(128) dequant_wide_kernel(const uint32_t *__restrict__ W, const __half *__restrict__ lut,
                                                           __half *__restrict__ O, uint32_t N, uint32_t K) {
    constexpr int RTB = WideCfg<BITS>::ROW_TBL_BYTES;
    constexpr int WTB = (RTB + 255) / 256 * 256;
    __shared__ __align__(1024) uint8_t tb[4 * WTB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t row = blockIdx.x * 4u + warp;
    if (row >= N) return;
    const uint32_t tbl = smem_u32(tb) + warp * WTB;
    if constexpr (BITS <= 3) {
        typename Tables<BITS, 1>::Regs lr;
        Tables<BITS, 1>::fetch(lr, lut, row, N, lane);
        Tables<BITS, 1>::store(lr, tbl, lane);
    } else {
        const uint4 *src = reinterpret_cast<const uint4 *>(lut + ((size_t)row << BITS));
        for (int e = lane; e < (2 << BITS) / 16; e += 32) {
            const uint4 v = __ldg(src + e);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tbl + e * 16), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                         : "memory");
        }
    }
    __syncwarp();
    const uint32_t words = K >> 5, nchunk = (K + 1023u) >> 10;
    uint32_t cur[BITS], nxt[BITS];
    auto load_planes = [&](uint32_t i, uint32_t (&dst)[BITS]) {
        if ((uint32_t)lane < chunk_eff(K, i)) {
#pragma unroll
            for (int j = 0; j < BITS; j++) dst[j] = ldg_stream_b32(W + ((size_t)j * N + row) * words + i * 32u + lane);
        }
    };
    load_planes(0, cur);
    for (uint32_t i = 0; i < nchunk; i++) {
        if (i + 1 < nchunk) load_planes(i + 1, nxt);
        const uint32_t eff = chunk_eff(K, i);
        if ((uint32_t)lane < eff) {
            uint32_t dq[16];
            WideDequant<BITS, 0>::run(cur, tbl, dq);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                __half *dst = O + (size_t)row * K + i * 1024u + c * 8u * eff + 8u * lane;
                asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(dq[4 * c + 0]), "r"(dq[4 * c + 1]),
                             "r"(dq[4 * c + 2]), "r"(dq[4 * c + 3])
                             : "memory");
            }
        }
#pragma unroll
        for (int j = 0; j < BITS; j++) cur[j] = nxt[j];
    }
}

}  // namespace apg
