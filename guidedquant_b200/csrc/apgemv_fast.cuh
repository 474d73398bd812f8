// Fast Any-Precision LUT GEMV for M = 1, bits in {2,3,4}, K % 128 == 0, K <= 32768  (sm_100a).
//
// Replaces matmul_kbit_32<1,bits,false> (reference inference/ap_gemv/anyprec.cu:372-542) with a
// different work decomposition:
//   * 128-bit coalesced loads of the bit-planes: lane L of a warp owns words 4L..4L+3 of a
//     128-word "slab" (4096 k) of a (row, plane) -> one LDG.128 per plane = 128 weights.
//   * the lane<->k map is row-independent, so the lane's 128 activations live in
//     64 half2 REGISTERS for the whole kernel (the reference re-reads x from L1/L2 for every row).
//   * per-row codebooks are expanded once per row-batch into conflict-free shared-memory tables
//     (<= 32 banks for 2/4-bit): 2-bit -> 16-entry pair table (one LDS = two weights),
//     3-bit -> 64-entry pair table, 4-bit -> 16-entry scalar table.
//   * table byte-offsets are produced four at a time (masked words) and turned into complete LDS
//     addresses by ONE PRMT each (byte insert into a 256-byte-aligned per-warp table base).
//   * fp16 HFMA2 chains of length 8 feed fp32 accumulators (the reference accumulates everything
//     in fp16); rows are reduced RB at a time with a transposing shuffle tree.
//   * K > 4096: the row's slabs are spread over the warps of a group (each warp keeps ITS slab's x
//     in registers) and combined through shared memory.
#pragma once
#include "apgemv_common.cuh"

namespace apg {

template <int BITS>
struct FastCfg;
template <>
struct FastCfg<2> {
    static constexpr int RB = 4;              // rows per batch (8 LDG.128 in flight per lane)
    static constexpr int ROW_TBL_BYTES = 64;  // 16 x half2
};
template <>
struct FastCfg<3> {
    static constexpr int RB = 4;               // 12 LDG.128 in flight per lane
    static constexpr int ROW_TBL_BYTES = 256;  // 64 x half2
};
template <>
struct FastCfg<4> {
    static constexpr int RB = 2;              // 8 LDG.128 in flight per lane
    static constexpr int ROW_TBL_BYTES = 32;  // 16 x half
};
template <int BITS>
struct FastWarpTbl {
    static constexpr int BYTES = (FastCfg<BITS>::RB * FastCfg<BITS>::ROW_TBL_BYTES + 255) / 256 * 256;
};

// ---------------------------------------------------------------------------------------------
// Codebook staging: expand lut rows [row0, row0+RB) into this warp's shared-memory tables.
// ---------------------------------------------------------------------------------------------
template <int BITS>
__device__ __forceinline__ void stage_tables(uint32_t tbl_addr, const __half *__restrict__ lut, uint32_t row0,
                                             uint32_t N, int lane);

// 2-bit pair table: entry p = (hA hB lA lB) -> half2( C[2hA+lA], C[2hB+lB] ); A = even k (low half).
template <>
__device__ __forceinline__ void stage_tables<2>(uint32_t tbl_addr, const __half *__restrict__ lut, uint32_t row0,
                                                uint32_t N, int lane) {
    const int p = lane & 15;
    const uint32_t a = ((p >> 3) & 1) * 2 + ((p >> 1) & 1), b = ((p >> 2) & 1) * 2 + (p & 1);
    const uint32_t sel = (2 * a) | ((2 * a + 1) << 4) | ((2 * b) << 8) | ((2 * b + 1) << 12);
#pragma unroll
    for (int it = 0; it < FastCfg<2>::RB / 2; it++) {
        const int r = it * 2 + (lane >> 4);
        const uint32_t row = min(row0 + r, N - 1);
        const uint2 c = __ldg(reinterpret_cast<const uint2 *>(lut + (size_t)row * 4));
        const uint32_t e = __byte_perm(c.x, c.y, sel);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tbl_addr + r * 64 + p * 4), "r"(e));
    }
}

// 3-bit pair table: entry p = (a2 b2 a1 b1 a0 b0) -> half2( C[a], C[b] ); A = even k.
template <>
__device__ __forceinline__ void stage_tables<3>(uint32_t tbl_addr, const __half *__restrict__ lut, uint32_t row0,
                                                uint32_t N, int lane) {
    const unsigned short *l16 = reinterpret_cast<const unsigned short *>(lut);
#pragma unroll
    for (int r = 0; r < FastCfg<3>::RB; r++) {
        const uint32_t row = min(row0 + r, N - 1);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int p = lane + 32 * h;
            const uint32_t a = ((p >> 5) & 1) * 4 + ((p >> 3) & 1) * 2 + ((p >> 1) & 1);
            const uint32_t b = ((p >> 4) & 1) * 4 + ((p >> 2) & 1) * 2 + (p & 1);
            const uint32_t ca = __ldg(l16 + (size_t)row * 8 + a), cb = __ldg(l16 + (size_t)row * 8 + b);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(tbl_addr + r * 256 + p * 4), "r"(ca | (cb << 16)));
        }
    }
}

// 4-bit scalar table: 16 halfs per row, copied verbatim.
template <>
__device__ __forceinline__ void stage_tables<4>(uint32_t tbl_addr, const __half *__restrict__ lut, uint32_t row0,
                                                uint32_t N, int lane) {
    const int r = lane >> 3, w = lane & 7;
    if (r < FastCfg<4>::RB) {
        const uint32_t row = min(row0 + r, N - 1);
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(lut + (size_t)row * 16) + w);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tbl_addr + r * 32 + w * 4), "r"(v));
    }
}

// ---------------------------------------------------------------------------------------------
// One row x one slab for one lane: 128 weights (BITS LDG.128 worth) against the lane's 64 half2 x.
// xr[(4c+q)*4 + e2] = half2( x[kb(c)+8q+2e2], x[kb(c)+8q+2e2+1] ).  ROW_OFF = r * ROW_TBL_BYTES.
// ---------------------------------------------------------------------------------------------
template <int BITS, int ROW_OFF>
struct RowDot;

template <int ROW_OFF>
struct RowDot<2, ROW_OFF> {
    __device__ __forceinline__ static float run(const uint4 (&pl)[2], const uint32_t (&xr)[64], uint32_t tbl_base) {
        const uint32_t Hq[4] = {pl[0].x, pl[0].y, pl[0].z, pl[0].w};
        const uint32_t Lq[4] = {pl[1].x, pl[1].y, pl[1].z, pl[1].w};
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t H = Hq[q], L = Lq[q];
            // nibble m of zh = (H[4m+3] H[4m+2] L[4m+3] L[4m+2]); of zl = (H[4m+1] H[4m] L[4m+1] L[4m])
            const uint32_t zh = bitsel(H, L >> 2, 0xCCCCCCCCu);
            const uint32_t zl = bitsel(H << 2, L, 0xCCCCCCCCu);
            // byte b of each word = 4 * pair-index (a complete table byte offset)
            const uint32_t a0 = (zh << 2) & 0x3C3C3C3Cu;  // even nibbles of zh -> e2 = 2
            const uint32_t a1 = (zh >> 2) & 0x3C3C3C3Cu;  // odd  nibbles of zh -> e2 = 0
            const uint32_t a2 = (zl << 2) & 0x3C3C3C3Cu;  // even nibbles of zl -> e2 = 3
            const uint32_t a3 = (zl >> 2) & 0x3C3C3C3Cu;  // odd  nibbles of zl -> e2 = 1
            uint32_t s0 = 0u, s1 = 0u;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int c = 3 - b, xb = (4 * c + q) * 4;
                const uint32_t sel = 0x7650u | b;
                const uint32_t w0 = lds_b32_imm<ROW_OFF>(__byte_perm(a1, tbl_base, sel));
                const uint32_t w1 = lds_b32_imm<ROW_OFF>(__byte_perm(a3, tbl_base, sel));
                const uint32_t w2 = lds_b32_imm<ROW_OFF>(__byte_perm(a0, tbl_base, sel));
                const uint32_t w3 = lds_b32_imm<ROW_OFF>(__byte_perm(a2, tbl_base, sel));
                s0 = hfma2_u32(w0, xr[xb + 0], s0);
                s1 = hfma2_u32(w1, xr[xb + 1], s1);
                s0 = hfma2_u32(w2, xr[xb + 2], s0);
                s1 = hfma2_u32(w3, xr[xb + 3], s1);
            }
            acc += h2_sum_f32(hadd2_u32(s0, s1));
        }
        return acc;
    }
};

template <int ROW_OFF>
struct RowDot<3, ROW_OFF> {
    __device__ __forceinline__ static float run(const uint4 (&pl)[3], const uint32_t (&xr)[64], uint32_t tbl_base) {
        const uint32_t P2q[4] = {pl[0].x, pl[0].y, pl[0].z, pl[0].w};
        const uint32_t P1q[4] = {pl[1].x, pl[1].y, pl[1].z, pl[1].w};
        const uint32_t P0q[4] = {pl[2].x, pl[2].y, pl[2].z, pl[2].w};
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t P2 = P2q[q], P1 = P1q[q], P0 = P0q[q];
            // target j: byte b = (P2[p+1] P2[p] P1[p+1] P1[p] P0[p+1] P0[p] 0 0), p = 8b+2j -> x pair e2 = 3-j
            uint32_t t[4];
            t[0] = bitsel(P2 << 6, bitsel(P1 << 4, P0 << 2, 0x30303030u), 0xC0C0C0C0u) & 0xFCFCFCFCu;
            t[1] = bitsel(P2 << 4, bitsel(P1 << 2, P0, 0x30303030u), 0xC0C0C0C0u) & 0xFCFCFCFCu;
            t[2] = bitsel(P2 << 2, bitsel(P1, P0 >> 2, 0x30303030u), 0xC0C0C0C0u) & 0xFCFCFCFCu;
            t[3] = bitsel(P2, bitsel(P1 >> 2, P0 >> 4, 0x30303030u), 0xC0C0C0C0u) & 0xFCFCFCFCu;
            uint32_t s0 = 0u, s1 = 0u;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int c = 3 - b, xb = (4 * c + q) * 4;
                const uint32_t sel = 0x7650u | b;
                const uint32_t w0 = lds_b32_imm<ROW_OFF>(__byte_perm(t[3], tbl_base, sel));
                const uint32_t w1 = lds_b32_imm<ROW_OFF>(__byte_perm(t[2], tbl_base, sel));
                const uint32_t w2 = lds_b32_imm<ROW_OFF>(__byte_perm(t[1], tbl_base, sel));
                const uint32_t w3 = lds_b32_imm<ROW_OFF>(__byte_perm(t[0], tbl_base, sel));
                s0 = hfma2_u32(w0, xr[xb + 0], s0);
                s1 = hfma2_u32(w1, xr[xb + 1], s1);
                s0 = hfma2_u32(w2, xr[xb + 2], s0);
                s1 = hfma2_u32(w3, xr[xb + 3], s1);
            }
            acc += h2_sum_f32(hadd2_u32(s0, s1));
        }
        return acc;
    }
};

template <int ROW_OFF>
struct RowDot<4, ROW_OFF> {
    __device__ __forceinline__ static float run(const uint4 (&pl)[4], const uint32_t (&xr)[64], uint32_t tbl_base) {
        const uint32_t P3q[4] = {pl[0].x, pl[0].y, pl[0].z, pl[0].w};
        const uint32_t P2q[4] = {pl[1].x, pl[1].y, pl[1].z, pl[1].w};
        const uint32_t P1q[4] = {pl[2].x, pl[2].y, pl[2].z, pl[2].w};
        const uint32_t P0q[4] = {pl[3].x, pl[3].y, pl[3].z, pl[3].w};
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t P3 = P3q[q], P2 = P2q[q], P1 = P1q[q], P0 = P0q[q];
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
                // byte b = 2*index.  ylo: bit position 8b+sft   -> k offset 31-8b-sft -> c = 3-b, e = 7-sft
                //                    yhi: bit position 8b+4+sft -> k offset 27-8b-sft -> c = 3-b, e = 3-sft
                const uint32_t ylo = (y << 1) & 0x1E1E1E1Eu, yhi = (y >> 3) & 0x1E1E1E1Eu;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int c = 3 - b, xb = (4 * c + q) * 4;
                    const uint32_t sel = 0x7650u | b;
                    const int e_lo = 7 - sft, e_hi = 3 - sft;
                    const __half wl = __ushort_as_half((unsigned short)lds_u16_imm<ROW_OFF>(__byte_perm(ylo, tbl_base, sel)));
                    const __half wh = __ushort_as_half((unsigned short)lds_u16_imm<ROW_OFF>(__byte_perm(yhi, tbl_base, sel)));
                    const __half2 xl = *reinterpret_cast<const __half2 *>(&xr[xb + e_lo / 2]);
                    const __half2 xh = *reinterpret_cast<const __half2 *>(&xr[xb + e_hi / 2]);
                    s[sft] = __hfma(wl, (e_lo & 1) ? __high2half(xl) : __low2half(xl), s[sft]);
                    s[(sft + 2) & 3] = __hfma(wh, (e_hi & 1) ? __high2half(xh) : __low2half(xh), s[(sft + 2) & 3]);
                }
            }
            acc += (__half2float(s[0]) + __half2float(s[1])) + (__half2float(s[2]) + __half2float(s[3]));
        }
        return acc;
    }
};

// ---------------------------------------------------------------------------------------------
// Kernel
// ---------------------------------------------------------------------------------------------
struct FastParams {
    const __half *x;      // [K]
    const uint4 *W;       // [bits][N][K/128] as uint4
    const __half *lut;    // [N][2^bits]
    __half *out;          // [N] or nullptr
    float *partial;       // [N] or nullptr
    uint32_t N, K;
    uint32_t nslab;       // ceil(K/4096)
    uint32_t groups;      // row groups per CTA (warps per CTA = groups * nslab)
};

// Shared memory: [ per-warp tables (256B aligned) | x slabs, swizzled 16B units | cross-slab reduction ]
template <int BITS>
__global__ void __launch_bounds__(256) gemv_fast_kernel(const FastParams p) {
    constexpr int RB = FastCfg<BITS>::RB;
    constexpr int ROWB = FastCfg<BITS>::ROW_TBL_BYTES;
    constexpr int WTB = FastWarpTbl<BITS>::BYTES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const uint32_t nslab = p.nslab;
    const uint32_t g = warp / nslab, s = warp - g * nslab;
    const uint32_t words = p.K >> 5;          // words per (row, plane)
    const uint32_t vecs = words >> 2;         // uint4 per (row, plane)

    const uint32_t smem_base = (smem_u32(smem_raw) + 255u) & ~255u;
    const uint32_t tbl_base = smem_base + warp * WTB;
    const uint32_t x_base = smem_base + nwarps * WTB;
    float *red = reinterpret_cast<float *>(smem_raw + (x_base - smem_u32(smem_raw)) + nslab * 8192u);

    // rows of this group
    const uint32_t total_groups = gridDim.x * p.groups;
    const uint32_t gid = blockIdx.x * p.groups + g;
    const uint32_t r_begin = (uint32_t)(((uint64_t)p.N * gid) / total_groups);
    const uint32_t r_end = (uint32_t)(((uint64_t)p.N * (gid + 1)) / total_groups);

    const uint32_t v0 = s * 32u + lane;       // this lane's uint4 index inside a (row, plane)
    const bool active = v0 < vecs;

    // ---- independent of x: first batch of bit-planes + codebooks (overlaps the previous kernel under PDL)
    uint4 pl[RB][BITS];
    auto load_batch = [&](uint32_t row0) {
#pragma unroll
        for (int r = 0; r < RB; r++) {
            const uint32_t row = min(row0 + r, p.N - 1);
#pragma unroll
            for (int j = 0; j < BITS; j++)
                if (active) pl[r][j] = ldg_stream_v4(p.W + ((size_t)j * p.N + row) * vecs + v0);
        }
    };
    if (r_begin < r_end) {
        load_batch(r_begin);
        stage_tables<BITS>(tbl_base, p.lut, r_begin, p.N, lane);
    }

    // ---- x: global -> swizzled smem units -> registers
    pdl_wait_prior_grid();
    {
        const uint32_t units = p.K >> 3;  // 16-byte units of x
        const uint32_t full = p.K >> 10;
        const uint32_t eff_tail = (p.K & 1023u) >> 5;
        for (uint32_t u = threadIdx.x; u < units; u += blockDim.x) {
            const uint32_t k = u << 3;
            const uint32_t i = k >> 10, r = k & 1023u;
            const uint32_t eff = (i < full) ? 32u : eff_tail;
            const uint32_t c = r / (8u * eff), t = (r - c * 8u * eff) >> 3;
            const uint32_t w = i * 32u + t;
            const uint32_t sl = w >> 7, L = (w & 127u) >> 2, q = w & 3u;
            const uint32_t j = 4u * c + q;
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p.x) + u);
            sts_v4(x_base + ((sl * 32u + L) * 16u + (j ^ (L & 7u))) * 16u, v);
        }
    }
    __syncthreads();
    pdl_launch_dependents();

    uint32_t xr[64];
    if (active) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const uint4 v = lds_v4(x_base + ((s * 32u + lane) * 16u + (j ^ (lane & 7))) * 16u);
            xr[4 * j + 0] = v.x, xr[4 * j + 1] = v.y, xr[4 * j + 2] = v.z, xr[4 * j + 3] = v.w;
        }
    }

    // ---- row batches
    uint32_t parity = 0;
    for (uint32_t row0 = r_begin; row0 < r_end; row0 += RB) {
        if (row0 != r_begin) {
            __syncwarp();
            load_batch(row0);
            stage_tables<BITS>(tbl_base, p.lut, row0, p.N, lane);
        }
        __syncwarp();

        float sums[RB];
#pragma unroll
        for (int r = 0; r < RB; r++) sums[r] = 0.f;
        if (active) {
            sums[0] = RowDot<BITS, 0 * ROWB>::run(pl[0], xr, tbl_base);
            if (RB > 1) sums[1 % RB] = RowDot<BITS, (1 % RB) * ROWB>::run(pl[1 % RB], xr, tbl_base);
            if (RB > 2) sums[2 % RB] = RowDot<BITS, (2 % RB) * ROWB>::run(pl[2 % RB], xr, tbl_base);
            if (RB > 3) sums[3 % RB] = RowDot<BITS, (3 % RB) * ROWB>::run(pl[3 % RB], xr, tbl_base);
        }
        float v = batch_reduce<RB>(sums, lane);
        const int rl = batch_row_of_lane<RB>(lane);
        const bool writer = batch_lane_is_writer<RB>(lane);
        const uint32_t row = row0 + rl;
        if (nslab > 1) {
            float *rbuf = red + ((g * 2 + parity) * nslab) * RB;
            if (writer) rbuf[s * RB + rl] = v;
            asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(nslab * 32u));
            if (s == 0 && writer) {
                v = rbuf[rl];
                for (uint32_t ss = 1; ss < nslab; ss++) v += rbuf[ss * RB + rl];
            }
            parity ^= 1;
        }
        if (s == 0 && writer && row < r_end) {
            if (p.out) p.out[row] = __float2half_rn(v);
            if (p.partial) p.partial[row] = v;
        }
    }
}

template <int BITS>
inline size_t fast_smem_bytes(uint32_t nslab, uint32_t groups) {
    const size_t nwarps = (size_t)nslab * groups;
    return 256 + nwarps * FastWarpTbl<BITS>::BYTES + (size_t)nslab * 8192 +
           (nslab > 1 ? groups * 2 * nslab * FastCfg<BITS>::RB * sizeof(float) : 0) + 16;
}

}  // namespace apg
