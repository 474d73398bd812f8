// Prefill path of an Any-Precision Linear: fused codebook dequantisation + tensor-core GEMM (sm_100a).
//
// Replaces the reference's two-kernel prefill (inference/ap_gemv/APLinear.py:35-38: `anyprec_dequant` writes the whole
// fp16 [N, K] weight to HBM — dequant_kbit_store, anyprec.cu:294-359 — and cuBLAS reads it back) for more than 8 tokens:
//
//     Y[t, n] = sum_k X[t, k] * lut[n, idx[n, k]]            X fp16 [T, K], Y fp16 [T, N], fp32 accumulation
//
// Machine mapping (one CTA = 128 weight rows x T_TILE tokens x a range of K):
//   * "swap-AB": the 128 weight rows are the M side of tcgen05.mma (cta_group::1, kind::f16, M = 128), the tokens are the
//     N side (N = T_TILE, 32..256), so a decode-sized token tile still uses the full 128-lane datapath.  The fp32
//     accumulator tile [128 lanes x T_TILE columns] lives in TMEM.
//   * A operand = dequantised weights, never in HBM and never in shared memory: 8 warps (two threads per weight row) read
//     the packed bit-plane words of their row straight from global memory (16 B per plane per 256 k), turn them into table
//     indices with the same LOP3 networks as the GEMV kernels, look the fp16 pairs up in bank-striped shared-memory tables
//     (one LDS = two weights, conflict-free: entry-major, lane-minor) and write them with tcgen05.st into a ring of A tiles
//     IN TMEM (tcgen05.mma with the A operand in tensor memory: lane = weight row, one 32-bit column = two consecutive k —
//     exactly the thread-per-row order the dequantiser produces, so there is no swizzle, no st.shared, no proxy fence, and
//     the MMA reads only the token tile from shared memory).  The packed layout makes this cheap: byte c of word t of a
//     chunk holds 8 CONSECUTIVE k (k = i*1024 + c*8*eff + 8t + e, pack.py:58-75), so a K block of 64 consecutive k is
//     "byte c of 8 consecutive words" and the token tile needs no permutation at all.
//   * B operand = the token tile [T_TILE x 64] of X, loaded by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) by one thread;
//     rows past T are zero-filled by the TMA unit.
//   * One elected thread issues the MMAs (4 x K=16 per stage) and releases stages with tcgen05.commit; a ring of
//     (full, empty) mbarrier pairs decouples the three roles.
//   * Epilogue: tcgen05.ld 32x32b (thread = weight row, 16 tokens per load) -> fp16 stores, or fp32 partial slabs when K is
//     split over CTAs (small T x N: deterministic split-K, summed in a fixed order by reduce_splits_kernel).
#pragma once
#include <cuda.h>

#include "apgemv_common.cuh"

namespace apg {
namespace ptc {

constexpr int ROWS = 128;                     // weight rows per CTA (UMMA M)
constexpr int BK = 64;                        // k per stage: one 128-byte swizzled row of fp16
constexpr int DQ_WARPS = 8;                   // dequantising + epilogue warps (2 threads per weight row)
constexpr int DQ_THREADS = DQ_WARPS * 32;
constexpr int X_WARP = DQ_WARPS;              // TMA producer of the token tiles
constexpr int MMA_WARP = DQ_WARPS + 1;        // MMA issuer, owns the TMEM allocation
constexpr int THREADS = (DQ_WARPS + 2) * 32;
constexpr int MAX_STAGES = 8;
constexpr uint32_t CTRL_BYTES = 256;          // barriers + TMEM base, at the very start of dynamic shared memory
constexpr uint32_t TBL_ABS = 2048;            // shared-WINDOW-relative address of the lookup tables: an LDS immediate
constexpr uint32_t A_COLS = BK / 2;            // TMEM columns of one A stage (two fp16 per 32-bit column)
constexpr uint32_t SMEM_LIMIT = 227u * 1024u;
constexpr long long WATCHDOG_CYCLES = 6000000000ll;  // ~3 s: a wait that long is a bug; trap instead of hanging the GPU

// entries are 256 B apart (two warps x 32 lanes x 4 B); index byte = entry | warp-pair << ENTRY_BITS
template <int BITS>
struct Cfg;
template <>
struct Cfg<2> {
    static constexpr int ENTRY_BITS = 4;  // pair table: (hA hB lA lB) -> half2
};
template <>
struct Cfg<3> {
    static constexpr int ENTRY_BITS = 6;  // pair table: (a2 b2 a1 b1 a0 b0) -> half2
};
template <>
struct Cfg<4> {
    static constexpr int ENTRY_BITS = 4;  // single table: 4-bit index -> half (low 16 bits)
};
template <int BITS>
struct Lay {
    static constexpr uint32_t TBL_BYTES = 2u * (1u << Cfg<BITS>::ENTRY_BITS) * 256u;
    static constexpr uint32_t STAGE_BASE = (TBL_ABS + TBL_BYTES + 1023u) & ~1023u;  // absolute shared address of token tile 0
};

struct Params {
    const uint32_t *W;   // [bits][N][K/32]
    const __half *lut;   // [N][2^bits]
    __half *out;         // [T][N]                 (splits == 1)
    float *partial;      // [splits][T][N] fp32    (splits > 1)
    uint32_t N, K, T;
    uint32_t t_tile;     // tokens per CTA: multiple of 32, <= 256  (UMMA N)
    uint32_t stages;     // ring depth
    uint32_t splits;     // K split over gridDim.z
    uint32_t sb_total;   // K / 256 "super blocks" (8 words per row and plane = 4 stages)
    uint32_t tmem_cols;  // power of two >= t_tile + stages * A_COLS: accumulator columns [0, t_tile), then the A ring
    uint32_t idesc;      // tcgen05 instruction descriptor
};

// ------------------------------------------------------------------------------------------------ primitives
__device__ __forceinline__ void mb_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mb_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mb_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mb_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}
__device__ __noinline__ void mb_spin(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mb_try(bar, parity)) {
        if (clock64() - t0 > WATCHDOG_CYCLES) {
            printf("prefill_tc watchdog: cta (%d,%d,%d) thread %d bar %u parity %u\n", blockIdx.x, blockIdx.y, blockIdx.z,
                   threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void mb_wait(uint32_t bar, uint32_t parity) {
    if (!mb_try(bar, parity)) mb_spin(bar, parity);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t c0, uint32_t c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1),
// descriptor version 1 (sm_100), layout type 2 = 128-byte swizzle.  Stepping K by 16 elements = +32 B on the start address.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x 8 columns (16 fp16 of k per row), B described by `db`
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane (warp-collective; taddr = lane-quadrant base | column)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint4 &a, const uint4 &b, const uint4 &c, const uint4 &d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w),
                 "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void sts_b32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// (a & m) | c in one LOP3
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t m, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(m), "r"(c));
    return d;
}
// table word of index byte B of `idx`: address = lanebase + TBL_ABS + index * 256 (lanebase = window base + a lane offset < 256:
// its byte 1 is zero), formed by ONE PRMT
template <int B>
__device__ __forceinline__ uint32_t lookup(uint32_t idx, uint32_t lanebase) {
    return lds_b32_imm<(int)TBL_ABS>(__byte_perm(idx, lanebase, 0x7604u | (B << 4)));
}
template <int S>
__device__ __forceinline__ uint32_t shl(uint32_t v) {  // shift left by S (right when negative)
    if constexpr (S >= 0)
        return v << S;
    else
        return v >> (-S);
}
__device__ __forceinline__ uint32_t comp(const uint4 &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// ------------------------------------------------------------------------------------------------ index networks
// Net<BITS>::run turns one word per plane (32 weights of one row) into NREG registers of table-index bytes; byte B of those
// registers belongs to byte B of the packed word, i.e. to c = 3 - B (bit 31 - (8c+e)).  Net<BITS>::gather<B> looks up
// the 8 weights e = 0..7 of that byte: four half2 (even e in the low half) = four consecutive A columns of the row.
template <int BITS>
struct Net;

template <>
struct Net<2> {
    static constexpr int NREG = 4;
    __device__ __forceinline__ static void run(uint32_t (&o)[4], const uint32_t (&pw)[2], uint32_t wc) {
        const uint32_t H = pw[0], L = pw[1];
        // nibble m of zh = (H[4m+3] H[4m+2] L[4m+3] L[4m+2]),  of zl = (H[4m+1] H[4m] L[4m+1] L[4m])
        const uint32_t zh = bitsel(H, L >> 2, 0xCCCCCCCCu), zl = bitsel(H << 2, L, 0xCCCCCCCCu);
        o[0] = and_or(zh >> 4, 0x0F0F0F0Fu, wc);  // byte bits 7,6 -> e = 0,1
        o[1] = and_or(zl >> 4, 0x0F0F0F0Fu, wc);  //           5,4 -> e = 2,3
        o[2] = and_or(zh, 0x0F0F0F0Fu, wc);       //           3,2 -> e = 4,5
        o[3] = and_or(zl, 0x0F0F0F0Fu, wc);       //           1,0 -> e = 6,7
    }
    template <int B>
    __device__ __forceinline__ static uint4 gather(const uint32_t (&o)[4], uint32_t lb) {
        return make_uint4(lookup<B>(o[0], lb), lookup<B>(o[1], lb), lookup<B>(o[2], lb), lookup<B>(o[3], lb));
    }
    // entry p = (hA hB lA lB) -> half2(C[2hA+lA], C[2hB+lB]); A = the higher bit position = the even e
    __device__ __forceinline__ static void build(const __half *lut_row, uint32_t tbl, uint32_t half_id) {
        const uint2 c = __ldg(reinterpret_cast<const uint2 *>(lut_row));
#pragma unroll
        for (int p = 0; p < 16; p++) {
            const uint32_t a = ((p >> 3) & 1) * 2 + ((p >> 1) & 1), b = ((p >> 2) & 1) * 2 + (p & 1);
            const uint32_t sel = (2 * a) | ((2 * a + 1) << 4) | ((2 * b) << 8) | ((2 * b + 1) << 12);
            if ((uint32_t)(p >> 3) == half_id) sts_b32(tbl + p * 256, __byte_perm(c.x, c.y, sel));
        }
    }
};

template <>
struct Net<3> {
    static constexpr int NREG = 4;
    template <int J>
    __device__ __forceinline__ static uint32_t one(uint32_t P2, uint32_t P1, uint32_t P0, uint32_t wc) {
        // byte b = (P2[q+1] P2[q] P1[q+1] P1[q] P0[q+1] P0[q]), q = 8b + 2J
        const uint32_t t = bitsel(shl<4 - 2 * J>(P2), bitsel(shl<2 - 2 * J>(P1), shl<-2 * J>(P0), 0x0C0C0C0Cu), 0x30303030u);
        return and_or(t, 0x3F3F3F3Fu, wc);
    }
    __device__ __forceinline__ static void run(uint32_t (&o)[4], const uint32_t (&pw)[3], uint32_t wc) {
        o[0] = one<3>(pw[0], pw[1], pw[2], wc);  // byte bits 7,6 -> e = 0,1
        o[1] = one<2>(pw[0], pw[1], pw[2], wc);
        o[2] = one<1>(pw[0], pw[1], pw[2], wc);
        o[3] = one<0>(pw[0], pw[1], pw[2], wc);  // byte bits 1,0 -> e = 6,7
    }
    template <int B>
    __device__ __forceinline__ static uint4 gather(const uint32_t (&o)[4], uint32_t lb) {
        return make_uint4(lookup<B>(o[0], lb), lookup<B>(o[1], lb), lookup<B>(o[2], lb), lookup<B>(o[3], lb));
    }
    // entry p = (a2 b2 a1 b1 a0 b0) -> half2(C[a], C[b])
    __device__ __forceinline__ static void build(const __half *lut_row, uint32_t tbl, uint32_t half_id) {
        const uint4 c4 = __ldg(reinterpret_cast<const uint4 *>(lut_row));
        const uint32_t c[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int p = 0; p < 64; p++) {
            const int a = ((p >> 5) & 1) * 4 + ((p >> 3) & 1) * 2 + ((p >> 1) & 1);
            const int b = ((p >> 4) & 1) * 4 + ((p >> 2) & 1) * 2 + (p & 1);
            const uint32_t sel = (2 * (a & 1)) | ((2 * (a & 1) + 1) << 4) | ((4 + 2 * (b & 1)) << 8) | ((5 + 2 * (b & 1)) << 12);
            if ((uint32_t)(p >> 5) == half_id) sts_b32(tbl + p * 256, __byte_perm(c[a >> 1], c[b >> 1], sel));
        }
    }
};

template <>
struct Net<4> {
    static constexpr int NREG = 8;  // o[e]: byte B = index of the weight (c = 3 - B, e)
    template <int SFT>
    __device__ __forceinline__ static uint32_t nib(uint32_t P3, uint32_t P2, uint32_t P1, uint32_t P0) {
        // nibble m = 4-bit index (P3 P2 P1 P0) of the weight at bit position 4m + SFT
        return bitsel(shl<3 - SFT>(P3), bitsel(shl<2 - SFT>(P2), bitsel(shl<1 - SFT>(P1), shl<-SFT>(P0), 0x22222222u), 0x44444444u),
                      0x88888888u);
    }
    template <int SFT>
    __device__ __forceinline__ static void two(uint32_t (&o)[8], const uint32_t (&pw)[4], uint32_t wc) {
        const uint32_t y = nib<SFT>(pw[0], pw[1], pw[2], pw[3]);
        o[7 - SFT] = and_or(y, 0x0F0F0F0Fu, wc);       // low nibble of a byte: bit SFT     -> e = 7 - SFT
        o[3 - SFT] = and_or(y >> 4, 0x0F0F0F0Fu, wc);  // high nibble:          bit 4 + SFT -> e = 3 - SFT
    }
    __device__ __forceinline__ static void run(uint32_t (&o)[8], const uint32_t (&pw)[4], uint32_t wc) {
        two<0>(o, pw, wc);
        two<1>(o, pw, wc);
        two<2>(o, pw, wc);
        two<3>(o, pw, wc);
    }
    template <int B>
    __device__ __forceinline__ static uint4 gather(const uint32_t (&o)[8], uint32_t lb) {
        uint4 v;
        v.x = __byte_perm(lookup<B>(o[0], lb), lookup<B>(o[1], lb), 0x5410u);
        v.y = __byte_perm(lookup<B>(o[2], lb), lookup<B>(o[3], lb), 0x5410u);
        v.z = __byte_perm(lookup<B>(o[4], lb), lookup<B>(o[5], lb), 0x5410u);
        v.w = __byte_perm(lookup<B>(o[6], lb), lookup<B>(o[7], lb), 0x5410u);
        return v;
    }
    __device__ __forceinline__ static void build(const __half *lut_row, uint32_t tbl, uint32_t half_id) {
        const uint4 c4 = __ldg(reinterpret_cast<const uint4 *>(lut_row) + half_id);  // halfs 8*half_id .. +7
        const uint32_t c[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int p = 0; p < 8; p++) sts_b32(tbl + (half_id * 8 + p) * 256, (p & 1) ? (c[p >> 1] >> 16) : (c[p >> 1] & 0xFFFFu));
    }
};

// ------------------------------------------------------------------------------------------------ kernel
// MINB = CTAs per SM the register allocation must allow: 2 for token tiles <= 128 (bits 2 / 3), where a CTA is a latency chain
// (lookup -> tcgen05.st -> wait::st per 64 k, IPC 0.14 per warp) and a second resident CTA fills the gaps; 1 otherwise
template <int BITS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) prefill_tc_kernel(const __grid_constant__ CUtensorMap map_x, const Params p) {
    extern __shared__ __align__(16) uint8_t ptc_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t s0 = smem_u32(ptc_smem);
    // control block: full[MAX_STAGES], empty[MAX_STAGES], accum, tmem base
    const uint32_t bar_full = s0, bar_empty = s0 + 8 * MAX_STAGES, bar_accum = s0 + 16 * MAX_STAGES, tmem_slot = bar_accum + 8;
    // base of this CTA's shared window: 0 for a plain launch, rank-dependent (far above 256 KB) inside a cluster (measured:
    // an absolute layout traps in cluster rank 1).  Tables and token tiles sit at FIXED offsets from it, so that a table
    // address is window base + LDS immediate + index byte.
    const uint32_t win = s0 & 0xFFFC0000u;
    const uint32_t b_base = win + Lay<BITS>::STAGE_BASE;
    const uint32_t b_tile = p.t_tile * (BK * 2);
    uint32_t dyn_size;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
    if (s0 - win + CTRL_BYTES > TBL_ABS || b_base + p.stages * b_tile > s0 + dyn_size) {
        if (threadIdx.x == 0) printf("prefill_tc: shared memory layout does not fit (base %u, size %u)\n", s0, dyn_size);
        __trap();
    }
    const uint32_t row0 = blockIdx.x * ROWS, tok0 = blockIdx.y * p.t_tile;
    const uint32_t sb0 = (uint32_t)((uint64_t)p.sb_total * blockIdx.z / p.splits);
    const uint32_t sb1 = (uint32_t)((uint64_t)p.sb_total * (blockIdx.z + 1) / p.splits);

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.stages; s++) {
            mb_init(bar_full + 8 * s, DQ_THREADS + 1);
            mb_init(bar_empty + 8 * s, 1);
        }
        mb_init(bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    uint4 cur[BITS];
    if (warp < DQ_WARPS) {
        // lookup tables of this CTA's 128 rows: warps w and w+4 share rows 32*(w&3)..+31 and build half of the entries each
        const uint32_t row = (warp & 3) * 32 + lane;
        const uint32_t grow = min(row0 + row, p.N - 1);
        // the first packed words of the row are requested before anything else: their DRAM latency hides the set-up
        const uint4 *wrow0 = reinterpret_cast<const uint4 *>(p.W) + (size_t)grow * (p.K >> 7) + (warp >> 2);
#pragma unroll
        for (int j = 0; j < BITS; j++) cur[j] = ldg_stream_v4(wrow0 + j * ((size_t)p.N * (p.K >> 7)) + 2 * (size_t)sb0);
        const uint32_t tbl = win + TBL_ABS + (((warp & 3) >> 1) << Cfg<BITS>::ENTRY_BITS) * 256 + (warp & 1) * 128 + lane * 4;
        Net<BITS>::build(p.lut + (size_t)grow * (1u << BITS), tbl, warp >> 2);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp < DQ_WARPS) {
        // ---------------------------------------------------------------- dequantise into the A tiles
        const uint32_t h = warp >> 2;  // which 4 of the 8 words of a super block
        const uint32_t row = (warp & 3) * 32 + lane;
        const uint32_t grow = min(row0 + row, p.N - 1);
        const uint32_t wc = ((warp & 3) >> 1) * (BITS == 3 ? 0x40404040u : 0x10101010u);
        const uint32_t lb = win + (warp & 1) * 128 + lane * 4;  // bytes 2,3 = window base, byte 1 = 0 (index goes there)
        const uint32_t wpr4 = p.K >> 7;  // uint4 per row and plane
        const uint4 *wrow = reinterpret_cast<const uint4 *>(p.W) + (size_t)grow * wpr4 + h;
        const size_t plane4 = (size_t)p.N * wpr4;
        // A ring in TMEM: stage s = columns t_tile + s*A_COLS .. +31 of all 128 lanes; this warp owns lanes 32*(warp&3)..+31
        // and, of a stage, the 16 columns of its 4 words
        const uint32_t a_tmem = tmem_base + (((warp & 3) * 32) << 16) + p.t_tile + h * 16;

        uint4 nxt[BITS];
        uint32_t s = 0, ph = 0;
        for (uint32_t sb = sb0; sb < sb1; sb++) {
            const uint32_t sbn = min(sb + 1, sb1 - 1);
#pragma unroll
            for (int j = 0; j < BITS; j++) nxt[j] = ldg_stream_v4(wrow + j * plane4 + 2 * (size_t)sbn);
            uint32_t o[4][Net<BITS>::NREG];
#pragma unroll
            for (int tt = 0; tt < 4; tt++) {
                uint32_t pw[BITS];
#pragma unroll
                for (int j = 0; j < BITS; j++) pw[j] = comp(cur[j], tt);
                Net<BITS>::run(o[tt], pw, wc);
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint4 v[4];
#pragma unroll
                for (int tt = 0; tt < 4; tt++) {
                    if (c == 0) v[tt] = Net<BITS>::template gather<3>(o[tt], lb);
                    if (c == 1) v[tt] = Net<BITS>::template gather<2>(o[tt], lb);
                    if (c == 2) v[tt] = Net<BITS>::template gather<1>(o[tt], lb);
                    if (c == 3) v[tt] = Net<BITS>::template gather<0>(o[tt], lb);
                }
                mb_wait(bar_empty + 8 * s, ph ^ 1);  // the MMAs that read this A stage (and token tile) have completed
                tc_fence_after();
                tmem_st16(a_tmem + s * A_COLS, v[0], v[1], v[2], v[3]);
                tmem_wait_st();
                tc_fence_before();
                mb_arrive(bar_full + 8 * s);
                if (++s == p.stages) s = 0, ph ^= 1;
            }
#pragma unroll
            for (int j = 0; j < BITS; j++) cur[j] = nxt[j];
        }
        // ---------------------------------------------------------------- epilogue: TMEM -> registers -> global
        mb_wait(bar_accum, 0);
        tc_fence_after();
        const uint32_t q = warp & 3;
        const uint32_t half_cols = p.t_tile >> 1;
        const uint32_t n = row0 + q * 32 + lane;
        for (uint32_t cc = 0; cc < half_cols; cc += 16) {
            const uint32_t col = h * half_cols + cc;
            uint32_t v[16];
            tmem_ld16(tmem_base + ((q * 32) << 16) + col, v);
            if (n < p.N) {
                if (p.splits == 1) {
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        const uint32_t tok = tok0 + col + u;
                        if (tok < p.T) p.out[(size_t)tok * p.N + n] = __float2half_rn(__uint_as_float(v[u]));
                    }
                } else {
                    float *dst = p.partial + (size_t)blockIdx.z * p.T * p.N;
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        const uint32_t tok = tok0 + col + u;
                        if (tok < p.T) dst[(size_t)tok * p.N + n] = __uint_as_float(v[u]);
                    }
                }
            }
        }
    } else if (warp == X_WARP) {
        // ---------------------------------------------------------------- token tiles by TMA
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (uint32_t sb = sb0; sb < sb1; sb++) {
                const uint32_t i = sb >> 2, j = sb & 3, eff = chunk_eff(p.K, i);
                for (uint32_t c = 0; c < 4; c++) {
                    mb_wait(bar_empty + 8 * s, ph ^ 1);
                    mb_arrive_expect_tx(bar_full + 8 * s, b_tile);
                    tma_load_2d(b_base + s * b_tile, &map_x, i * 1024 + c * 8 * eff + 64 * j, tok0, bar_full + 8 * s);
                    if (++s == p.stages) s = 0, ph ^= 1;
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- MMA issue
        if (lane == 0) {
            uint32_t s = 0, ph = 0, acc = 0;
            const uint32_t n_it = (sb1 - sb0) * 4;
            for (uint32_t it = 0; it < n_it; it++) {
                mb_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                const uint64_t db = smem_desc(b_base + s * b_tile);
                const uint32_t ta = tmem_base + p.t_tile + s * A_COLS;
#pragma unroll
                for (uint32_t kk = 0; kk < BK / 16; kk++) {
                    umma_f16_ts(tmem_base, ta + 8 * kk, db + 2 * kk, p.idesc, acc);
                    acc = 1;
                }
                umma_commit(bar_empty + 8 * s);
                if (++s == p.stages) s = 0, ph ^= 1;
            }
            umma_commit(bar_accum);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
}

// out[t][n] = fp16( sum_s partial[s][t][n] ), summed in split order (deterministic)
__global__ void __launch_bounds__(256) reduce_splits_kernel(const float *__restrict__ partial, __half *__restrict__ out, uint64_t total,
                                                            uint32_t splits) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float acc = 0.f;
    for (uint32_t s = 0; s < splits; s++) acc += partial[(uint64_t)s * total + i];
    out[i] = __float2half_rn(acc);
}

}  // namespace ptc
}  // namespace apg
