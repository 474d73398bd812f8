// Persistent token kernel: ONE cooperative launch runs a whole list of dependent jobs (the Any-Precision LUT GEMVs of a
// decode step with their fused RMSNorm / residual / SwiGLU, the attention of every block, the embedding row) — sm_100a.
//
// Why (measured on B200, tools/probes/*): at batch 1 every Linear needs the WHOLE output vector of the previous one, so a
// token is ~160 strictly dependent all-to-all hand-overs.  As separate PDL-chained launches one hand-over costs ~2.9 us
// (148 CTAs x 544 threads, 16 KB vector); a counter + fence grid barrier inside a persistent kernel costs the same.  A
// data-with-flag hand-over does not: every activation vector lives in global memory as 8-byte packets
// (half2 value, epoch) written with ONE store, consumers spin on the packets they need.  No fences, no counters, no grid
// barrier — the cost is one L2 write + one L2 read — and, because the CTAs never leave the SMs, the weight ring keeps
// streaming the NEXT Linears' bit-planes (which do not depend on x) across job boundaries.
//
//   * grid = #SMs, 1 CTA/SM, 16 warps (512 threads, 128 registers each); launched cooperatively (co-residency is required:
//     consumers spin on packets produced by other CTAs).
//   * PRODUCER thread: walks the job list, streams this CTA's rows of every GEMV job through a BYTE ring in shared memory
//     (variable-size stages: RS rows x BITS planes x K/8 bytes) with 1-D bulk async copies; 32 (full, empty) mbarrier
//     pairs used round-robin by stage sequence number; the stage's ring offset is published in shared memory before the
//     expect_tx arrive.  It runs ahead of the consumers by up to the ring size (~200 KB = several Linears' share).
//   * CONSUMERS: per GEMV job the same mapping as gemv_fast_kernel (apgemv_fast.cuh: warp = K chunk(s) of a row group,
//     x in registers, per-warp codebook tables, LOP3/PRMT index networks, fp16 chains -> fp32), rows of the job dealt to
//     the CTAs in units of RS/2 rows.  x is read from the previous job's packets; the epilogue writes packets.
//   * Math per row is IDENTICAL to gemv_fast_kernel (same chunk -> warp map, same summation order), so results are
//     bit-identical to the per-launch path; the attention job merges 16 warps instead of 4 (fp32 merge order differs).
//
// Reference semantics: inference/model.py:151-167 (block), 206-236 (attention), 261-266 (SwiGLU), 280-285 (RMSNorm),
// inference/ap_gemv/anyprec.cu:372-542 (GEMV).
#pragma once
#include <math_constants.h>

#include "apgemv_fast.cuh"

namespace apg {

constexpr int PK_NCW = 15;                       // consumer warps per CTA; + 1 producer warp = 16 warps = 512 threads -> 128
                                                 // registers per thread (a 17th warp caps every thread at 96, where ptxas
                                                 // serialises the row loop: see pk_gemv_job)
constexpr int PK_THREADS = (PK_NCW + 1) * 32;
constexpr uint32_t PK_NB = 32;                   // (full, empty) barrier pairs, used round-robin by stage number
constexpr uint32_t PK_EMPTY_COUNT = 720720u;     // lcm(1..16): a group of nwk warps arrives with 720720 / nwk each
constexpr uint32_t PK_SCRATCH_BYTES = 16384;     // per-chunk partial sums of a job / attention scratch
constexpr uint32_t PK_MAX_STAGE = 32768;
constexpr long long PK_WATCHDOG_CYCLES = 40000000000ll;  // ~20 s (ranks of a tensor-parallel group may start seconds apart): a wait
                                                         // that long is a bug -> trap instead of hanging the GPU

enum : uint32_t { PJ_END = 0, PJ_GEMV = 1, PJ_ATTN = 2, PJ_PACK = 3, PJ_REDUCE = 4 };
enum : uint32_t { PF_NORM = 1, PF_RESIDUAL = 4, PF_GLU = 8, PF_PUSH = 16 };

// One job.  Activation vectors are "LL buffers": uint2 packets (half2 bits, epoch), n / 2 packets for n halfs.
struct alignas(16) PJob {
    uint32_t type, flags, N, K;
    uint32_t nwk, groups, cpw, rs;
    uint32_t stage_bytes, unit_rows, units_q, units_rem;
    uint32_t inv_nwk;
    float eps;
    uint32_t a0, a1;      // ATTN: H, Hkv        PACK: vocab, -
    uint32_t a2;          // ATTN: S (cache length)
    float f0;             // ATTN: softmax scale
    uint32_t world, rank; // PUSH / REDUCE
    uint32_t tag_x, tag_res, tag_out, tag_pad;  // job indices: packets of job j carry the tag  epoch * n_jobs + j, so
                                                // buffers can be re-used by later jobs of the same token
    const void *x;        // GEMV/ATTN: LL input; PACK: plain fp16 source rows; REDUCE: uint2 [world][N] fp32 packets
    const void *W;        // GEMV: bit-planes [bits][N][K/8] bytes
    const void *lut;      // GEMV: fp16 [N][2^bits]
    const void *norm_w;   // GEMV + PF_NORM: fp16 [K]
    const void *residual; // GEMV/REDUCE + PF_RESIDUAL: LL buffer of the output length
    void *out;            // LL output (N halfs; N / 2 with PF_GLU)
    void *out_plain;      // optional plain fp16 copy of the output
    void *p0, *p1, *p2;   // ATTN: rope table fp16 [S][cos 64 | sin 64], k_cache, v_cache fp16 [Hkv][S][128];  PACK: p0 = token id (int*) or NULL
    void *peer[8];        // PF_PUSH: every rank's uint2 [world][N] receive buffer of this site
};

struct PParams {
    const PJob *jobs;
    uint32_t n_jobs;
    uint32_t ring_bytes;
    uint32_t xs_bytes;   // shared-memory staging area of a job's input vector (>= 2 * max K over the GEMV jobs)
    uint32_t *epoch;     // device counter; packets of this launch carry *epoch + 1; bumped at the end if bump_epoch
    const int *pos;      // device position (attention)
    uint32_t *err;       // device word: non-zero = a watchdog fired (code in the low byte)
    uint32_t *done;      // device counter (zero between launches): the last CTA to finish bumps the epoch
    uint32_t bump_epoch;
    long long *prof;     // optional (debug): [gridDim.x][n_jobs][4] clock64 stamps of consumer thread 0: job start, x in registers,
                         // stages done, epilogue done
};

// static shared memory at namespace scope: addresses are compile-time constants, so nothing here costs a register and the
// (non-inlined) job functions reach it without arguments; the consumers read the fields of their current job from here
// whenever they need them instead of keeping them live
struct PkShared {
    uint64_t full[PK_NB], empty[PK_NB];
    uint32_t stage_off[PK_NB];  // ring offset of the stage in each barrier slot (written by the producer)
    float ssq[PK_NCW];
    PJob jobs[2];               // the consumers' current job (slot j & 1) and the next one, copied from global memory
    uint32_t *err;
    uint32_t tag_base;          // epoch * n_jobs
    uint32_t q_base[2];         // stage sequence number of the first stage of job j (this CTA) in slot j & 1: the slot of the
                                // NEXT job is written while the current one runs, so no barrier separates update and use
    long long *prof;            // this CTA's 4 clock stamps of the current job (debug), or nullptr
    // cooperative weight producer (pk_produce): cursor over (job, stage), ring allocation state, try-lock
    const PJob *jobs_g;
    uint32_t n_jobs, ring_bytes, ring0, bits;   // ring0: shared address of the weight ring
    uint32_t p_lock, p_done, p_job, p_stage, p_q, p_qtail, p_head;
    uint32_t xs_bytes;
    uint32_t scratch_rel;       // offset of the 16 KB scratch area in the dynamic shared memory (depends on the bit-width)
};
static_assert(sizeof(PkShared) <= 2048, "the static shared header must stay within the 2 KB the host leaves for it");
static_assert(sizeof(PJob) % 16 == 0, "PJob is copied in 16-byte pieces");

__shared__ __align__(16) PkShared g_sh;
extern __shared__ __align__(1024) uint8_t g_dyn[];  // [tables NCW x WTB | scratch 16 KB | x staging xs_bytes | ring]


// ------------------------------------------------------------------------------------------------------------
// packet primitives
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ll_ld16(const void *p) {
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint2 ll_ld8(const void *p) {
    uint2 r;
    asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void ll_st8(void *p, uint32_t v, uint32_t ep) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v), "r"(ep) : "memory");
}
__device__ __forceinline__ void ll_st16(void *p, uint32_t v0, uint32_t v1, uint32_t ep) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v0), "r"(ep), "r"(v1), "r"(ep) : "memory");
}
__device__ __noinline__ void pk_die(uint32_t *err, uint32_t code) {
    atomicExch(err, code | (blockIdx.x << 8) | (threadIdx.x << 20));
    __threadfence_system();
    __trap();
}
// Slow paths of the packet waits live in their own (non-inlined) functions: they are cold, and keeping the kernel's code
// small is what keeps ptxas's good schedule of the row loop (see pk_gemv_job).
__device__ __noinline__ uint2 ll_spin16(const uint2 *p, uint32_t ep, uint32_t *err) {
    const long long t0 = clock64();
    uint4 v;
    do {
        __nanosleep(32);
        v = ll_ld16(p);
        if (clock64() - t0 > PK_WATCHDOG_CYCLES) pk_die(err, 1u);
    } while (v.y != ep || v.w != ep);
    return make_uint2(v.x, v.z);
}
__device__ __noinline__ uint32_t ll_spin8(const uint2 *p, uint32_t ep, uint32_t *err) {
    const long long t0 = clock64();
    uint2 v;
    do {
        __nanosleep(32);
        v = ll_ld8(p);
        if (clock64() - t0 > PK_WATCHDOG_CYCLES) pk_die(err, 2u);
    } while (v.y != ep);
    return v.x;
}
// two packets (4 halfs) at packet index pk (even): spin until both carry the epoch
__device__ __forceinline__ uint2 ll_wait16(const uint2 *base, uint32_t pk, uint32_t ep, uint32_t *err) {
    const uint4 v = ll_ld16(base + pk);
    if (v.y != ep || v.w != ep) return ll_spin16(base + pk, ep, err);
    return make_uint2(v.x, v.z);
}
__device__ __forceinline__ uint32_t ll_wait8(const uint2 *base, uint32_t pk, uint32_t ep, uint32_t *err) {
    const uint2 v = ll_ld8(base + pk);
    if (v.y != ep) return ll_spin8(base + pk, ep, err);
    return v.x;
}

__device__ __forceinline__ void mbar_arrive_cnt(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __noinline__ void mbar_spin(uint32_t bar, uint32_t parity, uint32_t *err, uint32_t code, bool backoff) {
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (backoff) __nanosleep(256);
        if (clock64() - t0 > PK_WATCHDOG_CYCLES) pk_die(err, code);
    }
}
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity, uint32_t *err, uint32_t code, bool backoff) {
    if (!mbar_try(bar, parity)) mbar_spin(bar, parity, err, code, backoff);
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 2, %0;" ::"n"(PK_NCW * 32) : "memory"); }

// rows [r_begin, r_end) of a job owned by this CTA (same dealing as gemv_fast_kernel: consecutive units)
__device__ __forceinline__ void pk_rows(const PJob &jb, uint32_t &r_begin, uint32_t &r_end) {
    const uint32_t u_begin = blockIdx.x * jb.units_q + min(blockIdx.x, jb.units_rem);
    const uint32_t nunits = jb.units_q + (blockIdx.x < jb.units_rem ? 1u : 0u);
    r_begin = min(u_begin * jb.unit_rows, jb.N);
    r_end = min((u_begin + nunits) * jb.unit_rows, jb.N);
}

// ------------------------------------------------------------------------------------------------------------
// GEMV job, consumer side: ONE function per bit-width for every shape (K <= 16384: one 1024-chunk per warp; stages of 8
// or 4 rows, all rows unrolled), with the SwiGLU / residual / push epilogues behind run-time flags.
// Why one small function: ptxas schedules the row loop well (table lookups hoisted ~35 instructions ahead of their FMAs,
// two rows interleaved) only while the kernel's whole call graph stays small; with several unrolled variants in the same
// kernel — inlined or not — it falls back to a schedule that puts every FMA right behind its lookup (measured: 1.9x
// slower stages).  NOT inlined, so that its registers are allocated on their own.
// ------------------------------------------------------------------------------------------------------------
template <int BITS>
__device__ __noinline__ void pk_gemv_job(const uint32_t slot) {
    constexpr int WTB = FastWarpTbl<BITS, 8>::BYTES;
    constexpr uint32_t scratch_rel = PK_NCW * WTB, xs_rel = scratch_rel + PK_SCRATCH_BYTES;
    using Tb = Tables<BITS, 8>;
    const PJob &jb = g_sh.jobs[slot];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tid = threadIdx.x;
    uint32_t r_begin, r_end;
    pk_rows(jb, r_begin, r_end);
    const uint32_t nrows = r_end - r_begin;
    const uint32_t RS = jb.rs;  // rows per stage: 8 or 4
    const uint32_t nstages = (nrows + RS - 1) / RS;
    float *red = reinterpret_cast<float *>(g_dyn + scratch_rel);
    const uint32_t smem0 = smem_u32(g_dyn);
    const bool active = (uint32_t)warp < jb.groups * jb.nwk && nrows > 0;  // a CTA without rows only passes the barriers
    const uint32_t K = jb.K;
    const bool do_norm = (jb.flags & PF_NORM) != 0;

    // codebook rows of this group's first stage: static data, pulled into L1 before x is waited for (a register prefetch
    // would stay live across the row loop, and the row loop needs every register it can get at 96 per thread)
    uint32_t g = 0, wk = 0;
    if (active) {
        g = ((uint32_t)warp * jb.inv_nwk) >> 16, wk = warp - g * jb.nwk;
        if (g < nstages) Tb::prefetch(static_cast<const __half *>(jb.lut), r_begin + g * RS, jb.N, lane);
    }

    // ---- x: packets -> shared memory, ONCE per CTA (every thread spins on its own 16-byte pieces), with the sum of
    //      squares of the fused RMSNorm on the way; then every warp takes its K chunk from shared memory
    if (nrows > 0) {
        const uint32_t ep = g_sh.tag_base + jb.tag_x;
        const uint2 *xin = static_cast<const uint2 *>(jb.x);
        const uint32_t nvec = K >> 2;  // 16-byte vectors = 2 packets = 4 halfs
        // lane 0 of each warp spins first, so that a CTA that is early polls with 16 loads, not 512
        if (lane == 0 && tid < nvec) (void)ll_wait16(xin, 2u * tid, ep, g_sh.err);
        __syncwarp();
        float ss = 0.f;
        for (uint32_t v0 = tid; v0 < nvec; v0 += 4u * PK_NCW * 32u) {
            uint4 q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t v = v0 + u * PK_NCW * 32u;
                if (v < nvec) q[u] = ll_ld16(xin + 2u * v);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t v = v0 + u * PK_NCW * 32u;
                if (v < nvec) {
                    uint2 d = make_uint2(q[u].x, q[u].z);
                    if (q[u].y != ep || q[u].w != ep) d = ll_spin16(xin + 2u * v, ep, g_sh.err);
                    *reinterpret_cast<uint2 *>(g_dyn + xs_rel + 8u * v) = d;
                    if (do_norm) {
                        const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&d.x));
                        const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&d.y));
                        ss = fmaf(f0.x, f0.x, ss), ss = fmaf(f0.y, f0.y, ss), ss = fmaf(f1.x, f1.x, ss), ss = fmaf(f1.y, f1.y, ss);
                    }
                }
            }
        }
        if (do_norm) {
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0) g_sh.ssq[warp] = ss;
        }
    }
    bar_consumers();
    if (tid == 0 && g_sh.prof) g_sh.prof[1] = clock64();

    if (active) {
        const uint32_t nwk = jb.nwk;
        const uint32_t tbl = smem0 + warp * WTB;
        const uint32_t eff = chunk_eff(K, wk);  // wk < nchunk by construction (nwk = number of chunks)
        const bool act = (uint32_t)lane < eff;
        uint32_t xr[16];
        if (act) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint4 v = *reinterpret_cast<const uint4 *>(g_dyn + xs_rel + 2u * (wk * 1024u + c * 8u * eff + 8u * lane));
                xr[4 * c + 0] = v.x, xr[4 * c + 1] = v.y, xr[4 * c + 2] = v.z, xr[4 * c + 3] = v.w;
            }
        }
        if (do_norm) {
            // fused RMSNorm (model.py:280-285), sums of squares added in a fixed order
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < PK_NCW; w++) tot += g_sh.ssq[w];
            const float rs = rsqrtf(tot / (float)K + jb.eps);
            if (act) {
                const __half *nw = static_cast<const __half *>(jb.norm_w) + wk * 1024u + 8u * lane;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint4 wv = __ldg(reinterpret_cast<const uint4 *>(nw + c * 8u * eff));
                    const uint32_t wn[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&xr[4 * c + e]));
                        const __half2 n = __floats2half2_rn(f.x * rs, f.y * rs);                    // .type_as(x)
                        const __half2 r = __hmul2(n, *reinterpret_cast<const __half2 *>(&wn[e]));  // * weight
                        xr[4 * c + e] = *reinterpret_cast<const uint32_t *>(&r);
                    }
                }
            }
        }

        // ---- stages of this group
        const uint32_t G = jb.groups;
        const uint32_t ring0 = g_sh.ring0;
        const uint32_t row_bytes = K >> 3;
        for (uint32_t s = g; s < nstages; s += G) {
            const uint32_t q = g_sh.q_base[slot] + s, b = q & (PK_NB - 1u), par = (q / PK_NB) & 1u;
            const uint32_t row0 = r_begin + s * RS;
            const uint32_t rows = min(RS, r_end - row0);
            __syncwarp();
            {
                typename Tb::Regs lr;  // L1 hits (prefetched one stage ahead)
                Tb::fetch(lr, static_cast<const __half *>(jb.lut), row0, jb.N, lane);
                Tb::store(lr, tbl, lane);
            }
            Tb::prefetch(static_cast<const __half *>(jb.lut), row0 + G * RS, jb.N, lane);
            __syncwarp();
            mbar_wait_wd(smem_u32(&g_sh.full[b]), par, g_sh.err, 3u, false);
            const uint32_t stage = ring0 + g_sh.stage_off[b] + (wk * 32u + lane) * 4u;
            // all rows of the stage unrolled, like gemv_fast_kernel (a rolled loop over groups of four rows was measured
            // 8.6 % slower there); rows past the end of a partial stage read stale-but-mapped ring bytes of the same stage
            // and are discarded; a stage of at most four rows (a 4-row stage, or the tail of the CTA's rows) runs the
            // half-length copy of the loop
            float acc[8];
#pragma unroll
            for (int r = 0; r < 8; r++) acc[r] = 0.f;
            if (act) {
                if (rows <= 4u) RowLoop<BITS, 8, 0, 4>::run(acc, stage, row_bytes, RS * row_bytes, xr, tbl);
                else RowLoop<BITS, 8, 0, 8>::run(acc, stage, row_bytes, RS * row_bytes, xr, tbl);
            }
            {
                const float v = BatchReduce<8>::run(acc, lane);
                const uint32_t rl = lane >> BatchReduce<8>::SH;
                if ((lane & ((1 << BatchReduce<8>::SH) - 1)) == 0 && rl < rows) red[(row0 - r_begin + rl) * nwk + wk] = v;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_cnt(smem_u32(&g_sh.empty[b]), PK_EMPTY_COUNT / nwk);
        }
    }

    // the residual packet of this thread's first epilogue row pair was produced at least two jobs ago: fetched before the
    // barrier so that its L2 latency is off the critical path (validated, and re-fetched if stale, below)
    const bool glu = (jb.flags & PF_GLU) != 0;
    uint2 res_pre = make_uint2(0u, 0u);
    if (!glu && (jb.flags & PF_RESIDUAL) && 2u * tid < nrows)
        res_pre = ll_ld8(static_cast<const uint2 *>(jb.residual) + ((r_begin + 2u * tid) >> 1));
    bar_consumers();  // every chunk partial of every row of this CTA is in `red`
    if (tid == 0 && g_sh.prof) g_sh.prof[2] = clock64();
    // ---- epilogue: fixed-order combination (deterministic), fused residual / SwiGLU, packets out
    const uint32_t nwk = jb.nwk;
    const uint32_t ep_out = g_sh.tag_base + jb.tag_out;
    if (glu) {
        // rows are interleaved (gate_i, up_i); out[i] = silu(y[2i]) * y[2i+1] with the roundings of FeedForward on half
        // tensors (model.py:261-266): y -> fp16, silu in fp32 -> fp16, fp16 product.  One packet = two outputs = 4 rows.
        uint2 *outp = static_cast<uint2 *>(jb.out);
        for (uint32_t r = 4u * tid; r < nrows; r += 4u * PK_NCW * 32u) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                v[e] = red[(r + e) * nwk];
                for (uint32_t w = 1; w < nwk; w++) v[e] += red[(r + e) * nwk + w];
            }
            __half o[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const float gf = __half2float(__float2half_rn(v[2 * e]));
                const __half sg = __float2half_rn(__fdividef(gf, 1.f + __expf(-gf)));
                o[e] = __hmul(sg, __float2half_rn(v[2 * e + 1]));
            }
            const uint32_t ov = (uint32_t)__half_as_ushort(o[0]) | ((uint32_t)__half_as_ushort(o[1]) << 16);
            const uint32_t oi = (r_begin + r) >> 2;  // packet index = output index / 2
            ll_st8(outp + oi, ov, ep_out);
            if (jb.out_plain) reinterpret_cast<uint32_t *>(jb.out_plain)[oi] = ov;
        }
    } else {
        uint2 *outp = static_cast<uint2 *>(jb.out);
        const uint2 *resp = static_cast<const uint2 *>(jb.residual);
        const bool push = (jb.flags & PF_PUSH) != 0;
        for (uint32_t r = 2u * tid; r < nrows; r += 2u * PK_NCW * 32u) {
            float v0 = red[r * nwk], v1 = red[(r + 1u) * nwk];
            for (uint32_t w = 1; w < nwk; w++) v0 += red[r * nwk + w], v1 += red[(r + 1u) * nwk + w];
            const uint32_t pk = (r_begin + r) >> 1;
            if (push) {
                // K-sharded Linear: fp32 partial sums pushed, with the tag, into slot `rank` of every rank's buffer over
                // NVLink; the PJ_REDUCE job that follows sums them in rank order
                for (uint32_t pr = 0; pr < jb.world; pr++) {
                    uint2 *dst = static_cast<uint2 *>(jb.peer[pr]) + (size_t)jb.rank * jb.N + r_begin + r;
                    ll_st16(dst, __float_as_uint(v0), __float_as_uint(v1), ep_out);
                }
                continue;
            }
            __half2 h = __floats2half2_rn(v0, v1);
            if (jb.flags & PF_RESIDUAL) {
                const uint32_t ep_res = g_sh.tag_base + jb.tag_res;
                uint32_t rv = res_pre.x;
                if (r != 2u * tid || res_pre.y != ep_res) rv = ll_wait8(resp, pk, ep_res, g_sh.err);
                h = __hadd2(h, *reinterpret_cast<const __half2 *>(&rv));
            }
            const uint32_t hv = *reinterpret_cast<const uint32_t *>(&h);
            ll_st8(outp + pk, hv, ep_out);
            if (jb.out_plain) reinterpret_cast<uint32_t *>(jb.out_plain)[pk] = hv;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// all-reduce finish of a K-sharded Linear: rows dealt to all CTAs; sum the `world` fp32 packets in rank order
// (identical on every rank), add the residual, round once, emit packets
// ------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void pk_reduce_job(const uint32_t slot) {
    const PJob &jb = g_sh.jobs[slot];
    const uint32_t N = jb.N, world = jb.world;
    const uint32_t ep = g_sh.tag_base + jb.tag_x, ep_out = g_sh.tag_base + jb.tag_out, ep_res = g_sh.tag_base + jb.tag_res;
    const uint2 *recv = static_cast<const uint2 *>(jb.x);
    const uint2 *resp = static_cast<const uint2 *>(jb.residual);
    uint2 *outp = static_cast<uint2 *>(jb.out);
    for (uint32_t pk = blockIdx.x * (PK_NCW * 32u) + threadIdx.x; pk < (N >> 1); pk += gridDim.x * (PK_NCW * 32u)) {
        uint32_t rv = 0;
        if (jb.flags & PF_RESIDUAL) rv = ll_wait8(resp, pk, ep_res, g_sh.err);  // old data: ready, fetched before the long wait
        float s0 = 0.f, s1 = 0.f;
        for (uint32_t w = 0; w < world; w++) {
            const uint2 v = ll_wait16(recv, w * N + 2u * pk, ep, g_sh.err);
            s0 += __uint_as_float(v.x), s1 += __uint_as_float(v.y);
        }
        __half2 h = __floats2half2_rn(s0, s1);
        if (jb.flags & PF_RESIDUAL) h = __hadd2(h, *reinterpret_cast<const __half2 *>(&rv));
        const uint32_t hv = *reinterpret_cast<const uint32_t *>(&h);
        ll_st8(outp + pk, hv, ep_out);
        if (jb.out_plain) reinterpret_cast<uint32_t *>(jb.out_plain)[pk] = hv;
    }
}

// plain fp16 row -> packets (embedding row, model.py:123; or the chain benchmark's input vector)
__device__ __noinline__ void pk_pack_job(const uint32_t slot) {
    const PJob &jb = g_sh.jobs[slot];
    const uint32_t n = jb.N;
    uint32_t row = 0;
    if (jb.p0) row = min((uint32_t)max(*static_cast<const volatile int *>(jb.p0), 0), jb.a0 - 1u);  // never read outside the table
    const uint32_t *src = reinterpret_cast<const uint32_t *>(static_cast<const __half *>(jb.x) + (size_t)row * n);
    uint2 *outp = static_cast<uint2 *>(jb.out);
    for (uint32_t i = blockIdx.x * (PK_NCW * 32u) + threadIdx.x; i < (n >> 1); i += gridDim.x * (PK_NCW * 32u)) {
        const uint32_t v = __ldcg(src + i);
        ll_st8(outp + i, v, g_sh.tag_base + jb.tag_out);
        if (jb.out_plain) reinterpret_cast<uint32_t *>(jb.out_plain)[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------
// attention job: CTA h < H attends for head h with its 16 consumer warps (RoPE + KV append + softmax(q.K^T).V,
// Attention.forward model.py:206-236; same arithmetic as apd::attn_decode_kernel, decode_kernels.cuh, in blocks of 4
// cached steps per warp: lane = (step of the block, 16-dim group) for K and V alike, so a block is one memory round trip
// of four 16-byte loads per lane and two short FMA runs — long unrolled FMA blocks anywhere in this kernel make ptxas
// schedule the GEMV row loop conservatively, see pk_gemv_job)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pk_rope4(const __half (&v)[4], const __half (&partner)[4], int lane, const __half (&c)[4],
                                         const __half (&s)[4], __half (&o)[4]) {
    const bool hi = lane >= 16;  // rotate_half: d < 64 -> -x[d+64], d >= 64 -> x[d-64]  (model.py:268-272)
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const __half rot = hi ? partner[j] : __hneg(partner[j]);
        o[j] = __hadd(__hmul(v[j], c[j]), __hmul(rot, s[j]));
    }
}

__device__ __noinline__ void pk_attn_job(const uint32_t slot, const int *pos_ptr) {
    constexpr int HD = 128, NW = PK_NCW, BS = 4;  // BS cached steps per warp iteration
    const PJob &jb = g_sh.jobs[slot];
    const uint32_t H = jb.a0, Hkv = jb.a1, S = jb.a2;
    if (blockIdx.x >= H) return;  // uniform per CTA: the other CTAs move on to the next job
    const float scale = jb.f0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float *scratch = reinterpret_cast<float *>(g_dyn + g_sh.scratch_rel);
    const uint32_t G = H / Hkv, h = blockIdx.x, kvh = h / G;
    const uint32_t ep = g_sh.tag_base + jb.tag_x, ep_out = g_sh.tag_base + jb.tag_out;
    float *qs = scratch;                                             // [128] roped q (fp32)
    float(*wacc)[HD + 4] = reinterpret_cast<float(*)[HD + 4]>(scratch + HD);  // [NW][132]: (m, l, -, -, acc[128])
    __half *k_cache = static_cast<__half *>(jb.p1), *v_cache = static_cast<__half *>(jb.p2);
    const uint2 *qkv = static_cast<const uint2 *>(jb.x);
    int pos = *static_cast<const volatile int *>(pos_ptr);
    const bool pos_ok = pos >= 0 && pos < (int)S;  // the host API refuses to step past the cache; never write outside it
    if (!pos_ok) pos = 0;

    __half kr[4], vn[4];
    float qf[4];
    if (w == 0) {
        // cos / sin of this position, computed in fp32 and rounded to fp16 on the host (model.py:396-405): a table row
        // [cos(64) | sin(64)] instead of sincosf keeps this kernel small (see pk_gemv_job on why that matters)
        const __half *cs = static_cast<const __half *>(jb.p0) + (size_t)pos * 128;
        __half rc[4], rs[4];
#pragma unroll
        for (int j = 0; j < 4; j++) rc[j] = cs[(4 * lane + j) & 63], rs[j] = cs[64 + ((4 * lane + j) & 63)];
        __half q[4], kn[4], qp[4], kp[4], qr[4];
        const uint2 qv = ll_wait16(qkv, (h * HD + 4 * lane) >> 1, ep, g_sh.err);
        const uint2 kv = ll_wait16(qkv, ((H + kvh) * HD + 4 * lane) >> 1, ep, g_sh.err);
        const uint2 vv = ll_wait16(qkv, ((H + Hkv + kvh) * HD + 4 * lane) >> 1, ep, g_sh.err);
        *reinterpret_cast<uint2 *>(q) = qv, *reinterpret_cast<uint2 *>(kn) = kv, *reinterpret_cast<uint2 *>(vn) = vv;
        uint2 qpv, kpv;
        qpv.x = __shfl_xor_sync(0xffffffffu, qv.x, 16), qpv.y = __shfl_xor_sync(0xffffffffu, qv.y, 16);
        kpv.x = __shfl_xor_sync(0xffffffffu, kv.x, 16), kpv.y = __shfl_xor_sync(0xffffffffu, kv.y, 16);
        *reinterpret_cast<uint2 *>(qp) = qpv, *reinterpret_cast<uint2 *>(kp) = kpv;
        pk_rope4(q, qp, lane, rc, rs, qr);
        pk_rope4(kn, kp, lane, rc, rs, kr);
        if (pos_ok && h == kvh * G) {  // cache append (KVCache.update, model.py:70-79), once per KV head
            *reinterpret_cast<uint2 *>(k_cache + ((size_t)kvh * S + pos) * HD + 4 * lane) = *reinterpret_cast<uint2 *>(kr);
            *reinterpret_cast<uint2 *>(v_cache + ((size_t)kvh * S + pos) * HD + 4 * lane) = *reinterpret_cast<uint2 *>(vn);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) qf[j] = __half2float(qr[j]);
        *reinterpret_cast<float4 *>(qs + 4 * lane) = make_float4(qf[0], qf[1], qf[2], qf[3]);
    }
    bar_consumers();

    const __half *Kb = k_cache + (size_t)kvh * S * HD;
    const __half *Vb = v_cache + (size_t)kvh * S * HD;
    const int nblk = (pos + BS - 1) / BS;  // blocks of BS = 4 cached steps t < pos
    const int tg = lane >> 3, dg = lane & 7;   // lane = (step of the block, 16-dim group) for K and for V alike
    float m = -CUDART_INF_F, l = 0.f, acc[16];
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0.f;
    float qv[16];  // this lane's 16 dims of the roped q
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
        const float4 t = *reinterpret_cast<const float4 *>(qs + 16 * dg + j);
        qv[j] = t.x, qv[j + 1] = t.y, qv[j + 2] = t.z, qv[j + 3] = t.w;
    }
#pragma unroll 1
    for (int b = w; b < nblk; b += NW) {
        const int t = b * BS + tg;
        const bool ok = t < pos;
        // ---- the block's loads first (one memory round trip): 32 bytes of the K row and of the V row of this lane's step
        uint4 k0 = make_uint4(0, 0, 0, 0), k1 = k0, v0 = k0, v1 = k0;
        if (ok) {
            const uint4 *kr4 = reinterpret_cast<const uint4 *>(Kb + (size_t)t * HD + 16 * dg);
            const uint4 *vr4 = reinterpret_cast<const uint4 *>(Vb + (size_t)t * HD + 16 * dg);
            k0 = __ldcg(kr4), k1 = __ldcg(kr4 + 1), v0 = __ldcg(vr4), v1 = __ldcg(vr4 + 1);
        }
        // ---- score of the step: 16 dims per lane, summed over the 8 lanes of the step
        const uint32_t kw[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&kw[j]));
            a0 = fmaf(qv[2 * j], f.x, a0), a1 = fmaf(qv[2 * j + 1], f.y, a1);
        }
        float s = a0 + a1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s = ok ? s * scale : -CUDART_INF_F;
        // ---- online softmax update over the block's 4 steps (every lane of a step holds its score)
        float bm = fmaxf(s, __shfl_xor_sync(0xffffffffu, s, 8));
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
        const float mn = fmaxf(m, bm);
        const float corr = __expf(m - mn);
        const float pe = ok ? __expf(s - mn) : 0.f;
        float bl = pe + __shfl_xor_sync(0xffffffffu, pe, 8);
        bl += __shfl_xor_sync(0xffffffffu, bl, 16);
        l = l * corr + bl;
        m = mn;
        // ---- acc = acc * corr + p * V (this lane: its step, its 16 dims; the 4 steps are summed after the loop)
        const uint32_t vw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&vw[j]));
            acc[2 * j] = fmaf(pe, f.x, acc[2 * j] * corr);
            acc[2 * j + 1] = fmaf(pe, f.y, acc[2 * j + 1] * corr);
        }
    }
    // sum the four step groups of the warp; lanes with tg == 0 publish the warp's state
#pragma unroll
    for (int j = 0; j < 16; j++) {
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
    }
    if (lane == 0) wacc[w][0] = m, wacc[w][1] = l;
    if (tg == 0) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4 *>(&wacc[w][4 + 16 * dg + j]) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
    bar_consumers();
    if (w == 0) {
        // merge the warps' states + the current step (whose k, v are still in registers); lane owns dims 4*lane..4*lane+3
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) d = fmaf(qf[j], __half2float(kr[j]), d);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        const float s_cur = d * scale;
        float M = s_cur;
#pragma unroll
        for (int i = 0; i < NW; i++) M = fmaxf(M, wacc[i][0]);
        float Lsum = 0.f, a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int i = 0; i < NW; i++) {
            if (wacc[i][1] > 0.f) {
                const float c = __expf(wacc[i][0] - M);
                Lsum += wacc[i][1] * c;
                const float4 v = *reinterpret_cast<const float4 *>(&wacc[i][4 + 4 * lane]);
                a4[0] = fmaf(v.x, c, a4[0]), a4[1] = fmaf(v.y, c, a4[1]), a4[2] = fmaf(v.z, c, a4[2]), a4[3] = fmaf(v.w, c, a4[3]);
            }
        }
        const float pc = __expf(s_cur - M);
        Lsum += pc;
#pragma unroll
        for (int j = 0; j < 4; j++) a4[j] = fmaf(pc, __half2float(vn[j]), a4[j]);
        const float inv = 1.f / Lsum;
        __half r4[4];
#pragma unroll
        for (int j = 0; j < 4; j++) r4[j] = __float2half_rn(a4[j] * inv);
        const uint2 rv = *reinterpret_cast<uint2 *>(r4);
        ll_st16(static_cast<uint2 *>(jb.out) + ((h * HD + 4 * lane) >> 1), rv.x, rv.y, ep_out);
        if (jb.out_plain) *reinterpret_cast<uint2 *>(static_cast<__half *>(jb.out_plain) + h * HD + 4 * lane) = rv;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Weight producer (one thread of the producer warp calls it in a loop): advances a cursor over (job, stage) and issues the
// bulk copies of as many stages as the byte ring and the 32 barrier slots take right now, never blocking.  The cursor runs
// ahead of the consumers across job boundaries, so the next Linears' bit-planes (which do not depend on x) stream while a
// hand-over is in flight.  (Feeding the ring cooperatively from the consumer warps — 16 consumer warps, no producer warp — was
// measured 1.9x slower: the stream starves, profiles/r2_persist_variants.txt.)
// FIFO byte allocation: a stage takes `stage_bytes` contiguous ring bytes; stages are retired in issue order by testing
// their empty barriers (completed by the PK_EMPTY_COUNT arrivals of the consuming group).
// ------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void pk_produce() {
    if (*static_cast<volatile uint32_t *>(&g_sh.p_done)) return;
    if (atomicCAS(&g_sh.p_lock, 0u, 1u) != 0u) return;  // somebody else is producing
    __threadfence_block();
    volatile PkShared &vs = g_sh;
    uint32_t j = vs.p_job, s = vs.p_stage, q = vs.p_q, q_tail = vs.p_qtail, head = vs.p_head;
    const uint32_t n_jobs = vs.n_jobs, RB = vs.ring_bytes, ring0 = vs.ring0, bits = vs.bits;
    const PJob *jobs = g_sh.jobs_g;
    const uint64_t pol = l2_policy_evict_first();
    bool done = false;
    while (true) {
        if (j >= n_jobs) {
            done = true;
            break;
        }
        const PJob &jb = jobs[j];
        if (jb.type != PJ_GEMV) {
            j++, s = 0;
            continue;
        }
        uint32_t r_begin, r_end;
        pk_rows(jb, r_begin, r_end);
        const uint32_t RS = jb.rs, row_bytes = jb.K >> 3, size = jb.stage_bytes, N = jb.N;
        const uint32_t nstages = (r_end - r_begin + RS - 1) / RS;
        if (s >= nstages) {
            j++, s = 0;
            continue;
        }
        // FIFO allocation of `size` contiguous ring bytes + a free barrier slot, without blocking
        uint32_t off = 0;
        bool ok = false;
        while (true) {
            if (q - q_tail < PK_NB) {
                if (q == q_tail) {
                    off = 0, ok = true;  // nothing in flight: restart at the ring base
                } else {
                    const uint32_t tail = vs.stage_off[q_tail & (PK_NB - 1u)];
                    if (head > tail) {  // live bytes = [tail, head)
                        if (head + size <= RB) off = head, ok = true;
                        else if (size <= tail) off = 0, ok = true;
                    } else if (head + size <= tail) {  // live bytes wrap: free = [head, tail)
                        off = head, ok = true;
                    }
                }
            }
            if (ok) break;
            // retire the oldest stage in flight if its consumers are done with it; otherwise come back later
            if (!mbar_try(smem_u32(&g_sh.empty[q_tail & (PK_NB - 1u)]), (q_tail / PK_NB) & 1u)) break;
            q_tail++;
        }
        if (!ok) break;
        head = off + size;
        const uint32_t b = q & (PK_NB - 1u);
        vs.stage_off[b] = off;
        const uint32_t row0 = r_begin + s * RS;
        const uint32_t rows = min(RS, r_end - row0);
        const uint32_t bytes = rows * row_bytes;
        const uint32_t fb = smem_u32(&g_sh.full[b]);
        const uint8_t *Wp = static_cast<const uint8_t *>(jb.W);
        mbar_arrive_expect_tx(fb, bytes * bits);
        for (uint32_t pl = 0; pl < bits; pl++)
            bulk_g2s(ring0 + off + pl * RS * row_bytes, Wp + ((size_t)pl * N + row0) * row_bytes, bytes, fb, pol);
        q++, s++;
    }
    vs.p_job = j, vs.p_stage = s, vs.p_q = q, vs.p_qtail = q_tail, vs.p_head = head;
    if (done) vs.p_done = 1u;
    __threadfence_block();
    atomicExch(&g_sh.p_lock, 0u);
}

// ------------------------------------------------------------------------------------------------------------
// the kernel
// static smem: PkShared;  dynamic smem: [tables NCW x WTB | scratch 16 KB | x staging | ring]
// ------------------------------------------------------------------------------------------------------------
template <int BITS>
__global__ void __launch_bounds__(PK_THREADS, 1) decode_persistent_kernel(const PParams p) {
    constexpr int WTB = FastWarpTbl<BITS, 8>::BYTES;
    constexpr uint32_t scratch_rel = PK_NCW * WTB;
    const uint32_t smem0 = smem_u32(g_dyn);

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < PK_NB; s++) {
            mbar_init(smem_u32(&g_sh.full[s]), 1u);
            mbar_init(smem_u32(&g_sh.empty[s]), PK_EMPTY_COUNT);
        }
        mbar_fence_init();
        g_sh.err = p.err;
        g_sh.xs_bytes = p.xs_bytes;
        g_sh.scratch_rel = scratch_rel;
        g_sh.prof = nullptr;
        g_sh.jobs_g = p.jobs, g_sh.n_jobs = p.n_jobs, g_sh.ring_bytes = p.ring_bytes, g_sh.bits = BITS;
        g_sh.ring0 = smem0 + scratch_rel + PK_SCRATCH_BYTES + p.xs_bytes;
        g_sh.p_lock = 0u, g_sh.p_done = 0u, g_sh.p_job = 0u, g_sh.p_stage = 0u, g_sh.p_q = 0u, g_sh.p_qtail = 0u, g_sh.p_head = 0u;
    }
    __syncthreads();
    uint32_t ep;  // token counter of this launch (volatile read: written by the previous launch)
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(ep) : "l"(p.epoch) : "memory");
    ep += 1u;
    if (smem0 & 255u) pk_die(p.err, 6u);  // the table addressing (PRMT byte insert) needs a 256-byte aligned base

    if ((threadIdx.x >> 5) == PK_NCW) {
        // ===================== producer warp: one thread feeds the weight ring until every stage has been issued =====================
        if ((threadIdx.x & 31) == 0) {
            const long long t0 = clock64();
            while (!*static_cast<volatile uint32_t *>(&g_sh.p_done)) {
                pk_produce();
                __nanosleep(128);
                if (clock64() - t0 > 8 * PK_WATCHDOG_CYCLES) pk_die(p.err, 4u);
            }
        }
    } else {
        // ===================== consumers: the job list in order =====================
        // Loop state lives in shared memory and the job functions are not inlined: a callee only gets the registers its
        // caller leaves alone, and the row loops need all 96 (measured: four registers held across the call are enough
        // to make ptxas serialise every table lookup with its FMA).
        constexpr uint32_t NPC = sizeof(PJob) / 16;  // 16-byte pieces of a descriptor, one per thread
        if (threadIdx.x < NPC) reinterpret_cast<uint4 *>(&g_sh.jobs[0])[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(p.jobs) + threadIdx.x);
        if (threadIdx.x == 0) g_sh.tag_base = ep * p.n_jobs, g_sh.q_base[0] = 0;
#pragma unroll 1
        for (uint32_t j = 0; j < p.n_jobs; j++) {
            bar_consumers();  // descriptor j is in its slot; everyone is done with job j-1 (`red`, x staging, scratch, slot (j+1)&1)
            // the next descriptor goes straight into the other slot (an L1 hit: the producer thread walked the table already)
            if (threadIdx.x < NPC && j + 1 < p.n_jobs)
                reinterpret_cast<uint4 *>(&g_sh.jobs[(j + 1) & 1u])[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(p.jobs + j + 1) + threadIdx.x);
            if (threadIdx.x == 0) {
                g_sh.prof = p.prof ? p.prof + ((size_t)blockIdx.x * p.n_jobs + j) * 4 : nullptr;
                if (p.prof) g_sh.prof[0] = clock64();
            }
            const PJob &jb = g_sh.jobs[j & 1u];
            // stage numbering of the next job, written into the OTHER slot now: its previous readers (job j-1) are behind the
            // barrier above, its next readers (job j+1) behind the next one — no barrier of its own
            if (threadIdx.x == 0) {
                uint32_t nst = 0;
                if (jb.type == PJ_GEMV) {
                    uint32_t r_begin, r_end;
                    pk_rows(jb, r_begin, r_end);
                    nst = (r_end - r_begin + jb.rs - 1) / jb.rs;
                }
                g_sh.q_base[(j + 1) & 1u] = g_sh.q_base[j & 1u] + nst;
            }
            switch (jb.type) {
                case PJ_GEMV: pk_gemv_job<BITS>(j & 1u); break;
                case PJ_ATTN: pk_attn_job(j & 1u, p.pos); break;
                case PJ_PACK: pk_pack_job(j & 1u); break;
                case PJ_REDUCE: pk_reduce_job(j & 1u); break;
                default: break;
            }
            if (p.prof && threadIdx.x == 0) p.prof[((size_t)blockIdx.x * p.n_jobs + j) * 4 + 3] = clock64();
        }
        if (p.bump_epoch && threadIdx.x == 0) {  // the last CTA to finish advances the token counter (all have read it by then)
            if (atomicAdd(p.done, 1u) == gridDim.x - 1u) {
                *p.done = 0u;
                *p.epoch = ep;
            }
        }
    }
}

}  // namespace apg
