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
//   * grid = #SMs, 1 CTA/SM, 16 consumer warps + 1 producer warp; launched cooperatively (co-residency is required:
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

constexpr int PK_NCW = 16;                       // consumer warps per CTA
constexpr int PK_THREADS = (PK_NCW + 1) * 32;    // + 1 producer warp
constexpr uint32_t PK_NB = 32;                   // (full, empty) barrier pairs, used round-robin by stage number
constexpr uint32_t PK_EMPTY_COUNT = 720720u;     // lcm(1..16): a group of nwk warps arrives with 720720 / nwk each
constexpr uint32_t PK_SCRATCH_BYTES = 16384;     // per-chunk partial sums of a job / attention scratch
constexpr uint32_t PK_MAX_STAGE = 32768;
constexpr long long PK_WATCHDOG_CYCLES = 6000000000ll;  // ~3 s: a wait that long is a bug -> trap instead of hanging the GPU

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
    void *p0, *p1, *p2;   // ATTN: inv_freq fp32 [64], k_cache, v_cache fp16 [Hkv][S][128];  PACK: p0 = token id (int*) or NULL
    void *peer[8];        // PF_PUSH: every rank's uint2 [world][N] receive buffer of this site
};

struct PParams {
    const PJob *jobs;
    uint32_t n_jobs;
    uint32_t ring_bytes;
    uint32_t *epoch;     // device counter; packets of this launch carry *epoch + 1; bumped at the end if bump_epoch
    const int *pos;      // device position (attention)
    uint32_t *err;       // device word: non-zero = a watchdog fired (code in the low byte)
    uint32_t *done;      // device counter (zero between launches): the last CTA to finish bumps the epoch
    uint32_t bump_epoch;
};

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
// two packets (4 halfs) at packet index pk (even): spin until both carry the epoch
__device__ __forceinline__ uint2 ll_wait16(const uint2 *base, uint32_t pk, uint32_t ep, uint32_t *err) {
    const uint2 *p = base + pk;
    uint4 v = ll_ld16(p);
    if (v.y != ep || v.w != ep) {
        const long long t0 = clock64();
        do {
            __nanosleep(32);
            v = ll_ld16(p);
            if (clock64() - t0 > PK_WATCHDOG_CYCLES) pk_die(err, 1u);
        } while (v.y != ep || v.w != ep);
    }
    return make_uint2(v.x, v.z);
}
__device__ __forceinline__ uint32_t ll_wait8(const uint2 *base, uint32_t pk, uint32_t ep, uint32_t *err) {
    const uint2 *p = base + pk;
    uint2 v = ll_ld8(p);
    if (v.y != ep) {
        const long long t0 = clock64();
        do {
            __nanosleep(32);
            v = ll_ld8(p);
            if (clock64() - t0 > PK_WATCHDOG_CYCLES) pk_die(err, 2u);
        } while (v.y != ep);
    }
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
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity, uint32_t *err, uint32_t code, bool backoff) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (backoff) __nanosleep(256);
        if (clock64() - t0 > PK_WATCHDOG_CYCLES) pk_die(err, code);
    }
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 2, %0;" ::"n"(PK_NCW * 32) : "memory"); }

// rows [r_begin, r_end) of a job owned by this CTA (same dealing as gemv_fast_kernel: consecutive units)
__device__ __forceinline__ void pk_rows(const PJob &jb, uint32_t &r_begin, uint32_t &r_end) {
    const uint32_t u_begin = blockIdx.x * jb.units_q + min(blockIdx.x, jb.units_rem);
    const uint32_t nunits = jb.units_q + (blockIdx.x < jb.units_rem ? 1u : 0u);
    r_begin = min(u_begin * jb.unit_rows, jb.N);
    r_end = min((u_begin + nunits) * jb.unit_rows, jb.N);
}

// static shared memory: addresses are compile-time constants, so nothing here costs a register; the consumers read the
// fields of their current job from here whenever they need them instead of keeping them live
struct PkShared {
    uint64_t full[PK_NB], empty[PK_NB];
    uint32_t stage_off[PK_NB];  // ring offset of the stage in each barrier slot (written by the producer)
    float ssq[PK_NCW];
    PJob job;                   // the consumers' current job (copied from global memory once per job)
    uint32_t *err;
};

static_assert(sizeof(PkShared) <= 1024, "the static shared header must stay within the 1 KB the host leaves for it");
static_assert(sizeof(PJob) % 16 == 0, "PJob is copied in 16-byte pieces");

struct PCtx {
    uint32_t tbl0, ring0;  // shared addresses of the codebook tables and of the ring
    float *scratch;
    uint32_t tag_base;     // epoch * n_jobs
    int lane, warp;
};

// ------------------------------------------------------------------------------------------------------------
// GEMV job, consumer side
// ------------------------------------------------------------------------------------------------------------
template <int BITS, int CPW, int RS, bool GLU>
__device__ __forceinline__ void pk_gemv_job(PkShared &sh, const PCtx &cx, uint32_t q_base) {
    constexpr int WTB = FastWarpTbl<BITS, 8>::BYTES;
    using Tb = Tables<BITS, RS>;
    const PJob &jb = sh.job;
    const int lane = cx.lane, warp = cx.warp;
    uint32_t r_begin, r_end;
    pk_rows(jb, r_begin, r_end);
    const uint32_t nrows = r_end - r_begin;
    const uint32_t nstages = (nrows + RS - 1) / RS;
    float *red = cx.scratch;
    const bool active = (uint32_t)warp < jb.groups * jb.nwk && nrows > 0;  // a CTA without rows only passes the barriers

    if (active) {
        const uint32_t nwk = jb.nwk;
        const uint32_t g = ((uint32_t)warp * jb.inv_nwk) >> 16, wk = warp - g * nwk;
        const uint32_t tbl = cx.tbl0 + warp * WTB;
        const uint32_t K = jb.K;
        const uint32_t nchunk = (K + 1023u) >> 10;
        const bool do_norm = CPW == 1 && (jb.flags & PF_NORM) != 0;
        const uint32_t ep = cx.tag_base + jb.tag_x;
        const uint2 *xin = static_cast<const uint2 *>(jb.x);

        // codebook rows of this group's first stage: static data, fetched before x is waited for
        typename Tb::Regs lr;
        if (g < nstages) Tb::fetch(lr, static_cast<const __half *>(jb.lut), r_begin + g * RS, jb.N, lane);

        // ---- x -> registers from the producer job's packets (spin until they carry this token's tag)
        uint32_t xr[CPW][16];
        bool act[CPW];
        float ss = 0.f;
#pragma unroll
        for (int cc = 0; cc < CPW; cc++) {
            const uint32_t i = wk * CPW + cc;
            const uint32_t eff = (i < nchunk) ? chunk_eff(K, i) : 0u;
            act[cc] = (uint32_t)lane < eff;
            // lane 0 of the warp spins first, so that a CTA that is early polls with 16 loads, not 512
            if (act[cc] && lane == 0) (void)ll_wait16(xin, (i * 1024u) >> 1, ep, sh.err);
        }
        __syncwarp();
#pragma unroll
        for (int cc = 0; cc < CPW; cc++) {
            const uint32_t i = wk * CPW + cc;
            const uint32_t eff = (i < nchunk) ? chunk_eff(K, i) : 0u;
            if (act[cc]) {
#pragma unroll
                for (int c2 = 0; c2 < 4; c2 += 2) {  // four 16-byte loads in flight
                    uint4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) v[u] = ll_ld16(xin + ((i * 1024u + (c2 + (u >> 1)) * 8u * eff + 8u * lane) >> 1) + 2u * (u & 1));
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        uint2 d = make_uint2(v[u].x, v[u].z);
                        if (v[u].y != ep || v[u].w != ep)
                            d = ll_wait16(xin, ((i * 1024u + (c2 + (u >> 1)) * 8u * eff + 8u * lane) >> 1) + 2u * (u & 1), ep, sh.err);
                        xr[cc][4 * (c2 + (u >> 1)) + 2 * (u & 1) + 0] = d.x;
                        xr[cc][4 * (c2 + (u >> 1)) + 2 * (u & 1) + 1] = d.y;
                    }
                }
                if (do_norm) {
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&xr[cc][e]));
                        ss = fmaf(f.x, f.x, ss);
                        ss = fmaf(f.y, f.y, ss);
                    }
                }
            }
        }
        if (do_norm) {
            // fused RMSNorm (model.py:280-285): the group's warps exchange their sums of squares (fixed summation order)
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0) sh.ssq[warp] = ss;
            asm volatile("bar.sync 1, %0;" ::"r"(jb.groups * nwk * 32u) : "memory");
            float tot = 0.f;
            for (uint32_t w = 0; w < nwk; w++) tot += sh.ssq[g * nwk + w];
            const float rs = rsqrtf(tot / (float)K + jb.eps);
            if (act[0]) {
                const uint32_t eff = chunk_eff(K, wk);
                const __half *nw = static_cast<const __half *>(jb.norm_w) + wk * 1024u + 8u * lane;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint4 wv = __ldg(reinterpret_cast<const uint4 *>(nw + c * 8u * eff));
                    const uint32_t wn[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&xr[0][4 * c + e]));
                        const __half2 n = __floats2half2_rn(f.x * rs, f.y * rs);                    // .type_as(x)
                        const __half2 r = __hmul2(n, *reinterpret_cast<const __half2 *>(&wn[e]));  // * weight
                        xr[0][4 * c + e] = *reinterpret_cast<const uint32_t *>(&r);
                    }
                }
            }
        }

        // ---- stages of this group
        const uint32_t G = jb.groups;
        for (uint32_t s = g; s < nstages; s += G) {
            const uint32_t q = q_base + s, b = q & (PK_NB - 1u), par = (q / PK_NB) & 1u;
            const uint32_t row0 = r_begin + s * RS;
            const uint32_t rows = min((uint32_t)RS, r_end - row0);
            __syncwarp();
            Tb::store(lr, tbl, lane);
            Tb::fetch(lr, static_cast<const __half *>(jb.lut), row0 + G * RS, jb.N, lane);
            __syncwarp();
            mbar_wait_wd(smem_u32(&sh.full[b]), par, sh.err, 3u, false);
            const uint32_t stage = cx.ring0 + sh.stage_off[b];
            const uint32_t row_bytes = K >> 3;
            float acc[RS];
#pragma unroll
            for (int r = 0; r < RS; r++) acc[r] = 0.f;
            if (RS >= 2 && rows <= (uint32_t)(RS / 2)) {
#pragma unroll
                for (int cc = 0; cc < CPW; cc++)
                    if (act[cc])
                        RowLoop<BITS, RS, 0, (RS >= 2 ? RS / 2 : RS)>::run(acc, stage + ((wk * CPW + cc) * 32u + lane) * 4u, row_bytes,
                                                                           RS * row_bytes, xr[cc], tbl);
            } else {
#pragma unroll
                for (int cc = 0; cc < CPW; cc++)
                    if (act[cc])
                        RowLoop<BITS, RS, 0, RS>::run(acc, stage + ((wk * CPW + cc) * 32u + lane) * 4u, row_bytes, RS * row_bytes, xr[cc], tbl);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_cnt(smem_u32(&sh.empty[b]), PK_EMPTY_COUNT / nwk);
            const float v = BatchReduce<RS>::run(acc, lane);
            const int rl = lane >> BatchReduce<RS>::SH;
            if ((lane & ((1 << BatchReduce<RS>::SH) - 1)) == 0 && (uint32_t)rl < rows) red[(row0 - r_begin + rl) * nwk + wk] = v;
        }
    }

    bar_consumers();  // every chunk partial of every row of this CTA is in `red`
    // ---- epilogue: fixed-order combination (deterministic), fused residual / SwiGLU, packets out
    const uint32_t tid = threadIdx.x, nwk = jb.nwk;
    const uint32_t ep_out = cx.tag_base + jb.tag_out;
    if constexpr (GLU) {
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
                const uint32_t rv = ll_wait8(resp, pk, cx.tag_base + jb.tag_res, sh.err);
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
__device__ __forceinline__ void pk_reduce_job(PkShared &sh, const PCtx &cx) {
    const PJob &jb = sh.job;
    const uint32_t N = jb.N, world = jb.world;
    const uint32_t ep = cx.tag_base + jb.tag_x, ep_out = cx.tag_base + jb.tag_out, ep_res = cx.tag_base + jb.tag_res;
    const uint2 *recv = static_cast<const uint2 *>(jb.x);
    const uint2 *resp = static_cast<const uint2 *>(jb.residual);
    uint2 *outp = static_cast<uint2 *>(jb.out);
    for (uint32_t pk = blockIdx.x * (PK_NCW * 32u) + threadIdx.x; pk < (N >> 1); pk += gridDim.x * (PK_NCW * 32u)) {
        float s0 = 0.f, s1 = 0.f;
        for (uint32_t w = 0; w < world; w++) {
            const uint2 v = ll_wait16(recv, w * N + 2u * pk, ep, sh.err);
            s0 += __uint_as_float(v.x), s1 += __uint_as_float(v.y);
        }
        __half2 h = __floats2half2_rn(s0, s1);
        if (jb.flags & PF_RESIDUAL) {
            const uint32_t rv = ll_wait8(resp, pk, ep_res, sh.err);
            h = __hadd2(h, *reinterpret_cast<const __half2 *>(&rv));
        }
        const uint32_t hv = *reinterpret_cast<const uint32_t *>(&h);
        ll_st8(outp + pk, hv, ep_out);
        if (jb.out_plain) reinterpret_cast<uint32_t *>(jb.out_plain)[pk] = hv;
    }
}

// plain fp16 row -> packets (embedding row, model.py:123; or the chain benchmark's input vector)
__device__ __forceinline__ void pk_pack_job(PkShared &sh, const PCtx &cx) {
    const PJob &jb = sh.job;
    const uint32_t n = jb.N;
    uint32_t row = 0;
    if (jb.p0) row = min((uint32_t)max(*static_cast<const volatile int *>(jb.p0), 0), jb.a0 - 1u);  // never read outside the table
    const uint32_t *src = reinterpret_cast<const uint32_t *>(static_cast<const __half *>(jb.x) + (size_t)row * n);
    uint2 *outp = static_cast<uint2 *>(jb.out);
    for (uint32_t i = blockIdx.x * (PK_NCW * 32u) + threadIdx.x; i < (n >> 1); i += gridDim.x * (PK_NCW * 32u)) {
        const uint32_t v = __ldcg(src + i);
        ll_st8(outp + i, v, cx.tag_base + jb.tag_out);
        if (jb.out_plain) reinterpret_cast<uint32_t *>(jb.out_plain)[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------
// attention job: CTA h < H attends for head h with its 16 consumer warps (RoPE + KV append + softmax(q.K^T).V,
// Attention.forward model.py:206-236; same arithmetic as apd::attn_decode_kernel, decode_kernels.cuh, in blocks of 8
// cached steps per warp so that the loads in flight fit the 96-register budget of this kernel:
//   scores: 4 lanes per step, each 32 dims of the K row (4 x 16 B);  P.V: lane = (step mod 4, 16-dim group), 2 passes)
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

__device__ __noinline__ void pk_attn_job(PkShared &sh, const PCtx &cx, const int *pos_ptr) {
    constexpr int HD = 128, NW = PK_NCW, BS = 8;  // BS cached steps per warp iteration
    const PJob &jb = sh.job;
    const uint32_t H = jb.a0, Hkv = jb.a1, S = jb.a2;
    if (blockIdx.x >= H) return;  // uniform per CTA: the other CTAs move on to the next job
    const float scale = jb.f0;
    const int lane = cx.lane, w = cx.warp;
    const uint32_t G = H / Hkv, h = blockIdx.x, kvh = h / G;
    const uint32_t ep = cx.tag_base + jb.tag_x, ep_out = cx.tag_base + jb.tag_out;
    float *qs = cx.scratch;                                             // [128] roped q (fp32)
    float(*sc)[BS] = reinterpret_cast<float(*)[BS]>(cx.scratch + HD);   // [NW][8] probabilities of the block in flight
    float(*wacc)[HD + 4] = reinterpret_cast<float(*)[HD + 4]>(cx.scratch + HD + NW * BS);  // [NW][132]: (m, l, -, -, acc[128])
    __half *k_cache = static_cast<__half *>(jb.p1), *v_cache = static_cast<__half *>(jb.p2);
    const uint2 *qkv = static_cast<const uint2 *>(jb.x);
    int pos = *static_cast<const volatile int *>(pos_ptr);
    const bool pos_ok = pos >= 0 && pos < (int)S;  // the host API refuses to step past the cache; never write outside it
    if (!pos_ok) pos = 0;

    __half kr[4], vn[4];
    float qf[4];
    if (w == 0) {
        const float *inv_freq = static_cast<const float *>(jb.p0);
        const float fpos = (float)pos;
        __half rc[4], rs[4];  // cos / sin in fp32, rounded to fp16 (model.py:396-405), shared by q and k
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float sn, cs;
            sincosf(fpos * inv_freq[(4 * lane + j) & 63], &sn, &cs);
            rc[j] = __float2half_rn(cs), rs[j] = __float2half_rn(sn);
        }
        __half q[4], kn[4], qp[4], kp[4], qr[4];
        const uint2 qv = ll_wait16(qkv, (h * HD + 4 * lane) >> 1, ep, sh.err);
        const uint2 kv = ll_wait16(qkv, ((H + kvh) * HD + 4 * lane) >> 1, ep, sh.err);
        const uint2 vv = ll_wait16(qkv, ((H + Hkv + kvh) * HD + 4 * lane) >> 1, ep, sh.err);
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
    const int nblk = (pos + BS - 1) / BS;  // blocks of cached steps t < pos
    const int st = lane >> 2, qd = lane & 3;   // scores: step within the block, 32-dim quarter
    const int tg = lane >> 3, dg = lane & 7;   // P.V:    step mod 4, 16-dim group
    float m = -CUDART_INF_F, l = 0.f, acc[16];
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0.f;
    for (int b = w; b < nblk; b += NW) {
        const int tb = b * BS, nt = min(BS, pos - tb);
        // ---- all loads of the block first (one memory round trip): this lane's quarter K row and its two V pieces
        uint4 kv[4], vv[2][2];
        if (st < nt) {
            const uint4 *kr4 = reinterpret_cast<const uint4 *>(Kb + (size_t)(tb + st) * HD + 32 * qd);
#pragma unroll
            for (int i = 0; i < 4; i++) kv[i] = __ldcg(kr4 + i);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int tt = 4 * u + tg;
            vv[u][0] = vv[u][1] = make_uint4(0, 0, 0, 0);
            if (tt < nt) {
                const uint4 *vr = reinterpret_cast<const uint4 *>(Vb + (size_t)(tb + tt) * HD + 16 * dg);
                vv[u][0] = __ldcg(vr), vv[u][1] = __ldcg(vr + 1);
            }
        }
        // ---- scores
        float a0 = 0.f, a1 = 0.f;
        if (st < nt) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 qa = *reinterpret_cast<const float4 *>(qs + 32 * qd + 8 * i);
                const float4 qb = *reinterpret_cast<const float4 *>(qs + 32 * qd + 8 * i + 4);
                const float2 k0 = __half22float2(*reinterpret_cast<const __half2 *>(&kv[i].x));
                const float2 k1 = __half22float2(*reinterpret_cast<const __half2 *>(&kv[i].y));
                const float2 k2 = __half22float2(*reinterpret_cast<const __half2 *>(&kv[i].z));
                const float2 k3 = __half22float2(*reinterpret_cast<const __half2 *>(&kv[i].w));
                a0 = fmaf(qa.x, k0.x, a0), a1 = fmaf(qa.y, k0.y, a1);
                a0 = fmaf(qa.z, k1.x, a0), a1 = fmaf(qa.w, k1.y, a1);
                a0 = fmaf(qb.x, k2.x, a0), a1 = fmaf(qb.y, k2.y, a1);
                a0 = fmaf(qb.z, k3.x, a0), a1 = fmaf(qb.w, k3.y, a1);
            }
        }
        float s = a0 + a1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s = (st < nt) ? s * scale : -CUDART_INF_F;
        // ---- online softmax update (every lane of a step's quad holds the same score)
        float bm = s;
#pragma unroll
        for (int o = 16; o >= 4; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
        const float mn = fmaxf(m, bm);
        const float corr = __expf(m - mn);
        const float pe = (st < nt) ? __expf(s - mn) : 0.f;
        float bl = pe;
#pragma unroll
        for (int o = 16; o >= 4; o >>= 1) bl += __shfl_xor_sync(0xffffffffu, bl, o);
        l = l * corr + bl;
        m = mn;
        __syncwarp();
        if (qd == 0) sc[w][st] = pe;
        __syncwarp();
        // ---- acc = acc * corr + P.V of the block
#pragma unroll
        for (int j = 0; j < 16; j++) acc[j] *= corr;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int tt = 4 * u + tg;
            const float pw = (tt < nt) ? sc[w][tt] : 0.f;
            const uint32_t wv[8] = {vv[u][0].x, vv[u][0].y, vv[u][0].z, vv[u][0].w, vv[u][1].x, vv[u][1].y, vv[u][1].z, vv[u][1].w};
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&wv[j]));
                acc[2 * j] = fmaf(pw, f.x, acc[2 * j]);
                acc[2 * j + 1] = fmaf(pw, f.y, acc[2 * j + 1]);
            }
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
#pragma unroll 4
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
// the kernel
// static smem: PkShared;  dynamic smem: [tables NCW x WTB | scratch 16 KB | ring]
// ------------------------------------------------------------------------------------------------------------
template <int BITS>
__global__ void __launch_bounds__(PK_THREADS, 1) decode_persistent_kernel(const PParams p) {
    constexpr int WTB = FastWarpTbl<BITS, 8>::BYTES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(16) PkShared sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t smem0 = smem_u32(smem_raw);
    constexpr uint32_t scratch_rel = PK_NCW * WTB;
    constexpr uint32_t ring_rel = scratch_rel + PK_SCRATCH_BYTES;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < PK_NB; s++) {
            mbar_init(smem_u32(&sh.full[s]), 1u);
            mbar_init(smem_u32(&sh.empty[s]), PK_EMPTY_COUNT);
        }
        mbar_fence_init();
        sh.err = p.err;
    }
    __syncthreads();
    uint32_t ep;  // token counter of this launch (volatile read: written by the previous launch)
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(ep) : "l"(p.epoch) : "memory");
    ep += 1u;
    if (smem0 & 255u) pk_die(p.err, 6u);  // the table addressing (PRMT byte insert) needs a 256-byte aligned base

    if (warp == PK_NCW) {
        // ===================== producer: stream this CTA's rows of every GEMV job through the byte ring =====================
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            const uint32_t RB = p.ring_bytes;
            uint32_t q = 0, q_tail = 0, head = 0;
            for (uint32_t j = 0; j < p.n_jobs; j++) {
                const PJob &jb = p.jobs[j];
                if (jb.type != PJ_GEMV) continue;
                uint32_t r_begin, r_end;
                pk_rows(jb, r_begin, r_end);
                const uint32_t RS = jb.rs, row_bytes = jb.K >> 3, size = jb.stage_bytes, N = jb.N;
                const uint8_t *Wp = static_cast<const uint8_t *>(jb.W);
                const uint32_t nstages = (r_end - r_begin + RS - 1) / RS;
                for (uint32_t s = 0; s < nstages; s++) {
                    uint32_t off = 0;
                    while (true) {  // FIFO allocation of `size` contiguous ring bytes + a free barrier slot
                        bool ok = false;
                        if (q - q_tail < PK_NB) {
                            if (q == q_tail) {
                                off = 0, ok = true;  // nothing in flight: restart at the ring base
                            } else {
                                const uint32_t tail = sh.stage_off[q_tail & (PK_NB - 1u)];
                                if (head > tail) {  // live bytes = [tail, head)
                                    if (head + size <= RB) off = head, ok = true;
                                    else if (size <= tail) off = 0, ok = true;
                                } else if (head + size <= tail) {  // live bytes wrap: free = [head, tail)
                                    off = head, ok = true;
                                }
                            }
                        }
                        if (ok) break;
                        // retire the oldest stage in flight (its consumers arrive on the empty barrier when done)
                        mbar_wait_wd(smem_u32(&sh.empty[q_tail & (PK_NB - 1u)]), (q_tail / PK_NB) & 1u, p.err, 4u, true);
                        q_tail++;
                    }
                    head = off + size;
                    const uint32_t b = q & (PK_NB - 1u);
                    sh.stage_off[b] = off;
                    const uint32_t row0 = r_begin + s * RS;
                    const uint32_t rows = min(RS, r_end - row0);
                    const uint32_t bytes = rows * row_bytes;
                    const uint32_t fb = smem_u32(&sh.full[b]);
                    mbar_arrive_expect_tx(fb, bytes * BITS);
#pragma unroll
                    for (int pl = 0; pl < BITS; pl++)
                        bulk_g2s(smem0 + ring_rel + off + pl * RS * row_bytes, Wp + ((size_t)pl * N + row0) * row_bytes, bytes, fb, pol);
                    q++;
                }
            }
        }
    } else {
        // ===================== consumers: the job list in order =====================
        PCtx cx;
        cx.tbl0 = smem0, cx.ring0 = smem0 + ring_rel;
        cx.scratch = reinterpret_cast<float *>(smem_raw + scratch_rel);
        cx.tag_base = ep * p.n_jobs, cx.lane = lane, cx.warp = warp;
        uint32_t q_base = 0;
        for (uint32_t j = 0; j < p.n_jobs; j++) {
            // the job descriptor -> shared memory (constant addresses): 14 threads x 16 bytes
            bar_consumers();  // everyone is done with the previous job (its descriptor, `red`, the attention scratch)
            if (threadIdx.x < sizeof(PJob) / 16)
                reinterpret_cast<uint4 *>(&sh.job)[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(p.jobs + j) + threadIdx.x);
            bar_consumers();
            switch (sh.job.type) {
                case PJ_GEMV: {
                    const uint32_t key = sh.job.cpw * 16u + sh.job.rs + ((sh.job.flags & PF_GLU) ? 256u : 0u);
                    switch (key) {
                        case 16 + 8: pk_gemv_job<BITS, 1, 8, false>(sh, cx, q_base); break;
                        case 16 + 4: pk_gemv_job<BITS, 1, 4, false>(sh, cx, q_base); break;
                        case 16 + 2: pk_gemv_job<BITS, 1, 2, false>(sh, cx, q_base); break;
                        case 32 + 8: pk_gemv_job<BITS, 2, 8, false>(sh, cx, q_base); break;
                        case 32 + 4: pk_gemv_job<BITS, 2, 4, false>(sh, cx, q_base); break;
                        case 32 + 2: pk_gemv_job<BITS, 2, 2, false>(sh, cx, q_base); break;
                        case 256 + 16 + 8: pk_gemv_job<BITS, 1, 8, true>(sh, cx, q_base); break;
                        default: pk_die(p.err, 5u);
                    }
                    uint32_t r_begin, r_end;
                    pk_rows(sh.job, r_begin, r_end);
                    q_base += (r_end - r_begin + sh.job.rs - 1) / sh.job.rs;
                    break;
                }
#ifndef PK_NO_ATTN
                case PJ_ATTN: pk_attn_job(sh, cx, p.pos); break;
#endif
                case PJ_PACK: pk_pack_job(sh, cx); break;
                case PJ_REDUCE: pk_reduce_job(sh, cx); break;
                default: break;
            }
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
