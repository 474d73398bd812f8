// Decode-step kernels around the Any-Precision GEMV: the non-Linear ops of the reference's
// gpt-fast Transformer at batch 1, sequence 1 (inference/model.py:121-131, 151-167, 206-236, 268-285, 381-405;
// sampling inference/generate.py:55-73).  RMSNorm, SiLU*mul and the residual adds are fused into the GEMV
// (apgemv_fast.cuh); what remains is: embedding row copy, RoPE + KV-cache append + attention, the fp16 lm_head
// GEMV (with the final RMSNorm fused) and greedy sampling.  Everything reads the token id and the position from
// DEVICE memory so that one CUDA graph replays for every token.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "apgemv_common.cuh"

namespace apd {

// ------------------------------------------------------------------------------------------------------------
// x = tok_embeddings[token]   (model.py:123)
// ------------------------------------------------------------------------------------------------------------
__global__ void embed_kernel(const __half *__restrict__ emb, const int *__restrict__ token, __half *__restrict__ x,
                             uint32_t dim, uint32_t vocab) {
    apg::pdl_wait_prior_grid();
    apg::pdl_launch_dependents();  // the next kernel's weight prefetch may start now; it waits for us before reading x
    const uint32_t tok = min((uint32_t)max(*token, 0), vocab - 1u);  // never read outside the table
    const uint4 *src = reinterpret_cast<const uint4 *>(emb + (size_t)tok * dim);
    uint4 *dst = reinterpret_cast<uint4 *>(x);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < dim / 8; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------------------
// RoPE + KV append + attention for one new token (Attention.forward, model.py:206-236).
//   qkv      fp16 [(H + 2*Hkv) * 128]   output of the fused wqkv Linear: q | k | v          (model.py:211)
//   inv_freq fp32 [64]                  1 / base^(2i/128)                                    (LlamaRotaryEmbedding)
//   k_cache, v_cache fp16 [Hkv, S, 128]                                                      (KVCache, model.py:63-79)
//   out      fp16 [H * 128]             attention output, input of wo
// head_dim is fixed at 128: in the 4-dims-per-lane layout a lane owns dims 4*lane..4*lane+3.
// cos/sin are computed in fp32 and rounded to fp16, and q*cos + rotate_half(q)*sin is evaluated in fp16 exactly as the
// reference's half tensors do (model.py:309-314, 396-405); scores, softmax and P.V accumulate in fp32.
// nsplit > 1: each (head, split) writes an un-normalised partial (m, l, acc[128]) and attn_merge_kernel combines them.
// ------------------------------------------------------------------------------------------------------------
constexpr int kHeadDim = 128;
constexpr int kPartStride = kHeadDim + 4;

__device__ __forceinline__ void rope4(const __half (&v)[4], const __half (&partner)[4], int lane, const __half (&c)[4],
                                      const __half (&s)[4], __half (&o)[4]) {
    // dims d = 4*lane + j; rotate_half: d < 64 -> -x[d+64], d >= 64 -> x[d-64]  (model.py:268-272)
    const bool hi = lane >= 16;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const __half rot = hi ? partner[j] : __hneg(partner[j]);
        o[j] = __hadd(__hmul(v[j], c[j]), __hmul(rot, s[j]));
    }
}

// grid = (H, nsplit), block = 32 * kAttnWarps.  The cached time steps are cut into blocks of 32 which are dealt
// round-robin to the nsplit * kAttnWarps workers (warps), so the work is balanced at every position.  Per block a warp
// runs three phases, all with many independent loads in flight (a first version walked the steps one by one and was
// bound by one L2 round trip per step, 0.44 us/step):
//   1. scores: lane <-> time step; each lane reads a whole K row (16 x 16 B, independent) against q held in smem
//   2. running softmax state (m, l) per warp
//   3. P.V: lane = (step mod 4, 16-dim group); 4 steps per instruction, 4x unrolled
// The warps' (m, l, acc[128]) meet in shared memory; warp 0 merges them, adds the current step (whose k, v are still in
// registers: the cache row written by this launch is never read by it) and writes the head's output or split partial.
constexpr int kAttnWarps = 8;

__global__ void __launch_bounds__(32 * kAttnWarps) attn_decode_kernel(
    const __half *__restrict__ qkv, const float *__restrict__ inv_freq, __half *__restrict__ k_cache,
    __half *__restrict__ v_cache, const int *__restrict__ pos_ptr, __half *__restrict__ out, float *__restrict__ part,
    uint32_t H, uint32_t Hkv, uint32_t S, float scale) {
    __shared__ __align__(16) float qs[kHeadDim];                    // roped q (fp32)
    __shared__ __align__(16) float sc[kAttnWarps][32];              // scores / probabilities of the block in flight
    __shared__ __align__(16) float wacc[kAttnWarps][kHeadDim + 4];  // per-warp (m, l, -, -, acc[128])
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t G = H / Hkv, h = blockIdx.x, kvh = h / G, split = blockIdx.y, nsplit = gridDim.y;
    float fr4[4];
#pragma unroll
    for (int j = 0; j < 4; j++) fr4[j] = inv_freq[(4 * lane + j) & 63];  // static table: before the dependency wait
    apg::pdl_wait_prior_grid();
    apg::pdl_launch_dependents();  // wo's weight stream may start while we attend; it waits for us before reading `out`
    // the new q / k / v rows and the position are fetched together (one L2 round trip, not two): none of the addresses
    // depends on the position
    uint2 qv = make_uint2(0u, 0u), kvn = qv, vvn = qv;
    if (w == 0) {
        qv = *reinterpret_cast<const uint2 *>(qkv + (size_t)h * kHeadDim + 4 * lane);
        kvn = *reinterpret_cast<const uint2 *>(qkv + (size_t)(H + kvh) * kHeadDim + 4 * lane);
        vvn = *reinterpret_cast<const uint2 *>(qkv + (size_t)(H + Hkv + kvh) * kHeadDim + 4 * lane);
    }
    const int pos = *pos_ptr;
    if (pos < 0 || pos >= (int)S) return;  // cache full: never write outside it (the host API refuses to get here)

    // warp 0: RoPE of q and of the new k, KV append, q -> smem
    __half kr[4], vn[4];
    float qf[4];
    if (w == 0) {
        const float fpos = (float)pos;
        __half rc[4], rs[4];  // cos / sin in fp32, rounded to fp16 (model.py:396-405), shared by q and k
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float sn, cs;
            sincosf(fpos * fr4[j], &sn, &cs);
            rc[j] = __float2half_rn(cs), rs[j] = __float2half_rn(sn);
        }
        __half q[4], kn[4], qp[4], kp[4], qr[4];
        *reinterpret_cast<uint2 *>(q) = qv, *reinterpret_cast<uint2 *>(kn) = kvn, *reinterpret_cast<uint2 *>(vn) = vvn;
        uint2 qpv, kpv;
        qpv.x = __shfl_xor_sync(0xffffffffu, qv.x, 16), qpv.y = __shfl_xor_sync(0xffffffffu, qv.y, 16);
        kpv.x = __shfl_xor_sync(0xffffffffu, kvn.x, 16), kpv.y = __shfl_xor_sync(0xffffffffu, kvn.y, 16);
        *reinterpret_cast<uint2 *>(qp) = qpv, *reinterpret_cast<uint2 *>(kp) = kpv;
        rope4(q, qp, lane, rc, rs, qr);
        rope4(kn, kp, lane, rc, rs, kr);
        if (split == 0 && h == kvh * G) {  // cache append (KVCache.update, model.py:70-79), once per KV head
            *reinterpret_cast<uint2 *>(k_cache + ((size_t)kvh * S + pos) * kHeadDim + 4 * lane) = *reinterpret_cast<uint2 *>(kr);
            *reinterpret_cast<uint2 *>(v_cache + ((size_t)kvh * S + pos) * kHeadDim + 4 * lane) = *reinterpret_cast<uint2 *>(vn);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) qf[j] = __half2float(qr[j]);
        *reinterpret_cast<float4 *>(qs + 4 * lane) = make_float4(qf[0], qf[1], qf[2], qf[3]);
    }
    __syncthreads();

    const __half *Kb = k_cache + (size_t)kvh * S * kHeadDim;
    const __half *Vb = v_cache + (size_t)kvh * S * kHeadDim;
    const int nblk = (pos + 31) >> 5;  // blocks of cached steps t < pos
    const int tg = lane >> 3, dg = lane & 7;
    float m = -CUDART_INF_F, l = 0.f, acc[16];
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0.f;

    for (int b = (int)(split * kAttnWarps) + w; b < nblk; b += (int)(nsplit * kAttnWarps)) {
        const int tb = b << 5, nt = min(32, pos - tb);
        // ---- issue ALL loads of the block first: this lane's K row (phase 1) and its 8 V half-rows (phase 3), so the
        //      block costs one memory round trip instead of two
        uint4 kv[16];
        if (lane < nt) {
            const uint4 *kr4 = reinterpret_cast<const uint4 *>(Kb + (size_t)(tb + lane) * kHeadDim);
#pragma unroll
            for (int i = 0; i < 16; i++) kv[i] = kr4[i];
        }
        uint4 va[8], vb[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int tt = 4 * u + tg;
            va[u] = vb[u] = make_uint4(0, 0, 0, 0);
            if (tt < nt) {
                const uint4 *vr = reinterpret_cast<const uint4 *>(Vb + (size_t)(tb + tt) * kHeadDim + 16 * dg);
                va[u] = vr[0], vb[u] = vr[1];
            }
        }
        // ---- phase 1: scores of the block
        float s = -CUDART_INF_F;
        if (lane < nt) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const float4 qa = *reinterpret_cast<const float4 *>(qs + 8 * i);
                const float4 qb = *reinterpret_cast<const float4 *>(qs + 8 * i + 4);
                const float2 k0 = __half22float2(*reinterpret_cast<const __half2 *>(&kv[i].x));
                const float2 k1 = __half22float2(*reinterpret_cast<const __half2 *>(&kv[i].y));
                const float2 k2 = __half22float2(*reinterpret_cast<const __half2 *>(&kv[i].z));
                const float2 k3 = __half22float2(*reinterpret_cast<const __half2 *>(&kv[i].w));
                a0 = fmaf(qa.x, k0.x, a0), a1 = fmaf(qa.y, k0.y, a1);
                a0 = fmaf(qa.z, k1.x, a0), a1 = fmaf(qa.w, k1.y, a1);
                a0 = fmaf(qb.x, k2.x, a0), a1 = fmaf(qb.y, k2.y, a1);
                a0 = fmaf(qb.z, k3.x, a0), a1 = fmaf(qb.w, k3.y, a1);
            }
            s = (a0 + a1) * scale;
        }
        // ---- phase 2: online softmax update
        float bm = s;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
        const float mn = fmaxf(m, bm);
        const float corr = __expf(m - mn);
        const float pe = (lane < nt) ? __expf(s - mn) : 0.f;
        float bl = pe;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) bl += __shfl_xor_sync(0xffffffffu, bl, o);
        l = l * corr + bl;
        m = mn;
        __syncwarp();
        sc[w][lane] = pe;
        __syncwarp();
        // ---- phase 3: acc = acc * corr + P.V of the block (V already in registers)
#pragma unroll
        for (int j = 0; j < 16; j++) acc[j] *= corr;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int tt = 4 * u + tg;
            const float pw = (tt < nt) ? sc[w][tt] : 0.f;
            const uint32_t wv[8] = {va[u].x, va[u].y, va[u].z, va[u].w, vb[u].x, vb[u].y, vb[u].z, vb[u].w};
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
    __syncthreads();
    if (w != 0) return;

    // ---- warp 0: merge the warps' states (+ the current step on split 0); lane owns dims 4*lane..4*lane+3
    float s_cur = -CUDART_INF_F;
    if (split == 0) {
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) d = fmaf(qf[j], __half2float(kr[j]), d);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        s_cur = d * scale;
    }
    float M = s_cur;
#pragma unroll
    for (int i = 0; i < kAttnWarps; i++) M = fmaxf(M, wacc[i][0]);
    float Lsum = 0.f, a4[4] = {0.f, 0.f, 0.f, 0.f};
    if (M > -CUDART_INF_F) {
#pragma unroll
        for (int i = 0; i < kAttnWarps; i++) {
            if (wacc[i][1] > 0.f) {
                const float c = __expf(wacc[i][0] - M);
                Lsum += wacc[i][1] * c;
                const float4 v = *reinterpret_cast<const float4 *>(&wacc[i][4 + 4 * lane]);
                a4[0] = fmaf(v.x, c, a4[0]), a4[1] = fmaf(v.y, c, a4[1]), a4[2] = fmaf(v.z, c, a4[2]), a4[3] = fmaf(v.w, c, a4[3]);
            }
        }
        if (split == 0) {
            const float pc = __expf(s_cur - M);
            Lsum += pc;
#pragma unroll
            for (int j = 0; j < 4; j++) a4[j] = fmaf(pc, __half2float(vn[j]), a4[j]);
        }
    }
    if (nsplit == 1) {
        const float inv = 1.f / Lsum;
        __half r4[4];
#pragma unroll
        for (int j = 0; j < 4; j++) r4[j] = __float2half_rn(a4[j] * inv);
        *reinterpret_cast<uint2 *>(out + (size_t)h * kHeadDim + 4 * lane) = *reinterpret_cast<uint2 *>(r4);
    } else {
        float *pp = part + ((size_t)h * nsplit + split) * kPartStride;  // (m, l, -, -, acc[128]): 16-byte aligned rows
        if (lane == 0) pp[0] = M, pp[1] = Lsum;
        *reinterpret_cast<float4 *>(pp + 4 + 4 * lane) = make_float4(a4[0], a4[1], a4[2], a4[3]);
    }
}

// combine nsplit partial softmax states per head: grid = H, block = 128
__global__ void attn_merge_kernel(const float *__restrict__ part, __half *__restrict__ out, uint32_t nsplit) {
    apg::pdl_wait_prior_grid();
    apg::pdl_launch_dependents();
    const uint32_t h = blockIdx.x, d = threadIdx.x;
    const float *pp = part + (size_t)h * nsplit * kPartStride;
    float m = -CUDART_INF_F;
    for (uint32_t s = 0; s < nsplit; s++) m = fmaxf(m, pp[s * kPartStride]);
    float l = 0.f, a = 0.f;
    for (uint32_t s = 0; s < nsplit; s++) {
        const float *q = pp + s * kPartStride;
        if (q[1] > 0.f) {  // a split with no time step has l == 0 (and m == -inf)
            const float c = __expf(q[0] - m);
            l += q[1] * c;
            a += q[4 + d] * c;
        }
    }
    out[(size_t)h * kHeadDim + d] = __float2half_rn(a / l);
}

// ------------------------------------------------------------------------------------------------------------
// logits = output( norm(x) )   (model.py:128-129): fp16 weight [V, D] streamed once, final RMSNorm fused into the
// x load, fp32 accumulation, fp16 logits (the reference's nn.Linear is fp16).  One warp per row, D % 256 == 0.
// HBM-bound: 2*V*D bytes (1.05 GB for Llama-3: 37% of all bytes of a 2-bit token).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_evict_first_v4(const uint4 *p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}

// total order on (value, index): larger value first, then smaller index; NaN never wins
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

template <int NV>  // NV = D / 256: uint4 loads per lane per row
__global__ void __launch_bounds__(256) lm_head_kernel(const __half *__restrict__ x, const __half *__restrict__ norm_w, float eps,
                                                      const __half *__restrict__ W, __half *__restrict__ logits, uint32_t V,
                                                      uint32_t D, float *__restrict__ best_val, int *__restrict__ best_idx,
                                                      uint32_t row_offset) {
    extern __shared__ __align__(16) float xs[];  // [D] normalised activations as fp32
    __shared__ float red[8];
    __shared__ int redi[8];
    uint64_t pol;  // the 2*V*D-byte weight is read once per token: keep it from evicting LUTs / KV / activations in L2
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    apg::pdl_wait_prior_grid();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float ss = 0.f;
    for (uint32_t i = threadIdx.x; i < D; i += blockDim.x) {
        const float f = __half2float(x[i]);
        xs[i] = f;
        ss = fmaf(f, f, ss);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) red[w] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < nw; i++) tot += red[i];
    const float rs = rsqrtf(tot / (float)D + eps);
    for (uint32_t i = threadIdx.x; i < D; i += blockDim.x) {
        const __half n = __float2half_rn(xs[i] * rs);          // .type_as(x)
        xs[i] = __half2float(__hmul(n, norm_w[i]));            // * weight (fp16)
    }
    __syncthreads();
    apg::pdl_launch_dependents();

    const uint32_t warps_total = gridDim.x * nw;
    float bv = -CUDART_INF_F;
    int bi = 0;
    for (uint32_t row = blockIdx.x * nw + w; row < V; row += warps_total) {
        const uint4 *wr = reinterpret_cast<const uint4 *>(W + (size_t)row * D);
        uint4 v[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) v[i] = ldg_evict_first_v4(wr + i * 32 + lane, pol);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; i++) {
            const float4 xa = *reinterpret_cast<const float4 *>(xs + (i * 32 + lane) * 8);
            const float4 xb = *reinterpret_cast<const float4 *>(xs + (i * 32 + lane) * 8 + 4);
            const float2 w0 = __half22float2(*reinterpret_cast<const __half2 *>(&v[i].x));
            const float2 w1 = __half22float2(*reinterpret_cast<const __half2 *>(&v[i].y));
            const float2 w2 = __half22float2(*reinterpret_cast<const __half2 *>(&v[i].z));
            const float2 w3 = __half22float2(*reinterpret_cast<const __half2 *>(&v[i].w));
            a0 = fmaf(w0.x, xa.x, a0), a1 = fmaf(w0.y, xa.y, a1);
            a0 = fmaf(w1.x, xa.z, a0), a1 = fmaf(w1.y, xa.w, a1);
            a0 = fmaf(w2.x, xb.x, a0), a1 = fmaf(w2.y, xb.y, a1);
            a0 = fmaf(w3.x, xb.z, a0), a1 = fmaf(w3.y, xb.w, a1);
        }
        float a = a0 + a1;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        const __half lg = __float2half_rn(a);
        if (lane == 0) logits[row] = lg;
        const float lf = __half2float(lg);  // sampling sees the fp16 logits, like the reference (generate.py:69)
        if (better(lf, (int)(row + row_offset), bv, bi)) bv = lf, bi = (int)(row + row_offset);
    }
    // per-CTA arg-max partial for the greedy sampler (rows ascend per warp, so ties keep the smaller index)
    if (best_val) {
        __syncthreads();
        if (lane == 0) red[w] = bv, redi[w] = bi;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < nw; i++)
                if (better(red[i], redi[i], bv, bi)) bv = red[i], bi = redi[i];
            best_val[blockIdx.x] = bv, best_idx[blockIdx.x] = bi;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// greedy sampling (generate.py:55-73 with temperature 0 -> logits / 1e-5 -> softmax is one-hot -> argmax; first
// index wins ties like torch.argmax).  One CTA.  Writes the next token, appends it to the output ring and advances
// the device-side position so the same graph can be replayed for the next token.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) argmax_advance_kernel(const float *__restrict__ best_val, const int *__restrict__ best_idx,
                                                              uint32_t n, int *token, int *pos, int *history,
                                                              uint32_t history_len) {
    apg::pdl_wait_prior_grid();
    apg::pdl_launch_dependents();
    __shared__ float bv[32];
    __shared__ int bi[32];
    float best = -CUDART_INF_F;
    int idx = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        if (better(best_val[i], best_idx[i], best, idx)) best = best_val[i], idx = best_idx[i];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (better(ob, oi, best, idx)) best = ob, idx = oi;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) bv[w] = best, bi[w] = idx;
    __syncthreads();
    if (w == 0) {
        best = lane < (int)(blockDim.x >> 5) ? bv[lane] : -CUDART_INF_F;
        idx = lane < (int)(blockDim.x >> 5) ? bi[lane] : 0;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (better(ob, oi, best, idx)) best = ob, idx = oi;
        }
        if (lane == 0) {
            const int p = *pos;
            *token = idx;
            if (history && (uint32_t)(p + 1) < history_len) history[p + 1] = idx;
            *pos = p + 1;
        }
    }
}

// Vocab-sharded greedy sampling (tensor parallel lm_head): every rank reduces its own arg-max partials, pushes
// (value, epoch) and (index, epoch) packets into slot `rank` of every peer's exchange buffer over NVLink (same
// data-with-flag scheme as the fused all-reduce), polls the `world` slots of its own buffer and picks the global winner
// — identical on every rank (larger value, then smaller index).
__global__ void __launch_bounds__(1024) argmax_advance_tp_kernel(const float *__restrict__ best_val, const int *__restrict__ best_idx,
                                                                 uint32_t n, uint32_t world, uint32_t rank, uint2 *p0, uint2 *p1,
                                                                 uint2 *p2, uint2 *p3, uint2 *p4, uint2 *p5, uint2 *p6, uint2 *p7,
                                                                 uint32_t *epoch, int *token, int *pos, int *history,
                                                                 uint32_t history_len) {
    apg::pdl_wait_prior_grid();
    apg::pdl_launch_dependents();
    uint2 *peers[8] = {p0, p1, p2, p3, p4, p5, p6, p7};
    __shared__ float bv[32];
    __shared__ int bi[32];
    float best = -CUDART_INF_F;
    int idx = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        if (better(best_val[i], best_idx[i], best, idx)) best = best_val[i], idx = best_idx[i];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (better(ob, oi, best, idx)) best = ob, idx = oi;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) bv[w] = best, bi[w] = idx;
    __syncthreads();
    if (w != 0) return;
    best = lane < (int)(blockDim.x >> 5) ? bv[lane] : -CUDART_INF_F;
    idx = lane < (int)(blockDim.x >> 5) ? bi[lane] : 0;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (better(ob, oi, best, idx)) best = ob, idx = oi;
    }
    const uint32_t ep = *epoch + 1u;
    if (lane < (int)world) {  // lane p pushes this rank's winner to peer p
        uint2 *dst = peers[lane] + 2 * rank;
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(best)), "r"(ep) : "memory");
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst + 1), "r"((uint32_t)idx), "r"(ep) : "memory");
    }
    float gv = -CUDART_INF_F;
    int gi = 0;
    if (lane < (int)world) {  // lane r polls slot r of our own buffer
        const uint2 *src = peers[rank] + 2 * lane;
        uint32_t v, t0, i2, t1;
        do {
            asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(t0) : "l"(src) : "memory");
            asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(i2), "=r"(t1) : "l"(src + 1) : "memory");
        } while (t0 != ep || t1 != ep);
        gv = __uint_as_float(v), gi = (int)i2;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, gv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, gi, o);
        if (better(ob, oi, gv, gi)) gv = ob, gi = oi;
    }
    if (lane == 0) {
        const int p = *pos;
        *token = gi;
        if (history && (uint32_t)(p + 1) < history_len) history[p + 1] = gi;
        *pos = p + 1;
        *epoch = ep;
    }
}

// ------------------------------------------------------------------------------------------------------------
// temperature / top-k sampling (logits_to_probs + multinomial_sample_one_no_sync, generate.py:55-73):
//     l = logits / max(T, 1e-5);  pivot = k-th largest l;  l[l < pivot] = -inf;  p = softmax(l);  q_i ~ Exp(1);
//     token = argmax_i p_i / q_i   ==   argmax_i ( l_i - log q_i )      (the softmax normaliser is common to all i)
// One CTA.  The k-th largest value is found EXACTLY on order-preserving 16-bit keys of the fp16 logits (dividing by a
// positive T keeps the order and keeps distinct fp16 values distinct in fp32): block max, then the elements within
// kSampleWindow key steps of the max are collected into shared memory and the pivot is bisected there; if that window
// holds fewer than k (or more than kSampleCap) elements the bisection runs over the whole array instead (slower, same
// result).  Ties with the pivot are kept, like the reference's `logits < pivot` mask.  q_i = -log(u_i) with u_i from a
// counter hash of (*seed, *pos, i) (sample_uniform below; restated in oracle/decode_oracle.py), so a replayed CUDA
// graph draws fresh noise every token.  NaN logits are never selected.  top_k == 0 or >= V: no filter.
// ------------------------------------------------------------------------------------------------------------
constexpr int kSampleThreads = 1024;
constexpr uint32_t kSampleCap = 6144, kSampleWindow = 1024;

__device__ __forceinline__ uint32_t f16_order_key(uint32_t h) {  // ascending with the value; NaN -> 0; -0 == +0
    h &= 0xffffu;
    if ((h & 0x7fffu) > 0x7c00u) return 0u;
    if (h == 0x8000u) h = 0u;
    return (h & 0x8000u) ? (~h & 0xffffu) : (h | 0x8000u);
}
__device__ __forceinline__ float f16_key_value(uint32_t key) {
    const uint32_t h = (key & 0x8000u) ? (key & 0x7fffu) : (~key & 0xffffu);
    return __half2float(__ushort_as_half((unsigned short)h));
}
__device__ __forceinline__ float sample_uniform(unsigned long long seed, uint32_t pos, uint32_t i) {  // in (0, 1), 23 bits
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * ((((unsigned long long)pos) << 32) | i);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return ((float)(uint32_t)(z >> 41) + 0.5f) * (1.0f / 8388608.0f);
}

// warp-uniform walk over logits[0:V]: every lane of every warp calls f(i, bits, valid) the same number of times
template <typename F>
__device__ __forceinline__ void for_each_logit(const __half *__restrict__ logits, uint32_t V, F f) {
    const uint4 *v4 = reinterpret_cast<const uint4 *>(logits);
    const uint32_t nv = V >> 3, lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u;
    for (uint32_t j0 = wbase; j0 < nv; j0 += blockDim.x) {
        const uint32_t j = j0 + lane;
        const bool ok = j < nv;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (ok) q = v4[j];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            f(8u * j + 2u * e, w[e] & 0xffffu, ok);
            f(8u * j + 2u * e + 1u, w[e] >> 16, ok);
        }
    }
    for (uint32_t i0 = (nv << 3) + wbase; i0 < V; i0 += blockDim.x) {
        const uint32_t i = i0 + lane;
        const bool ok = i < V;
        f(i, ok ? (uint32_t)__half_as_ushort(logits[i]) : 0u, ok);
    }
}

__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t *red) {  // every thread gets the total
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    uint32_t t = lane < (blockDim.x >> 5) ? red[lane] : 0u;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}
__device__ __forceinline__ uint32_t block_max_u32(uint32_t v, uint32_t *red) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    uint32_t t = lane < (blockDim.x >> 5) ? red[lane] : 0u;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) t = max(t, __shfl_xor_sync(0xffffffffu, t, o));
    return t;
}

__global__ void __launch_bounds__(kSampleThreads) sample_topk_advance_kernel(const __half *__restrict__ logits, uint32_t V,
                                                                             float temperature, uint32_t top_k,
                                                                             const unsigned long long *__restrict__ seed,
                                                                             int *token, int *pos, int *history,
                                                                             uint32_t history_len) {
    apg::pdl_wait_prior_grid();
    apg::pdl_launch_dependents();
    __shared__ uint32_t red[32];
    __shared__ float bv[32];
    __shared__ int bi[32];
    __shared__ uint32_t cand_idx[kSampleCap];
    __shared__ unsigned short cand_key[kSampleCap];
    __shared__ uint32_t n_cand;
    const uint32_t lane = threadIdx.x & 31u;
    const int p = *pos;
    const unsigned long long sd = *seed;
    const float t = fmaxf(temperature, 1e-5f);

    uint32_t pivot = 1u, ncand = 0u;  // key 0 is NaN: never kept
    bool use_list = false;
    if (top_k > 0 && top_k < V) {
        uint32_t mk = 0;
        for_each_logit(logits, V, [&](uint32_t, uint32_t h, bool ok) {
            if (ok) mk = max(mk, f16_order_key(h));
        });
        mk = block_max_u32(mk, red);
        const uint32_t wlo = mk > kSampleWindow ? mk - kSampleWindow : 1u;
        if (threadIdx.x == 0) n_cand = 0;
        __syncthreads();
        for_each_logit(logits, V, [&](uint32_t i, uint32_t h, bool ok) {
            const uint32_t key = f16_order_key(h);
            const bool in = ok && key >= wlo;
            const uint32_t m = __ballot_sync(0xffffffffu, in);
            if (m) {  // warp-aggregated append
                const int leader = __ffs(m) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(&n_cand, (uint32_t)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
                if (in && slot < kSampleCap) cand_idx[slot] = i, cand_key[slot] = (unsigned short)key;
            }
        });
        __syncthreads();
        ncand = n_cand;
        use_list = ncand >= top_k && ncand <= kSampleCap;
        // largest `lo` with count(key >= lo) >= top_k; the invariant holds at the start on both paths
        uint32_t lo = use_list ? wlo : 0u, hi = mk;
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1u) >> 1;
            uint32_t c = 0;
            if (use_list) {
                for (uint32_t j = threadIdx.x; j < ncand; j += blockDim.x) c += (uint32_t)cand_key[j] >= mid;
            } else {
                for_each_logit(logits, V, [&](uint32_t, uint32_t h, bool ok) { c += (ok && f16_order_key(h) >= mid) ? 1u : 0u; });
            }
            c = block_sum_u32(c, red);
            if (c >= top_k) lo = mid;
            else hi = mid - 1u;
        }
        pivot = max(lo, 1u);
    }

    float best = -CUDART_INF_F;
    int idx = 0;
    auto consider = [&](uint32_t i, uint32_t key) {
        const float u = sample_uniform(sd, (uint32_t)p, i);
        const float s = f16_key_value(key) / t - logf(-logf(u));
        if (better(s, (int)i, best, idx)) best = s, idx = (int)i;
    };
    if (use_list) {
        for (uint32_t j = threadIdx.x; j < ncand; j += blockDim.x)
            if ((uint32_t)cand_key[j] >= pivot) consider(cand_idx[j], cand_key[j]);
    } else {
        for_each_logit(logits, V, [&](uint32_t i, uint32_t h, bool ok) {
            const uint32_t key = f16_order_key(h);
            if (ok && key >= pivot) consider(i, key);
        });
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (better(ob, oi, best, idx)) best = ob, idx = oi;
    }
    const uint32_t w = threadIdx.x >> 5;
    if (lane == 0) bv[w] = best, bi[w] = idx;
    __syncthreads();
    if (w == 0) {
        best = lane < (blockDim.x >> 5) ? bv[lane] : -CUDART_INF_F;
        idx = lane < (blockDim.x >> 5) ? bi[lane] : 0;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (better(ob, oi, best, idx)) best = ob, idx = oi;
        }
        if (lane == 0) {
            *token = idx;
            if (history && (uint32_t)(p + 1) < history_len) history[p + 1] = idx;
            *pos = p + 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// The same sampler for a VOCAB-SHARDED lm_head (tensor parallel): every rank holds logits[rank*V : (rank+1)*V].  Two
// packet exchanges over the peers' buffers (data-with-flag stores, like argmax_advance_tp_kernel):
//   1. top-k pivot: every rank pushes the keys of ITS k largest logits (k packets); the k-th largest of the union of those
//      world*k keys is the k-th largest of the whole vocabulary, so every rank finds the same global pivot;
//   2. winner: every rank pushes its best (score, global index) among its logits >= pivot; the global winner is picked
//      with the same tie rule on every rank.
// The noise u_i is hashed from the GLOBAL index, so the token equals what sample_topk_advance_kernel draws from the
// gathered logits (tested in tests/dist_check.py).  Slot layout of a rank's buffer: uint2 [world][slot_stride]; packets
// 0..k-1 = keys, k and k+1 = (score, index).  Requires top_k <= kSampleTpMaxK, top_k < V, top_k + 2 <= slot_stride.
// ------------------------------------------------------------------------------------------------------------
constexpr uint32_t kSampleTpMaxK = 256;

__device__ __forceinline__ void ll_push(uint2 *dst, uint32_t v, uint32_t tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(v), "r"(tag) : "memory");
}
__device__ __forceinline__ uint32_t ll_poll(const uint2 *src, uint32_t tag) {
    uint32_t v, t;
    do {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(t) : "l"(src) : "memory");
    } while (t != tag);
    return v;
}

__global__ void __launch_bounds__(kSampleThreads) sample_topk_advance_tp_kernel(
    const __half *__restrict__ logits, uint32_t V, float temperature, uint32_t top_k, const unsigned long long *__restrict__ seed,
    uint32_t world, uint32_t rank, uint2 *p0, uint2 *p1, uint2 *p2, uint2 *p3, uint2 *p4, uint2 *p5, uint2 *p6, uint2 *p7,
    uint32_t slot_stride, uint32_t *epoch, int *token, int *pos, int *history, uint32_t history_len) {
    apg::pdl_wait_prior_grid();
    apg::pdl_launch_dependents();
    uint2 *peers[8] = {p0, p1, p2, p3, p4, p5, p6, p7};
    __shared__ uint32_t red[32];
    __shared__ float bv[32];
    __shared__ int bi[32];
    __shared__ uint32_t cand_idx[kSampleCap];
    __shared__ unsigned short cand_key[kSampleCap];
    __shared__ unsigned short top_keys[kSampleTpMaxK];
    __shared__ unsigned short all_keys[8 * kSampleTpMaxK];
    __shared__ uint32_t n_cand, n_top;
    const uint32_t lane = threadIdx.x & 31u;
    const int p = *pos;
    const unsigned long long sd = *seed;
    const float t = fmaxf(temperature, 1e-5f);
    const uint32_t ep = *epoch + 1u;
    const uint32_t goff = rank * V;  // global index of local logit 0

    uint32_t pivot = 1u, ncand = 0u;
    bool use_list = false;
    if (top_k > 0) {
        // ---- local k-th largest key (as on one GPU) ----
        uint32_t mk = 0;
        for_each_logit(logits, V, [&](uint32_t, uint32_t h, bool ok) {
            if (ok) mk = max(mk, f16_order_key(h));
        });
        mk = block_max_u32(mk, red);
        const uint32_t wlo = mk > kSampleWindow ? mk - kSampleWindow : 1u;
        if (threadIdx.x == 0) n_cand = 0, n_top = 0;
        __syncthreads();
        for_each_logit(logits, V, [&](uint32_t i, uint32_t h, bool ok) {
            const uint32_t key = f16_order_key(h);
            const bool in = ok && key >= wlo;
            const uint32_t m = __ballot_sync(0xffffffffu, in);
            if (m) {
                const int leader = __ffs(m) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(&n_cand, (uint32_t)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
                if (in && slot < kSampleCap) cand_idx[slot] = i, cand_key[slot] = (unsigned short)key;
            }
        });
        __syncthreads();
        ncand = n_cand;
        use_list = ncand >= top_k && ncand <= kSampleCap;
        uint32_t lo = use_list ? wlo : 0u, hi = mk;
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1u) >> 1;
            uint32_t c = 0;
            if (use_list) {
                for (uint32_t j = threadIdx.x; j < ncand; j += blockDim.x) c += (uint32_t)cand_key[j] >= mid;
            } else {
                for_each_logit(logits, V, [&](uint32_t, uint32_t h, bool ok) { c += (ok && f16_order_key(h) >= mid) ? 1u : 0u; });
            }
            c = block_sum_u32(c, red);
            if (c >= top_k) lo = mid;
            else hi = mid - 1u;
        }
        // ---- the k largest local keys: every key above the local pivot, the rest of the k slots = the pivot itself ----
        auto keep = [&](uint32_t key, bool ok) {
            const bool in = ok && key > lo;
            const uint32_t m = __ballot_sync(0xffffffffu, in);
            if (m) {
                const int leader = __ffs(m) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(&n_top, (uint32_t)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
                if (in && slot < top_k) top_keys[slot] = (unsigned short)key;
            }
        };
        if (use_list) {
            for (uint32_t j0 = threadIdx.x & ~31u; j0 < ncand; j0 += blockDim.x) {
                const uint32_t j = j0 + lane;
                keep(j < ncand ? (uint32_t)cand_key[j] : 0u, j < ncand);
            }
        } else {
            for_each_logit(logits, V, [&](uint32_t, uint32_t h, bool ok) { keep(f16_order_key(h), ok); });
        }
        __syncthreads();
        const uint32_t above = min(n_top, top_k);  // < top_k by the definition of the pivot
        for (uint32_t j = above + threadIdx.x; j < top_k; j += blockDim.x) top_keys[j] = (unsigned short)lo;
        __syncthreads();
        // ---- exchange 1: keys ----
        for (uint32_t e = threadIdx.x; e < world * top_k; e += blockDim.x) {
            const uint32_t pr = e / top_k, j = e % top_k;
            ll_push(peers[pr] + (size_t)rank * slot_stride + j, top_keys[j], ep);
        }
        for (uint32_t e = threadIdx.x; e < world * top_k; e += blockDim.x) {
            const uint32_t r = e / top_k, j = e % top_k;
            all_keys[e] = (unsigned short)ll_poll(peers[rank] + (size_t)r * slot_stride + j, ep);
        }
        __syncthreads();
        uint32_t glo = 0u, ghi = 0xffffu;
        while (glo < ghi) {  // largest g with count(all_keys >= g) >= top_k
            const uint32_t mid = (glo + ghi + 1u) >> 1;
            uint32_t c = 0;
            for (uint32_t j = threadIdx.x; j < world * top_k; j += blockDim.x) c += (uint32_t)all_keys[j] >= mid;
            c = block_sum_u32(c, red);
            if (c >= top_k) glo = mid;
            else ghi = mid - 1u;
        }
        pivot = max(glo, 1u);  // >= the local pivot, so the candidate list (keys >= wlo) still covers everything kept
    }

    float best = -CUDART_INF_F;
    int idx = 0x7fffffff;
    auto consider = [&](uint32_t i, uint32_t key) {
        const float u = sample_uniform(sd, (uint32_t)p, goff + i);
        const float s = f16_key_value(key) / t - logf(-logf(u));
        if (better(s, (int)(goff + i), best, idx)) best = s, idx = (int)(goff + i);
    };
    if (use_list) {
        for (uint32_t j = threadIdx.x; j < ncand; j += blockDim.x)
            if ((uint32_t)cand_key[j] >= pivot) consider(cand_idx[j], cand_key[j]);
    } else {
        for_each_logit(logits, V, [&](uint32_t i, uint32_t h, bool ok) {
            const uint32_t key = f16_order_key(h);
            if (ok && key >= pivot) consider(i, key);
        });
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (better(ob, oi, best, idx)) best = ob, idx = oi;
    }
    const uint32_t w = threadIdx.x >> 5;
    if (lane == 0) bv[w] = best, bi[w] = idx;
    __syncthreads();
    if (w != 0) return;
    best = lane < (blockDim.x >> 5) ? bv[lane] : -CUDART_INF_F;
    idx = lane < (blockDim.x >> 5) ? bi[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (better(ob, oi, best, idx)) best = ob, idx = oi;
    }
    // ---- exchange 2: (score, global index) of this rank's winner ----
    if (lane < world) {
        uint2 *dst = peers[lane] + (size_t)rank * slot_stride + top_k;
        ll_push(dst, __float_as_uint(best), ep);
        ll_push(dst + 1, (uint32_t)idx, ep);
    }
    float gv = -CUDART_INF_F;
    int gi = 0x7fffffff;
    if (lane < world) {
        const uint2 *src = peers[rank] + (size_t)lane * slot_stride + top_k;
        gv = __uint_as_float(ll_poll(src, ep));
        gi = (int)ll_poll(src + 1, ep);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, gv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, gi, o);
        if (better(ob, oi, gv, gi)) gv = ob, gi = oi;
    }
    if (lane == 0) {
        *token = gi;
        if (history && (uint32_t)(p + 1) < history_len) history[p + 1] = gi;
        *pos = p + 1;
        *epoch = ep;
    }
}

}  // namespace apd
