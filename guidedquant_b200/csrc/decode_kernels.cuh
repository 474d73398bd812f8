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
                             uint32_t dim) {
    apg::pdl_wait_prior_grid();
    const uint4 *src = reinterpret_cast<const uint4 *>(emb + (size_t)(*token) * dim);
    uint4 *dst = reinterpret_cast<uint4 *>(x);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < dim / 8; i += gridDim.x * blockDim.x) dst[i] = src[i];
    apg::pdl_launch_dependents();
}

// ------------------------------------------------------------------------------------------------------------
// RoPE + KV append + attention for one new token (Attention.forward, model.py:206-236).
//   qkv      fp16 [(H + 2*Hkv) * 128]   output of the fused wqkv Linear: q | k | v          (model.py:211)
//   inv_freq fp32 [64]                  1 / base^(2i/128)                                    (LlamaRotaryEmbedding)
//   k_cache, v_cache fp16 [Hkv, S, 128]                                                      (KVCache, model.py:63-79)
//   out      fp16 [H * 128]             attention output, input of wo
// grid = (Hkv, nsplit); block = 32 * G threads, G = H / Hkv query heads per KV head; warp w serves query head
// kvh*G + w and the time steps t = split, split + nsplit, ... <= pos.  head_dim is fixed at 128: a lane owns 4 dims.
// cos/sin are computed in fp32 and rounded to fp16, and q*cos + rotate_half(q)*sin is evaluated in fp16 exactly as the
// reference's half tensors do (model.py:309-314, 396-405); scores, softmax and P.V accumulate in fp32.
// nsplit > 1: each (head, split) writes an un-normalised partial (m, l, acc[128]) and attn_merge_kernel combines them.
// ------------------------------------------------------------------------------------------------------------
constexpr int kHeadDim = 128;
constexpr int kPartStride = kHeadDim + 4;

__device__ __forceinline__ void rope4(const __half (&v)[4], const __half (&partner)[4], int lane, float pos,
                                      const float *__restrict__ inv_freq, __half (&o)[4]) {
    // dims d = 4*lane + j; rotate_half: d < 64 -> -x[d+64], d >= 64 -> x[d-64]  (model.py:268-272)
    const bool hi = lane >= 16;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int d = 4 * lane + j;
        const float fr = pos * inv_freq[d & 63];
        const __half c = __float2half_rn(cosf(fr)), s = __float2half_rn(sinf(fr));
        const __half rot = hi ? partner[j] : __hneg(partner[j]);
        o[j] = __hadd(__hmul(v[j], c), __hmul(rot, s));
    }
}

__global__ void __launch_bounds__(256) attn_decode_kernel(const __half *__restrict__ qkv, const float *__restrict__ inv_freq,
                                                          __half *__restrict__ k_cache, __half *__restrict__ v_cache,
                                                          const int *__restrict__ pos_ptr, __half *__restrict__ out,
                                                          float *__restrict__ part, uint32_t H, uint32_t Hkv, uint32_t S,
                                                          float scale) {
    apg::pdl_wait_prior_grid();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t G = H / Hkv, kvh = blockIdx.x, split = blockIdx.y, nsplit = gridDim.y;
    const uint32_t h = kvh * G + w;
    const int pos = *pos_ptr;
    const float fpos = (float)pos;

    // q, k_new, v_new of this head / kv head: 4 dims per lane
    __half q[4], kn[4], vn[4], qp[4], kp[4];
    {
        const uint2 qv = *reinterpret_cast<const uint2 *>(qkv + (size_t)h * kHeadDim + 4 * lane);
        const uint2 kv = *reinterpret_cast<const uint2 *>(qkv + (size_t)(H + kvh) * kHeadDim + 4 * lane);
        const uint2 vv = *reinterpret_cast<const uint2 *>(qkv + (size_t)(H + Hkv + kvh) * kHeadDim + 4 * lane);
        *reinterpret_cast<uint2 *>(q) = qv, *reinterpret_cast<uint2 *>(kn) = kv, *reinterpret_cast<uint2 *>(vn) = vv;
        uint2 qpv, kpv;
        qpv.x = __shfl_xor_sync(0xffffffffu, qv.x, 16), qpv.y = __shfl_xor_sync(0xffffffffu, qv.y, 16);
        kpv.x = __shfl_xor_sync(0xffffffffu, kv.x, 16), kpv.y = __shfl_xor_sync(0xffffffffu, kv.y, 16);
        *reinterpret_cast<uint2 *>(qp) = qpv, *reinterpret_cast<uint2 *>(kp) = kpv;
    }
    __half qr[4], kr[4];
    rope4(q, qp, lane, fpos, inv_freq, qr);
    rope4(kn, kp, lane, fpos, inv_freq, kr);
    if (split == 0 && w == 0) {  // cache append (KVCache.update, model.py:70-79); readers of t = pos use registers
        *reinterpret_cast<uint2 *>(k_cache + ((size_t)kvh * S + pos) * kHeadDim + 4 * lane) = *reinterpret_cast<uint2 *>(kr);
        *reinterpret_cast<uint2 *>(v_cache + ((size_t)kvh * S + pos) * kHeadDim + 4 * lane) = *reinterpret_cast<uint2 *>(vn);
    }
    float qf[4];
#pragma unroll
    for (int j = 0; j < 4; j++) qf[j] = __half2float(qr[j]);

    float m = -CUDART_INF_F, l = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = (int)split; t <= pos; t += (int)nsplit) {
        __half kk[4], vv[4];
        if (t == pos) {
#pragma unroll
            for (int j = 0; j < 4; j++) kk[j] = kr[j], vv[j] = vn[j];
        } else {
            *reinterpret_cast<uint2 *>(kk) = *reinterpret_cast<const uint2 *>(k_cache + ((size_t)kvh * S + t) * kHeadDim + 4 * lane);
            *reinterpret_cast<uint2 *>(vv) = *reinterpret_cast<const uint2 *>(v_cache + ((size_t)kvh * S + t) * kHeadDim + 4 * lane);
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) s = fmaf(qf[j], __half2float(kk[j]), s);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s *= scale;
        const float mn = fmaxf(m, s);
        const float corr = __expf(m - mn), pe = __expf(s - mn);
        l = l * corr + pe;
#pragma unroll
        for (int j = 0; j < 4; j++) acc[j] = acc[j] * corr + pe * __half2float(vv[j]);
        m = mn;
    }
    if (nsplit == 1) {
        const float inv = 1.f / l;
        __half o4[4];
#pragma unroll
        for (int j = 0; j < 4; j++) o4[j] = __float2half_rn(acc[j] * inv);
        *reinterpret_cast<uint2 *>(out + (size_t)h * kHeadDim + 4 * lane) = *reinterpret_cast<uint2 *>(o4);
    } else {
        float *pp = part + ((size_t)h * nsplit + split) * kPartStride;  // (m, l, -, -, acc[128]): 16-byte aligned rows
        if (lane == 0) pp[0] = m, pp[1] = l;
        *reinterpret_cast<float4 *>(pp + 4 + 4 * lane) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
    apg::pdl_launch_dependents();
}

// combine nsplit partial softmax states per head: grid = H, block = 128
__global__ void attn_merge_kernel(const float *__restrict__ part, __half *__restrict__ out, uint32_t nsplit) {
    apg::pdl_wait_prior_grid();
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
    apg::pdl_launch_dependents();
}

// ------------------------------------------------------------------------------------------------------------
// logits = output( norm(x) )   (model.py:128-129): fp16 weight [V, D] streamed once, final RMSNorm fused into the
// x load, fp32 accumulation, fp16 logits (the reference's nn.Linear is fp16).  One warp per row, D % 256 == 0.
// HBM-bound: 2*V*D bytes (1.05 GB for Llama-3: 37% of all bytes of a 2-bit token).
// ------------------------------------------------------------------------------------------------------------
template <int NV>  // NV = D / 256: uint4 loads per lane per row
__global__ void __launch_bounds__(256) lm_head_kernel(const __half *__restrict__ x, const __half *__restrict__ norm_w, float eps,
                                                      const __half *__restrict__ W, __half *__restrict__ logits, uint32_t V,
                                                      uint32_t D) {
    extern __shared__ __align__(16) float xs[];  // [D] normalised activations as fp32
    __shared__ float red[8];
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
    for (uint32_t row = blockIdx.x * nw + w; row < V; row += warps_total) {
        const uint4 *wr = reinterpret_cast<const uint4 *>(W + (size_t)row * D);
        uint4 v[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) v[i] = apg::ldg_stream_v4(wr + i * 32 + lane);
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
        if (lane == 0) logits[row] = __float2half_rn(a);
    }
}

// ------------------------------------------------------------------------------------------------------------
// greedy sampling (generate.py:55-73 with temperature 0 -> logits / 1e-5 -> softmax is one-hot -> argmax; first
// index wins ties like torch.argmax).  One CTA.  Writes the next token, appends it to the output ring and advances
// the device-side position so the same graph can be replayed for the next token.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) argmax_advance_kernel(const __half *__restrict__ logits, uint32_t V, int *token,
                                                              int *pos, int *history, uint32_t history_len) {
    apg::pdl_wait_prior_grid();
    __shared__ float bv[32];
    __shared__ int bi[32];
    float best = -CUDART_INF_F;
    int idx = 0x7fffffff;
    for (uint32_t i = threadIdx.x; i < V; i += blockDim.x) {
        const float f = __half2float(logits[i]);
        if (f > best || (f == best && (int)i < idx)) best = f, idx = (int)i;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ob > best || (ob == best && oi < idx)) best = ob, idx = oi;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) bv[w] = best, bi[w] = idx;
    __syncthreads();
    if (w == 0) {
        best = lane < (int)(blockDim.x >> 5) ? bv[lane] : -CUDART_INF_F;
        idx = lane < (int)(blockDim.x >> 5) ? bi[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ob > best || (ob == best && oi < idx)) best = ob, idx = oi;
        }
        if (lane == 0) {
            const int p = *pos;
            *token = idx;
            if (history && (uint32_t)(p + 1) < history_len) history[p + 1] = idx;
            *pos = p + 1;
        }
    }
    apg::pdl_launch_dependents();
}

}  // namespace apd
