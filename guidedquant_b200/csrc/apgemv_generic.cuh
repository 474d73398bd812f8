// Generic (untuned but complete) kernels: every bits in 2..8, M in 1..8, any K % 32 == 0, any N.
//   * gemv_generic_kernel<EXACT>: one warp per output row, lane t owns word t of every 1024-chunk
//     (the reference's ownership, anyprec.cu:432-448).  EXACT = true reproduces the reference's
//     fp16 arithmetic and accumulation order bit for bit (anyprec.cu:495-505, 362-370, 532-541;
//     SURVEY.md Appendix A) — the APG_FLAG_REF_ORDER parity mode.  EXACT = false accumulates in fp32.
//   * dequant_kernel: W[n,k] = lut[n, idx[n,k]] (replaces dequant_kbit_store, anyprec.cu:294-359);
//     a pure gather, so it is bit-identical to the reference by construction.
#pragma once
#include "apgemv_common.cuh"

namespace apg {

// 2^bits-entry index of the weight at k-offset o (0..31) of a chunk word, planes MSB first.
__device__ __forceinline__ uint32_t word_index(const uint32_t (&q)[8], int bits, int o) {
    const int beta = 31 - o;
    uint32_t idx = 0;
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j < bits) idx = (idx << 1) | ((q[j] >> beta) & 1u);
    return idx;
}

template <bool EXACT>
__global__ void __launch_bounds__(128) gemv_generic_kernel(const __half *__restrict__ x, const uint32_t *__restrict__ W,
                                                           const __half *__restrict__ lut, __half *__restrict__ out,
                                                           float *__restrict__ partial, uint32_t M, uint32_t N,
                                                           uint32_t K, int bits) {
    const int lane = threadIdx.x & 31;
    const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;  // whole warp exits together
    const uint32_t words = K >> 5, nchunk = (K + 1023u) >> 10, nc = 1u << bits;
    const __half *lrow = lut + (size_t)row * nc;

    float accf[8];
    __half acch[8];
#pragma unroll
    for (int l = 0; l < 8; l++) accf[l] = 0.f, acch[l] = __ushort_as_half(0);

    for (uint32_t i = 0; i < nchunk; i++) {
        const uint32_t eff = chunk_eff(K, i);
        if ((uint32_t)lane >= eff) break;  // tail chunk: lanes >= eff are done (anyprec.cu:433-436)
        uint32_t q[8];
#pragma unroll
        for (int j = 0; j < 8; j++) q[j] = (j < bits) ? __ldg(W + ((size_t)j * N + row) * words + i * 32u + lane) : 0u;
        __half2 dq[16];  // dq[4c+m] = ( W[k0+2m], W[k0+2m+1] ), k0 = i*1024 + c*8*eff + 8*lane
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const __half w0 = lrow[word_index(q, bits, 8 * c + 2 * m)];
                const __half w1 = lrow[word_index(q, bits, 8 * c + 2 * m + 1)];
                dq[4 * c + m] = __halves2half2(w0, w1);
            }
#pragma unroll
        for (int l = 0; l < 8; l++) {
            if ((uint32_t)l >= M) break;
            const __half *xl = x + (size_t)l * K + i * 1024u + 8u * lane;
            if (EXACT) {
                __half2 s = __halves2half2(__ushort_as_half(0), __ushort_as_half(0));
#pragma unroll
                for (int c = 3; c >= 0; c--) {  // j loop runs downward (anyprec.cu:497)
                    const uint4 xv = __ldg(reinterpret_cast<const uint4 *>(xl + c * 8u * eff));
                    const __half2 *xh = reinterpret_cast<const __half2 *>(&xv);
#pragma unroll
                    for (int m = 0; m < 4; m++) s = __hfma2(dq[4 * c + m], xh[m], s);
                }
                acch[l] = __hadd(acch[l], __hadd(__low2half(s), __high2half(s)));  // (anyprec.cu:505)
            } else {
                float a = 0.f;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint4 xv = __ldg(reinterpret_cast<const uint4 *>(xl + c * 8u * eff));
                    const __half2 *xh = reinterpret_cast<const __half2 *>(&xv);
#pragma unroll
                    for (int m = 0; m < 4; m++) {
                        const float2 wf = __half22float2(dq[4 * c + m]), xf = __half22float2(xh[m]);
                        a = fmaf(wf.x, xf.x, a);
                        a = fmaf(wf.y, xf.y, a);
                    }
                }
                accf[l] += a;
            }
        }
    }
    // every lane of the warp reaches this point (lanes that broke out of the tail chunk included)
#pragma unroll
    for (int l = 0; l < 8; l++) {
        if ((uint32_t)l >= M) break;
        if (EXACT) {
            __half v = acch[l];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v = __hadd(v, __shfl_down_sync(0xffffffffu, v, o));  // (anyprec.cu:362-370)
            if (lane == 0) {
                if (out) out[(size_t)l * N + row] = v;
                if (partial) partial[(size_t)l * N + row] = __half2float(v);
            }
        } else {
            float v = accf[l];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) {
                if (out) out[(size_t)l * N + row] = __float2half_rn(v);
                if (partial) partial[(size_t)l * N + row] = v;
            }
        }
    }
}

__global__ void __launch_bounds__(128) dequant_kernel(const uint32_t *__restrict__ W, const __half *__restrict__ lut,
                                                      __half *__restrict__ O, uint32_t N, uint32_t K, int bits) {
    const int lane = threadIdx.x & 31;
    const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;
    const uint32_t words = K >> 5, nchunk = (K + 1023u) >> 10, nc = 1u << bits;
    const unsigned short *lrow = reinterpret_cast<const unsigned short *>(lut) + (size_t)row * nc;
    for (uint32_t i = 0; i < nchunk; i++) {
        const uint32_t eff = chunk_eff(K, i);
        if ((uint32_t)lane >= eff) break;
        uint32_t q[8];
#pragma unroll
        for (int j = 0; j < 8; j++) q[j] = (j < bits) ? __ldg(W + ((size_t)j * N + row) * words + i * 32u + lane) : 0u;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t v[4];
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const uint32_t w0 = __ldg(lrow + word_index(q, bits, 8 * c + 2 * m));
                const uint32_t w1 = __ldg(lrow + word_index(q, bits, 8 * c + 2 * m + 1));
                v[m] = w0 | (w1 << 16);
            }
            // lanes t, t+1 write adjacent 16-byte units: 512 B per warp-store, fully coalesced
            *reinterpret_cast<uint4 *>(O + (size_t)row * K + i * 1024u + c * 8u * eff + 8u * lane) =
                make_uint4(v[0], v[1], v[2], v[3]);
        }
    }
}

// Receiver side of the fused one-shot all-reduce: every element of recv[world][n] is an 8-byte (value, epoch) packet
// stored by the owning rank's GEMV epilogue; poll each packet until it carries this use's epoch, add the `world` values
// in RANK ORDER (deterministic, identical on every GPU), add the optional fp16 residual and round once.  One thread per
// pair of elements over several CTAs (a single 1024-thread CTA walking four elements per thread in turn was 2 us longer
// per site); the last CTA through advances the site's epoch counter for the next graph replay (all have read it by then).
__global__ void __launch_bounds__(256) allreduce_finish_kernel(const uint2 *recv, uint32_t *epoch, uint32_t *done,
                                                               const __half *__restrict__ residual,
                                                               __half *__restrict__ out, uint32_t n, uint32_t world) {
    pdl_wait_prior_grid();
    pdl_launch_dependents();
    uint32_t ep;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(ep) : "l"(epoch) : "memory");
    ep += 1u;
    const uint32_t i = 2u * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i + 1u < n) {  // n is even
        float v0 = 0.f, v1 = 0.f;
        for (uint32_t r = 0; r < world; r++) {
            const uint4 *src = reinterpret_cast<const uint4 *>(recv + (size_t)r * n + i);  // two packets
            uint4 q;
            do {
                asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(src) : "memory");
            } while (q.y != ep || q.w != ep);
            v0 += __uint_as_float(q.x), v1 += __uint_as_float(q.z);
        }
        __half2 h = __floats2half2_rn(v0, v1);
        if (residual) h = __hadd2(h, *reinterpret_cast<const __half2 *>(residual + i));
        *reinterpret_cast<__half2 *>(out + i) = h;
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1u) {
        *done = 0u;
        *epoch = ep;
    }
}

__global__ void round_f32_to_f16_kernel(const float *__restrict__ in, __half *__restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2half_rn(in[i]);
}

}  // namespace apg
