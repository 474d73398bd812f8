// Shared device helpers for the Any-Precision LUT GEMV kernels (sm_100a).
//
// Packed layout recap (reference: any_precision/quantization/pack.py:12-83):
//   qweight[j][n][w], w = i*32 + t; bit (31 - (8c+e)) of that word is plane-j's bit of
//   k = i*1024 + c*8*eff + 8t + e,   eff = 32 for full 1024-chunks, (K%1024)/32 for the tail chunk.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace apg {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// (a & m) | (b & ~m) — exactly one LOP3 (immLut 0xE4); written as asm because the C expression is
// lowered to two LOP3s when m is an immediate (two different immediates m and ~m).
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t m) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
}

// Streaming 128-bit load of weight bit-planes: read once, keep out of L1.
__device__ __forceinline__ uint4 ldg_stream_v4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t lds_b32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <int IMM>
__device__ __forceinline__ uint32_t lds_b32_imm(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}
template <int IMM>
__device__ __forceinline__ uint32_t lds_u16_imm(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1+%2];" : "=h"(v) : "r"(addr), "n"(IMM));
    return v;
}

__device__ __forceinline__ uint32_t hfma2_u32(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t hadd2_u32(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// acc + lo(v) + hi(v) with the adds done in fp32: two FHADD (mixed-precision add, PTX add.rn.f32.f16, sm_100+)
__device__ __forceinline__ float acc_add_h2(float acc, uint32_t v) {
    asm("{\n\t.reg .b16 lo, hi;\n\t"
        "mov.b32 {lo, hi}, %1;\n\t"
        "add.rn.f32.f16 %0, lo, %0;\n\t"
        "add.rn.f32.f16 %0, hi, %0;\n\t}"
        : "+f"(acc)
        : "r"(v));
    return acc;
}

// Programmatic dependent launch (PDL) controls.  No-ops when the kernel was launched without the
// programmatic-stream-serialization attribute.
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

// eff of chunk i (see header comment).
__device__ __forceinline__ uint32_t chunk_eff(uint32_t K, uint32_t i) {
    return (i < (K >> 10)) ? 32u : ((K & 1023u) >> 5);
}

// Transposing warp reduction: every lane holds RB partial sums (one per row of the batch); on return
// lanes [8r', 8r'+8) ... hold the full sum of one row.  Returns the sum of row `row_of_lane(lane)`.
//   RB = 4: 6 shuffles (instead of 20);  row_of_lane = lane >> 3
//   RB = 2: 5 shuffles;                  row_of_lane = lane >> 4
//   RB = 1: 5 shuffles;                  row_of_lane = 0
template <int RB>
__device__ __forceinline__ float batch_reduce(const float (&s)[RB], int lane);

template <>
__device__ __forceinline__ float batch_reduce<2>(const float (&s)[2], int lane) {
    const bool hi = lane & 16;
    float keep = hi ? s[1] : s[0], send = hi ? s[0] : s[1];
    keep += __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
    return keep;
}
template <>
__device__ __forceinline__ float batch_reduce<4>(const float (&s)[4], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8;
    float k0 = b4 ? s[2] : s[0], t0 = b4 ? s[0] : s[2];
    float k1 = b4 ? s[3] : s[1], t1 = b4 ? s[1] : s[3];
    k0 += __shfl_xor_sync(0xffffffffu, t0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, t1, 16);
    float keep = b3 ? k1 : k0, send = b3 ? k0 : k1;
    keep += __shfl_xor_sync(0xffffffffu, send, 8);
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
    return keep;
}

}  // namespace apg
