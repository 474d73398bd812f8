// Test infrastructure only (never shipped, never on the product path).
// Thin extern "C" shim over the UNMODIFIED reference translation unit
// /root/reference/inference/ap_gemv/anyprec.cu, which is compiled from where it lies
// (see oracle/build_ref.sh).  The two prototypes below are the reference's own public
// entry points (inference/ap_gemv/anyprec.h:8-27); this file adds nothing but C linkage
// so that tests/bench can call the reference kernels through ctypes on the GPU box.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

void anyprec_matmul(__half *in, __half *out, uint32_t *qweight, __half *lut,
                    uint32_t M, uint32_t N, uint32_t K, int w_bits, cudaStream_t stream);
void anyprec_dequant_kbit(const uint32_t *qweight, const uint32_t N, const uint32_t K,
                          const __half *lut, __half *weight, int w_bits, cudaStream_t stream);

extern "C" int ref_anyprec_gemv(const void *x, void *out, const void *qweight, const void *lut,
                                uint32_t M, uint32_t N, uint32_t K, int bits, void *stream) {
    anyprec_matmul((__half *)x, (__half *)out, (uint32_t *)qweight, (__half *)lut, M, N, K, bits,
                   (cudaStream_t)stream);
    return (int)cudaGetLastError();
}

extern "C" int ref_anyprec_dequant(const void *qweight, const void *lut, void *w_out, uint32_t N,
                                   uint32_t K, int bits, void *stream) {
    anyprec_dequant_kbit((const uint32_t *)qweight, N, K, (const __half *)lut, (__half *)w_out,
                         bits, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}
