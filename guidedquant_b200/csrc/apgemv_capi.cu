// C-ABI of libapgemv_b200.so (see include/apgemv_b200.h).  Host-side dispatch only: argument
// validation (the checks the reference does with TORCH_CHECK in inference/ap_gemv/gemv.cu:64-90, plus
// the ones it omits), kernel selection (replaces anyprec_matmul, anyprec.cu:587-620) and launch.
#include "apgemv_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include "apgemv_fast.cuh"
#include "apgemv_generic.cuh"

namespace {

thread_local int g_last_cuda_error = 0;

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return APG_ERR_CUDA;
}
#define APG_CUDA(call)                                   \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return cuda_fail(e__);   \
    } while (0)

struct DevInfo {
    int sms = 0;
    bool fast_attr_set[5] = {false, false, false, false, false};
};
DevInfo g_dev[64];

int device_info(DevInfo **out) {
    int dev = 0;
    APG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return APG_ERR_UNSUPPORTED;
    DevInfo &d = g_dev[dev];
    if (d.sms == 0) {
        int sms = 0;
        APG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        d.sms = sms;
    }
    *out = &d;
    return APG_OK;
}

inline bool aligned(const void *p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

constexpr size_t kFastMaxSmem = 100 * 1024;

template <int BITS>
int launch_fast(const void *x, void *out, float *partial, const void *qweight, const void *lut, uint32_t N,
                uint32_t K, uint32_t flags, int ctas_per_sm, cudaStream_t stream, DevInfo *dev) {
    using namespace apg;
    const uint32_t nslab = (K + 4095u) / 4096u;
    const uint32_t groups = nslab >= 4 ? 1u : (4u / nslab);
    const uint32_t warps = nslab * groups;
    const size_t smem = fast_smem_bytes<BITS>(nslab, groups);
    if (smem > kFastMaxSmem) return APG_ERR_UNSUPPORTED;
    if (!dev->fast_attr_set[BITS]) {
        APG_CUDA(cudaFuncSetAttribute(gemv_fast_kernel<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kFastMaxSmem));
        dev->fast_attr_set[BITS] = true;
    }
    // grid: ctas_per_sm CTAs on every SM, one wave; rows are split evenly over all row groups.
    int c = ctas_per_sm;
    if (c <= 0) {
        const uint32_t c_max = warps >= 8 ? 2u : (warps >= 6 ? 2u : (16u / warps));
        const uint32_t target_rows = 2u * FastCfg<BITS>::RB;
        uint32_t want = (N + dev->sms * groups * target_rows - 1) / (dev->sms * groups * target_rows);
        if (want < 1) want = 1;
        if (want > c_max) want = c_max;
        c = (int)want;
    }
    uint32_t grid = (uint32_t)dev->sms * (uint32_t)c;
    const uint32_t max_useful = (N + groups - 1) / groups;  // at least one row per group
    if (grid > max_useful) grid = max_useful;
    if (grid < 1) grid = 1;

    FastParams p;
    p.x = static_cast<const __half *>(x);
    p.W = static_cast<const uint4 *>(qweight);
    p.lut = static_cast<const __half *>(lut);
    p.out = static_cast<__half *>(out);
    p.partial = partial;
    p.N = N;
    p.K = K;
    p.nslab = nslab;
    p.groups = groups;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(warps * 32u);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (flags & APG_FLAG_PDL) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    APG_CUDA(cudaLaunchKernelEx(&cfg, gemv_fast_kernel<BITS>, p));
    return APG_OK;
}

}  // namespace

extern "C" {

int apg_version(void) { return APG_VERSION; }

const char *apg_status_string(int status) {
    switch (status) {
        case APG_OK: return "ok";
        case APG_ERR_NULL: return "null pointer argument";
        case APG_ERR_BITS: return "Bitwidth must be between 2 and 8.";
        case APG_ERR_BATCH: return "batch size M must be between 1 and 8";
        case APG_ERR_SHAPE: return "bad shape: need N >= 1, K >= 32 and K % 32 == 0";
        case APG_ERR_ALIGN: return "pointer not sufficiently aligned (x/out/lut: 2 B; x, qweight: 16 B when K % 128 == 0)";
        case APG_ERR_CUDA: return "CUDA runtime error (see apg_last_cuda_error)";
        case APG_ERR_MODE: return "unknown flag or mode";
        case APG_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}

int apg_last_cuda_error(void) { return g_last_cuda_error; }

int apg_gemv_ex(const void *x, void *out, float *partial_f32, const void *qweight, const void *lut, uint32_t M,
                uint32_t N, uint32_t K, int bits, uint32_t flags, int ctas_per_sm, void *stream_) {
    using namespace apg;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !qweight || !lut || (!out && !partial_f32)) return APG_ERR_NULL;
    if (bits < 2 || bits > 8) return APG_ERR_BITS;
    if (M < 1 || M > 8) return APG_ERR_BATCH;
    if (N < 1 || K < 32 || (K % 32u) != 0) return APG_ERR_SHAPE;
    if (flags & ~(APG_FLAG_REF_ORDER | APG_FLAG_GENERIC | APG_FLAG_PDL)) return APG_ERR_MODE;
    if (!aligned(x, 16) || !aligned(qweight, 4) || !aligned(lut, 2) || (out && !aligned(out, 2)) ||
        (partial_f32 && !aligned(partial_f32, 4)))
        return APG_ERR_ALIGN;

    DevInfo *dev = nullptr;
    int rc = device_info(&dev);
    if (rc != APG_OK) return rc;

    const bool fast_ok = M == 1 && bits <= 4 && (K % 128u) == 0 && K <= 32768u && aligned(qweight, 16) &&
                         aligned(lut, 16) && !(flags & (APG_FLAG_REF_ORDER | APG_FLAG_GENERIC));
    if (fast_ok) {
        if (bits == 2) rc = launch_fast<2>(x, out, partial_f32, qweight, lut, N, K, flags, ctas_per_sm, stream, dev);
        if (bits == 3) rc = launch_fast<3>(x, out, partial_f32, qweight, lut, N, K, flags, ctas_per_sm, stream, dev);
        if (bits == 4) rc = launch_fast<4>(x, out, partial_f32, qweight, lut, N, K, flags, ctas_per_sm, stream, dev);
        if (rc != APG_ERR_UNSUPPORTED) return rc;
    }
    const dim3 block(128), grid((N + 3) / 4);
    if (flags & APG_FLAG_REF_ORDER)
        gemv_generic_kernel<true><<<grid, block, 0, stream>>>(
            static_cast<const __half *>(x), static_cast<const uint32_t *>(qweight), static_cast<const __half *>(lut),
            static_cast<__half *>(out), partial_f32, M, N, K, bits);
    else
        gemv_generic_kernel<false><<<grid, block, 0, stream>>>(
            static_cast<const __half *>(x), static_cast<const uint32_t *>(qweight), static_cast<const __half *>(lut),
            static_cast<__half *>(out), partial_f32, M, N, K, bits);
    APG_CUDA(cudaGetLastError());
    return APG_OK;
}

int apg_gemv(const void *x, void *out, const void *qweight, const void *lut, uint32_t M, uint32_t N, uint32_t K,
             int bits, void *stream) {
    if (!out) return APG_ERR_NULL;
    return apg_gemv_ex(x, out, nullptr, qweight, lut, M, N, K, bits, 0u, 0, stream);
}

int apg_dequant(const void *qweight, const void *lut, void *w_out, uint32_t N, uint32_t K, int bits, void *stream_) {
    using namespace apg;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!qweight || !lut || !w_out) return APG_ERR_NULL;
    if (bits < 2 || bits > 8) return APG_ERR_BITS;
    if (N < 1 || K < 32 || (K % 32u) != 0) return APG_ERR_SHAPE;
    if (!aligned(qweight, 4) || !aligned(lut, 2) || !aligned(w_out, 16)) return APG_ERR_ALIGN;
    dequant_kernel<<<dim3((N + 3) / 4), dim3(128), 0, stream>>>(static_cast<const uint32_t *>(qweight),
                                                               static_cast<const __half *>(lut),
                                                               static_cast<__half *>(w_out), N, K, bits);
    APG_CUDA(cudaGetLastError());
    return APG_OK;
}

int apg_round_f32_to_f16(const float *in, void *out, uint32_t n, void *stream_) {
    using namespace apg;
    if (!in || !out) return APG_ERR_NULL;
    if (n == 0) return APG_OK;
    round_f32_to_f16_kernel<<<dim3((n + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream_)>>>(
        in, static_cast<__half *>(out), n);
    APG_CUDA(cudaGetLastError());
    return APG_OK;
}

}  // extern "C"
