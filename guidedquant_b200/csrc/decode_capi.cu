// C-ABI of the decode-step kernels (include/apdecode_b200.h): validation + launch only.
#include "apdecode_b200.h"

#include <cuda_runtime.h>
#include <stdlib.h>

#include "apgemv_b200.h"
#include "decode_kernels.cuh"

int apg_internal_cuda_fail(int e);  // apgemv_capi.cu: records the error for apg_last_cuda_error()

namespace {
template <typename... KArgs, typename... Args>
int launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, uint32_t flags, void *stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    if (flags & APG_FLAG_PDL) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
    return e == cudaSuccess ? APG_OK : apg_internal_cuda_fail((int)e);
}
inline bool al(const void *p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }
}  // namespace

extern "C" {

int apd_embed(const void *emb, const int *token, void *x, uint32_t dim, uint32_t vocab, uint32_t flags, void *stream) {
    if (!emb || !token || !x) return APG_ERR_NULL;
    if (dim == 0 || dim % 8 || vocab == 0) return APG_ERR_SHAPE;
    if (!al(emb, 16) || !al(x, 16)) return APG_ERR_ALIGN;
    return launch(apd::embed_kernel, dim3(1), dim3(256), 0, flags, stream, static_cast<const __half *>(emb), token,
                  static_cast<__half *>(x), dim, vocab);
}

int apd_attn_decode(const void *qkv, const float *inv_freq, void *k_cache, void *v_cache, const int *pos, void *out,
                    float *part_ws, uint32_t H, uint32_t Hkv, uint32_t S, uint32_t nsplit, float scale, uint32_t flags,
                    void *stream) {
    if (!qkv || !inv_freq || !k_cache || !v_cache || !pos || !out) return APG_ERR_NULL;
    if (H == 0 || Hkv == 0 || H % Hkv || H / Hkv > 8 || S == 0 || nsplit == 0) return APG_ERR_SHAPE;
    if (nsplit > 1 && !part_ws) return APG_ERR_NULL;
    if (!al(qkv, 8) || !al(k_cache, 8) || !al(v_cache, 8) || !al(out, 8) || (part_ws && !al(part_ws, 16))) return APG_ERR_ALIGN;
    int rc = launch(apd::attn_decode_kernel, dim3(H, nsplit), dim3(32 * apd::kAttnWarps), 0, flags, stream,
                    static_cast<const __half *>(qkv), inv_freq, static_cast<__half *>(k_cache),
                    static_cast<__half *>(v_cache), pos, static_cast<__half *>(out), part_ws, H, Hkv, S, scale);
    if (rc != APG_OK || nsplit == 1) return rc;
    return launch(apd::attn_merge_kernel, dim3(H), dim3(apd::kHeadDim), 0, flags, stream,
                  static_cast<const float *>(part_ws), static_cast<__half *>(out), nsplit);
}

int apd_lm_head(const void *x, const void *norm_w, float eps, const void *W, void *logits, uint32_t V, uint32_t D,
                float *best_val, int *best_idx, uint32_t *n_partials, uint32_t row_offset, uint32_t flags, void *stream) {
    if (!x || !norm_w || !W || !logits) return APG_ERR_NULL;
    if (V == 0 || D == 0 || D % 256 || D > 8192) return APG_ERR_SHAPE;
    if (!al(W, 16) || !al(x, 2) || !al(norm_w, 2)) return APG_ERR_ALIGN;
    int dev = 0, sms = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (ce != cudaSuccess) return apg_internal_cuda_fail((int)ce);
    const uint32_t nv = D / 256;
    const uint32_t per_sm = 6u;  // CTAs per SM, measured on B200: 2 -> 5.1 TB/s, 4 -> 6.5, 6 -> 6.6
    uint32_t grid = (uint32_t)sms * per_sm;
    if (grid * 8u > V) grid = (V + 7) / 8;
    if (n_partials) *n_partials = grid;
    const size_t smem = (size_t)D * sizeof(float);
    auto go = [&](auto kern) {
        return launch(kern, dim3(grid), dim3(256), smem, flags, stream, static_cast<const __half *>(x),
                      static_cast<const __half *>(norm_w), eps, static_cast<const __half *>(W),
                      static_cast<__half *>(logits), V, D, best_val, best_idx, row_offset);
    };
    switch (nv) {
        case 1: return go(apd::lm_head_kernel<1>);
        case 2: return go(apd::lm_head_kernel<2>);
        case 4: return go(apd::lm_head_kernel<4>);
        case 8: return go(apd::lm_head_kernel<8>);
        case 16: return go(apd::lm_head_kernel<16>);
        case 32: return go(apd::lm_head_kernel<32>);
        default: return APG_ERR_UNSUPPORTED;
    }
}

int apd_argmax_advance(const float *best_val, const int *best_idx, uint32_t n, int *token, int *pos, int *history,
                       uint32_t history_len, uint32_t flags, void *stream) {
    if (!best_val || !best_idx || !token || !pos) return APG_ERR_NULL;
    if (n == 0) return APG_ERR_SHAPE;
    return launch(apd::argmax_advance_kernel, dim3(1), dim3(n >= 512 ? 1024 : 256), 0, flags, stream, best_val, best_idx, n,
                  token, pos, history, history_len);
}

int apd_argmax_advance_tp(const float *best_val, const int *best_idx, uint32_t n, uint32_t world, uint32_t rank,
                          void *const *peer_slots, uint32_t *epoch, int *token, int *pos, int *history,
                          uint32_t history_len, uint32_t flags, void *stream) {
    if (!best_val || !best_idx || !token || !pos || !peer_slots || !epoch) return APG_ERR_NULL;
    if (n == 0 || world < 2 || world > 8 || rank >= world) return APG_ERR_SHAPE;
    uint2 *pp[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (uint32_t i = 0; i < world; i++) {
        if (!peer_slots[i] || !al(peer_slots[i], 8)) return APG_ERR_ALIGN;
        pp[i] = static_cast<uint2 *>(peer_slots[i]);
    }
    return launch(apd::argmax_advance_tp_kernel, dim3(1), dim3(n >= 512 ? 1024 : 256), 0, flags, stream, best_val, best_idx,
                  n, world, rank, pp[0], pp[1], pp[2], pp[3], pp[4], pp[5], pp[6], pp[7], epoch, token, pos, history,
                  history_len);
}

int apd_sample_topk_advance(const void *logits, uint32_t V, float temperature, uint32_t top_k,
                            const unsigned long long *seed, int *token, int *pos, int *history, uint32_t history_len,
                            uint32_t flags, void *stream) {
    if (!logits || !seed || !token || !pos) return APG_ERR_NULL;
    if (V == 0 || !(temperature >= 0.f)) return APG_ERR_SHAPE;
    if (!al(logits, 16) || !al(seed, 8)) return APG_ERR_ALIGN;
    return launch(apd::sample_topk_advance_kernel, dim3(1), dim3(apd::kSampleThreads), 0, flags, stream,
                  static_cast<const __half *>(logits), V, temperature, top_k, seed, token, pos, history, history_len);
}

int apd_sample_topk_advance_tp(const void *logits, uint32_t V_local, float temperature, uint32_t top_k,
                               const unsigned long long *seed, uint32_t world, uint32_t rank, void *const *peer_slots,
                               uint32_t slot_stride, uint32_t *epoch, int *token, int *pos, int *history, uint32_t history_len,
                               uint32_t flags, void *stream) {
    if (!logits || !seed || !token || !pos || !peer_slots || !epoch) return APG_ERR_NULL;
    if (V_local == 0 || !(temperature >= 0.f) || world < 2 || world > 8 || rank >= world) return APG_ERR_SHAPE;
    if (!al(logits, 16) || !al(seed, 8)) return APG_ERR_ALIGN;
    if ((uint64_t)top_k >= (uint64_t)V_local * world) top_k = 0;  // no filter, like the single-GPU sampler
    if (top_k > apd::kSampleTpMaxK || top_k >= V_local || top_k + 2 > slot_stride) return APG_ERR_UNSUPPORTED;
    uint2 *pp[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (uint32_t i = 0; i < world; i++) {
        if (!peer_slots[i] || !al(peer_slots[i], 8)) return APG_ERR_ALIGN;
        pp[i] = static_cast<uint2 *>(peer_slots[i]);
    }
    return launch(apd::sample_topk_advance_tp_kernel, dim3(1), dim3(apd::kSampleThreads), 0, flags, stream,
                  static_cast<const __half *>(logits), V_local, temperature, top_k, seed, world, rank, pp[0], pp[1], pp[2], pp[3],
                  pp[4], pp[5], pp[6], pp[7], slot_stride, epoch, token, pos, history, history_len);
}

}  // extern "C"
