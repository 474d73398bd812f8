// C-ABI of the prefill path (include/apgemv_b200.h, "Prefill" section): planning, TMA descriptor, launch.
#include "apgemv_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "prefill_tc.cuh"

int apg_internal_cuda_fail(int e);  // apgemv_capi.cu

namespace {
using namespace apg::ptc;

#define PTC_CUDA(call)                                                    \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) return apg_internal_cuda_fail((int)e__);  \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is not linked: the driver's tensor-map encoder is looked up through the runtime
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

struct Plan {
    uint32_t t_tile, tok_tiles, row_tiles, splits, stages, smem_bytes, tmem_cols, sb_total;
    uint64_t workspace_bytes;
};

uint32_t stage_base(int bits) { return bits == 2 ? Lay<2>::STAGE_BASE : bits == 3 ? Lay<3>::STAGE_BASE : Lay<4>::STAGE_BASE; }

int make_plan(uint32_t T, uint32_t N, uint32_t K, int bits, int sms, int allow_split, Plan *pl) {
    if (bits < 2 || bits > 4) return APG_ERR_UNSUPPORTED;
    if (T < 1 || N < 1 || K < 256 || (K % 256u) != 0) return APG_ERR_UNSUPPORTED;
    if ((uint64_t)T * N >= (1ull << 40)) return APG_ERR_SHAPE;
    const uint32_t tok_tiles = (T + 255) / 256;
    uint32_t t_tile = ((T + tok_tiles - 1) / tok_tiles + 31) / 32 * 32;
    pl->t_tile = t_tile;
    pl->tok_tiles = (T + t_tile - 1) / t_tile;
    pl->row_tiles = (N + ROWS - 1) / ROWS;
    pl->sb_total = K / 256;
    const uint64_t ctas = (uint64_t)pl->tok_tiles * pl->row_tiles;
    uint32_t splits = 1;
    if (allow_split && ctas < (uint64_t)sms) {
        splits = (uint32_t)((uint64_t)sms / ctas);
        if (splits > pl->sb_total / 2) splits = pl->sb_total / 2;  // at least two super blocks (512 k) per split
        if (splits < 1) splits = 1;
    }
    pl->splits = splits;
    pl->workspace_bytes = splits > 1 ? (uint64_t)splits * T * N * sizeof(float) : 0;
    const uint32_t base = stage_base(bits);
    const uint32_t per_stage = t_tile * (BK * 2);  // token tile in shared memory; the weight tile of a stage lives in TMEM
    uint32_t stages = (SMEM_LIMIT - 1024u - base) / per_stage;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages > (512u - t_tile) / A_COLS) stages = (512u - t_tile) / A_COLS;
    // up to 128 tokens per tile: keep accumulator + A ring within 256 TMEM columns, so that TWO CTAs run on an SM at once
    // (a second CTA would otherwise sit in tcgen05.alloc until the first one exits) and overlap their set-up / epilogue
    if (t_tile <= 128u && stages > (256u - t_tile) / A_COLS) stages = (256u - t_tile) / A_COLS;
    if (stages < 2) return APG_ERR_UNSUPPORTED;
    pl->stages = stages;
    pl->smem_bytes = base + stages * per_stage;  // dynamic shared memory starts at (or near) shared address 0
    uint32_t cols = 32;
    while (cols < t_tile + stages * A_COLS) cols *= 2;
    pl->tmem_cols = cols;
    return APG_OK;
}

template <int BITS, int MINB>
int launch(const CUtensorMap &map, const Params &p, const Plan &pl, cudaStream_t stream) {
    static bool attr_set[64] = {};
    int dev = 0;
    PTC_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        PTC_CUDA(cudaFuncSetAttribute(prefill_tc_kernel<BITS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_LIMIT - 1024u)));
        attr_set[dev] = true;
    }
    const dim3 grid(pl.row_tiles, pl.tok_tiles, pl.splits);
    prefill_tc_kernel<BITS, MINB><<<grid, THREADS, pl.smem_bytes, stream>>>(map, p);
    PTC_CUDA(cudaGetLastError());
    return APG_OK;
}
}  // namespace

extern "C" {

int apg_prefill_plan(uint32_t T, uint32_t N, uint32_t K, int bits, int sms, uint32_t plan[8], uint64_t *workspace_bytes) {
    if (!plan) return APG_ERR_NULL;
    Plan pl;
    const int st = make_plan(T, N, K, bits, sms > 0 ? sms : 148, 1, &pl);
    if (st != APG_OK) return st;
    plan[0] = pl.t_tile, plan[1] = pl.tok_tiles, plan[2] = pl.row_tiles, plan[3] = pl.splits;
    plan[4] = pl.stages, plan[5] = pl.smem_bytes, plan[6] = pl.tmem_cols, plan[7] = pl.sb_total;
    if (workspace_bytes) *workspace_bytes = pl.workspace_bytes;
    return APG_OK;
}

int apg_prefill_gemm(const void *x, void *out, const void *qweight, const void *lut, uint32_t T, uint32_t N, uint32_t K, int bits,
                     void *workspace, uint64_t workspace_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !out || !qweight || !lut) return APG_ERR_NULL;
    if (bits < 2 || bits > 8) return APG_ERR_BITS;
    if (T < 1 || N < 1 || K < 32 || (K % 32u) != 0) return APG_ERR_SHAPE;
    if (((uintptr_t)x & 15) || ((uintptr_t)qweight & 15) || ((uintptr_t)lut & 15) || ((uintptr_t)out & 1)) return APG_ERR_ALIGN;
    int dev = 0, sms = 0;
    PTC_CUDA(cudaGetDevice(&dev));
    PTC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    Plan pl;
    int st = make_plan(T, N, K, bits, sms, 1, &pl);
    if (st != APG_OK) return st;
    if (pl.splits > 1 && (!workspace || workspace_bytes < pl.workspace_bytes || ((uintptr_t)workspace & 15))) {
        st = make_plan(T, N, K, bits, sms, 0, &pl);  // no (or too small a) workspace: one CTA walks the whole K
        if (st != APG_OK) return st;
    }
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return APG_ERR_UNSUPPORTED;
    CUtensorMap map;
    const cuuint64_t gdim[2] = {K, T};
    const cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {BK, pl.t_tile};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(x), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return apg_internal_cuda_fail((int)cudaErrorInvalidValue);
    Params p;
    p.W = static_cast<const uint32_t *>(qweight);
    p.lut = static_cast<const __half *>(lut);
    p.out = static_cast<__half *>(out);
    p.partial = static_cast<float *>(workspace);
    p.N = N, p.K = K, p.T = T;
    p.t_tile = pl.t_tile, p.stages = pl.stages, p.splits = pl.splits, p.sb_total = pl.sb_total, p.tmem_cols = pl.tmem_cols;
    // instruction descriptor (kind::f16): D fp32, A/B fp16, both K-major, N = t_tile, M = 128
    p.idesc = (1u << 4) | ((pl.t_tile >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
    // small token tiles on a grid of more CTAs than SMs: the 96-register build, two CTAs per SM (measured: 28672x4096 3-bit,
    // 16 tokens 63 -> 38 us; on grids that give every SM one CTA anyway the tighter register budget only costs)
    const bool two = pl.t_tile <= 128u && (uint64_t)pl.row_tiles * pl.tok_tiles * pl.splits > (uint64_t)sms;
    switch (bits) {
        case 2: st = two ? launch<2, 2>(map, p, pl, stream) : launch<2, 1>(map, p, pl, stream); break;
        case 3: st = two ? launch<3, 2>(map, p, pl, stream) : launch<3, 1>(map, p, pl, stream); break;
        default: st = launch<4, 1>(map, p, pl, stream); break;
    }
    if (st != APG_OK) return st;
    if (pl.splits > 1) {
        const uint64_t total = (uint64_t)T * N;
        reduce_splits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p.partial, p.out, total, pl.splits);
        PTC_CUDA(cudaGetLastError());
    }
    return APG_OK;
}

}  // extern "C"
