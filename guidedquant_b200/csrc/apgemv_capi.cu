// C-ABI of libapgemv_b200.so (see include/apgemv_b200.h).  Host-side dispatch only: argument
// validation (the checks the reference does with TORCH_CHECK in inference/ap_gemv/gemv.cu:64-90, plus
// the ones it omits), kernel selection (replaces anyprec_matmul, anyprec.cu:587-620) and launch.
#include "apgemv_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "apgemv_fast.cuh"
#include "apgemv_generic.cuh"
#include "apgemv_wide.cuh"

namespace {

thread_local int g_last_cuda_error = 0;
struct Fusion {
    const void *norm_w = nullptr;
    float eps = 0.f;
    int silu_mul = 0;
    const void *residual = nullptr;
    uint32_t world = 1, rank = 0;
    void *const *peer_recv = nullptr;  // host array [world] of device pointers
    const void *epoch = nullptr;
};
thread_local const void *g_prefetch_ptr = nullptr;  // one-shot hint consumed by the next fast GEMV launch
thread_local uint64_t g_prefetch_bytes = 0;

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return APG_ERR_CUDA;
}
}  // namespace
// shared with decode_capi.cu: every failing CUDA call of the library lands in apg_last_cuda_error()
int apg_internal_cuda_fail(int e) { return cuda_fail((cudaError_t)e); }
namespace {
#define APG_CUDA(call)                                   \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return cuda_fail(e__);   \
    } while (0)

struct DevInfo {
    int sms = 0;
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

constexpr size_t kFastMaxSmem = 200 * 1024;

struct FastPlan {
    uint32_t cpw, nwk, groups, rs, nslots, stage_bytes, grid, rows_per_cta, threads, unit_rows, tot_units;
    size_t smem;
};

// Work decomposition of the fast kernel: consumer warps per CTA, rows per stage, ring depth, grid.
template <int BITS>
bool plan_fast(uint32_t N, uint32_t K, int ctas_per_sm, int sms, FastPlan *pl) {
    const uint32_t nchunk = (K + 1023u) / 1024u;
    if (nchunk > 32) return false;
    pl->cpw = nchunk > 8u ? 2u : 1u;                             // more than 8 chunks: two chunks per consumer warp
    pl->nwk = (nchunk + pl->cpw - 1) / pl->cpw;                  // <= 16
    pl->groups = pl->nwk >= 8 ? 1u : (8u / pl->nwk);             // ~8 consumer warps per CTA
    const uint32_t ncons = pl->groups * pl->nwk;
    pl->threads = (ncons + 1) * 32u;
    const uint32_t row_bytes = K / 8u * BITS;                     // all planes of one row
    uint32_t rs = 8;
    while (rs > 2 && rs * row_bytes > 32u * 1024u) rs >>= 1;      // stage <= 32 KB (measured best; 16 KB stages are 5-12 % slower)
    pl->rs = rs;
    pl->stage_bytes = rs * row_bytes;
    int c = ctas_per_sm > 0 ? ctas_per_sm : (ncons <= 8 ? 2 : 1);
    uint32_t grid = (uint32_t)sms * (uint32_t)c;
    const uint32_t max_grid = (N + rs - 1) / rs;                  // at least one stage per CTA
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    pl->grid = grid;
    // rows are dealt in half stages: with whole stages an SM ends up with e.g. 4 stages against an average of 3.46
    // (N = 4096, RS = 8, 296 CTAs); a trailing half stage runs the half-length row loop
    pl->unit_rows = rs / 2u;
    pl->tot_units = (N + pl->unit_rows - 1) / pl->unit_rows;
    const uint32_t units_per_cta = (pl->tot_units + grid - 1) / grid;
    const uint32_t stages_per_cta = (units_per_cta * pl->unit_rows + rs - 1) / rs;
    pl->rows_per_cta = stages_per_cta * rs;
    // ring: as deep as the CTA's share needs, bounded so that c CTAs (+ a dependent kernel's) fit in 227 KB
    const size_t ring_budget = (c >= 2 ? 56u : 96u) * 1024u;
    uint32_t ns = (uint32_t)(ring_budget / pl->stage_bytes);
    if (ns > stages_per_cta + pl->groups - 1) ns = stages_per_cta + pl->groups - 1;
    ns = ns / pl->groups * pl->groups;
    if (ns < pl->groups) ns = pl->groups;
    pl->nslots = ns;
    const uint32_t wtb = rs == 8 ? apg::FastWarpTbl<BITS, 8>::BYTES : (rs == 4 ? apg::FastWarpTbl<BITS, 4>::BYTES : apg::FastWarpTbl<BITS, 2>::BYTES);
    pl->smem = 2 * (size_t)ns * 8 + 512 + (size_t)ncons * wtb + (size_t)ns * pl->stage_bytes +
               (size_t)pl->rows_per_cta * pl->nwk * sizeof(float) + 16;
    return pl->smem <= kFastMaxSmem;
}

template <int BITS, int CPW, int RS, bool GLU = false>
int launch_fast_inst(const apg::FastParams &p, const FastPlan &pl, uint32_t flags, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    int dev = 0;
    APG_CUDA(cudaGetDevice(&dev));
    if (!attr_set[dev & 63]) {
        APG_CUDA(cudaFuncSetAttribute(apg::gemv_fast_kernel<BITS, CPW, RS, GLU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kFastMaxSmem));
        attr_set[dev & 63] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.grid);
    cfg.blockDim = dim3(pl.threads);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (flags & APG_FLAG_PDL) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    APG_CUDA(cudaLaunchKernelEx(&cfg, apg::gemv_fast_kernel<BITS, CPW, RS, GLU>, p));
    return APG_OK;
}

template <int BITS>
int launch_fast(const void *x, void *out, float *partial, const void *qweight, const void *lut, uint32_t N,
                uint32_t K, uint32_t flags, int ctas_per_sm, cudaStream_t stream, DevInfo *dev,
                const void *prefetch, uint64_t prefetch_bytes, const Fusion &fu) {
    FastPlan pl;
    if (!plan_fast<BITS>(N, K, ctas_per_sm, dev->sms, &pl)) return APG_ERR_UNSUPPORTED;
    apg::FastParams p;
    p.x = static_cast<const __half *>(x);
    p.W = static_cast<const uint8_t *>(qweight);
    p.lut = static_cast<const __half *>(lut);
    p.out = static_cast<__half *>(out);
    p.partial = partial;
    p.N = N;
    p.K = K;
    p.nwk = pl.nwk;
    p.groups = pl.groups;
    p.nslots = pl.nslots;
    p.stage_bytes = pl.stage_bytes;
    p.unit_rows = pl.unit_rows;
    p.units_q = pl.tot_units / pl.grid;
    p.units_rem = pl.tot_units % pl.grid;
    p.inv_nwk = (65536u + pl.nwk - 1) / pl.nwk;
    p.norm_w = static_cast<const __half *>(fu.norm_w);
    p.norm_eps = fu.eps;
    p.act_silu_mul = fu.silu_mul == 1 ? 1u : 0u;
    if (fu.silu_mul == 2) {  // experimental GLU epilogue: interleaved (gate, up) rows, out has N / 2 elements
        if ((N & 1u) || (pl.unit_rows & 1u) || pl.cpw != 1 || pl.rs != 8 || !out || partial || fu.residual || fu.world > 1)
            return APG_ERR_UNSUPPORTED;
    }
    p.residual = static_cast<const __half *>(fu.residual);
    p.world = fu.world;
    p.rank = fu.rank;
    for (uint32_t i = 0; i < 8; i++)
        p.peer_recv[i] = (fu.world > 1 && i < fu.world) ? static_cast<uint2 *>(fu.peer_recv[i]) : nullptr;
    p.epoch = static_cast<const uint32_t *>(fu.epoch);
    p.prefetch = static_cast<const uint8_t *>(prefetch);
    p.prefetch_bytes = (prefetch && aligned(prefetch, 16)) ? prefetch_bytes : 0;
    if (fu.silu_mul == 2) return launch_fast_inst<BITS, 1, 8, true>(p, pl, flags, stream);
#define APG_FAST_CASE(CPW_, RS_) \
    if (pl.cpw == CPW_ && pl.rs == RS_) return launch_fast_inst<BITS, CPW_, RS_>(p, pl, flags, stream);
    APG_FAST_CASE(1, 8)
    APG_FAST_CASE(1, 4)
    APG_FAST_CASE(1, 2)
    APG_FAST_CASE(2, 8)
    APG_FAST_CASE(2, 4)
    APG_FAST_CASE(2, 2)
#undef APG_FAST_CASE
    return APG_ERR_UNSUPPORTED;
}

// every (bits, M) the fast kernel does not take: bits 5..8, M 2..8, K % 128 != 0, K > 32768 (apgemv_wide.cuh)
template <int BITS, int R>
int launch_wide_rows(const void *x, void *out, float *partial, const void *qweight, const void *lut, uint32_t M,
                     uint32_t N, uint32_t K, cudaStream_t stream) {
    const dim3 block(128), grid((N + 4 * R - 1) / (4 * R));  // 4 warps x R rows
    const __half *xp = static_cast<const __half *>(x), *lp = static_cast<const __half *>(lut);
    const uint32_t *wp = static_cast<const uint32_t *>(qweight);
    __half *op = static_cast<__half *>(out);
    if (M == 1) apg::gemv_wide_kernel<BITS, 1, R><<<grid, block, 0, stream>>>(xp, wp, lp, op, partial, M, N, K);
    else if (M == 2) apg::gemv_wide_kernel<BITS, 2, R><<<grid, block, 0, stream>>>(xp, wp, lp, op, partial, M, N, K);
    else if (M <= 4) apg::gemv_wide_kernel<BITS, 4, R><<<grid, block, 0, stream>>>(xp, wp, lp, op, partial, M, N, K);
    else apg::gemv_wide_kernel<BITS, 8, R><<<grid, block, 0, stream>>>(xp, wp, lp, op, partial, M, N, K);
    APG_CUDA(cudaGetLastError());
    return APG_OK;
}

template <int BITS>
int launch_wide_bits(const void *x, void *out, float *partial, const void *qweight, const void *lut, uint32_t M,
                     uint32_t N, uint32_t K, cudaStream_t stream) {
    // two rows per warp share every activation load; with few rows one row per warp keeps more warps in flight —
    // except where the activation traffic dominates anyway (cheap pair-table dequant against 5..8 batch rows).
    // Measured on B200, N = 4096, K = 4096 (us, R=1 / R=2): 2-bit M=2 5.3 / 6.9, 4-bit M=8 14.5 / 18.2, 8-bit M=1 9.6 / 11.5,
    // 2-bit M=8 16.5 / 13.1.
    if (N <= 8192u && !(BITS <= 3 && M > 4))
        return launch_wide_rows<BITS, 1>(x, out, partial, qweight, lut, M, N, K, stream);
    return launch_wide_rows<BITS, 2>(x, out, partial, qweight, lut, M, N, K, stream);
}

int launch_wide(const void *x, void *out, float *partial, const void *qweight, const void *lut, uint32_t M, uint32_t N,
                uint32_t K, int bits, cudaStream_t stream) {
    switch (bits) {
        case 2: return launch_wide_bits<2>(x, out, partial, qweight, lut, M, N, K, stream);
        case 3: return launch_wide_bits<3>(x, out, partial, qweight, lut, M, N, K, stream);
        case 4: return launch_wide_bits<4>(x, out, partial, qweight, lut, M, N, K, stream);
        case 5: return launch_wide_bits<5>(x, out, partial, qweight, lut, M, N, K, stream);
        case 6: return launch_wide_bits<6>(x, out, partial, qweight, lut, M, N, K, stream);
        case 7: return launch_wide_bits<7>(x, out, partial, qweight, lut, M, N, K, stream);
        case 8: return launch_wide_bits<8>(x, out, partial, qweight, lut, M, N, K, stream);
        default: return APG_ERR_BITS;
    }
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

static int gemv_impl(const void *x, void *out, float *partial_f32, const void *qweight, const void *lut, uint32_t M,
                     uint32_t N, uint32_t K, int bits, uint32_t flags, int ctas_per_sm, void *stream_, const Fusion &fu,
                     bool fused) {
    using namespace apg;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !qweight || !lut || (!out && !partial_f32)) return APG_ERR_NULL;
    if (bits < 2 || bits > 8) return APG_ERR_BITS;
    if (M < 1 || M > 8) return APG_ERR_BATCH;
    if (N < 1 || K < 32 || (K % 32u) != 0) return APG_ERR_SHAPE;
    if (flags & ~(APG_FLAG_REF_ORDER | APG_FLAG_GENERIC | APG_FLAG_PDL)) return APG_ERR_MODE;
    if (!aligned(x, 16) || !aligned(qweight, 4) || !aligned(lut, 2) || (out && !aligned(out, 2)) ||
        (partial_f32 && !aligned(partial_f32, 4)) || (fu.norm_w && !aligned(fu.norm_w, 16)) ||
        (fu.residual && !aligned(fu.residual, 2)))
        return APG_ERR_ALIGN;

    DevInfo *dev = nullptr;
    int rc = device_info(&dev);
    if (rc != APG_OK) return rc;

    const bool fast_ok = M == 1 && bits <= 4 && (K % 128u) == 0 && K <= 32768u && aligned(qweight, 16) &&
                         aligned(lut, 16) && !(flags & (APG_FLAG_REF_ORDER | APG_FLAG_GENERIC));
    const void *pf = g_prefetch_ptr;
    const uint64_t pfb = g_prefetch_bytes;
    g_prefetch_ptr = nullptr, g_prefetch_bytes = 0;
    if (fast_ok) {
        if (bits == 2) rc = launch_fast<2>(x, out, partial_f32, qweight, lut, N, K, flags, ctas_per_sm, stream, dev, pf, pfb, fu);
        if (bits == 3) rc = launch_fast<3>(x, out, partial_f32, qweight, lut, N, K, flags, ctas_per_sm, stream, dev, pf, pfb, fu);
        if (bits == 4) rc = launch_fast<4>(x, out, partial_f32, qweight, lut, N, K, flags, ctas_per_sm, stream, dev, pf, pfb, fu);
        if (rc != APG_ERR_UNSUPPORTED) return rc;
    }
    if (fused) return APG_ERR_UNSUPPORTED;  // the fused prologue/epilogue exists in the fast kernel only
    if (aligned(lut, 16) && !(flags & (APG_FLAG_REF_ORDER | APG_FLAG_GENERIC)))
        return launch_wide(x, out, partial_f32, qweight, lut, M, N, K, bits, stream);
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

int apg_gemv_ex(const void *x, void *out, float *partial_f32, const void *qweight, const void *lut, uint32_t M,
                uint32_t N, uint32_t K, int bits, uint32_t flags, int ctas_per_sm, void *stream) {
    return gemv_impl(x, out, partial_f32, qweight, lut, M, N, K, bits, flags, ctas_per_sm, stream, Fusion(), false);
}

int apg_gemv_fused(const void *x, void *out, float *partial_f32, const void *qweight, const void *lut, uint32_t N,
                   uint32_t K, int bits, const void *norm_w, float norm_eps, int silu_mul, const void *residual,
                   uint32_t flags, void *stream) {
    Fusion fu;
    fu.norm_w = norm_w, fu.eps = norm_eps, fu.silu_mul = silu_mul, fu.residual = residual;
    return gemv_impl(x, out, partial_f32, qweight, lut, 1, N, K, bits, flags, 0, stream, fu, true);
}

int apg_gemv_fused_push(const void *x, const void *qweight, const void *lut, uint32_t N, uint32_t K, int bits,
                        const void *norm_w, float norm_eps, int silu_mul, uint32_t world, uint32_t rank,
                        void *const *peer_recv, const void *epoch, float *scratch_f32, uint32_t flags, void *stream) {
    if (world < 2 || world > 8 || rank >= world || !peer_recv || !epoch || !scratch_f32) return APG_ERR_MODE;
    for (uint32_t i = 0; i < world; i++)
        if (!peer_recv[i] || !aligned(peer_recv[i], 8)) return APG_ERR_NULL;
    Fusion fu;
    fu.norm_w = norm_w, fu.eps = norm_eps, fu.silu_mul = silu_mul;
    fu.world = world, fu.rank = rank, fu.peer_recv = peer_recv, fu.epoch = epoch;
    return gemv_impl(x, nullptr, scratch_f32, qweight, lut, 1, N, K, bits, flags, 0, stream, fu, true);
}

int apg_allreduce_finish(const void *recv, uint32_t *epoch, uint32_t *done_counter, const void *residual, void *out, uint32_t n,
                         uint32_t world, uint32_t flags, void *stream) {
    if (!recv || !epoch || !done_counter || !out) return APG_ERR_NULL;
    if (n == 0 || (n & 1u) || world < 2 || world > 8) return APG_ERR_SHAPE;
    if (!aligned(recv, 16) || !aligned(out, 4) || (residual && !aligned(residual, 4))) return APG_ERR_ALIGN;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((n / 2u + 255u) / 256u);
    cfg.blockDim = dim3(256);
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    if (flags & APG_FLAG_PDL) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    APG_CUDA(cudaLaunchKernelEx(&cfg, apg::allreduce_finish_kernel, static_cast<const uint2 *>(recv), epoch, done_counter,
                                static_cast<const __half *>(residual), static_cast<__half *>(out), n, world));
    return APG_OK;
}

int apg_plan_fast(uint32_t N, uint32_t K, int bits, int ctas_per_sm, int sms, uint32_t plan[16]) {
    if (!plan) return APG_ERR_NULL;
    if (bits < 2 || bits > 4) return APG_ERR_UNSUPPORTED;
    if (N < 1 || K < 128 || (K % 128u) != 0 || K > 32768u || sms < 1) return APG_ERR_UNSUPPORTED;
    FastPlan pl;
    const bool ok = bits == 2 ? plan_fast<2>(N, K, ctas_per_sm, sms, &pl)
                              : (bits == 3 ? plan_fast<3>(N, K, ctas_per_sm, sms, &pl) : plan_fast<4>(N, K, ctas_per_sm, sms, &pl));
    if (!ok) return APG_ERR_UNSUPPORTED;
    const uint32_t v[16] = {pl.cpw, pl.nwk, pl.groups, pl.rs, pl.nslots, pl.stage_bytes, pl.grid, pl.rows_per_cta,
                            pl.threads, pl.unit_rows, pl.tot_units / pl.grid, pl.tot_units % pl.grid, (uint32_t)pl.smem,
                            0u, 0u, 0u};
    for (int i = 0; i < 16; i++) plan[i] = v[i];
    return APG_OK;
}

int apg_prefetch_hint(const void *next_weights, uint64_t bytes) {
    g_prefetch_ptr = next_weights;
    g_prefetch_bytes = next_weights ? bytes : 0;
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
    if (aligned(lut, 16)) {
        const dim3 grid((N + 3) / 4), block(128);
        const uint32_t *wp = static_cast<const uint32_t *>(qweight);
        const __half *lp = static_cast<const __half *>(lut);
        __half *op = static_cast<__half *>(w_out);
        switch (bits) {
            case 2: dequant_wide_kernel<2><<<grid, block, 0, stream>>>(wp, lp, op, N, K); break;
            case 3: dequant_wide_kernel<3><<<grid, block, 0, stream>>>(wp, lp, op, N, K); break;
            case 4: dequant_wide_kernel<4><<<grid, block, 0, stream>>>(wp, lp, op, N, K); break;
            case 5: dequant_wide_kernel<5><<<grid, block, 0, stream>>>(wp, lp, op, N, K); break;
            case 6: dequant_wide_kernel<6><<<grid, block, 0, stream>>>(wp, lp, op, N, K); break;
            case 7: dequant_wide_kernel<7><<<grid, block, 0, stream>>>(wp, lp, op, N, K); break;
            default: dequant_wide_kernel<8><<<grid, block, 0, stream>>>(wp, lp, op, N, K); break;
        }
        APG_CUDA(cudaGetLastError());
        return APG_OK;
    }
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
