// C-ABI of the persistent token kernel (include/apgemv_b200.h, "persistent engine"): host-side planning of the job
// descriptors + the cooperative launch.  No torch types; the Python runtime (guidedquant_b200/persist.py) assembles a job
// list in host memory with apg_persist_job_*, copies it to the device once and replays apg_persist_launch per token.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "apgemv_b200.h"
#include "apgemv_persist.cuh"

int apg_internal_cuda_fail(int e);  // apgemv_capi.cu

namespace {

inline bool al(const void *p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

constexpr uint32_t kSmemTotal = 227u * 1024u - 2048u;  // dynamic part; 2 KB are left for the kernel's static shared header (PkShared)

template <int BITS>
constexpr uint32_t ring_rel() {
    return apg::PK_NCW * apg::FastWarpTbl<BITS, 8>::BYTES + apg::PK_SCRATCH_BYTES;
}
uint32_t ring_rel_bits(int bits) { return bits == 2 ? ring_rel<2>() : (bits == 3 ? ring_rel<3>() : ring_rel<4>()); }

template <int BITS>
int launch_bits(const apg::PParams &p, int sms, bool cooperative, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return apg_internal_cuda_fail((int)e);
    if (!attr_set[dev & 63]) {
        e = cudaFuncSetAttribute(apg::decode_persistent_kernel<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
        if (e != cudaSuccess) return apg_internal_cuda_fail((int)e);
        attr_set[dev & 63] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)sms);
    cfg.blockDim = dim3(apg::PK_THREADS);
    cfg.dynamicSmemBytes = kSmemTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (cooperative) {  // co-residency of all CTAs is part of the kernel's contract (consumers spin on other CTAs' packets)
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    e = cudaLaunchKernelEx(&cfg, apg::decode_persistent_kernel<BITS>, p);
    if (e != cudaSuccess) return apg_internal_cuda_fail((int)e);
    return APG_OK;
}

}  // namespace

extern "C" {

uint32_t apg_persist_job_bytes(void) { return (uint32_t)sizeof(apg::PJob); }

int apg_persist_smem(int bits, uint32_t max_k, uint32_t *total_bytes, uint32_t *ring_bytes) {
    if (bits < 2 || bits > 4) return APG_ERR_UNSUPPORTED;
    const uint32_t xs = (2u * max_k + 1023u) & ~1023u;
    if (ring_rel_bits(bits) + xs + 2u * apg::PK_MAX_STAGE > kSmemTotal) return APG_ERR_UNSUPPORTED;
    if (total_bytes) *total_bytes = kSmemTotal;
    if (ring_bytes) *ring_bytes = (kSmemTotal - ring_rel_bits(bits) - xs) & ~1023u;
    return APG_OK;
}

int apg_persist_job_gemv(void *job, uint32_t N, uint32_t K, int bits, int sms, uint32_t flags, const void *x, const void *qweight,
                         const void *lut, void *out, void *out_plain, const void *norm_w, float norm_eps, const void *residual,
                         uint32_t world, uint32_t rank, void *const *peers, uint32_t tag_x, uint32_t tag_res, uint32_t tag_out) {
    using namespace apg;
    if (!job || !x || !qweight || !lut) return APG_ERR_NULL;
    if (bits < 2 || bits > 4) return APG_ERR_UNSUPPORTED;
    if (N < 2 || (N & 1u) || K < 128 || (K % 128u) != 0 || K > 32768u || sms < 1) return APG_ERR_SHAPE;
    if (flags & ~(PF_NORM | PF_RESIDUAL | PF_GLU | PF_PUSH)) return APG_ERR_MODE;
    if (!al(x, 16) || !al(qweight, 16) || !al(lut, 16) || (out && !al(out, 16)) || (out_plain && !al(out_plain, 4)) ||
        (norm_w && !al(norm_w, 16)) || (residual && !al(residual, 8)))
        return APG_ERR_ALIGN;
    if ((flags & PF_NORM) && !norm_w) return APG_ERR_NULL;
    if ((flags & PF_RESIDUAL) && !residual) return APG_ERR_NULL;
    if (!(flags & PF_PUSH) && !out) return APG_ERR_NULL;
    PJob jb;
    memset(&jb, 0, sizeof(jb));
    jb.type = PJ_GEMV, jb.flags = flags, jb.N = N, jb.K = K;
    const uint32_t nchunk = (K + 1023u) / 1024u;
    if (nchunk > PK_NCW) return APG_ERR_UNSUPPORTED;  // one 1024-chunk per consumer warp: K <= 16384
    jb.cpw = 1u;
    jb.nwk = nchunk;                                  // <= 16
    jb.groups = PK_NCW / jb.nwk;                      // >= 1
    jb.inv_nwk = (65536u + jb.nwk - 1) / jb.nwk;
    const uint32_t row_bytes = K / 8u * (uint32_t)bits;
    uint32_t rs = 8;                                  // rows per stage: 8, or 4 when 8 rows exceed a 32 KB stage
    if (rs * row_bytes > PK_MAX_STAGE) rs = 4;
    if (rs * row_bytes > PK_MAX_STAGE) return APG_ERR_UNSUPPORTED;
    jb.rs = rs;
    jb.stage_bytes = rs * row_bytes;
    if (flags & PF_GLU) {
        if ((N & 3u) || (flags & (PF_RESIDUAL | PF_PUSH))) return APG_ERR_UNSUPPORTED;
    }
    if (flags & PF_PUSH) {
        if (world < 2 || world > 8 || rank >= world || !peers || (flags & PF_RESIDUAL)) return APG_ERR_MODE;
        for (uint32_t i = 0; i < world; i++) {
            if (!peers[i] || !al(peers[i], 16)) return APG_ERR_ALIGN;
            jb.peer[i] = peers[i];
        }
        jb.world = world, jb.rank = rank;
    }
    jb.unit_rows = (flags & PF_GLU) ? 4u : rs / 2u;   // 4 or 2 rows: even, so an output packet never straddles two CTAs
    const uint32_t tot_units = (N + jb.unit_rows - 1) / jb.unit_rows;
    jb.units_q = tot_units / (uint32_t)sms, jb.units_rem = tot_units % (uint32_t)sms;
    const uint32_t units_per_cta = (tot_units + sms - 1) / sms;
    const uint32_t rows_per_cta = ((units_per_cta * jb.unit_rows + rs - 1) / rs) * rs;
    if ((size_t)rows_per_cta * jb.nwk * sizeof(float) > PK_SCRATCH_BYTES) return APG_ERR_UNSUPPORTED;
    jb.eps = norm_eps;
    jb.x = x, jb.W = qweight, jb.lut = lut, jb.norm_w = norm_w, jb.residual = residual, jb.out = out, jb.out_plain = out_plain;
    jb.tag_x = tag_x, jb.tag_res = tag_res, jb.tag_out = tag_out;
    memcpy(job, &jb, sizeof(jb));
    return APG_OK;
}

int apg_persist_job_attn(void *job, const void *qkv, const void *rope_cs, void *k_cache, void *v_cache, void *out, void *out_plain,
                         uint32_t H, uint32_t Hkv, uint32_t S, float scale, uint32_t tag_x, uint32_t tag_out) {
    using namespace apg;
    if (!job || !qkv || !rope_cs || !k_cache || !v_cache || !out) return APG_ERR_NULL;
    if (H == 0 || Hkv == 0 || H % Hkv || S == 0) return APG_ERR_SHAPE;
    if (!al(qkv, 16) || !al(k_cache, 16) || !al(v_cache, 16) || !al(out, 16) || (out_plain && !al(out_plain, 8))) return APG_ERR_ALIGN;
    PJob jb;
    memset(&jb, 0, sizeof(jb));
    jb.type = PJ_ATTN, jb.a0 = H, jb.a1 = Hkv, jb.a2 = S, jb.f0 = scale;
    jb.x = qkv, jb.out = out, jb.out_plain = out_plain;
    jb.p0 = const_cast<void *>(rope_cs), jb.p1 = k_cache, jb.p2 = v_cache;
    jb.tag_x = tag_x, jb.tag_out = tag_out;
    memcpy(job, &jb, sizeof(jb));
    return APG_OK;
}

int apg_persist_job_pack(void *job, const void *src_rows, const int *row_index, uint32_t n, uint32_t n_rows, void *out, void *out_plain,
                         uint32_t tag_out) {
    using namespace apg;
    if (!job || !src_rows || !out) return APG_ERR_NULL;
    if (n < 2 || (n & 1u) || n_rows == 0) return APG_ERR_SHAPE;
    if (!al(src_rows, 4) || !al(out, 16) || (out_plain && !al(out_plain, 4))) return APG_ERR_ALIGN;
    PJob jb;
    memset(&jb, 0, sizeof(jb));
    jb.type = PJ_PACK, jb.N = n, jb.a0 = n_rows;
    jb.x = src_rows, jb.p0 = const_cast<int *>(row_index), jb.out = out, jb.out_plain = out_plain;
    jb.tag_out = tag_out;
    memcpy(job, &jb, sizeof(jb));
    return APG_OK;
}

int apg_persist_job_reduce(void *job, const void *recv, uint32_t N, uint32_t world, const void *residual, void *out, void *out_plain,
                           uint32_t tag_x, uint32_t tag_res, uint32_t tag_out) {
    using namespace apg;
    if (!job || !recv || !out) return APG_ERR_NULL;
    if (N < 2 || (N & 1u) || world < 2 || world > 8) return APG_ERR_SHAPE;
    if (!al(recv, 16) || !al(out, 16) || (residual && !al(residual, 8)) || (out_plain && !al(out_plain, 4))) return APG_ERR_ALIGN;
    PJob jb;
    memset(&jb, 0, sizeof(jb));
    jb.type = PJ_REDUCE, jb.N = N, jb.world = world, jb.flags = residual ? PF_RESIDUAL : 0u;
    jb.x = recv, jb.residual = residual, jb.out = out, jb.out_plain = out_plain;
    jb.tag_x = tag_x, jb.tag_res = tag_res, jb.tag_out = tag_out;
    memcpy(job, &jb, sizeof(jb));
    return APG_OK;
}

int apg_persist_launch(const void *jobs_dev, uint32_t n_jobs, int bits, uint32_t max_k, uint32_t *epoch, const int *pos,
                       uint32_t *err_word, uint32_t *done_counter, int bump_epoch, uint32_t flags, void *prof, void *stream) {
    if (!jobs_dev || !epoch || !err_word || !done_counter) return APG_ERR_NULL;
    if (bits < 2 || bits > 4) return APG_ERR_UNSUPPORTED;
    if (n_jobs == 0) return APG_ERR_SHAPE;
    if (!al(jobs_dev, 16)) return APG_ERR_ALIGN;
    int dev = 0, sms = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (ce != cudaSuccess) return apg_internal_cuda_fail((int)ce);
    apg::PParams p;
    p.jobs = static_cast<const apg::PJob *>(jobs_dev);
    p.n_jobs = n_jobs;
    p.xs_bytes = (2u * max_k + 1023u) & ~1023u;  // staging area of a job's input vector
    if (max_k < 128u || ring_rel_bits(bits) + p.xs_bytes + 2u * apg::PK_MAX_STAGE > kSmemTotal) return APG_ERR_UNSUPPORTED;
    p.ring_bytes = (kSmemTotal - ring_rel_bits(bits) - p.xs_bytes) & ~1023u;
    p.epoch = epoch, p.pos = pos, p.err = err_word, p.done = done_counter, p.bump_epoch = bump_epoch ? 1u : 0u;
    p.prof = static_cast<long long *>(prof);
    const bool coop = !(flags & 1u);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (bits == 2) return launch_bits<2>(p, sms, coop, s);
    if (bits == 3) return launch_bits<3>(p, sms, coop, s);
    return launch_bits<4>(p, sms, coop, s);
}

}  // extern "C"
