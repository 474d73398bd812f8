/*
 * apgemv_b200 — C-ABI of the B200-native Any-Precision LUT GEMV (sm_100a).
 *
 * This header is the drop-in boundary for the reference's native extension `ap_gemv`
 * (snu-mllab/GuidedQuant, inference/ap_gemv).  Each entry point names the reference interface
 * it replaces; paths are relative to the reference root.  Plain pointers and sizes only: all
 * pointers are DEVICE pointers on the current CUDA device, `stream` is a cudaStream_t / CUstream
 * handle passed as void* (NULL = legacy default stream).  Every function returns 0 on success or
 * one of the APG_ERR_* codes below; nothing here calls exit() or assert() (the reference does:
 * inference/ap_gemv/gemv.cu:20-27, anyprec.cu:602).  Launches are asynchronous on `stream`
 * and are CUDA-graph capturable; launch errors are reported (the reference does not check them,
 * anyprec.cu:617-619).
 *
 * Tensor layouts (identical to the reference):
 *   x        fp16 [M, K]            row-major (the reference's input [M,1,K], gemv.cu:41-43)
 *   qweight  int32 [bits, N, K/32]  bit-plane major, MSB plane first, warp-permuted words
 *                                   (any_precision/quantization/pack.py:12-83,304-321)
 *   lut      fp16 [N, 2^bits]       per-output-row centroids
 *   out      fp16 [M, N]            overwritten (anyprec.cu:538), never accumulated
 */
#ifndef APGEMV_B200_H
#define APGEMV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APG_VERSION 202 /* major*100 + minor */

enum apg_status {
    APG_OK = 0,
    APG_ERR_NULL = 1,     /* a required pointer is NULL */
    APG_ERR_BITS = 2,     /* bits outside 2..8            (gemv.cu:64) */
    APG_ERR_BATCH = 3,    /* M outside 1..8               (anyprec.cu:602) */
    APG_ERR_SHAPE = 4,    /* K == 0, K % 32 != 0 or N == 0 (gemv.cu:76, APLinear.py:17) */
    APG_ERR_ALIGN = 5,    /* a pointer is not aligned to its element/vector size */
    APG_ERR_CUDA = 6,     /* the CUDA runtime reported an error; see apg_last_cuda_error() */
    APG_ERR_MODE = 7,     /* unknown flag / mode / world layout */
    APG_ERR_UNSUPPORTED = 8
};

/* flags for apg_gemv_ex */
#define APG_FLAG_REF_ORDER 0x1u /* reproduce the reference kernel's fp16 accumulation order bit for
                                   bit (matmul_kbit_32, anyprec.cu:424-541); slow, for parity tests */
#define APG_FLAG_GENERIC 0x2u   /* force the generic (untuned) kernel */
#define APG_FLAG_PDL 0x4u       /* launch with programmatic dependent launch: weight/LUT prefetch
                                   overlaps the tail of the previous kernel on the stream */

int apg_version(void);
const char *apg_status_string(int status);
/* cudaError_t of the last failing runtime call made by this library on the calling thread. */
int apg_last_cuda_error(void);

/*
 * Replaces ap_gemv.anyprec_gemv(input, output, qweight, lut, bitwidth)
 *   (inference/ap_gemv/gemv.cu:96-107 -> anyprec_matmul, anyprec.cu:591-620).
 * out[m, n] = sum_k lut[n, idx[n, k]] * x[m, k]  for m < M (1..8), bits in 2..8.
 * Requirements: K % 32 == 0 (layout), any N >= 1 (the reference silently drops rows when
 * N % 4 != 0, anyprec.cu:614; this library computes them).
 */
int apg_gemv(const void *x, void *out, const void *qweight, const void *lut,
             uint32_t M, uint32_t N, uint32_t K, int bits, void *stream);

/*
 * apg_gemv with options.
 *   flags        APG_FLAG_*.
 *   partial_f32  optional fp32 [M, N] output: if non-NULL the un-rounded fp32 sums are written
 *                there as well (used by the K-sharded multi-GPU path, where partial sums are
 *                all-reduced in fp32 and rounded once); `out` may then be NULL.
 *   ctas_per_sm  0 = heuristic; otherwise forces the grid to ctas_per_sm * SM count (tuning).
 */
int apg_gemv_ex(const void *x, void *out, float *partial_f32, const void *qweight, const void *lut,
                uint32_t M, uint32_t N, uint32_t K, int bits, uint32_t flags, int ctas_per_sm,
                void *stream);

/*
 * M = 1 GEMV fused with the element-wise ops that surround a Linear in the reference's decode step
 * (inference/model.py), so that a transformer block is 5 launches instead of 11:
 *   norm_w   != NULL: x := (fp16)(x * rsqrt(mean(x^2) + norm_eps)) * norm_w       RMSNorm.forward, model.py:280-285
 *   silu_mul == 1   : x := silu(x[0:K]) * x[K:2K]  (x holds 2K halfs)             FeedForward.forward, model.py:261-266
 *   silu_mul == 2   : EXPERIMENTAL (built, not yet measured): the rows of THIS Linear are interleaved (gate_0, up_0, gate_1,
 *                     ...) and out[N/2] receives silu(y[2i]) * y[2i+1] with the roundings of the silu_mul == 1 prologue it
 *                     replaces; needs N even, K <= 8192 (one chunk per warp), no residual / partial / multi-GPU push,
 *                     else APG_ERR_UNSUPPORTED
 *   residual != NULL: out := (fp16)y + residual  (fp16 add)                       TransformerBlock.forward, model.py:151-167
 * Fast-path shapes only (bits 2..4, K % 128 == 0, K <= 32768); otherwise APG_ERR_UNSUPPORTED.
 */
int apg_gemv_fused(const void *x, void *out, float *partial_f32, const void *qweight, const void *lut,
                   uint32_t N, uint32_t K, int bits, const void *norm_w, float norm_eps, int silu_mul,
                   const void *residual, uint32_t flags, void *stream);

/*
 * K-sharded multi-GPU Linear with the all-reduce FUSED into the GEMV (no NCCL on the data path).  Every rank calls
 *   apg_gemv_fused_push  : GEMV over its K shard (same optional RMSNorm / SiLU*mul prologue as apg_gemv_fused); the
 *                          epilogue writes each row's fp32 partial sum TOGETHER WITH THE EPOCH as one 8-byte store into
 *                          slot `rank` of EVERY peer's receive buffer peer_recv[p] (uint2 [world][N], peer-mapped device
 *                          memory such as torch symmetric memory) over NVLink: data and flag travel together, so there
 *                          are no fences or atomics and the cost is one one-way NVLink latency.
 *   apg_allreduce_finish : one small kernel (one thread per pair of elements) that polls this rank's `world` x N packets for
 *                          the epoch, adds the `world` values in rank order (deterministic), adds the optional fp16
 *                          residual, rounds to fp16; its last CTA (ticket in *done_counter, a zeroed device uint32 per
 *                          site) advances *epoch.  N even.
 * peer_recv is a HOST array of `world` device pointers; `epoch` is a device uint32 (zero-initialised, one per all-reduce
 * site, advanced identically on every rank) so the same CUDA graph can be replayed; scratch_f32 [N] receives the local
 * un-rounded sums.  world in 2..8.  New functionality: the reference has no multi-GPU inference path.
 */
int apg_gemv_fused_push(const void *x, const void *qweight, const void *lut, uint32_t N, uint32_t K, int bits,
                        const void *norm_w, float norm_eps, int silu_mul, uint32_t world, uint32_t rank,
                        void *const *peer_recv, const void *epoch, float *scratch_f32, uint32_t flags, void *stream);
int apg_allreduce_finish(const void *recv, uint32_t *epoch, uint32_t *done_counter, const void *residual, void *out, uint32_t n, uint32_t world,
                         uint32_t flags, void *stream);

/*
 * Optional one-shot hint (per calling thread), consumed by the NEXT apg_gemv / apg_gemv_ex launch that takes the fast
 * path: `next_weights[0, bytes)` (16-byte aligned; normally the qweight tensor of the Linear that follows on the
 * stream) is pulled into L2 by that launch's producer threads while its warps compute.  Purely a performance hint:
 * results never depend on it.  No reference counterpart (the reference issues one isolated launch per Linear).
 */
int apg_prefetch_hint(const void *next_weights, uint64_t bytes);

/*
 * Introspection of the fast kernel's work decomposition (host only, no GPU needed): how apg_gemv would cut an
 * N x K, `bits`-bit Linear on a device with `sms` SMs.  Returns APG_ERR_UNSUPPORTED when the fast kernel does not take
 * the shape.  plan[0..15] = { chunks_per_warp, chunk_warps_per_group, groups, rows_per_stage, ring_slots, stage_bytes,
 * grid, red_rows_per_cta, threads, unit_rows, units_q, units_rem, smem_bytes, 0, 0, 0 }: CTA b owns the rows
 * [u*unit_rows, (u + n)*unit_rows) with u = b*units_q + min(b, units_rem), n = units_q + (b < units_rem), walked in
 * stages of rows_per_stage rows dealt round-robin to the groups.  Used by the CPU tests to check that every row is
 * covered exactly once and that the shared-memory budget holds for every shape.
 */
int apg_plan_fast(uint32_t N, uint32_t K, int bits, int ctas_per_sm, int sms, uint32_t plan[16]);

/*
 * Replaces ap_gemv.anyprec_dequant(qweight, lut, bitwidth) -> fp16 [N, K]
 *   (inference/ap_gemv/gemv.cu:109-134 -> dequant_kbit_store, anyprec.cu:294-359, 622-645).
 * w_out[n, k] = lut[n, idx[n, k]] — a pure gather, bit-identical to the reference.
 * The caller allocates w_out (N*K fp16); the torch layer allocates it like the reference does.
 */
int apg_dequant(const void *qweight, const void *lut, void *w_out,
                uint32_t N, uint32_t K, int bits, void *stream);

/*
 * Prefill: replaces the reference's dequantise-then-matmul branch for more than 8 tokens
 *   (inference/ap_gemv/APLinear.py:35-38, any_precision/modules/AnyPrecisionLinear.py forward: anyprec_dequant + torch.matmul).
 * out[t, n] = sum_k x[t, k] * lut[n, idx[n, k]], x fp16 [T, K] row-major, out fp16 [T, N], fp32 accumulation, one rounding.
 * ONE kernel: the packed weights are dequantised straight into the shared-memory operand tiles of tcgen05.mma (the fp16 [N, K]
 * weight never exists in HBM), token tiles arrive by TMA, accumulators live in TMEM (csrc/prefill_tc.cuh).
 * Supported: bits 2..4, K % 256 == 0, x / qweight / lut 16-byte aligned; anything else returns APG_ERR_UNSUPPORTED and the
 * caller keeps the apg_dequant + library-GEMM route.  Small T*N splits K over CTAs and needs `workspace`
 * (apg_prefill_plan reports the bytes; without it the call still succeeds, with one CTA per tile walking the whole K).
 * plan[8] = {tokens per CTA, token tiles, row tiles, K splits, ring stages, dynamic smem bytes, TMEM columns, K/256}.
 */
int apg_prefill_plan(uint32_t T, uint32_t N, uint32_t K, int bits, int sms, uint32_t plan[8], uint64_t *workspace_bytes);
int apg_prefill_gemm(const void *x, void *out, const void *qweight, const void *lut, uint32_t T, uint32_t N, uint32_t K, int bits,
                     void *workspace, uint64_t workspace_bytes, void *stream);

/* fp32 [n] -> fp16 [n] round-to-nearest-even; the epilogue of the K-sharded path after the fp32 all-reduce. */
int apg_round_f32_to_f16(const float *in, void *out, uint32_t n, void *stream);

/*
 * ---------------------------------------------------------------------------------------------------------------
 * Persistent token engine: ONE cooperative kernel launch runs a whole list of dependent jobs (the fused GEMVs of a decode
 * step, the attention of every block, the embedding row, all-reduce finishers).  At batch 1 every Linear consumes the whole
 * output vector of the previous one; as separate launches each of the ~160 hand-overs of a token costs ~3 us on B200 (PDL),
 * inside this kernel activations travel as 8-byte (half2, tag) packets that consumers spin on — no grid barrier, no kernel
 * boundary — while a producer thread per CTA keeps the packed weights of the NEXT Linears streaming into a shared-memory
 * ring.  Replaces the reference's per-op launches of APLinear.forward + Inductor-fused glue under CUDA graphs
 * (inference/generate.py:330-336, inference/model.py:151-167, 206-236, 261-285).  GEMV arithmetic is that of apg_gemv_fused.
 *
 * Usage: fill `n_jobs` descriptors of apg_persist_job_bytes() bytes each in HOST memory with apg_persist_job_*, copy the
 * table to the device, then apg_persist_launch once per token.  "LL buffer" = device array of n/2 uint2 packets for a vector
 * of n halfs (zero-initialised).  tag_* = index of the job that produces the packets being read / written (a job writes
 * its own index); `sms` = SM count of the device that will run the table (rows are dealt to that many CTAs).
 */
uint32_t apg_persist_job_bytes(void);
/* shared-memory budget for a job list whose largest GEMV input has max_k elements (dynamic bytes, weight-ring bytes) */
int apg_persist_smem(int bits, uint32_t max_k, uint32_t *total_bytes, uint32_t *ring_bytes);
/* flags: 1 = RMSNorm prologue (norm_w, norm_eps), 4 = residual add (LL buffer), 8 = SwiGLU epilogue (rows interleaved
 * (gate_i, up_i), out has N/2 halfs), 16 = K-sharded push: fp32 partial sums + tag to slot `rank` of every rank's receive
 * buffer peers[r] (uint2 [world][N]); bits 2..4, K % 128 == 0, K <= 16384, 4 rows of all planes <= 32 KB, N even. */
int apg_persist_job_gemv(void *job, uint32_t N, uint32_t K, int bits, int sms, uint32_t flags, const void *x, const void *qweight,
                         const void *lut, void *out, void *out_plain, const void *norm_w, float norm_eps, const void *residual,
                         uint32_t world, uint32_t rank, void *const *peers, uint32_t tag_x, uint32_t tag_res, uint32_t tag_out);
/* RoPE + KV append at *pos + attention for head h on CTA h (Attention.forward, model.py:206-236); head_dim 128, H <= SMs;
 * rope_cs: fp16 [S][128] = cos(pos * inv_freq[0..63]) | sin(...) computed in fp32 and rounded (model.py:396-405) */
int apg_persist_job_attn(void *job, const void *qkv, const void *rope_cs, void *k_cache, void *v_cache, void *out, void *out_plain,
                         uint32_t H, uint32_t Hkv, uint32_t S, float scale, uint32_t tag_x, uint32_t tag_out);
/* out := packets of the fp16 row src_rows[clamp(*row_index)] (row 0 if row_index is NULL): tok_embeddings, model.py:123 */
int apg_persist_job_pack(void *job, const void *src_rows, const int *row_index, uint32_t n, uint32_t n_rows, void *out, void *out_plain,
                         uint32_t tag_out);
/* out := fp16( sum over ranks of the fp32 packets in recv [world][N] ) + residual: finisher of a K-sharded push job */
int apg_persist_job_reduce(void *job, const void *recv, uint32_t N, uint32_t world, const void *residual, void *out, void *out_plain,
                           uint32_t tag_x, uint32_t tag_res, uint32_t tag_out);
/* epoch: device uint32 token counter (zero-initialised; advanced by the launch when bump_epoch != 0); pos: device int
 * position read by attention jobs; max_k: largest K over the GEMV jobs (sizes the shared-memory x staging); err_word: device uint32, non-zero after a launch = a device-side watchdog fired;
 * done_counter: device uint32 scratch (zero).  flags bit 0: launch WITHOUT the cooperative attribute (debug).
 * prof: NULL, or device int64 [SMs][n_jobs][4] receiving clock64 stamps per CTA and job (start, x loaded, stages done, end). */
int apg_persist_launch(const void *jobs_dev, uint32_t n_jobs, int bits, uint32_t max_k, uint32_t *epoch, const int *pos,
                       uint32_t *err_word, uint32_t *done_counter, int bump_epoch, uint32_t flags, void *prof, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* APGEMV_B200_H */
