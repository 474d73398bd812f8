/*
 * apdecode_b200 — C-ABI of the decode-step kernels around the Any-Precision GEMV.
 * Exported by the same library as apgemv_b200.h (libapgemv_b200.so).  They restate, at batch 1 / sequence 1, the
 * non-Linear ops of the reference's gpt-fast model (inference/model.py) and its sampling (inference/generate.py);
 * RMSNorm, SiLU*mul and the residual adds are NOT here — they are fused into apg_gemv_fused.
 * All pointers are device pointers; `token` and `pos` live in device memory so one CUDA graph serves every token.
 * flags: APG_FLAG_PDL (apgemv_b200.h) launches with programmatic dependent launch.  Returns apg_status codes.
 */
#ifndef APDECODE_B200_H
#define APDECODE_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* x[0:dim] = emb[clamp(*token, 0, vocab-1), :]            tok_embeddings, model.py:123.  dim % 8 == 0. */
int apd_embed(const void *emb, const int *token, void *x, uint32_t dim, uint32_t vocab, uint32_t flags, void *stream);

/* RoPE(q,k) + KV-cache append at *pos + softmax(q.K^T/sqrt(128)).V over t <= *pos     Attention.forward, model.py:206-236
 *   qkv fp16 [(H+2*Hkv)*128] (q|k|v, model.py:211); inv_freq fp32 [64]; k_cache/v_cache fp16 [Hkv, S, 128];
 *   out fp16 [H*128]; part_ws fp32 [H*nsplit*132] (only when nsplit > 1).  head_dim = 128, H/Hkv <= 8.
 *   A position outside [0, S) makes the kernel return without touching the cache or `out`. */
int apd_attn_decode(const void *qkv, const float *inv_freq, void *k_cache, void *v_cache, const int *pos, void *out,
                    float *part_ws, uint32_t H, uint32_t Hkv, uint32_t S, uint32_t nsplit, float scale, uint32_t flags,
                    void *stream);

/* logits[V] (fp16) = W[V,D] . (RMSNorm(x) * norm_w)      Transformer.forward tail, model.py:128-129.  D % 256 == 0, D <= 8192.
 * best_val/best_idx (optional, >= 8*SMs entries each): per-CTA arg-max partials of the fp16 logits for apd_argmax_advance;
 * *n_partials (optional, host) receives how many entries were written (= the grid size). */
int apd_lm_head(const void *x, const void *norm_w, float eps, const void *W, void *logits, uint32_t V, uint32_t D,
                float *best_val, int *best_idx, uint32_t *n_partials, uint32_t row_offset, uint32_t flags, void *stream);
/* row_offset: index of W's first row in the full vocabulary (0 unless lm_head is vocab-sharded): added to best_idx. */

/* greedy sampling (generate.py:55-73 at temperature 0) from apd_lm_head's partials: *token = argmax(logits) (first index
 * on ties; NaN never wins); history[*pos + 1] = *token (if history != NULL and in range); *pos += 1. */
int apd_argmax_advance(const float *best_val, const int *best_idx, uint32_t n, int *token, int *pos, int *history,
                       uint32_t history_len, uint32_t flags, void *stream);

/* Vocab-sharded variant: every rank reduces its partials, pushes (value|index, epoch) packets into slot `rank` of every
 * peer's exchange buffer peer_slots[p] (uint2 [2*world], peer-mapped memory) and polls its own buffer for all `world`
 * winners; the global pick is identical on every rank.  *epoch: device uint32, zero-initialised, advanced per call. */
int apd_argmax_advance_tp(const float *best_val, const int *best_idx, uint32_t n, uint32_t world, uint32_t rank,
                          void *const *peer_slots, uint32_t *epoch, int *token, int *pos, int *history,
                          uint32_t history_len, uint32_t flags, void *stream);

/* temperature / top-k sampling (logits_to_probs + multinomial_sample_one_no_sync, generate.py:55-73) on the fp16 logits
 * of apd_lm_head: *token = argmax_i( logits[i]/max(T,1e-5) - log q_i ) over the top_k largest logits (ties with the k-th
 * kept, like the reference's `logits < pivot` mask; top_k == 0 or >= V: no filter), q_i = -log(u_i), u_i a counter hash
 * of (*seed, *pos, i) - a fixed documented stream, not torch's Philox (restated in oracle/decode_oracle.py).  Then
 * history[*pos + 1] = *token and *pos += 1 like apd_argmax_advance.  seed: device uint64; logits 16-byte aligned.
 * Single GPU; apd_sample_topk_advance_tp is the vocab-sharded form. */
int apd_sample_topk_advance(const void *logits, uint32_t V, float temperature, uint32_t top_k,
                            const unsigned long long *seed, int *token, int *pos, int *history, uint32_t history_len,
                            uint32_t flags, void *stream);

/* The same sampler for a vocab-sharded lm_head (tensor parallel, `logits` = this rank's V_local logits of global indices
 * rank*V_local ..): two data-with-flag packet exchanges through the peers' buffers (peer_slots[r] = rank r's uint2
 * [world][slot_stride] exchange buffer, zero-initialised; *epoch advanced by the call) — the k largest keys of every rank
 * give the global top-k pivot, then every rank's best (score, global index) gives the winner.  The noise is hashed from the
 * GLOBAL index, so the token is the one apd_sample_topk_advance draws from the gathered logits.  top_k <= 256,
 * top_k < V_local (or >= the whole vocabulary = no filter), top_k + 2 <= slot_stride; else APG_ERR_UNSUPPORTED. */
int apd_sample_topk_advance_tp(const void *logits, uint32_t V_local, float temperature, uint32_t top_k,
                               const unsigned long long *seed, uint32_t world, uint32_t rank, void *const *peer_slots,
                               uint32_t slot_stride, uint32_t *epoch, int *token, int *pos, int *history, uint32_t history_len,
                               uint32_t flags, void *stream);

#ifdef __cplusplus
}
#endif
#endif
