/* decaf_b200.h — C ABI of libdecaf_b200.so (hand-written sm_100a kernels for the
 * DeCaf-Grounder inference hot path of ZijiaLewisLu/CVPR2025-DeCafNet).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant,
 *     allocates nothing (callers pass workspaces) and returns 0 on success, non-zero on
 *     error (decaf_last_error() gives the message; no exceptions cross the ABI);
 *   - activations are channels-last: a logical tensor (n_seq, rows, C) is addressed as
 *     base + (seq * seq_stride + row) * ld + c.  `act dtype` is DECAF_F32 or DECAF_BF16;
 *     the residual stream, LayerNorm statistics, softmax and all accumulators are fp32;
 *   - masks are uint8 (torch.bool compatible), 1 = valid.
 *
 * "replaces:" cites the reference code each entry point stands in for (paths relative to the
 * reference repository root).
 */
#ifndef DECAF_B200_H
#define DECAF_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DECAF_F32  0
#define DECAF_BF16 1

#define DECAF_ACT_NONE 0
#define DECAF_ACT_RELU 1
#define DECAF_ACT_GELU 2   /* exact erf GELU (nn.GELU default) */

#define DECAF_MAX_LEVELS 16

const char *decaf_last_error(void);
int decaf_version(void);
/* 1 when the running device is compute capability 10.x (tcgen05 path usable) */
int decaf_device_is_sm100(void);
/* Launch width of the persistent tensor-core kernels (decaf_gemm's tcgen05 path, decaf_ffn): at most n_sms CTAs per launch
 * (even; 0 = every SM, the default; environment DECAF_GEMM_SMS presets it).  A caller that keeps several videos in flight
 * on different streams sets a fraction of the device: launches of different videos then run side by side, and the
 * fixed start-up / drain time of a launch idles that fraction of the SMs instead of all of them (measured on the NLQ
 * shape: 13.7k -> 15.8k pairs/s with 48 of 148 SMs per launch and 8 videos in flight).  Results do not depend on it.
 * Process-wide; returns the previous value.  No reference counterpart (scheduling only). */
int decaf_set_gemm_sms(int32_t n_sms);

/* Geometry of the level-major, zero-row-padded point layout used by the heads:
 * per query Pp = 1 + sum_l (len[l] + 1) rows; level l occupies rows [off[l], off[l]+len[l]),
 * every level is preceded and followed by a row that is kept zero, so a k=3 convolution over
 * the flat row axis sees the reference's per-level zero padding. */
typedef struct {
    int32_t n_levels;
    int32_t Pp;
    int32_t off[DECAF_MAX_LEVELS];
    int32_t len[DECAF_MAX_LEVELS];
} decaf_levels_t;

/* ------------------------------------------------------------------ GEMM / conv1d
 * out[seq, t, n] = epi( sum_{tap,k} A[seq, t + (tap - taps/2) * dil, k] * W[n, tap, k] )
 * rows outside [0, rows_per_seq) contribute zero (conv zero padding).
 *   v = acc + bias[n];  [v = LN_n(v) * ln_w[n] + ln_b[n]];  v = act(v);  v *= colscale[n];
 *   v += resid[seq,t,n];  v += pe[t,n];  v *= rowmask[seq,t];  out_f32 <- v;  out_act <- (act dtype) v
 * replaces: every nn.Conv1d with groups == 1 on the path — MaskedConv1D
 * (libs/modeling/blocks.py:87-106), MaskedMHA query/key/value/proj (:182-185,348-350,392),
 * FFN fc/proj (:530-538), the residual/LayerScale/mask glue of TransformerEncoder (:584-590)
 * and TransformerDecoder (:648-649).
 * impl: 0 = auto (tcgen05 when dtype is bf16 and shapes allow, else SIMT), 1 = SIMT fp32-FMA,
 *       2 = tcgen05 (error if not applicable). */
typedef struct {
    const void *A; int32_t dtype; int64_t lda; int64_t a_seq_stride;   /* rows */
    int32_t n_seq, rows_per_seq;
    const void *W;               /* (N, taps, K) contiguous, act dtype */
    int32_t N, K, taps, dil;
    const float *bias;           /* (N) or NULL */
    int32_t act;
    const float *colscale;       /* (N) or NULL */
    const float *resid; int64_t ldr; int64_t r_seq_stride;             /* fp32 or NULL */
    const uint8_t *rowmask; int64_t m_seq_stride;                      /* or NULL */
    float *out_f32; int64_t ldo; int64_t o_seq_stride;                 /* or NULL */
    void *out_act; int64_t ldo2; int64_t o2_seq_stride;                /* or NULL; act dtype */
    /* grouped launch (e.g. q/k/v): group g adds these element strides */
    int32_t n_group;
    int64_t g_stride_a, g_stride_w, g_stride_bias, g_stride_out_f32, g_stride_out_act;
    int32_t impl;
    /* fused channel LayerNorm of the conv output (tcgen05 path only, N <= 512, n_group == 1):
     * ln != 0 inserts  v = LN(acc + bias) [* ln_w + ln_b]  (two-pass, biased variance, eps inside
     * the sqrt) before the activation.  pe (rows_per_seq, N) fp32 is added after resid, before
     * the row mask.  replaces: the MaskedConv1D -> LayerNorm -> ReLU (-> +PE) chains of
     * libs/modeling/head.py:55-58,97-100 and libs/modeling/video_net.py:139-152. */
    int32_t ln; const float *ln_w, *ln_b; float ln_eps;
    const float *pe;
} decaf_gemm_t;
int decaf_gemm(const decaf_gemm_t *p, void *stream);
/* debug only: per-role clock64 stamps of CTA 0 of the following tcgen05 launches into
 * buf[3][2048] (device memory); NULL switches the trace off. */
int decaf_debug_gemm_trace(unsigned long long *buf);

/* ------------------------------------------------------------------ row-wise kernels
 * Channel LayerNorm of each row (two-pass, biased variance, eps inside sqrt):
 *   y = LN(x) [* w + b]; y = relu(y) if relu; y += pe[t]; y *= rowmask -> out_f32 / out_act
 * replaces: LayerNorm.forward (libs/modeling/blocks.py:125-131) and the F.relu / abs-PE /
 * mask glue around it (video_net.py:139-152, head.py:57-58,99-100, fusion.py:59). */
typedef struct {
    const float *x; int64_t ldx; int64_t x_seq_stride;
    int32_t n_seq, rows_per_seq, C;
    const float *w, *b;          /* (C) or NULL,NULL (affine=False) */
    float eps; int32_t relu;
    const float *pe;             /* (rows_per_seq, C) or NULL */
    const uint8_t *rowmask; int64_t m_seq_stride;
    float *out_f32; int64_t ldo; int64_t o_seq_stride;
    void *out_act; int32_t dtype; int64_t ldo2; int64_t o2_seq_stride;
} decaf_layernorm_t;
int decaf_layernorm(const decaf_layernorm_t *p, void *stream);

/* Pre-attention block of ConvAttNLayer / ConvXAttNLayer:
 *   ln  = LN(x; w_pre, b_pre) * mask_in                       (x is stored masked)
 *   br_j = LN( depthwise_conv3(ln; wd_j, stride) ; w_j, b_j ) for j < n_branch  -> out_act[j]
 *   skip = masked_max_pool(x, k=3, stride) * mask_out          (stride 2 only)  -> skip_out
 *   mask_out[t] = mask_in[stride * t]                          (stride 2 only)
 * replaces: TransformerEncoder.forward's ln_attn + attn_skip (libs/modeling/blocks.py:581-585,
 * masked_max_pool1d :31-47), ConvAttNLayer.forward q/k/v conv + norm (:462-469),
 * ConvXAttNLayer.forward (:513-516) with TransformerDecoder's ln_xattn_q (:638-639). */
typedef struct {
    const float *x; int32_t n_seq, T_in, C, stride;
    const uint8_t *mask_in; int64_t mi_seq_stride;
    const float *w_pre, *b_pre;
    int32_t n_branch;
    const float *wd;             /* (n_branch, C, 3) */
    const float *w_br, *b_br;    /* (n_branch, C) */
    float eps;
    void *out_act; int32_t dtype; int64_t out_branch_stride;   /* elements between branches */
    float *skip_out;             /* (n_seq, T_out, C) or NULL */
    uint8_t *mask_out; int64_t mo_seq_stride;                  /* or NULL (stride 1) */
} decaf_preattn_t;
int decaf_preattn(const decaf_preattn_t *p, void *stream);

/* AdaLN + ln_ffn of TransformerDecoder:  q' = (LN_noaffine(q) * ss[:, :C] + ss[:, C:]) * mask;
 * out_q <- q' (fp32, may alias q);  out_act <- LN(q'; w_ffn, b_ffn)
 * replaces: libs/modeling/blocks.py:643-648. */
typedef struct {
    const float *q; int32_t rows, C;
    const void *ss; int32_t ss_dtype;      /* (rows, 2C) */
    const uint8_t *rowmask;                /* (rows) */
    const float *w_ffn, *b_ffn; float eps;
    float *out_q; void *out_act; int32_t dtype;
} decaf_adaln_t;
int decaf_adaln(const decaf_adaln_t *p, void *stream);

/* ------------------------------------------------------------------ attention
 * Banded (local-window) multi-head self-attention: key j in [t-s, t+s], -inf outside the
 * sequence, additive -1e4 on masked keys, zero rows for masked queries; scale d^-1/2 total.
 * q,k,v,out: (n_seq, T, C) act dtype, C = n_heads * d.
 * replaces: MaskedMHA.forward local branch (libs/modeling/blocks.py:357-373) and its chunked
 * helpers _query_key_matmul / _attn_normalize / _attn_value_matmul (:224-325). */
int decaf_local_attn(const void *q, const void *k, const void *v, void *out, int32_t dtype,
                     int32_t n_seq, int32_t T, int32_t C, int32_t n_heads, int32_t window,
                     const uint8_t *mask, int64_t m_seq_stride, void *stream);
/* The same with the 16-row tiles of the tensor-core kernel shifted to start at step -phase (0..15): a time shard whose
 * window starts at global step w0 (at this level's resolution) passes phase = w0 mod 16, which makes every row's fp32
 * summation order — hence its bf16-rounded output — identical to the unsharded run's. */
int decaf_local_attn_phase(const void *q, const void *k, const void *v, void *out, int32_t dtype,
                     int32_t n_seq, int32_t T, int32_t C, int32_t n_heads, int32_t window,
                     const uint8_t *mask, int64_t m_seq_stride, int32_t phase, void *stream);

/* FP32 configuration on the tensor cores (the reference disables TF32, eval.py:40-41, so its convolutions are fp32):
 * every fp32 operand row of K values becomes 3K bf16 values, x = hi + lo with hi = bf16(x), lo = bf16(x - hi);
 * order 0 = [hi | hi | lo] (activations), order 1 = [hi | lo | hi] (weights).  decaf_gemm over the K-concatenated
 * operands (K' = 3K, bf16) then accumulates hi.hi + hi.lo + lo.hi in fp32: the fp32 product up to 2^-16 relative per
 * term (measured ~1e-5 on whole outputs, against the 1e-3 bar of the FP32 configuration), with every fused epilogue of
 * the tcgen05 kernel available.  src: rows x K fp32 with pitch ld_src; dst: rows x 3K bf16, dense.
 * replaces: the operand side of nn.Conv1d in fp32 (libs/modeling/blocks.py MaskedConv1D, :60-110). */
int decaf_split_bf16x3(const float *src, int64_t rows, int32_t K, int64_t ld_src, void *dst, int32_t order, void *stream);

/* Global attention of Tq queries over a short key/value set (text tokens):
 * softmax over keys j < kv_len[seq] (-inf on the rest), no query masking.
 * q/out: (n_seq, Tq, C) q_dtype/out_dtype; k, v: (n_seq, Lk, C) fp32.
 * replaces: MaskedMHA.forward global branch (libs/modeling/blocks.py:374-389) as used by the
 * fusion cross-attention (ConvXAttNLayer :517) and the text encoder self-attention. */
int decaf_xattn(const void *q, int32_t q_dtype, const float *k, const float *v, void *out,
                int32_t out_dtype, int32_t n_seq, int32_t Tq, int32_t Lk, int32_t C,
                int32_t n_heads, const int32_t *kv_len, void *stream);

/* The same attention with the keys/values of every sequence converted ONCE into the bf16 shared-memory image the
 * tensor-core kernel reads (K rows [16 ceil(Lk/16)][C + 8], then V transposed [C][16 ceil(Lk/16) + 8]; keys >= kv_len
 * zero), instead of by every CTA of every launch: the text keys/values of a video are produced once by the text side
 * (blocks.py:640-641) and consumed by n_query * T / 128 CTAs per fusion layer.
 *   decaf_xattn_packed_supported(Lk, C, n_heads): 1 when the tensor-core kernel covers the shape (bf16 queries,
 *       head dim 32 or 64, Lk <= 64); decaf_xattn_packed_elems: bf16 elements of the packed buffer for n_seq sequences.
 *   decaf_xattn_pack_kv: k, v (n_seq, Lk, C) fp32 -> packed;  decaf_xattn_packed: q/out (n_seq, Tq, C) bf16.
 * replaces: the same lines as decaf_xattn (MaskedMHA.forward global branch, libs/modeling/blocks.py:374-389). */
int64_t decaf_xattn_packed_elems(int32_t n_seq, int32_t Lk, int32_t C);
int decaf_xattn_packed_supported(int32_t Lk, int32_t C, int32_t n_heads);
int decaf_xattn_pack_kv(const float *k, const float *v, const int32_t *kv_len, void *packed, int32_t n_seq,
                        int32_t Lk, int32_t C, void *stream);
int decaf_xattn_packed(const void *q, const void *packed, void *out, int32_t n_seq, int32_t Tq, int32_t Lk,
                       int32_t C, int32_t n_heads, const int32_t *kv_len, void *stream);

/* ------------------------------------------------------------------ saliency / selection / merge
 * correl[q, t] = sum_h v^[h,t] * t^[q,h]; with norm: x / (||x||_2 + 1e-4) on both sides.
 * shallow: (Cs, T) fp32 with T contiguous (the reference's layout), text_cls: (n_query, Cs).
 * replaces: libs/modeling/model.py:500-505. */
int decaf_saliency(const float *shallow, const float *text_cls, float *correl,
                   int32_t Cs, int32_t T, int32_t n_query, int32_t norm, void *stream);

/* Exact top-k block selection per query (libs/modeling/model.py:531-541):
 * vid_len = sum(vid_mask); M = ceil(vid_len / sn) block means (last block over its valid
 * count); k = (int)(sratio * M) in double, k == 0 selects all; stable ascending rank (ties:
 * higher index ranks higher); sel[q, t] = block_selected[min((int)floorf(t * (float)M /
 * (float)vid_len), M-1)] for t < vid_len else 0.  pooled (n_query, max_blocks) is optional.
 * out_mask[q,t] = vid_mask[t] & sel[q,t] when and_mask (msf == False) else vid_mask[t]. */
int decaf_select(const float *correl, const uint8_t *vid_mask, uint8_t *sel, uint8_t *out_mask,
                 float *pooled, int32_t max_blocks, int32_t T, int32_t n_query, int32_t sn,
                 double sratio, int32_t and_mask, int32_t *vid_len_out, void *stream);

/* Merge into the dense per-query timeline, transposing to channels-last:
 * x0[q, t, :] = [ vid[:, t] * sel[q,t] | shallow[:, t] | correl[q,t] ] * out_mask[q,t]
 * (expert part present iff Ce > 0, sidekick iff Cs > 0, correl channel iff scat), zero
 * padded to ldx columns.  vid/shallow: (C, T) fp32, T contiguous.
 * replaces: libs/modeling/model.py:543-554 (vid * all_weight, cat, MaskedConv1D's x * mask). */
int decaf_merge(const float *vid, int32_t Ce, const float *shallow, int32_t Cs,
                const float *correl, int32_t scat, const uint8_t *sel, const uint8_t *out_mask,
                void *x0, int32_t dtype, int64_t ldx, int32_t T, int32_t n_query, void *stream);

/* vid_map by linearity: X[q,t,:] = mask[q,t] * ( sel[q,t] * E[t,:] + S[t,:] + bias (+ correl[q,t] * wc) ) with
 * E = W_e vid and S = W_s shallow (T, C) fp32 computed once per video by decaf_gemm (either may be NULL: msf == False has
 * no sidekick part, sfonly no expert part).  Equals decaf_merge + the n_query * T-row vid_map GEMM up to fp32 summation
 * order.  replaces: libs/modeling/model.py:543-555 (vid * all_weight, cat, vid_map) for all queries of a video. */
int decaf_map_combine(const float *E, const float *S, const float *bias, const float *correl,
                      const float *wc, const uint8_t *sel, const uint8_t *mask, float *X, int32_t T,
                      int32_t C, int32_t n_query, void *stream);

/* Compact expert-feature ingest: dense[c, index[k]] = compact[c, k] for k < K, all other steps zero.  compact: (Ce, K)
 * fp32 with row pitch ld; index: (K,) clip positions (out-of-range entries are ignored); dense: (Ce, T) fp32.
 * Expert features are only needed at the clips some query selected (libs/modeling/model.py:543 multiplies the others by
 * zero), so a deployment computes / ships only those. */
int decaf_scatter_clips(const float *compact, int64_t ld, const int32_t *index, int32_t Ce, int32_t K,
                        float *dense, int32_t T, void *stream);

/* fp32 -> bf16 (round to nearest even) of n contiguous values, n % 4 == 0: the token features of the bf16 configuration's
 * tensor-core text encoder (the operands of its first GEMM, libs/modeling/text_net.py:163-166). */
int decaf_cast_bf16(const float *in, void *out, int64_t n, void *stream);

/* ------------------------------------------------------------------ pyramid masks / heads
 * hmask (n_query, Pp): level 0 rows <- mask0[q, t]; level l rows <- level l-1 mask at 2t;
 * pad rows <- 0.  replaces: the nearest mask down-sampling of MaskedConv1D (blocks.py:101-105).*/
int decaf_build_masks(const uint8_t *mask0, int64_t m0_seq_stride, uint8_t *hmask,
                      const decaf_levels_t *lv, int32_t n_query, void *stream);

/* Final k=3 conv of a head tower over the padded flat layout, N_out in {1, 2}:
 *   v = conv3(x)[row] + bias;  mode 0: out = v;  mode 1: out = relu(level_scale[level] * v)
 * x: (rows, C) act dtype (stored masked, pad rows zero); out: (rows, n_out) fp32.
 * replaces: ClsHead.cls_head / RegHead.reg_head + Scale + relu (libs/modeling/head.py:59-60,
 * 101-103). */
int decaf_head_out(const void *x, int32_t dtype, int64_t ldx, int32_t rows_total, int32_t C,
                   const float *w /* (n_out, 3, C) */, const float *bias, int32_t n_out,
                   int32_t mode, const float *level_scale, const decaf_levels_t *lv,
                   float *out, void *stream);

/* Iterative refinement (TCN) between the first and second heads.
 * tcn_in:   r0[q,t,:] = W_in * [l0[t], l1[t>>1]*m, ..., lL[t>>L]*m] + b_in   (nearest expand)
 * tcn_layer: r' = LN32(( r + W1 * relu(Wd (*)dil r + bd) + b1 ) * m)
 * tcn_out:  y = (W_out r + b_out) * m  -> cat[q, off0 + t, col0 : col0+R]  (act dtype)
 * refine_pool: level l columns <- masked max-pool(k3,s2) of level l-1 columns
 * replaces: fuse_and_predict's expand / refine / down-sample / cat (libs/modeling/model.py:
 * 449-467), TCN.forward and DilatedResidualLayer.forward (libs/modeling/tcn.py:21-38,66-84). */
int decaf_tcn_in(const float *logits1, const uint8_t *hmask, const decaf_levels_t *lv,
                 const float *w_in /* (R, L) */, const float *b_in, int32_t R, float *r0,
                 int32_t n_query, void *stream);
int decaf_tcn_layer(const float *r_in, float *r_out, const uint8_t *mask0, int64_t m_seq_stride,
                    const float *wd /* (R,R,3) */, const float *bd, const float *w1 /* (R,R) */,
                    const float *b1, const float *ln_w, const float *ln_b, float eps,
                    int32_t R, int32_t dil, int32_t n_query, int32_t T, void *stream);
int decaf_tcn_out(const float *r_in, const uint8_t *mask0, int64_t m_seq_stride,
                  const float *w_out, const float *b_out, int32_t R, void *cat, int32_t dtype,
                  int64_t ldc, int32_t col0, const decaf_levels_t *lv, int32_t n_query,
                  void *stream);
int decaf_refine_pool(void *cat, int32_t dtype, int64_t ldc, int32_t col0, int32_t R,
                      const uint8_t *hmask, const decaf_levels_t *lv, int32_t level,
                      int32_t n_query, void *stream);

/* The same refinement stage for the bf16 configuration, in two launches instead of 10 + (L - 1).
 * decaf_tcn_fused: tcn_in -> n_layers x tcn_layer (dilation 2^i) -> tcn_out for tiles of 256 level-0 steps with a
 *   recompute halo of 2^n_layers - 1 steps, state in shared memory, contractions on mma.sync (bf16 operands, fp32
 *   accumulation / residual / LayerNorm).  wblob (bf16): per layer Wd^T [R cout][3R = tap * R + cin] followed by
 *   W1^T [R cout][R cin]; vblob (fp32): per layer bd[R], b1[R], ln_w[R], ln_b[R]; w_out (bf16) [R cout][R cin].
 *   Writes cat[q, off0 + t, col0 : col0 + R] (bf16).  n_layers <= 8.  scratch: NULL, or (n_query, len[0], R) fp32 - with it
 *   a stack of >= 6 layers runs as two launches (layers [0, n - 3) and [n - 3, n)) that hand the fp32 state over through
 *   it: each launch's halo is the receptive field of its own layers only (31 + 224 steps instead of 255 per side of
 *   every 256-step tile, 2884 instead of 5124 row-layers per tile); results are identical.
 * decaf_refine_pyramid: every level l >= 1 of the masked max-pool pyramid (== decaf_refine_pool for l = 1..L-1) in
 *   one launch; level lengths must halve exactly, L <= 9.
 * replaces: the same reference code as decaf_tcn_in / _layer / _out / decaf_refine_pool above. */
int decaf_tcn_fused_supported(int32_t n_layers, int32_t n_levels);
int decaf_tcn_fused(const float *logits1, const uint8_t *hmask, const decaf_levels_t *lv,
                    const float *w_in /* (R, L) */, const float *b_in, const void *wblob,
                    const float *vblob, int32_t n_layers, const void *w_out, const float *b_out,
                    int32_t R, float eps, void *cat, int64_t ldc, int32_t col0, int32_t n_query,
                    float *scratch, void *stream);
int decaf_refine_pyramid_supported(int32_t n_levels);
int decaf_refine_pyramid(void *cat, int32_t dtype, int64_t ldc, int32_t col0, int32_t R,
                         const uint8_t *hmask, const decaf_levels_t *lv, int32_t n_query,
                         void *stream);

/* Plumbing: strided host -> device upload (cudaMemcpy2DAsync) of `height` rows of `width_bytes` each; src should be pinned
 * host memory for the copy to be asynchronous.  Used to upload a column window of a (C, t) feature matrix without a
 * contiguous staging copy (the reference uploads whole padded tensors, libs/worker_v2.py:997-1003). */
int decaf_upload_2d(void *dst, int64_t dst_pitch_bytes, const void *src, int64_t src_pitch_bytes,
                    int64_t width_bytes, int64_t height, void *stream);

/* Fused transformer FFN (one tcgen05 launch, CTA pairs; the hidden tensor stays on the SM):
 *   out[seq, t, :] = ((GELU(A[seq, t, :] W1^T + b1) W2^T + b2) * colscale + resid[seq, t, :]) * rowmask[seq, t]
 * A (n_seq, rows_per_seq, C) bf16 row-contiguous over sequences (a_seq_stride 0 or rows_per_seq), pitch lda; W1 (4C, C) and
 * W2 (C, 4C) bf16 row-major (the Conv1d weights with the kernel dimension squeezed); b1 (4C), b2 (C), colscale (C) fp32 or
 * NULL; resid / out_f32 fp32 and out_act bf16 (either output optional) with their own row pitch and per-sequence stride
 * (0 = rows_per_seq); rowmask u8 or NULL.  GELU is the tanh form of the bf16 configuration (see decaf_gemm).  C in {128, 256}.
 * Same arithmetic, in the same accumulation order, as decaf_gemm(act = GELU, out_act) followed by decaf_gemm(colscale,
 * resid, rowmask) — without the 2 x rows x 4C x 2 bytes round trip of the hidden tensor.
 * replaces: FFN.forward (libs/modeling/blocks.py:523-538) + the LayerScale / residual / mask of TransformerEncoder.forward
 * (:587-590) and TransformerDecoder's FFN half (:646-648). */
typedef struct {
    const void *A; int32_t dtype; int64_t lda; int64_t a_seq_stride;
    int32_t n_seq, rows_per_seq, C;
    const void *W1; const float *b1;
    const void *W2; const float *b2;
    const float *colscale;
    const float *resid; int64_t ldr; int64_t r_seq_stride;
    const uint8_t *rowmask; int64_t m_seq_stride;
    float *out_f32; int64_t ldo; int64_t o_seq_stride;
    void *out_act; int64_t ldo2; int64_t o2_seq_stride;
} decaf_ffn_t;
int decaf_ffn_supported(int32_t C, int32_t dtype);
int decaf_ffn(const decaf_ffn_t *p, void *stream);
/* debug (not on the product path): clock64 stamps of CTA 0 of the next decaf_ffn launches into buf[3][512]; NULL = off */
int decaf_debug_ffn_trace(unsigned long long *buf);

/* First step of the composed text encoder for a padded query batch: x (n_query, L1, C) = 0, kv_len[q] = len[q] + 1
 * (the background token, libs/modeling/text_net.py:176-180), tmask[q, r] = r < kv_len[q] (uint8).
 * replaces: the per-query tensor construction of TextTransformer.forward (text_net.py:158-181) for a batch. */
int decaf_text_init(float *x, int32_t n_query, int32_t L1, int32_t C, const int32_t *len, int32_t *kv_len,
                    uint8_t *tmask, void *stream);

/* text encoder glue: x[q, 0, :] <- bkgd;  x[q, 1+i, :] += PE_q[i, :] * (i < len[q]), where PE_q is the raw sinusoid
 * table pe (pe_rows = max_seq_len, C) when len[q] <= pe_rows and its linear (align_corners) interpolation to len[q]
 * rows otherwise — per QUERY, as the reference encodes every query alone (libs/worker_v2.py:945-955).
 * replaces: libs/modeling/text_net.py:163-183. */
int decaf_text_prep(float *x, int32_t n_query, int32_t L1 /* Lmax+1 */, int32_t C,
                    const float *bkgd, const float *pe /* or NULL */, int32_t pe_rows, const int32_t *len,
                    void *stream);

/* Whole text encoder of a batch of queries in one launch (a cluster of 8 CTAs per query) + the fusion
 * layers' key/value projections of the encoded text:
 *   x = [bkgd ; (embd_w tok + embd_b + pe) * mask];  n_layers x { x += ls_attn * proj(MHA_global(LN(x))) ;
 *   x += ls_ffn * proj2(GELU(fc(LN(x)))) } with x *= mask after each residual update;
 *   kv_out[f][0|1] = LN(x; lnkv_f) kv_w_f^T + kv_b_f  for every fusion layer f.
 * tokens (n_query, Lmax, Ctok) fp32 zero padded, lens (n_query); text_out (n_query, Lmax+1, Ct) or NULL;
 * kv_out (n_fusion, 2, n_query * (Lmax+1), C) fp32; kv_len_out (n_query) = lens + 1 or NULL.
 * Weights arrive as two fp32 blobs packed once by the caller (16-byte aligned).  With cpc = Ct / 8, H = 4 Ct
 * and W[rows] a row block of the reference's Conv1d weight viewed as (N, K):
 *   wblob = embd: for r in 0..7: (embd_w[r cpc : (r+1) cpc])^T                      (Ctok, cpc)
 *           per layer: for r: [Wq[r-block]; Wk[r-block]; Wv[r-block]]^T              (Ct, 3 cpc)
 *                      for r: (proj_w[r-block])^T (Ct, cpc);  for r: (fc_w[r H/8 : (r+1) H/8])^T (Ct, H/8);
 *                      for r: (proj2_w[r-block])^T (H, cpc)
 *           per fusion layer: for r: ([Wk; Wv][r C/4 : (r+1) C/4])^T                 (Ct, C/4)
 *   pblob = [embd_b | bkgd]  then per layer [ln_attn_w | ln_attn_b | q_b | k_b | v_b | proj_b | ls_attn |
 *           ln_ffn_w | ln_ffn_b | fc_b (4 Ct) | proj2_b | ls_ffn]  then per fusion layer [lnkv_w | lnkv_b | k_b | v_b]
 * i.e. every (stage, CTA) weight slice is contiguous and transposed ([k][column]), which is the layout the
 * kernel's register-tiled dot product reads from shared memory.
 * decaf_text_encoder_supported() tells whether the shape fits this kernel (Lmax + 1 <= 32, Ct <= 128, ...);
 * otherwise callers compose the same computation from decaf_gemm / decaf_layernorm / decaf_xattn.
 * replaces: TextTransformer.forward (libs/modeling/text_net.py:158-188) as called per query by
 * Evaluator._forward (libs/worker_v2.py:940-955), TransformerEncoder.forward with stride 0
 * (libs/modeling/blocks.py:578-591), and TransformerDecoder's ln_xattn_kv + key/value projections
 * (blocks.py:640-641, 348-350). */
typedef struct {
    const float *tokens; const int32_t *lens;
    int32_t n_query, Lmax, Ctok, Ct, n_heads, n_layers, n_fusion, C;
    const float *wblob, *pblob;
    const float *pe;                 /* (pe_rows, Ct) raw absolute-PE table of the words (interpolated per query when
                                        len > pe_rows, see decaf_text_prep), or NULL */
    int32_t pe_rows;
    float eps;
    float *text_out; float *kv_out; int32_t *kv_len_out;
} decaf_text_encoder_t;
int decaf_text_encoder_supported(int32_t Lmax, int32_t Ct, int32_t Ctok, int32_t n_heads, int32_t n_layers,
                                 int32_t C, int32_t n_fusion);
int64_t decaf_text_encoder_wblob_floats(int32_t Ct, int32_t Ctok, int32_t n_layers, int32_t C, int32_t n_fusion);
int64_t decaf_text_encoder_pblob_floats(int32_t Ct, int32_t n_layers, int32_t C, int32_t n_fusion);
int decaf_text_encoder(const decaf_text_encoder_t *p, void *stream);
/* debug only: clock64 stamps of cluster 0 after every stage of the following launches; NULL = off */
int decaf_debug_text_trace(unsigned long long *buf);
int decaf_debug_text_max_clusters(void);

/* ------------------------------------------------------------------ decode
 * Per query: score = sigmoid(logit) * mask (or the given score when from_logits == 0);
 * candidates = level rows with score > pre_nms_thresh; exact top-k (descending score, ties by
 * ascending level-major flat index); seg = [c - o0 * stride, c + o1 * stride]; keep
 * seg_len > seg_len_thresh.  Outputs per query: cand_segs (topk,2), cand_scores (topk),
 * cand_idx (topk, level-major flat point index without pad rows), cand_count.
 * topk <= 4096.  replaces: Evaluator._collect_segments (libs/worker_v2.py:1131-1187) and
 * PtGenerator's (coordinate, stride) columns (libs/modeling/model.py:703-743). */
int decaf_decode(const float *logits, const float *offsets, const uint8_t *hmask,
                 const decaf_levels_t *lv, int32_t n_query, int32_t from_logits,
                 float pre_nms_thresh, int32_t topk, float seg_len_thresh,
                 float *cand_segs, float *cand_scores, int32_t *cand_idx, int32_t *cand_count,
                 void *stream);

/* Eval-time loss statistics of one video: for every query, over its valid points, the focal classification loss (smoothed
 * targets, alpha) and the 1 - IoU regression loss of the positive points against the ground-truth segment `targets[q]`
 * (level-0 steps), plus the number of positives: out (n_query, 3) = {cls_sum, reg_sum, n_pos}.  reg_range (n_levels, 2):
 * PtGenerator's per-level regression range; center_sampling 1 = 'radius' (radius_mul x stride around the segment centre).
 * logits / offsets / hmask: the padded flat layout of decaf_decode.  Deterministic (one CTA per query).
 * replaces: Evaluator._calc_loss (libs/worker_v2.py:1029-1061) with annotate_points_per_video (:93-133), calc_focal_loss /
 * calc_iou_loss (:85-91) and sigmoid_focal_loss / ctr_giou_loss (libs/modeling/loss.py:6-108). */
int decaf_eval_loss(const float *logits, const float *offsets, const uint8_t *hmask, const decaf_levels_t *lv,
                    int32_t n_query, const float *targets, const float *reg_range, int32_t center_sampling,
                    float radius_mul, float smoothing, float alpha, float *out, void *stream);

/* Time-sharded decode (hour-long videos split along time across GPUs, decaf_b200/time_shard.py): the level
 * buffers describe a WINDOW of the timeline starting at level-0 step t0; only points whose level-0 position
 * lies in [own_lo, own_hi) (window coordinates; own_hi <= 0: all) become candidates, point coordinates are
 * global (t * stride + t0, exact in fp32) and cand_idx is the level-major flat index in the whole timeline of
 * T_global steps (T_global = 0: window-local index).  t0 / own_lo / own_hi: multiples of 2^(levels-1);
 * sum_l (T_global >> l) < 2^18 (checked: the merge key below carries the index in 18 bits). */
typedef struct { int32_t t0, own_lo, own_hi, T_global; } decaf_decode_window_t;
int decaf_decode_window(const float *logits, const float *offsets, const uint8_t *hmask,
                        const decaf_levels_t *lv, int32_t n_query, int32_t from_logits,
                        float pre_nms_thresh, int32_t topk, float seg_len_thresh,
                        const decaf_decode_window_t *win, float *cand_segs, float *cand_scores,
                        int32_t *cand_idx, int32_t *cand_count, void *stream);
/* Global top-k over the per-shard candidate lists after their all-gather: inputs (n_src, n_query, topk[,2]) +
 * count (n_src, n_query); output = the topk best by (score descending, global flat index ascending), i.e. the
 * order Evaluator._collect_segments' argsort gives on the whole timeline (libs/worker_v2.py:1169-1173) under
 * the stable tie rule of SURVEY.md A.6.  n_src * topk <= 16384, idx < 2^18. */
int decaf_merge_candidates(const float *segs, const float *scores, const int32_t *idx, const int32_t *count,
                           int32_t n_src, int32_t n_query, int32_t topk, float *out_segs, float *out_scores,
                           int32_t *out_idx, int32_t *out_count, void *stream);

/* ------------------------------------------------------------------ NMS
 * These two are the drop-in for the reference's only native FFI, the pybind module
 * nms_1d_cpu_vg (libs/nms/src/nms_cpu.cpp:184-194), batched over queries:
 *   decaf_softnms_1d  replaces softnms(segs, scores, dets, iou_thresh, sigma, min_score, method)
 *                     (nms_cpu.cpp:72-181): dets (n,3) rows written in selection order, inds =
 *                     original indices; returns per-query count in n_out.  max_iters > 0 stops
 *                     after that many outer steps (libs/nms/nms.py:54-59 only consumes the first
 *                     max_num_segs rows); max_iters <= 0 runs to completion.
 *   decaf_nms_1d      replaces nms(segs, scores, iou_thresh) (nms_cpu.cpp:20-70): kept indices
 *                     in descending score order (stable), at most max_keep when > 0.
 * Batch layout: query q reads n[q] candidates at segs + q * cand_stride * 2 etc.
 * workspace: decaf_nms_workspace_bytes(n_query, max_n) bytes. */
int64_t decaf_nms_workspace_bytes(int32_t n_query, int32_t max_n);
int decaf_softnms_1d(const float *segs, const float *scores, const int32_t *n, int32_t n_query,
                     int32_t cand_stride, float *dets, int32_t *inds, int32_t *n_out,
                     float iou_thresh, float sigma, float min_score, int32_t method,
                     int32_t max_iters, void *workspace, void *stream);
int decaf_nms_1d(const float *segs, const float *scores, const int32_t *n, int32_t n_query,
                 int32_t cand_stride, int32_t *keep, int32_t *n_out, float iou_thresh,
                 float min_score, int32_t max_keep, void *workspace, void *stream);

/* Fused post-processing of one batch of queries = batched_nms + the seconds conversion:
 * (soft-)NMS -> first max_num_segs -> segment voting over ALL input candidates -> stable
 * descending sort -> seg = clamp(((seg * vid_stride) * clip_stride + 0.5 * clip_size) / fps,
 * 0, duration).  mode: 0 none, 1 hard nms, 2 soft-nms (gaussian).
 * out_segs (n_query, max_num_segs, 2), out_scores (n_query, max_num_segs), out_count.
 * replaces: batched_nms / NMSop / SoftNMSop / segment_voting (libs/nms/nms.py:6-148) and
 * Evaluator._generate_proposals' conversion (libs/worker_v2.py:1113-1122). */
typedef struct {
    int32_t mode; float iou_thresh, sigma, min_score; int32_t max_num_segs; float voting_thresh;
    int32_t to_seconds; float vid_stride, clip_stride, half_clip_size, fps, duration;
    /* optional DEVICE pointer to {vid_stride, clip_stride, half_clip_size, fps, duration}: when set it
     * overrides the five by-value fields, so a captured CUDA graph can be replayed for another video */
    const float *video_meta;
} decaf_nms_params_t;
int decaf_batched_nms(const float *segs, const float *scores, const int32_t *n, int32_t n_query,
                      int32_t cand_stride, const decaf_nms_params_t *prm, float *out_segs,
                      float *out_scores, int32_t *out_count, void *workspace, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DECAF_B200_H */
