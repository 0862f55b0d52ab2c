/*
 * videoblip_b200 — C ABI of the sm_100a kernels behind the VideoBLIP hot path.
 *
 * The reference (yukw777/EILEV) has no native layer: its hot path is
 * eilev/model/v2.py calling HuggingFace modules, i.e. torch/ATen/cuBLAS/cuDNN
 * library kernels.  Each entry point below replaces the library call sites the
 * reference reaches (cited per function as  <file>:<line>, paths relative to the
 * reference checkout, "HF:" = transformers/models/...).
 *
 * Conventions
 *   - plain C types only: device pointers, int64 sizes/strides (in ELEMENTS),
 *     floats, and the CUDA stream as an opaque void* (cudaStream_t).
 *   - every function returns 0 on success, non-zero on failure; the message of
 *     the last failure on the calling thread is vb_last_error().
 *   - no allocation, no device synchronisation, no global mutable state;
 *     all work is enqueued on `stream`.  Pointers must be device pointers.
 *   - bf16 tensors are raw uint16 storage (`__nv_bfloat16`).
 */
#ifndef VIDEOBLIP_B200_H
#define VIDEOBLIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB_ABI_VERSION 6

/* dtype tags */
#define VB_BF16 0
#define VB_F32 1
#define VB_F16 2

/* GEMM epilogue activation */
#define VB_EPI_NONE 0
#define VB_EPI_GELU 1 /* exact erf GELU  (HF:blip_2/modeling_blip_2.py:365-369, :686-689) */
#define VB_EPI_RELU 2 /* OPT FFN        (HF:opt/modeling_opt.py:238-239) */
/* Backward of an activation fused into the dgrad GEMM that produces its output gradient:
 * C = (alpha * A.B^T) * act'(S), S = `residual` (the saved forward tensor, bf16, same shape as C; for GELU the
 * pre-activation, for ReLU the activation output).  Replaces vb_act_bwd after the GEMM. */
#define VB_EPI_GELU_BWD 3
#define VB_EPI_RELU_BWD 4

/* GEMM backend selector (vb_gemm_args.backend) */
#define VB_GEMM_AUTO 0    /* tcgen05 when the shape allows, else generic */
#define VB_GEMM_TCGEN05 1 /* fail if the shape is not eligible */
#define VB_GEMM_GENERIC 2 /* CUDA-core kernel, any shape/alignment */

int vb_abi_version(void);
const char* vb_last_error(void);
/* Compute capability major*10+minor of the current device, or <0 on error. */
int vb_device_arch(void);

/* ------------------------------------------------------------------------
 * C[m, n] = act(alpha_n * (sum_k A[m,k] * B[n,k] + bias[n])) + residual[m',n] + beta*C[m,n]
 *
 * A: (M,K) bf16 row-major (lda), B: (N,K) bf16 row-major (ldb) — i.e. exactly an
 * nn.Linear weight — C: (M,N) bf16 or f32 (ldc).  bias: f32 (N) or NULL.
 * residual: bf16 (ldr) or NULL.  alpha is applied to columns < alpha_cols only
 * (alpha_cols <= 0: all columns) — OPT scales q but not k,v of a fused QKV.
 * row_group P > 0 selects the patch-embedding store: GEMM row r is stored at
 * row r + r/P + 1 (one CLS slot in front of every P patches) and the residual
 * row is 1 + r%P (the position table).
 *
 * Replaces every nn.Linear / Conv2d(k=s=patch) on the path:
 *   HF:blip_2/modeling_blip_2.py:246-254 (patch embedding as GEMM), :326-353 (qkv,
 *   projection), :365-369 (fc1+GELU, fc2), :592-628 (Q-Former q/k/v), :644-648,
 *   :686-689, :700-704; eilev/model/v2.py:201-203 (language_projection);
 *   HF:opt/modeling_opt.py:151-181 (q,k,v,out_proj), :238-241 (fc1+ReLU, fc2), :512 (lm_head);
 *   and their autograd backward (dgrad / wgrad with pre-transposed operands).
 * ---------------------------------------------------------------------- */
typedef struct vb_gemm_args {
  const void* a;
  const void* b;
  void* c;
  const float* bias;
  const void* residual;
  int64_t m, n, k;
  int64_t lda, ldb, ldc, ldr;
  float alpha;
  float beta;
  int64_t alpha_cols;
  int64_t row_group;
  int32_t epilogue;  /* VB_EPI_* */
  int32_t out_dtype; /* VB_BF16 or VB_F32 */
  int32_t backend;   /* VB_GEMM_* */
  int32_t reserved;
  /* Dropout on act(...) BEFORE the residual is added (HF: dropout(dense(x)) + residual):
   * kept elements are scaled by 1/(1-p).  The mask is a counter hash of
   * (*dropout_seed + dropout_salt, row*n + col) — vb_dropout with the same seed/salt
   * regenerates it in the backward pass.  dropout_p == 0 or dropout_seed == NULL: off. */
  float dropout_p;
  /* operand_layout: 0 = a is A (M, K) and b is B (N, K), both row-major (nn.Linear forward / dgrad).
   * 1 = a is A^T stored (K, M) row-major with row stride lda, b is B^T stored (K, N) with row stride ldb: C = A B^T
   * = a^T b, the weight-gradient product dW (N_out, N_in) = dY^T X taken straight from the (tokens, features)
   * activations (M = N_out, N = N_in, K = tokens); tcgen05 path only (M, N, lda, ldb multiples of 8), no
   * LayerNorm fold / row statistics. */
  int32_t reserved2;
  const uint64_t* dropout_seed; /* device pointer */
  uint64_t dropout_salt;
  /* LayerNorm folded into the GEMM (ABI 5; tcgen05 path only).  For y = LN(x) W^T + b with
   * LN(x) = (x - mean) rstd * gamma + beta, the caller passes A = x (un-normalised), B = W * gamma
   * (column-scaled), bias = b + W beta, and
   *   ln_stats  (M, 2) f64: [sum_k x[m,k], sum_k x[m,k]^2] of every A row (over all K columns),
   *   ln_colsum (N) f32:    sum_k B[n,k],
   * and the epilogue forms  rstd_m * acc - rstd_m * mean_m * ln_colsum[n] + bias[n]  before alpha /
   * activation — the normalised activations are never written to memory
   * (HF:blip_2/modeling_blip_2.py:388-402: layer_norm1 -> qkv, layer_norm2 -> fc1).  NULL: off. */
  const double* ln_stats;
  const float* ln_colsum;
  float ln_eps;
  int32_t reserved3;
  /* Row statistics of the OUTPUT for the next folded LayerNorm: when non-NULL, the epilogue adds
   * [sum_n C[m,n], sum_n C[m,n]^2] (of the bf16 values it stores) to stats_out[2 m'], [2 m' + 1]
   * with f64 atomics (m' = the stored row).  f64 because the partial sums arrive in no fixed order: adding a few
   * dozen f32 partials in f64 is exact, so the statistics — and with them the whole step — are reproducible bit
   * for bit from run to run (ABI 6; they were f32 in ABI 5).  The buffer must be zeroed by the caller. */
  double* stats_out;
  /* When non-NULL: rows [0, M) of this (M, 2) f64 buffer are set to zero by the epilogue (by the tiles of
   * the first column block).  Lets the two statistics buffers of a transformer layer re-arm each other
   * without memset launches: the GEMM that fills one buffer clears the other, which its predecessor in
   * the stream has already consumed. */
  double* stats_zero;
} vb_gemm_args;

int vb_gemm(const vb_gemm_args* args, void* stream);
/* 1 if vb_gemm would take the tcgen05 path for these args. */
int vb_gemm_uses_tcgen05(const vb_gemm_args* args);

/* ------------------------------------------------------------------------
 * y = LayerNorm(x (+ residual)) * gamma + beta   (rows x cols, fp32 statistics)
 * x, residual, y: bf16 (row strides ldx/ldr/ldy); gamma/beta: f32.
 * mean/rstd (f32, rows) are written when non-NULL (saved for backward).
 * HF:blip_2/modeling_blip_2.py:388-402, :522-526, :644-648, :700-704, :984-986;
 * HF:opt/modeling_opt.py:206-207, :232-233, :370-371.
 * ---------------------------------------------------------------------- */
int vb_layernorm(const void* x, const void* residual, const float* gamma, const float* beta,
                 void* y, float* mean, float* rstd, int64_t rows, int64_t cols, int64_t ldx,
                 int64_t ldr, int64_t ldy, float eps, void* stream);

/* stats[r] = [sum_c x[r,c], sum_c x[r,c]^2] (f64, rows x 2) of a bf16 matrix: the statistics a
 * LayerNorm folded into the consuming GEMM needs (vb_gemm_args.ln_stats) when x was not produced
 * by a GEMM epilogue (vb_gemm_args.stats_out).  HF:blip_2/modeling_blip_2.py:388-389 (layer_norm1
 * of the first encoder layer, fed by the embeddings). */
int vb_row_stats(const void* x, double* stats, int64_t rows, int64_t cols, int64_t ldx, void* stream);

/* dx (+ optional dgamma/dbeta accumulation, f32 atomics) of the LayerNorm above.
 * xin is the tensor that was normalised (x + residual), bf16.  dx is bf16; when
 * dx_add != NULL it is added to the result (gradient joining a residual branch). */
int vb_layernorm_bwd(const void* dy, const void* xin, const float* gamma, const float* mean,
                     const float* rstd, const void* dx_add, void* dx, float* dgamma, float* dbeta,
                     int64_t rows, int64_t cols, float eps_unused, void* stream);
/* The same dx plus a second output dx_drop = vb_dropout(dx, dropout_p, seed, salt) written in the same pass: the
 * gradient that enters the dgrad GEMM of the linear layer in front of a residual dropout (HF:opt/modeling_opt.py
 * :230, :246 — `dropout(hidden) + residual`; in the backward the residual stream's gradient feeds both the branch,
 * through the mask, and the previous residual, unmasked).  Saves the separate dropout pass over the gradient. */
int vb_layernorm_bwd_dropout(const void* dy, const void* xin, const float* gamma, const float* mean,
                             const float* rstd, const void* dx_add, void* dx, void* dx_drop, float dropout_p,
                             const uint64_t* dropout_seed, uint64_t dropout_salt, int64_t rows, int64_t cols,
                             void* stream);

/* ------------------------------------------------------------------------
 * Fused softmax attention, FlashAttention-style (online softmax, fp32 accumulate).
 * Element (b, s, h, d) of q lives at q + b*q_bs + s*q_rs + h*D + d  (elements); same
 * for k, v, o with their own strides, so the operands can alias a fused QKV buffer.
 * scale multiplies q.k^T.  causal: key j visible to query i iff j <= i + (Skv - Sq).
 * key_mask: optional uint8 (B, Skv), 0 = masked key.  lse: optional f32 (B,H,Sq).
 * HF:blip_2/modeling_blip_2.py:326-353 (ViT, S=257 d=88), :592-628 (Q-Former self /
 * cross, d=64), HF:opt/modeling_opt.py:135-181 (causal, d=80).
 * ---------------------------------------------------------------------- */
typedef struct vb_attn_args {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  float* lse;
  const uint8_t* key_mask;
  int64_t batch, heads, sq, skv, d;
  int64_t q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;
  float scale;
  int32_t causal;
  /* Dropout on the attention probabilities (HF:blip_2/modeling_blip_2.py:621-624): mask =
   * hash(*dropout_seed + dropout_salt, ((b*H + h)*Sq + i)*Skv + j); off when p == 0 / NULL. */
  float dropout_p;
  int32_t reserved;
  const uint64_t* dropout_seed;
  uint64_t dropout_salt;
  /* Additive relative-position bias (T5: HF:t5/modeling_t5.py compute_bias): f32 table
   * rel_bias[h * rel_bias_stride + (j - i) + (Sq - 1)] added to scale * q_i.k_j; NULL: none.
   * The table has Sq + Skv - 1 entries per head; no gradient is produced for it. */
  const float* rel_bias;
  int64_t rel_bias_stride;
} vb_attn_args;
int vb_attention_fwd(const vb_attn_args* args, void* stream);
/* The attention MAPS probs[b, h, i, j] = softmax_j(scale * q_i . k_j + masks), f32 or bf16, (batch, heads, sq, skv)
 * contiguous: the `attentions` of output_attentions=True (eilev/model/v2.py:87-95).  Only q, k, key_mask, the
 * shapes / strides, scale and causal of `args` are read.  A diagnostic output computed by a plain CUDA-core
 * kernel; the fused kernels of vb_attention_fwd never materialise it. */
int vb_attention_probs(const vb_attn_args* args, void* probs, int32_t probs_dtype, void* stream);
/* Which kernel vb_attention_fwd takes: 1 = the single-pass tcgen05 / TMEM kernel (non-causal, unmasked, no lse,
 * 64 <= S <= 272, d <= 128: the ViT shape class); 2 = the tcgen05 flash kernel (any Sq / Skv, causal, key mask,
 * dropout, relative bias, lse; d % 16 == 0, 16-byte aligned operands, batches stored back to back);
 * 0 = the mma.sync flash kernel. */
int vb_attention_uses_tcgen05(const vb_attn_args* args);

/* Backward of the above: dq, dk, dv (bf16, same addressing as q,k,v via the dq_, dk_, dv_
 * strides).  delta: f32 workspace (B,H,Sq).  dq_acc: f32 workspace (B,Sq,H*D) zeroed by
 * the call.  */
typedef struct vb_attn_bwd_args {
  vb_attn_args fwd; /* q,k,v,o,lse,key_mask and shapes as in the forward */
  const void* d_o;  /* same addressing as o */
  void* dq;
  void* dk;
  void* dv;
  int64_t dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs;
  float* delta;
  float* dq_acc;
  float dq_scale; /* extra factor on dq (OPT: q was pre-scaled by the projection epilogue) */
  int32_t reserved;
} vb_attn_bwd_args;
int vb_attention_bwd(const vb_attn_bwd_args* args, void* stream);
/* 1 if vb_attention_bwd takes the tcgen05 / TMEM kernels (d % 16 == 0, 16-byte aligned operands, batches stored
 * back to back), 0 for the mma.sync kernel. */
int vb_attention_bwd_uses_tcgen05(const vb_attn_bwd_args* args);

/* ------------------------------------------------------------------------
 * Patch gather for the ViT patch embedding (Conv2d k=s=P as a GEMM):
 * pixels (NV, C, T, H, W) of dtype px_dtype -> out (NV*T*gh*gw, kpad) bf16 where row
 * ((v*T+t)*gh+gy)*gw+gx holds the C*P*P patch in (c, py, px) order (the Conv2d
 * weight's flattening), zero padded to kpad.  Fuses the (N,C,T,H,W)->(N*T,C,H,W)
 * permute of eilev/model/v2.py:57 and the cast.  HF:blip_2/modeling_blip_2.py:246-248.
 * ---------------------------------------------------------------------- */
int vb_patch_gather(const void* pixels, int32_t px_dtype, void* out, int64_t nv, int64_t c,
                    int64_t t, int64_t h, int64_t w, int64_t patch, int64_t kpad, void* stream);
/* The same gather straight from decoded uint8 frames (NV, C, T, H, W), C <= 4, with the
 * image-processor arithmetic fused:  v = (f32(f64(u8) * rescale) - mean[c]) / stdv[c]  (the
 * float64 rescale then float32 normalize of the reference's pinned transformers 4.33.1)
 * (BlipImageProcessor rescale + normalize, HF:models/blip/image_processing_blip.py, behind
 * eilev/model/utils.py:5-26 `process`; the /255 + Normalize of scripts/general/train_v2.py:143-167).
 * mean / stdv are HOST arrays of c floats, read during the call.  SURVEY §8(f) rank 3. */
int vb_patch_gather_u8(const void* pixels_u8, void* out, int64_t nv, int64_t c, int64_t t, int64_t h,
                       int64_t w, int64_t patch, int64_t kpad, double rescale, const float* mean,
                       const float* stdv, void* stream);
/* ------------------------------------------------------------------------
 * Antialiased bicubic resize of uint8 planes, bit-exact with PIL.Image.resize(size, BICUBIC) —
 * the resize BlipImageProcessor performs for eilev/model/utils.py:5-26 `process` (transformers
 * 4.33.1 image_transforms.resize -> Pillow src/libImaging/Resample.c).  SURVEY §8(f) rank 3.
 *   vb_resize_bicubic_ksize  : taps per output sample of one axis (host only, no device work)
 *   vb_resize_bicubic_coeffs : HOST arrays bounds[2*out] = {first tap, tap count} and
 *                              kk[out*ksize] = 22-bit fixed-point weights (precompute_coeffs +
 *                              normalize_coeffs_8bpc, double precision, Pillow's expression order)
 *   vb_resize_u8_pass        : one pass along one axis on the device (bounds / kk are DEVICE copies):
 *       out[p, l, o] = clip8((2^21 + sum_k in[p, l, bounds[2o] + k] * kk[o*ksize + k]) >> 22)
 *     addressed through element strides, so the same kernel runs Pillow's horizontal pass (over
 *     the source rows the vertical pass needs) and then its vertical pass; lines_fastest selects
 *     which index is consecutive across a warp.
 * ---------------------------------------------------------------------- */
int vb_resize_bicubic_ksize(int64_t in_size, int64_t out_size);
int vb_resize_bicubic_coeffs(int64_t in_size, int64_t out_size, int32_t* bounds, int32_t* kk,
                             int64_t kk_capacity);
int vb_resize_u8_pass(const void* in, void* out, const int32_t* bounds, const int32_t* kk, int64_t planes,
                      int64_t lines, int64_t out_len, int64_t ksize, int64_t in_plane_stride,
                      int64_t in_line_stride, int64_t in_elem_stride, int64_t out_plane_stride,
                      int64_t out_line_stride, int64_t out_elem_stride, int32_t lines_fastest, void* stream);

/* Training-time frame transform of one clip on the device (scripts/general/train_v2.py:143-167: pytorchvideo
 * ConvertUint8ToFloat -> Normalize -> RandomResizedCrop(bicubic) -> RandomHorizontalFlip): reads the crop box
 * [crop_top, crop_top + crop_h) x [crop_left, crop_left + crop_w) of every plane of a decoded uint8 clip
 * (c, t, h, w), resizes it to (out_h, out_w) as torch.nn.functional.interpolate(mode="bicubic",
 * align_corners=False) does (cubic convolution, A = -0.75, taps clamped to the crop), mirrors the columns when
 * flip != 0 and writes (x * rescale - mean[c]) / std[c] as f32 or bf16 (c, t, out_h, out_w), contiguous.
 * The crop box and the flip are drawn on the host (integer bookkeeping). */
int vb_crop_resize_normalize_u8(const void* frames_u8, int64_t c, int64_t t, int64_t h, int64_t w, int64_t crop_top,
                                int64_t crop_left, int64_t crop_h, int64_t crop_w, int32_t flip, void* out,
                                int32_t out_dtype, int64_t out_h, int64_t out_w, double rescale, const float* mean,
                                const float* stdv, void* stream);

/* hidden[f, 0, :] = cls + pos[0]  for every frame f (HF:...:249-254). bf16. */
int vb_cls_rows(const void* cls, const void* pos, void* hidden, int64_t frames, int64_t tokens,
                int64_t dim, void* stream);

/* ------------------------------------------------------------------------
 * LM input assembly (eilev/model/v2.py:205-214 + HF:opt/modeling_opt.py:350-368):
 *   e[b,l] = video_mask[b,l] ? video_features[rank of (b,l) among masked slots]
 *                            : embed_tokens[input_ids[b,l]]
 *   pos[b,l] = cumsum(attention_mask)[b,l]*attention_mask[b,l] - 1 + pos_offset
 *   inputs_embeds = e ; hidden = e + pos_table[pos]   (pos_table NULL: hidden = e)
 * ids/masks are int64; embeddings bf16.  slot_index (int32, B*L) receives the
 * video-feature row of every position (-1 for text) for the backward gather, pos_ids
 * (int32, B*L) the position-table rows.  status (int32[2]): status[0] = 1 if the mask
 * count differs from n_features (the reference's index_put shape error), status[1] =
 * the mask count.  inputs_embeds / hidden may be NULL.
 * ---------------------------------------------------------------------- */
int vb_embed_splice(const int64_t* input_ids, const int64_t* attention_mask,
                    const int64_t* video_mask, const void* embed_tokens,
                    const void* video_features, const void* pos_table, int64_t pos_offset,
                    void* inputs_embeds, void* hidden, int32_t* slot_index, int32_t* pos_ids,
                    int32_t* status, int64_t batch, int64_t seq, int64_t dim, int64_t vocab, int64_t n_features,
                    void* stream);
/* d_video_features[slot] = d_inputs_embeds[b,l] for masked slots (bf16). */
int vb_splice_bwd(const void* d_embeds, const int32_t* slot_index, void* d_features,
                  int64_t positions, int64_t dim, int64_t n_features, void* stream);

/* ------------------------------------------------------------------------
 * Shifted causal-LM cross entropy (HF:loss/loss_utils.py:28-67): position l predicts
 * labels[b, l+1]; ignore_index -100; mean over valid targets; fp32 math.
 * logits (B, L, V) bf16|f32 (row stride ldl).  Outputs: loss (f32 scalar),
 * row_lse (f32, B*L), n_valid (int32 scalar).
 * ---------------------------------------------------------------------- */
int vb_cross_entropy(const void* logits, int32_t logits_dtype, const int64_t* labels, float* loss,
                     float* row_lse, int32_t* n_valid, int64_t batch, int64_t seq, int64_t vocab,
                     int64_t ldl, int32_t shift, void* stream);
/* dlogits (bf16, B*L x V, ldd) = grad_scale * (softmax - onehot) / n_valid on valid rows,
 * 0 elsewhere. grad_scale: f32 device scalar (upstream d loss) or NULL (=1). */
int vb_cross_entropy_bwd(const void* logits, int32_t logits_dtype, const int64_t* labels,
                         const float* row_lse, const int32_t* n_valid, const float* grad_scale,
                         void* dlogits, int64_t batch, int64_t seq, int64_t vocab, int64_t ldl,
                         int64_t ldd, int32_t shift, void* stream);

/* ------------------------------------------------------------------------ flan-T5 LM
 * (eilev/model/v2.py:228-238 -> HF:t5/modeling_t5.py)
 * y = gamma * x * rsqrt(mean(x^2) + eps)  — T5LayerNorm: no mean subtraction, no bias; fp32
 * statistics, bf16 in/out; rstd (f32, rows) is written when non-NULL (saved for backward). */
int vb_rmsnorm(const void* x, const float* gamma, void* y, float* rstd, int64_t rows, int64_t cols,
               int64_t ldx, int64_t ldy, float eps, void* stream);
/* dx = rstd * (g - x * rstd^2 * mean(g * x)), g = gamma * dy  (+ dx_add when non-NULL); contiguous bf16. */
int vb_rmsnorm_bwd(const void* dy, const void* x, const float* gamma, const float* rstd, const void* dx_add,
                   void* dx, int64_t rows, int64_t cols, void* stream);
/* T5DenseGatedActDense middle: h01 = [wi_0 x | wi_1 x] (rows, 2*dff) bf16 ->
 * out = gelu_new(h01[:, :dff]) * h01[:, dff:]  (tanh approximation), and its backward
 * d_h01 = [d_out * h1 * gelu_new'(h0) | d_out * gelu_new(h0)]. */
int vb_gated_gelu(const void* h01, void* out, int64_t rows, int64_t dff, void* stream);
int vb_gated_gelu_bwd(const void* d_out, const void* h01, void* d_h01, int64_t rows, int64_t dff, void* stream);
/* out[i, :] = table[ids[i], :] (bf16): decoder_input_ids -> shared embedding (ids clamped to the table). */
int vb_embedding(const int64_t* ids, const void* table, void* out, int64_t n, int64_t dim, int64_t vocab,
                 void* stream);

/* ------------------------------------------------------------------------ elementwise */
/* out(cols, rows) = in(rows, cols)^T, bf16 (operand staging for dgrad / wgrad GEMMs). */
int vb_transpose(const void* in, void* out, int64_t rows, int64_t cols, int64_t ld_in,
                 int64_t ld_out, void* stream);
/* dtype conversion src -> dst (VB_BF16 / VB_F32 / VB_F16), n elements. */
int vb_convert(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n,
               void* stream);
/* dx = dy * act'(.)  bf16.  GELU: `saved` is the pre-activation; RELU: `saved` is the output. */
int vb_act_bwd(const void* dy, const void* saved, void* dx, int32_t epilogue, int64_t n,
               void* stream);
/* out(f32, cols) (+)= column sums of x (bf16, rows x cols): bias gradients. */
int vb_colsum(const void* x, float* out, int64_t rows, int64_t cols, int64_t ldx, int32_t accumulate,
              void* stream);
/* y[r, c] = keep(*seed + salt, r*cols + c) ? x[r, c] / (1 - p) : 0   (bf16, row strides ldx/ldy).
 * The same counter hash as the GEMM-epilogue dropout: used for the backward mask and for
 * stand-alone dropouts (HF:blip_2/modeling_blip_2.py:985, HF:opt/modeling_opt.py:219,243). */
int vb_dropout(const void* x, void* y, int64_t rows, int64_t cols, int64_t ldx, int64_t ldy,
               float p, const uint64_t* seed, uint64_t salt, void* stream);
/* y = a + b (bf16) */
int vb_add(const void* a, const void* b, void* y, int64_t n, void* stream);

/* Fused AdamW over a flat f32 parameter/gradient/moment buffer (torch.optim.AdamW
 * semantics, scripts/general/train_v2.py:99-101).  grad_scale (device f32 scalar or NULL)
 * multiplies the gradient first: fold 1/world, 1/accum and the clip factor into it. */
int vb_adamw(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
             float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
             const float* grad_scale, void* stream);
/* out[0] += sum(x^2) over n f32 elements (global grad norm for clipping). */
int vb_sumsq(const float* x, int64_t n, float* out, void* stream);

/* ------------------------------------------------------------------------ classify
 * eilev/model/v2.py:326-501 scores every class continuation against ONE cached prompt.  The
 * reference replicates the prompt's K/V num_classes times (:457-460); here each class token
 * attends to the shared prompt K/V once (vb_attention_fwd, non-causal, key_mask, lse) and to
 * its own continuation (causal, lse), and the two partial softmaxes are merged:
 *   out = (e^{lse1} o1 + e^{lse2} o2) / (e^{lse1} + e^{lse2}).
 * o1, o2, out: (rows, heads*d) bf16 contiguous; lse_i: f32 laid out (rows / s_i, heads, s_i)
 * exactly as vb_attention_fwd writes it for a batch of rows/s_i sequences of s_i queries. */
int vb_attention_merge(const void* o1, const float* lse1, int64_t s1, const void* o2, const float* lse2,
                       int64_t s2, void* out, int64_t rows, int64_t heads, int64_t d, void* stream);
/* out[i] = log_softmax(logits[row_index ? row_index[i] : i, :])[targets[i]], or 0 when
 * targets[i] is outside [0, vocab) (ignore_index): nn.CrossEntropyLoss(reduction="none")
 * negated, eilev/model/v2.py:486-494.  logits bf16|f32 with row stride ldl. */
int vb_token_logprob(const void* logits, int32_t logits_dtype, const int64_t* row_index,
                     const int64_t* targets, float* out, int64_t n, int64_t vocab, int64_t ldl,
                     void* stream);

/* ------------------------------------------------------------------------ decode */
/* y[m, n] = act(alpha_n * (LN?(x)[m,:] . W[n,:] + bias[n])) (+ residual) for small m (<= 16):
 * weight-streaming kernel for token-by-token generation (HBM bound).  bf16 in/out,
 * out_dtype selects bf16|f32.  When ln_gamma/ln_beta (f32, k) are given, x is
 * LayerNorm-ed (eps ln_eps, rounded to bf16 like vb_layernorm) while it is staged in
 * shared memory — the pre-LN of an OPT block costs no launch.  ln_gamma without ln_beta selects
 * RMSNorm (T5LayerNorm: no mean subtraction, no bias).
 * HF:opt/modeling_opt.py:135-253 at tgt_len == 1. */
int vb_gemv(const void* x, const void* w, const float* bias, const void* residual, void* y,
            int64_t m, int64_t n, int64_t k, int64_t ldx, int64_t ldw, int64_t ldy, int64_t ldr,
            float alpha, int64_t alpha_cols, int32_t epilogue, int32_t out_dtype,
            const float* ln_gamma, const float* ln_beta, float ln_eps, void* stream);

/* Prologue of one generation step: x[b,:] = embed[tokens[b],:] + pos_table[n_valid[b] +
 * pos_offset,:] (bf16), then n_valid[b] += 1 and ctx_len[b] += 1 (device-resident per-sequence
 * counters, so the step is CUDA-graph replayable).  tokens: (B) int64.
 * HF:opt/modeling_opt.py:45-70 (learned positions, offset 2), :350-354, :363. */
int vb_decode_embed(const int64_t* tokens, const void* embed, const void* pos_table, int32_t* n_valid,
                    int32_t* ctx_len, void* x, int64_t batch, int64_t dim, int64_t vocab,
                    int64_t pos_rows, int64_t pos_offset, void* stream);

/* ---- one generated token as ONE launch ------------------------------------------------
 * A decode step is a short program of vb_decode_op records executed by a persistent
 * cooperative kernel (one CTA per SM) with grid barriers between the ops; while a CTA waits
 * at a barrier the first weight loads of the next projection are already in flight, so the
 * HBM stream does not drain at every op boundary (~160 per token otherwise).
 * HF:opt/modeling_opt.py:321-396 at tgt_len == 1, all layers.
 *
 * Slot meaning per op type (unused slots must be zero):
 *  VB_OP_GEMV   y = act(alpha_n (LN?(x) . W^T + bias)) (+ residual), as vb_gemv with m rows
 *    ptr: 0 W (n,k) bf16 | 1 bias f32 | 2 residual bf16 | 3 x bf16 | 4 y | 5 ln_gamma | 6 ln_beta
 *    i64: 0 n | 1 k | 2 ldw | 3 ldx | 4 ldy | 5 ldr | 6 alpha_cols     f32: 0 alpha | 1 ln_eps
 *    i32: 0 epilogue (VB_EPI_*) | 1 output is f32
 *  VB_OP_ATTN   as vb_paged_decode_attention with batch = m
 *    ptr: 0 qkv | 1 k_cache | 2 v_cache | 3 page_table | 4 ctx_len | 5 first_valid | 6 out
 *         | 7 workspace | 8 counters
 *    i32: 0 heads | 1 d | 2 page_size | 3 max_pages | 4 splits | 5 ceil(page_size*max_pages/splits)
 *    f32: 0 scale
 *  VB_OP_EMBED  as vb_decode_embed with batch = m
 *    ptr: 0 tokens | 1 embed | 2 pos_table | 3 n_valid | 4 ctx_len | 5 x
 *    i64: 0 dim | 1 vocab | 2 pos_rows | 3 pos_offset */
#define VB_OP_GEMV 1
#define VB_OP_ATTN 2
#define VB_OP_EMBED 3
typedef struct vb_decode_op {
  int32_t type;
  int32_t i32[7];
  const void* ptr[10];
  int64_t i64[8];
  float f32[4];
} vb_decode_op; /* 192 bytes */
/* Runs ops[0..n_ops) for m <= 8 sequences.  ops_host / ops_dev: the same records in host
 * and in device memory (the host copy sizes the launch; the kernel reads the device copy).
 * workspace: VB_DECODE_STEP_WS_BYTES of device memory owned by this program (barrier counter,
 * hand-over flags and partial-tile slots of projections split across CTAs; its header is
 * zeroed by the call).  Every GEMV op needs k % 64 == 0 and 16-byte aligned rows; returns an
 * error otherwise (callers fall back to the per-op entry points above).  Stream-ordered,
 * CUDA-graph capturable. */
#define VB_DECODE_STEP_WS_BYTES (4096 + 1008 * 512)
int vb_decode_step(const vb_decode_op* ops_host, const vb_decode_op* ops_dev, int32_t n_ops, int32_t m,
                   uint32_t* workspace, void* stream);

/* Tooling: when buffer != NULL every later vb_decode_step writes 6 %globaltimer stamps (ns)
 * per (op, CTA) into it — uint64 [n_ops][num_SMs][6] = op start, x staged, weights consumed,
 * next op primed, finalised, barrier passed.  NULL turns it off (default). */
int vb_debug_decode_trace(void* buffer);

/* Append new K/V rows into a paged cache and run one-query-per-sequence attention over
 * it.  Cache pages: (n_pages, page_size, H*D) bf16 for K and for V; page_table (B,
 * max_pages) int32; ctx_len (B) int32 = number of cached tokens INCLUDING the new one;
 * first_valid (B) int32 = index of the first non-padding token (left padding).
 * qkv: (B, 3*H*D) bf16 (q pre-scaled).  out: (B, H*D) bf16.
 * The context is processed in `splits` independent CTAs per (sequence, head)
 * (flash-decoding); workspace: f32 (B*H*splits*(D+2)); counters: int32 (B*H), zero on
 * entry and left zero on exit.
 * rel_bias (f32 or NULL): T5 decoder self-attention — rel_bias[h * rel_stride + rel_center +
 * (l - (ctx - 1))] is added to the score of cached token l (HF:t5/modeling_t5.py compute_bias
 * with the newest token as the only query).
 * HF:opt/modeling_opt.py:159-161 (DynamicCache.update) + :163-176. */
int vb_paged_decode_attention(const void* qkv, void* k_cache, void* v_cache,
                              const int32_t* page_table, const int32_t* ctx_len,
                              const int32_t* first_valid, void* out, float* workspace,
                              int32_t* counters, int64_t splits, int64_t batch, int64_t heads,
                              int64_t d, int64_t page_size, int64_t max_pages, float scale,
                              const float* rel_bias, int64_t rel_stride, int64_t rel_center,
                              void* stream);
/* One-query cross-attention of a decoder step over the encoder's projected keys / values
 * (HF:t5/modeling_t5.py EncDecAttention at tgt_len == 1): q (B, heads*d) bf16 with row stride
 * q_stride; k, v: token l of sequence b at k + (seq_ids[b] * max_ctx + l) * kv_stride + h*d (dense
 * (B, L, stride) buffers, K and V may be column slices of one GEMM output); tokens
 * [first_valid[b], ctx_len[b]) are attended (padding at either end).  Same flash-decoding splits /
 * workspace / counters contract as vb_paged_decode_attention; nothing is appended. */
int vb_decode_cross_attention(const void* q, int64_t q_stride, const void* k, const void* v, int64_t kv_stride,
                              const int32_t* seq_ids, const int32_t* ctx_len, const int32_t* first_valid,
                              void* out, float* workspace, int32_t* counters, int64_t splits, int64_t batch,
                              int64_t heads, int64_t d, int64_t max_ctx, float scale, void* stream);
/* Copy prefill K/V (B, L, ld) rows into the paged cache. */
int vb_paged_kv_write(const void* k, const void* v, int64_t ld, void* k_cache, void* v_cache,
                      const int32_t* page_table, int64_t batch, int64_t seq, int64_t hd,
                      int64_t page_size, int64_t max_pages, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDEOBLIP_B200_H */
