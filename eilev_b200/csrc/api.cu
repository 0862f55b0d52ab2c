// extern "C" surface of libvideoblip_b200.so — see include/videoblip_b200.h.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "gemm.h"
#include "internal.h"

namespace {
thread_local char g_err[512] = "";

int fail(const char* where, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return 1;
}
int fail_msg(const char* where, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, msg);
  return 1;
}
inline cudaStream_t st(void* s) { return reinterpret_cast<cudaStream_t>(s); }
#define VB_CHECK(where, expr)                       \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) return fail(where, _e);  \
    return 0;                                       \
  } while (0)
// Keys cubic (a = -0.5) exactly as Pillow's Resample.c bicubic_filter evaluates it (host, double).
inline double vb_bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
}  // namespace

extern "C" {

int vb_abi_version(void) { return VB_ABI_VERSION; }
const char* vb_last_error(void) { return g_err; }

int vb_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return major * 10 + minor;
}

int vb_gemm_uses_tcgen05(const vb_gemm_args* a) {
  if (a == nullptr) return 0;
  if (a->backend == VB_GEMM_GENERIC) return 0;
  return vb::gemm_tcgen05_eligible(*a) ? 1 : 0;
}

int vb_gemm(const vb_gemm_args* a, void* stream) {
  if (a == nullptr) return fail_msg("vb_gemm", "null args");
  if (a->m < 0 || a->n < 0 || a->k <= 0) return fail_msg("vb_gemm", "bad shape");
  if (a->m == 0 || a->n == 0) return 0;
  if (a->a == nullptr || a->b == nullptr || a->c == nullptr) return fail_msg("vb_gemm", "null operand");
  if (a->out_dtype != VB_BF16 && a->out_dtype != VB_F32) return fail_msg("vb_gemm", "bad out_dtype");
  const bool elig = vb::gemm_tcgen05_eligible(*a);
  if (a->reserved2 != 0) {
    if (a->reserved2 != 1) return fail_msg("vb_gemm", "bad operand_layout");
    if (!elig || a->backend == VB_GEMM_GENERIC || a->ln_stats != nullptr || a->stats_out != nullptr ||
        a->stats_zero != nullptr || a->row_group != 0)
      return fail_msg("vb_gemm", "transposed operands need the plain tcgen05 path (M, N, lda, ldb % 8 == 0)");
  }
  if (a->backend == VB_GEMM_TCGEN05 && !elig)
    return fail_msg("vb_gemm", "shape/alignment not eligible for the tcgen05 path");
  if ((a->epilogue == VB_EPI_GELU_BWD || a->epilogue == VB_EPI_RELU_BWD) &&
      (a->residual == nullptr || a->row_group != 0 || a->dropout_p > 0.0f))
    return fail_msg("vb_gemm", "the activation-backward epilogues take the saved forward tensor as `residual`");
  if ((a->ln_stats != nullptr) != (a->ln_colsum != nullptr))
    return fail_msg("vb_gemm", "ln_stats and ln_colsum go together");
  if (a->ln_stats != nullptr || a->stats_out != nullptr || a->stats_zero != nullptr) {
    if (!elig || a->backend == VB_GEMM_GENERIC)
      return fail_msg("vb_gemm", "the LayerNorm fold / row statistics need the tcgen05 path");
    if (a->stats_out != nullptr && (a->out_dtype != VB_BF16 || a->beta != 0.0f))
      return fail_msg("vb_gemm", "stats_out needs a bf16 output and beta == 0");
    if (a->stats_zero != nullptr && (a->stats_zero == a->stats_out || a->stats_zero == a->ln_stats || a->row_group != 0))
      return fail_msg("vb_gemm", "stats_zero must be a buffer this launch neither reads nor fills");
    if ((reinterpret_cast<uintptr_t>(a->stats_zero) & 15u) != 0 || (reinterpret_cast<uintptr_t>(a->ln_stats) & 15u) != 0 ||
        (reinterpret_cast<uintptr_t>(a->stats_out) & 7u) != 0 || (reinterpret_cast<uintptr_t>(a->ln_colsum) & 15u) != 0)
      return fail_msg("vb_gemm", "ln_stats / ln_colsum alignment");
  }
  if (a->backend != VB_GEMM_GENERIC && elig) VB_CHECK("vb_gemm[tcgen05]", vb::gemm_tcgen05_launch(*a, st(stream)));
  VB_CHECK("vb_gemm[generic]", vb::gemm_generic_launch(*a, st(stream)));
}

int vb_layernorm(const void* x, const void* residual, const float* gamma, const float* beta,
                 void* y, float* mean, float* rstd, int64_t rows, int64_t cols, int64_t ldx,
                 int64_t ldr, int64_t ldy, float eps, void* stream) {
  if (x == nullptr || y == nullptr || gamma == nullptr || beta == nullptr)
    return fail_msg("vb_layernorm", "null operand");
  VB_CHECK("vb_layernorm", vb::layernorm_fwd(x, residual, gamma, beta, y, mean, rstd, rows, cols,
                                             ldx, ldr, ldy, eps, st(stream)));
}

int vb_row_stats(const void* x, double* stats, int64_t rows, int64_t cols, int64_t ldx, void* stream) {
  if (x == nullptr || stats == nullptr || rows < 0 || cols <= 0 || ldx < cols)
    return fail_msg("vb_row_stats", "bad arguments");
  VB_CHECK("vb_row_stats", vb::row_stats_launch(x, stats, rows, cols, ldx, st(stream)));
}

int vb_layernorm_bwd(const void* dy, const void* xin, const float* gamma, const float* mean,
                     const float* rstd, const void* dx_add, void* dx, float* dgamma, float* dbeta,
                     int64_t rows, int64_t cols, float, void* stream) {
  if (dy == nullptr || xin == nullptr || gamma == nullptr || mean == nullptr || rstd == nullptr ||
      dx == nullptr)
    return fail_msg("vb_layernorm_bwd", "null operand");
  VB_CHECK("vb_layernorm_bwd", vb::layernorm_bwd(dy, xin, gamma, mean, rstd, dx_add, dx, dgamma,
                                                 dbeta, rows, cols, nullptr, 0.0f, nullptr, 0, st(stream)));
}

int vb_layernorm_bwd_dropout(const void* dy, const void* xin, const float* gamma, const float* mean,
                             const float* rstd, const void* dx_add, void* dx, void* dx_drop, float dropout_p,
                             const uint64_t* dropout_seed, uint64_t dropout_salt, int64_t rows, int64_t cols,
                             void* stream) {
  if (dy == nullptr || xin == nullptr || gamma == nullptr || mean == nullptr || rstd == nullptr ||
      dx == nullptr || dx_drop == nullptr || dropout_seed == nullptr)
    return fail_msg("vb_layernorm_bwd_dropout", "null operand");
  if (!(dropout_p > 0.0f) || dropout_p >= 1.0f) return fail_msg("vb_layernorm_bwd_dropout", "dropout_p must be in (0, 1)");
  VB_CHECK("vb_layernorm_bwd_dropout",
           vb::layernorm_bwd(dy, xin, gamma, mean, rstd, dx_add, dx, nullptr, nullptr, rows, cols, dx_drop, dropout_p,
                             reinterpret_cast<const unsigned long long*>(dropout_seed), dropout_salt, st(stream)));
}

int vb_attention_fwd(const vb_attn_args* a, void* stream) {
  if (a == nullptr || a->q == nullptr || a->k == nullptr || a->v == nullptr || a->o == nullptr)
    return fail_msg("vb_attention_fwd", "null operand");
  VB_CHECK("vb_attention_fwd", vb::attention_fwd_launch(*a, st(stream)));
}

int vb_attention_probs(const vb_attn_args* a, void* probs, int32_t probs_dtype, void* stream) {
  if (a == nullptr || a->q == nullptr || a->k == nullptr || probs == nullptr)
    return fail_msg("vb_attention_probs", "null operand");
  if (probs_dtype != VB_F32 && probs_dtype != VB_BF16) return fail_msg("vb_attention_probs", "bad probs_dtype");
  if (a->d <= 0 || (a->d + a->skv) * 4 > 48 * 1024) return fail_msg("vb_attention_probs", "d + skv too large");
  VB_CHECK("vb_attention_probs", vb::attention_probs_launch(*a, probs, probs_dtype == VB_BF16 ? 1 : 0, st(stream)));
}

int vb_attention_uses_tcgen05(const vb_attn_args* a) {
  if (a == nullptr) return 0;
  if (vb::attention_tcgen05_eligible(*a)) return 1;
  return vb::attention_flash_tcgen05_eligible(*a) ? 2 : 0;
}

int vb_attention_bwd(const vb_attn_bwd_args* a, void* stream) {
  if (a == nullptr || a->d_o == nullptr || a->dq == nullptr || a->dk == nullptr || a->dv == nullptr)
    return fail_msg("vb_attention_bwd", "null operand");
  VB_CHECK("vb_attention_bwd", vb::attention_bwd_launch(*a, st(stream)));
}

int vb_attention_bwd_uses_tcgen05(const vb_attn_bwd_args* a) {
  if (a == nullptr) return 0;
  return vb::attention_bwd_tcgen05_eligible(*a) ? 1 : 0;
}

int vb_patch_gather(const void* pixels, int32_t px_dtype, void* out, int64_t nv, int64_t c,
                    int64_t t, int64_t h, int64_t w, int64_t patch, int64_t kpad, void* stream) {
  if (pixels == nullptr || out == nullptr || patch <= 0 || kpad < c * patch * patch)
    return fail_msg("vb_patch_gather", "bad arguments");
  VB_CHECK("vb_patch_gather",
           vb::patch_gather_launch(pixels, px_dtype, out, nv, c, t, h, w, patch, kpad, st(stream)));
}

int vb_patch_gather_u8(const void* pixels_u8, void* out, int64_t nv, int64_t c, int64_t t, int64_t h,
                       int64_t w, int64_t patch, int64_t kpad, double rescale, const float* mean,
                       const float* stdv, void* stream) {
  if (pixels_u8 == nullptr || out == nullptr || mean == nullptr || stdv == nullptr || patch <= 0 ||
      kpad < c * patch * patch || c < 1 || c > 4)
    return fail_msg("vb_patch_gather_u8", "bad arguments");
  for (int64_t i = 0; i < c; ++i)
    if (!(stdv[i] > 0.0f)) return fail_msg("vb_patch_gather_u8", "std must be positive");
  VB_CHECK("vb_patch_gather_u8", vb::patch_gather_u8_launch(pixels_u8, out, nv, c, t, h, w, patch, kpad,
                                                           rescale, mean, stdv, st(stream)));
}

// ---- Pillow-exact antialiased bicubic resize of uint8 planes --------------------------------
// Host side: the window bounds and 22-bit fixed-point weights of one axis, in double precision
// with Pillow's own expression order (Resample.c precompute_coeffs + normalize_coeffs_8bpc; Keys
// cubic a = -0.5, support 2 stretched by the scale when down-sampling).
int vb_resize_bicubic_ksize(int64_t in_size, int64_t out_size) {
  if (in_size <= 0 || out_size <= 0) return 0;
  double filterscale = static_cast<double>(in_size) / static_cast<double>(out_size);
  if (filterscale < 1.0) filterscale = 1.0;
  return static_cast<int>(std::ceil(2.0 * filterscale)) * 2 + 1;
}

int vb_resize_bicubic_coeffs(int64_t in_size, int64_t out_size, int32_t* bounds, int32_t* kk,
                             int64_t kk_capacity) {
  const int64_t ksize = vb_resize_bicubic_ksize(in_size, out_size);
  if (ksize <= 0 || bounds == nullptr || kk == nullptr || kk_capacity < out_size * ksize)
    return fail_msg("vb_resize_bicubic_coeffs", "bad arguments");
  const double scale = static_cast<double>(in_size) / static_cast<double>(out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const double ss = 1.0 / filterscale;
  std::vector<double> w(static_cast<size_t>(ksize));
  for (int64_t xx = 0; xx < out_size; ++xx) {
    const double center = (static_cast<double>(xx) + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = static_cast<int>(in_size);
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = vb_bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    int32_t* k = kk + xx * ksize;
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? w[x] / ww : w[x];
      k[x] = v < 0 ? static_cast<int32_t>(-0.5 + v * (1 << 22)) : static_cast<int32_t>(0.5 + v * (1 << 22));
    }
    for (int64_t x = xmax; x < ksize; ++x) k[x] = 0;
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return 0;
}

int vb_resize_u8_pass(const void* in, void* out, const int32_t* bounds, const int32_t* kk, int64_t planes,
                      int64_t lines, int64_t out_len, int64_t ksize, int64_t in_plane_stride,
                      int64_t in_line_stride, int64_t in_elem_stride, int64_t out_plane_stride,
                      int64_t out_line_stride, int64_t out_elem_stride, int32_t lines_fastest, void* stream) {
  if (in == nullptr || out == nullptr || bounds == nullptr || kk == nullptr || ksize <= 0 || planes < 0 ||
      lines < 0 || out_len < 0)
    return fail_msg("vb_resize_u8_pass", "bad arguments");
  VB_CHECK("vb_resize_u8_pass",
           vb::resize_u8_pass_launch(in, out, bounds, kk, planes, lines, out_len, ksize, in_plane_stride,
                                     in_line_stride, in_elem_stride, out_plane_stride, out_line_stride,
                                     out_elem_stride, lines_fastest, st(stream)));
}

int vb_crop_resize_normalize_u8(const void* frames_u8, int64_t c, int64_t t, int64_t h, int64_t w, int64_t crop_top,
                                int64_t crop_left, int64_t crop_h, int64_t crop_w, int32_t flip, void* out,
                                int32_t out_dtype, int64_t out_h, int64_t out_w, double rescale, const float* mean,
                                const float* stdv, void* stream) {
  if (frames_u8 == nullptr || out == nullptr || mean == nullptr || stdv == nullptr || c < 1 || c > 4 || t < 1 ||
      crop_h < 1 || crop_w < 1 || crop_top < 0 || crop_left < 0 || crop_top + crop_h > h || crop_left + crop_w > w ||
      out_h < 1 || out_w < 1 || (out_dtype != VB_F32 && out_dtype != VB_BF16))
    return fail_msg("vb_crop_resize_normalize_u8", "bad arguments");
  for (int64_t i = 0; i < c; ++i)
    if (!(stdv[i] > 0.0f)) return fail_msg("vb_crop_resize_normalize_u8", "std must be positive");
  VB_CHECK("vb_crop_resize_normalize_u8",
           vb::crop_resize_normalize_launch(frames_u8, c, t, h, w, crop_top, crop_left, crop_h, crop_w, flip, out,
                                            out_dtype == VB_BF16 ? 1 : 0, out_h, out_w, static_cast<float>(rescale),
                                            mean, stdv, st(stream)));
}

int vb_cls_rows(const void* cls, const void* pos, void* hidden, int64_t frames, int64_t tokens,
                int64_t dim, void* stream) {
  VB_CHECK("vb_cls_rows", vb::cls_rows_launch(cls, pos, hidden, frames, tokens, dim, st(stream)));
}

int vb_embed_splice(const int64_t* input_ids, const int64_t* attention_mask,
                    const int64_t* video_mask, const void* embed_tokens,
                    const void* video_features, const void* pos_table, int64_t pos_offset,
                    void* inputs_embeds, void* hidden, int32_t* slot_index, int32_t* pos_ids,
                    int32_t* status, int64_t batch, int64_t seq, int64_t dim, int64_t vocab,
                    int64_t n_features, void* stream) {
  if (input_ids == nullptr || embed_tokens == nullptr || slot_index == nullptr || pos_ids == nullptr)
    return fail_msg("vb_embed_splice", "null operand");
  if (video_mask != nullptr && (video_features == nullptr || n_features <= 0))
    return fail_msg("vb_embed_splice", "video_mask without video_features (pass a null mask when there are no features)");
  VB_CHECK("vb_embed_splice",
           vb::embed_splice_launch(reinterpret_cast<const long long*>(input_ids),
                                   reinterpret_cast<const long long*>(attention_mask),
                                   reinterpret_cast<const long long*>(video_mask), embed_tokens,
                                   video_features, pos_table, pos_offset, inputs_embeds, hidden,
                                   slot_index, pos_ids, status, batch, seq, dim, vocab, n_features,
                                   st(stream)));
}

int vb_splice_bwd(const void* d_embeds, const int32_t* slot_index, void* d_features,
                  int64_t positions, int64_t dim, int64_t n_features, void* stream) {
  VB_CHECK("vb_splice_bwd",
           vb::splice_bwd_launch(d_embeds, slot_index, d_features, positions, dim, n_features, st(stream)));
}

int vb_cross_entropy(const void* logits, int32_t logits_dtype, const int64_t* labels, float* loss,
                     float* row_lse, int32_t* n_valid, int64_t batch, int64_t seq, int64_t vocab,
                     int64_t ldl, int32_t shift, void* stream) {
  if (logits == nullptr || labels == nullptr || loss == nullptr || row_lse == nullptr || n_valid == nullptr)
    return fail_msg("vb_cross_entropy", "null operand");
  VB_CHECK("vb_cross_entropy",
           vb::ce_launch(logits, logits_dtype, reinterpret_cast<const long long*>(labels), loss,
                         row_lse, n_valid, batch, seq, vocab, ldl, shift, st(stream)));
}

int vb_cross_entropy_bwd(const void* logits, int32_t logits_dtype, const int64_t* labels,
                         const float* row_lse, const int32_t* n_valid, const float* grad_scale,
                         void* dlogits, int64_t batch, int64_t seq, int64_t vocab, int64_t ldl,
                         int64_t ldd, int32_t shift, void* stream) {
  if (logits == nullptr || labels == nullptr || row_lse == nullptr || n_valid == nullptr || dlogits == nullptr)
    return fail_msg("vb_cross_entropy_bwd", "null operand");
  VB_CHECK("vb_cross_entropy_bwd",
           vb::ce_bwd_launch(logits, logits_dtype, reinterpret_cast<const long long*>(labels),
                             row_lse, n_valid, grad_scale, dlogits, batch, seq, vocab, ldl, ldd, shift,
                             st(stream)));
}

int vb_attention_merge(const void* o1, const float* lse1, int64_t s1, const void* o2, const float* lse2,
                       int64_t s2, void* out, int64_t rows, int64_t heads, int64_t d, void* stream) {
  VB_CHECK("vb_attention_merge",
           vb::attn_merge_launch(o1, lse1, s1, o2, lse2, s2, out, rows, heads, d, st(stream)));
}

int vb_token_logprob(const void* logits, int32_t logits_dtype, const int64_t* row_index,
                     const int64_t* targets, float* out, int64_t n, int64_t vocab, int64_t ldl,
                     void* stream) {
  VB_CHECK("vb_token_logprob",
           vb::token_logprob_launch(logits, logits_dtype, reinterpret_cast<const long long*>(row_index),
                                    reinterpret_cast<const long long*>(targets), out, n, vocab, ldl,
                                    st(stream)));
}

int vb_rmsnorm(const void* x, const float* gamma, void* y, float* rstd, int64_t rows, int64_t cols,
               int64_t ldx, int64_t ldy, float eps, void* stream) {
  VB_CHECK("vb_rmsnorm", vb::rmsnorm_fwd_launch(x, gamma, y, rstd, rows, cols, ldx, ldy, eps, st(stream)));
}

int vb_rmsnorm_bwd(const void* dy, const void* x, const float* gamma, const float* rstd, const void* dx_add,
                   void* dx, int64_t rows, int64_t cols, void* stream) {
  VB_CHECK("vb_rmsnorm_bwd", vb::rmsnorm_bwd_launch(dy, x, gamma, rstd, dx_add, dx, rows, cols, st(stream)));
}

int vb_gated_gelu(const void* h01, void* out, int64_t rows, int64_t dff, void* stream) {
  VB_CHECK("vb_gated_gelu", vb::gated_gelu_fwd_launch(h01, out, rows, dff, st(stream)));
}

int vb_gated_gelu_bwd(const void* d_out, const void* h01, void* d_h01, int64_t rows, int64_t dff, void* stream) {
  VB_CHECK("vb_gated_gelu_bwd", vb::gated_gelu_bwd_launch(d_out, h01, d_h01, rows, dff, st(stream)));
}

int vb_embedding(const int64_t* ids, const void* table, void* out, int64_t n, int64_t dim, int64_t vocab,
                 void* stream) {
  VB_CHECK("vb_embedding", vb::embedding_launch(reinterpret_cast<const long long*>(ids), table, out, n, dim,
                                                vocab, st(stream)));
}

int vb_transpose(const void* in, void* out, int64_t rows, int64_t cols, int64_t ld_in,
                 int64_t ld_out, void* stream) {
  VB_CHECK("vb_transpose", vb::transpose_launch(in, out, rows, cols, ld_in, ld_out, st(stream)));
}

int vb_convert(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n,
               void* stream) {
  VB_CHECK("vb_convert", vb::convert_launch(src, src_dtype, dst, dst_dtype, n, st(stream)));
}

int vb_act_bwd(const void* dy, const void* saved, void* dx, int32_t epilogue, int64_t n,
               void* stream) {
  VB_CHECK("vb_act_bwd", vb::act_bwd_launch(dy, saved, dx, epilogue, n, st(stream)));
}

int vb_colsum(const void* x, float* out, int64_t rows, int64_t cols, int64_t ldx,
              int32_t accumulate, void* stream) {
  VB_CHECK("vb_colsum", vb::colsum_launch(x, out, rows, cols, ldx, accumulate, st(stream)));
}

int vb_dropout(const void* x, void* y, int64_t rows, int64_t cols, int64_t ldx, int64_t ldy, float p,
               const uint64_t* seed, uint64_t salt, void* stream) {
  VB_CHECK("vb_dropout", vb::dropout_launch(x, y, rows, cols, ldx, ldy, p,
                                            reinterpret_cast<const unsigned long long*>(seed), salt, st(stream)));
}

int vb_add(const void* a, const void* b, void* y, int64_t n, void* stream) {
  VB_CHECK("vb_add", vb::add_launch(a, b, y, n, st(stream)));
}

int vb_adamw(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
             float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
             const float* grad_scale, void* stream) {
  if (step <= 0) return fail_msg("vb_adamw", "step must be >= 1");
  VB_CHECK("vb_adamw", vb::adamw_launch(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                        weight_decay, step, grad_scale, st(stream)));
}

int vb_sumsq(const float* x, int64_t n, float* out, void* stream) {
  VB_CHECK("vb_sumsq", vb::sumsq_launch(x, n, out, st(stream)));
}

int vb_gemv(const void* x, const void* w, const float* bias, const void* residual, void* y,
            int64_t m, int64_t n, int64_t k, int64_t ldx, int64_t ldw, int64_t ldy, int64_t ldr,
            float alpha, int64_t alpha_cols, int32_t epilogue, int32_t out_dtype,
            const float* ln_gamma, const float* ln_beta, float ln_eps, void* stream) {
  if (ln_gamma == nullptr && ln_beta != nullptr) return fail_msg("vb_gemv", "ln_beta without ln_gamma");
  VB_CHECK("vb_gemv", vb::gemv_launch(x, w, bias, residual, y, m, n, k, ldx, ldw, ldy, ldr, alpha,
                                      alpha_cols, epilogue, out_dtype, ln_gamma, ln_beta, ln_eps, st(stream)));
}

int vb_decode_embed(const int64_t* tokens, const void* embed, const void* pos_table, int32_t* n_valid,
                    int32_t* ctx_len, void* x, int64_t batch, int64_t dim, int64_t vocab,
                    int64_t pos_rows, int64_t pos_offset, void* stream) {
  VB_CHECK("vb_decode_embed",
           vb::decode_embed_launch(reinterpret_cast<const long long*>(tokens), embed, pos_table, n_valid,
                                   ctx_len, x, batch, dim, vocab, pos_rows, pos_offset, st(stream)));
}

int vb_decode_step(const vb_decode_op* ops_host, const vb_decode_op* ops_dev, int32_t n_ops, int32_t m,
                   uint32_t* workspace, void* stream) {
  VB_CHECK("vb_decode_step", vb::decode_step_launch(ops_host, ops_dev, n_ops, m, workspace, st(stream)));
}

int vb_debug_decode_trace(void* buffer) {
  vb::decode_step_set_trace(buffer);
  return 0;
}

int vb_paged_decode_attention(const void* qkv, void* k_cache, void* v_cache,
                              const int32_t* page_table, const int32_t* ctx_len,
                              const int32_t* first_valid, void* out, float* workspace,
                              int32_t* counters, int64_t splits, int64_t batch, int64_t heads,
                              int64_t d, int64_t page_size, int64_t max_pages, float scale,
                              const float* rel_bias, int64_t rel_stride, int64_t rel_center,
                              void* stream) {
  VB_CHECK("vb_paged_decode_attention",
           vb::paged_decode_attention_launch(qkv, k_cache, v_cache, page_table, ctx_len,
                                             first_valid, out, workspace, counters, splits, batch,
                                             heads, d, page_size, max_pages, scale, rel_bias, rel_stride,
                                             rel_center, st(stream)));
}

int vb_decode_cross_attention(const void* q, int64_t q_stride, const void* k, const void* v, int64_t kv_stride,
                              const int32_t* seq_ids, const int32_t* ctx_len, const int32_t* first_valid,
                              void* out, float* workspace, int32_t* counters, int64_t splits, int64_t batch,
                              int64_t heads, int64_t d, int64_t max_ctx, float scale, void* stream) {
  VB_CHECK("vb_decode_cross_attention",
           vb::decode_cross_attention_launch(q, q_stride, k, v, kv_stride, seq_ids, ctx_len, first_valid, out,
                                             workspace, counters, splits, batch, heads, d, max_ctx, scale,
                                             st(stream)));
}

int vb_paged_kv_write(const void* k, const void* v, int64_t ld, void* k_cache, void* v_cache,
                      const int32_t* page_table, int64_t batch, int64_t seq, int64_t hd,
                      int64_t page_size, int64_t max_pages, void* stream) {
  VB_CHECK("vb_paged_kv_write", vb::paged_kv_write_launch(k, v, ld, k_cache, v_cache, page_table,
                                                          batch, seq, hd, page_size, max_pages,
                                                          st(stream)));
}

}  // extern "C"
