// Internal launcher declarations (one per kernel family) used by api.cu.
#pragma once
#include <cuda_runtime.h>

#include "../../include/videoblip_b200.h"

namespace vb {

struct LnParams;
struct LnBwdParams;

cudaError_t layernorm_fwd(const void* x, const void* residual, const float* gamma, const float* beta,
                          void* y, float* mean, float* rstd, long long rows, long long cols,
                          long long ldx, long long ldr, long long ldy, float eps, cudaStream_t s);
cudaError_t layernorm_bwd(const void* dy, const void* xin, const float* gamma, const float* mean,
                          const float* rstd, const void* dx_add, void* dx, float* dgamma,
                          float* dbeta, long long rows, long long cols, void* dx_drop, float drop_p,
                          const unsigned long long* drop_seed, unsigned long long drop_salt, cudaStream_t s);

cudaError_t row_stats_launch(const void* x, double* stats, long long rows, long long cols, long long ldx,
                             cudaStream_t s);

cudaError_t attention_fwd_launch(const vb_attn_args& a, cudaStream_t stream);
cudaError_t attention_bwd_launch(const vb_attn_bwd_args& a, cudaStream_t stream);
cudaError_t attention_probs_launch(const vb_attn_args& a, void* probs, int out_bf16, cudaStream_t stream);
bool attention_flash_tcgen05_eligible(const vb_attn_args& a);
cudaError_t attention_flash_tcgen05_launch(const vb_attn_args& a, cudaStream_t stream);
bool attention_bwd_tcgen05_eligible(const vb_attn_bwd_args& a);
cudaError_t attention_bwd_tcgen05_launch(const vb_attn_bwd_args& a, cudaStream_t stream);
bool attention_tcgen05_eligible(const vb_attn_args& a);
cudaError_t attention_tcgen05_launch(const vb_attn_args& a, cudaStream_t stream);

cudaError_t patch_gather_launch(const void* px, int dtype, void* out, long long nv, long long c,
                                long long t, long long h, long long w, long long patch,
                                long long kpad, cudaStream_t s);
cudaError_t patch_gather_u8_launch(const void* px, void* out, long long nv, long long c, long long t,
                                   long long h, long long w, long long patch, long long kpad,
                                   double rescale, const float* mean, const float* stdv,
                                   cudaStream_t s);
cudaError_t resize_u8_pass_launch(const void* in, void* out, const int* bounds, const int* kk,
                                  long long planes, long long lines, long long out_len, long long ksize,
                                  long long ips, long long ils, long long ies, long long ops,
                                  long long ols, long long oes, int lines_fastest, cudaStream_t s);
cudaError_t crop_resize_normalize_launch(const void* in, long long c, long long t, long long h, long long w,
                                         long long top, long long left, long long ch, long long cw, int flip,
                                         void* out, int out_bf16, long long oh, long long ow, float rescale,
                                         const float* mean, const float* stdv, cudaStream_t s);
cudaError_t cls_rows_launch(const void* cls, const void* pos, void* hidden, long long frames,
                            long long tokens, long long dim, cudaStream_t s);
cudaError_t embed_splice_launch(const long long* ids, const long long* attn, const long long* vmask,
                                const void* embed, const void* feats, const void* pos_table,
                                long long pos_offset, void* inputs_embeds, void* hidden,
                                int* slot_index, int* pos_ids, int* status, long long batch,
                                long long seq, long long dim, long long vocab, long long n_features,
                                cudaStream_t s);
cudaError_t splice_bwd_launch(const void* d_embeds, const int* slot_index, void* d_feats,
                              long long positions, long long dim, long long n_features,
                              cudaStream_t s);
cudaError_t ce_launch(const void* logits, int dtype, const long long* labels, float* loss,
                      float* row_lse, int* n_valid, long long batch, long long seq,
                      long long vocab, long long ldl, int shift, cudaStream_t s);
cudaError_t ce_bwd_launch(const void* logits, int dtype, const long long* labels,
                          const float* row_lse, const int* n_valid, const float* grad_scale,
                          void* dlogits, long long batch, long long seq, long long vocab,
                          long long ldl, long long ldd, int shift, cudaStream_t s);
cudaError_t attn_merge_launch(const void* o1, const float* lse1, long long s1, const void* o2,
                              const float* lse2, long long s2, void* out, long long rows,
                              long long heads, long long d, cudaStream_t s);
cudaError_t token_logprob_launch(const void* logits, int dtype, const long long* row_index,
                                 const long long* targets, float* out, long long n, long long vocab,
                                 long long ldl, cudaStream_t s);
cudaError_t rmsnorm_fwd_launch(const void* x, const float* gamma, void* y, float* rstd, long long rows,
                               long long cols, long long ldx, long long ldy, float eps, cudaStream_t s);
cudaError_t rmsnorm_bwd_launch(const void* dy, const void* x, const float* gamma, const float* rstd,
                               const void* dx_add, void* dx, long long rows, long long cols, cudaStream_t s);
cudaError_t gated_gelu_fwd_launch(const void* h01, void* out, long long rows, long long dff, cudaStream_t s);
cudaError_t gated_gelu_bwd_launch(const void* d_out, const void* h01, void* d_h01, long long rows,
                                  long long dff, cudaStream_t s);
cudaError_t embedding_launch(const long long* ids, const void* table, void* out, long long n, long long dim,
                             long long vocab, cudaStream_t s);
cudaError_t transpose_launch(const void* in, void* out, long long rows, long long cols,
                             long long ld_in, long long ld_out, cudaStream_t s);
cudaError_t convert_launch(const void* src, int sd, void* dst, int dd, long long n, cudaStream_t s);
cudaError_t act_bwd_launch(const void* dy, const void* saved, void* dx, int epi, long long n,
                           cudaStream_t s);
cudaError_t colsum_launch(const void* x, float* out, long long rows, long long cols, long long ldx,
                          int accumulate, cudaStream_t s);
cudaError_t dropout_launch(const void* x, void* y, long long rows, long long cols, long long ldx,
                           long long ldy, float p, const unsigned long long* seed, unsigned long long salt,
                           cudaStream_t s);
cudaError_t add_launch(const void* a, const void* b, void* y, long long n, cudaStream_t s);
cudaError_t adamw_launch(float* p, const float* g, float* m, float* v, long long n, float lr,
                         float b1, float b2, float eps, float wd, long long step,
                         const float* grad_scale, cudaStream_t s);
cudaError_t sumsq_launch(const float* x, long long n, float* out, cudaStream_t s);

cudaError_t gemv_launch(const void* x, const void* w, const float* bias, const void* residual,
                        void* y, long long m, long long n, long long k, long long ldx,
                        long long ldw, long long ldy, long long ldr, float alpha,
                        long long alpha_cols, int epilogue, int out_dtype, const float* ln_gamma,
                        const float* ln_beta, float ln_eps, cudaStream_t s);
cudaError_t decode_embed_launch(const long long* tokens, const void* embed, const void* pos_table,
                                int* n_valid, int* ctx_len, void* x, long long batch, long long dim,
                                long long vocab, long long pos_rows, long long pos_offset,
                                cudaStream_t s);
void decode_step_set_trace(void* buffer);
cudaError_t decode_step_launch(const vb_decode_op* ops_host, const vb_decode_op* ops_dev, int n_ops, int m,
                               unsigned* workspace, cudaStream_t s);
cudaError_t paged_decode_attention_launch(const void* qkv, void* k_cache, void* v_cache,
                                          const int* page_table, const int* ctx_len,
                                          const int* first_valid, void* out, float* workspace,
                                          int* counters, long long splits, long long batch,
                                          long long heads, long long d, long long page_size,
                                          long long max_pages, float scale, const float* rel_bias,
                                          long long rel_stride, long long rel_center, cudaStream_t s);
cudaError_t decode_cross_attention_launch(const void* q, long long q_stride, const void* k, const void* v,
                                          long long kv_stride, const int* seq_ids, const int* ctx_len,
                                          const int* first_valid, void* out, float* workspace, int* counters,
                                          long long splits, long long batch, long long heads, long long d,
                                          long long max_ctx, float scale, cudaStream_t s);
cudaError_t paged_kv_write_launch(const void* k, const void* v, long long ld, void* k_cache,
                                  void* v_cache, const int* page_table, long long batch,
                                  long long seq, long long hd, long long page_size,
                                  long long max_pages, cudaStream_t s);
}  // namespace vb
