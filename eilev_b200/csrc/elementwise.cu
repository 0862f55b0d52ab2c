// HBM-bound glue kernels of the VideoBLIP path: patch gather (im2col fused with the
// (N,C,T,H,W)->(N*T,C,H,W) permute and the bf16 cast), CLS rows, LM input assembly
// (embedding gather + video-feature splice + OPT learned positions), shifted cross
// entropy forward/backward, transposes / conversions for the backward GEMM operands,
// activation backward, column sums, fused AdamW.
#include "common.cuh"
#include "internal.h"

namespace vb {

// ------------------------------------------------------------------ dtype helpers
template <typename T> VB_DEVICE float to_f32(T v);
template <> VB_DEVICE float to_f32<float>(float v) { return v; }
template <> VB_DEVICE float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> VB_DEVICE float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> VB_DEVICE T from_f32(float v);
template <> VB_DEVICE float from_f32<float>(float v) { return v; }
template <> VB_DEVICE __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16(v); }
template <> VB_DEVICE __half from_f32<__half>(float v) { return __float2half(v); }

// ------------------------------------------------------------------ patch gather
template <typename T>
__global__ void __launch_bounds__(256)
patch_gather_kernel(const T* __restrict__ px, __nv_bfloat16* __restrict__ out, long long nv,
                    long long C, long long T_, long long H, long long W, long long P, long long gh,
                    long long gw, long long kpad, long long total) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long col = idx % kpad;
  const long long row = idx / kpad;
  float v = 0.0f;
  if (col < C * P * P) {
    const long long pxx = col % P, pyy = (col / P) % P, c = col / (P * P);
    const long long gx = row % gw, gy = (row / gw) % gh, f = row / (gw * gh);
    const long long t = f % T_, vid = f / T_;
    const long long y = gy * P + pyy, x = gx * P + pxx;
    v = to_f32<T>(px[(((vid * C + c) * T_ + t) * H + y) * W + x]);
  }
  out[idx] = __float2bfloat16(v);
}

cudaError_t patch_gather_launch(const void* px, int dtype, void* out, long long nv, long long c,
                                long long t, long long h, long long w, long long patch,
                                long long kpad, cudaStream_t s) {
  const long long gh = h / patch, gw = w / patch;
  const long long total = nv * t * gh * gw * kpad;
  if (total <= 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (dtype == VB_F32)
    launch_pdl(patch_gather_kernel<float>, dim3(grid), dim3(256), 0, s, reinterpret_cast<const float*>(px), o, nv, c, t,
                                                    h, w, patch, gh, gw, kpad, total);
  else if (dtype == VB_BF16)
    launch_pdl(patch_gather_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, 
        reinterpret_cast<const __nv_bfloat16*>(px), o, nv, c, t, h, w, patch, gh, gw, kpad, total);
  else if (dtype == VB_F16)
    launch_pdl(patch_gather_kernel<__half>, dim3(grid), dim3(256), 0, s, reinterpret_cast<const __half*>(px), o, nv, c,
                                                     t, h, w, patch, gh, gw, kpad, total);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// uint8 frames: the BlipImageProcessor arithmetic (rescale by 1/255 in fp64, then subtract the
// channel mean and divide by the channel std in fp32) fused into the same gather, so decoded frames go from
// one byte per sample in HBM straight to the bf16 patch matrix.  A warp covers 32 consecutive
// columns of one patch row group: its byte loads fall into one or two 32-byte sectors.
struct FrameNorm {
  float mean[4];
  float stdv[4];
  double rescale;
};

__global__ void __launch_bounds__(256)
patch_gather_u8_kernel(const unsigned char* __restrict__ px, __nv_bfloat16* __restrict__ out,
                       long long C, long long T_, long long H, long long W, long long P,
                       long long gh, long long gw, long long kpad, long long total, FrameNorm nrm) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long col = idx % kpad;
  const long long row = idx / kpad;
  float v = 0.0f;
  if (col < C * P * P) {
    const long long pxx = col % P, pyy = (col / P) % P, c = col / (P * P);
    const long long gx = row % gw, gy = (row / gw) % gh, f = row / (gw * gh);
    const long long t = f % T_, vid = f / T_;
    const long long y = gy * P + pyy, x = gx * P + pxx;
    const unsigned char raw = px[(((vid * C + c) * T_ + t) * H + y) * W + x];
    // transformers 4.33.1 (the reference's pin): rescale in float64 (uint8 array * python float),
    // cast to float32, then (x - mean) / std in float32
    const float r = static_cast<float>(static_cast<double>(raw) * nrm.rescale);
    // selects instead of a dynamic index: kernel-parameter arrays indexed at run time get copied to local memory
    const int ci = static_cast<int>(c);
    const float mean = ci == 0 ? nrm.mean[0] : ci == 1 ? nrm.mean[1] : ci == 2 ? nrm.mean[2] : nrm.mean[3];
    const float stdv = ci == 0 ? nrm.stdv[0] : ci == 1 ? nrm.stdv[1] : ci == 2 ? nrm.stdv[2] : nrm.stdv[3];
    v = __fdiv_rn(__fsub_rn(r, mean), stdv);
  }
  out[idx] = __float2bfloat16(v);
}

cudaError_t patch_gather_u8_launch(const void* px, void* out, long long nv, long long c, long long t,
                                   long long h, long long w, long long patch, long long kpad,
                                   double rescale, const float* mean, const float* stdv,
                                   cudaStream_t s) {
  if (c < 1 || c > 4) return cudaErrorInvalidValue;
  const long long gh = h / patch, gw = w / patch;
  const long long total = nv * t * gh * gw * kpad;
  if (total <= 0) return cudaSuccess;
  FrameNorm nrm;
  for (int i = 0; i < 4; ++i) {
    nrm.mean[i] = i < c ? mean[i] : 0.0f;
    nrm.stdv[i] = i < c ? stdv[i] : 1.0f;
  }
  nrm.rescale = rescale;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  launch_pdl(patch_gather_u8_kernel, dim3(grid), dim3(256), 0, s, reinterpret_cast<const unsigned char*>(px),
                                              reinterpret_cast<__nv_bfloat16*>(out), c, t, h, w,
                                              patch, gh, gw, kpad, total, nrm);
  return cudaGetLastError();
}

// One pass (horizontal or vertical) of Pillow's 8-bit two-pass resample (Resample.c
// ImagingResampleHorizontal_8bpc / Vertical_8bpc): out = clip8((2^21 + sum_k in[xmin + k] * kk[k]) >> 22)
// with 22-bit fixed-point weights and 32-bit integer accumulation, bit for bit.  The pass runs along
// an arbitrary axis through element strides; `lines_fastest` picks which index varies fastest across a
// warp so that both passes read and write consecutive bytes (horizontal: outputs of one row; vertical:
// neighbouring columns of one output row).
__global__ void __launch_bounds__(256)
resize_u8_pass_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                      const int* __restrict__ bounds, const int* __restrict__ kk, long long planes,
                      long long lines, long long out_len, int ksize, long long ips, long long ils,
                      long long ies, long long ops, long long ols, long long oes, int lines_fastest) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per_plane = lines * out_len;
  if (idx >= planes * per_plane) return;
  const long long p = idx / per_plane;
  const long long r = idx % per_plane;
  const long long l = lines_fastest ? r % lines : r / out_len;
  const long long o = lines_fastest ? r / lines : r % out_len;
  const int xmin = bounds[2 * o], cnt = bounds[2 * o + 1];
  const int* __restrict__ k = kk + o * ksize;
  const unsigned char* __restrict__ src = in + p * ips + l * ils + static_cast<long long>(xmin) * ies;
  int acc = 1 << 21;
  for (int x = 0; x < cnt; ++x) acc += static_cast<int>(src[x * ies]) * k[x];
  acc >>= 22;
  out[p * ops + l * ols + o * oes] = static_cast<unsigned char>(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
}

cudaError_t resize_u8_pass_launch(const void* in, void* out, const int* bounds, const int* kk,
                                  long long planes, long long lines, long long out_len, long long ksize,
                                  long long ips, long long ils, long long ies, long long ops,
                                  long long ols, long long oes, int lines_fastest, cudaStream_t s) {
  const long long total = planes * lines * out_len;
  if (total <= 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  launch_pdl(resize_u8_pass_kernel, dim3(grid), dim3(256), 0, s, reinterpret_cast<const unsigned char*>(in),
                                             reinterpret_cast<unsigned char*>(out), bounds, kk, planes,
                                             lines, out_len, static_cast<int>(ksize), ips, ils, ies, ops,
                                             ols, oes, lines_fastest);
  return cudaGetLastError();
}

__global__ void cls_rows_kernel(const __nv_bfloat16* cls, const __nv_bfloat16* pos,
                                __nv_bfloat16* hidden, long long frames, long long tokens,
                                long long dim) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= frames * dim) return;
  const long long f = idx / dim, c = idx % dim;
  hidden[f * tokens * dim + c] =
      __float2bfloat16(__bfloat162float(cls[c]) + __bfloat162float(pos[c]));
}

cudaError_t cls_rows_launch(const void* cls, const void* pos, void* hidden, long long frames,
                            long long tokens, long long dim, cudaStream_t s) {
  const long long total = frames * dim;
  if (total <= 0) return cudaSuccess;
  launch_pdl(cls_rows_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, 
      reinterpret_cast<const __nv_bfloat16*>(cls), reinterpret_cast<const __nv_bfloat16*>(pos),
      reinterpret_cast<__nv_bfloat16*>(hidden), frames, tokens, dim);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ LM input assembly
// Exclusive prefix sum of one int per thread over a 1024-thread block (warp shuffles + one shared-memory hop);
// `total` = the block-wide sum.  All threads must call it.
VB_DEVICE int block_scan_1024(int v, int* warp_sums, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += n;
  }
  __syncthreads();  // warp_sums may still be read by a previous call
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, w, off);
      if (lane >= off) w += n;
    }
    warp_sums[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  total = warp_sums[31];
  return inc - v + (warp > 0 ? warp_sums[warp - 1] : 0);
}

// Single block: slot_index = exclusive rank among masked positions (row-major over
// (batch, seq), the order of a boolean-mask assignment), pos_ids = OPT positions (cumsum(mask) * mask - 1 +
// offset per batch row).  Both are block-wide scans over contiguous per-thread chunks (round 1 walked each batch
// row with ONE thread: 88 us at batch 1).
__global__ void __launch_bounds__(1024)
splice_index_kernel(const long long* attn, const long long* vmask, int* slot_index, int* pos_ids,
                    int* status, long long batch, long long seq, long long pos_offset,
                    long long n_features) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  __shared__ int warp_sums[32];
  const long long n = batch * seq;
  const int tid = threadIdx.x;
  {
    const long long chunk = (n + 1023) / 1024;
    const long long lo = tid * chunk < n ? tid * chunk : n, hi = (lo + chunk < n) ? lo + chunk : n;
    int cnt = 0;
    if (vmask != nullptr)
      for (long long i = lo; i < hi; ++i) cnt += vmask[i] != 0;
    int total = 0;
    int base = block_scan_1024(cnt, warp_sums, total);
    for (long long i = lo; i < hi; ++i) {
      if (vmask != nullptr && vmask[i] != 0) slot_index[i] = base++;
      else slot_index[i] = -1;
    }
    if (tid == 0 && status != nullptr) {
      status[0] = (vmask != nullptr && total != n_features) ? 1 : 0;
      status[1] = total;
    }
  }
  // positions: a block-wide scan per batch row
  const long long chunk = (seq + 1023) / 1024;
  for (long long b = 0; b < batch; ++b) {
    const long long lo = tid * chunk < seq ? tid * chunk : seq, hi = (lo + chunk < seq) ? lo + chunk : seq;
    int cnt = 0;
    for (long long l = lo; l < hi; ++l) cnt += attn != nullptr ? (attn[b * seq + l] != 0) : 1;
    int total = 0;
    long long run = block_scan_1024(cnt, warp_sums, total);
    for (long long l = lo; l < hi; ++l) {
      const long long m = attn != nullptr ? (attn[b * seq + l] != 0) : 1;
      run += m;
      pos_ids[b * seq + l] = static_cast<int>(run * m - 1 + pos_offset);
    }
  }
}

__global__ void __launch_bounds__(128)
splice_gather_kernel(const long long* ids, const int* slot_index, const int* pos_ids,
                     const __nv_bfloat16* embed, const __nv_bfloat16* feats,
                     const __nv_bfloat16* pos_table, __nv_bfloat16* inputs_embeds,
                     __nv_bfloat16* hidden, long long dim, long long vocab, long long n_features) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long pos = blockIdx.x;
  const int slot = slot_index[pos];
  const __nv_bfloat16* src;
  if (slot >= 0) {
    src = feats + static_cast<long long>(slot < n_features ? slot : 0) * dim;
  } else {
    long long id = ids[pos];
    if (id < 0 || id >= vocab) id = 0;
    src = embed + id * dim;
  }
  const __nv_bfloat16* pt =
      pos_table != nullptr ? pos_table + static_cast<long long>(pos_ids[pos]) * dim : nullptr;
  for (long long c = threadIdx.x; c < dim; c += blockDim.x) {
    const __nv_bfloat16 e = src[c];
    if (inputs_embeds != nullptr) inputs_embeds[pos * dim + c] = e;
    if (hidden != nullptr) {
      float h = __bfloat162float(e);
      if (pt != nullptr) h += __bfloat162float(pt[c]);
      hidden[pos * dim + c] = __float2bfloat16(h);
    }
  }
}

__global__ void __launch_bounds__(128)
splice_bwd_kernel(const __nv_bfloat16* d_embeds, const int* slot_index, __nv_bfloat16* d_feats,
                  long long dim, long long n_features) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long pos = blockIdx.x;
  const int slot = slot_index[pos];
  if (slot < 0 || slot >= n_features) return;
  for (long long c = threadIdx.x; c < dim; c += blockDim.x)
    d_feats[static_cast<long long>(slot) * dim + c] = d_embeds[pos * dim + c];
}

// ------------------------------------------------------------------ cross entropy
template <typename T>
__global__ void __launch_bounds__(256)
ce_row_kernel(const T* logits, const long long* labels, float* row_lse, long long seq,
              long long vocab, long long ldl, int shift) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  // block r handles logits row (b, l); target = labels[b, l+shift] (1: causal LM, 0: seq2seq)
  const long long r = blockIdx.x;
  const long long l = r % seq;
  bool valid = (l + shift < seq);
  long long target = -100;
  if (valid) {
    target = labels[r + shift];
    valid = target >= 0 && target < vocab;
  }
  if (!valid) {
    if (threadIdx.x == 0) row_lse[r] = 0.0f;
    return;
  }
  const T* row = logits + r * ldl;
  __shared__ float red[8];
  float mx = -INFINITY;
  for (long long c = threadIdx.x; c < vocab; c += 256) mx = fmaxf(mx, to_f32<T>(row[c]));
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.0f;
  for (long long c = threadIdx.x; c < vocab; c += 256) sum += __expf(to_f32<T>(row[c]) - mx);
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    row_lse[r] = mx + logf(tot);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
ce_finalize_kernel(const T* logits, const long long* labels, const float* row_lse, float* loss,
                   int* n_valid, long long rows, long long seq, long long vocab, long long ldl, int shift) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  __shared__ float ssum[256];
  __shared__ int scnt[256];
  float acc = 0.0f;
  int cnt = 0;
  for (long long r = threadIdx.x; r < rows; r += 256) {
    const long long l = r % seq;
    if (l + shift >= seq) continue;
    const long long target = labels[r + shift];
    if (target < 0 || target >= vocab) continue;
    acc += row_lse[r] - to_f32<T>(logits[r * ldl + target]);
    cnt += 1;
  }
  ssum[threadIdx.x] = acc;
  scnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      ssum[threadIdx.x] += ssum[threadIdx.x + off];
      scnt[threadIdx.x] += scnt[threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *n_valid = scnt[0];
    // torch: mean over zero valid targets is NaN
    *loss = scnt[0] > 0 ? ssum[0] / static_cast<float>(scnt[0]) : __int_as_float(0x7fc00000);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const T* logits, const long long* labels, const float* row_lse, const int* n_valid,
              const float* grad_scale, __nv_bfloat16* dlogits, long long seq, long long vocab,
              long long ldl, long long ldd, int shift) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long r = blockIdx.x;
  const long long l = r % seq;
  bool valid = (l + shift < seq);
  long long target = -100;
  if (valid) {
    target = labels[r + shift];
    valid = target >= 0 && target < vocab;
  }
  __nv_bfloat16* out = dlogits + r * ldd;
  if (!valid) {
    for (long long c = threadIdx.x; c < vocab; c += 256) out[c] = __float2bfloat16(0.0f);
    return;
  }
  const float gs = (grad_scale != nullptr ? *grad_scale : 1.0f) / static_cast<float>(*n_valid);
  const float lse = row_lse[r];
  const T* row = logits + r * ldl;
  for (long long c = threadIdx.x; c < vocab; c += 256) {
    float pr = __expf(to_f32<T>(row[c]) - lse);
    if (c == target) pr -= 1.0f;
    out[c] = __float2bfloat16(pr * gs);
  }
}

cudaError_t ce_launch(const void* logits, int dtype, const long long* labels, float* loss,
                      float* row_lse, int* n_valid, long long batch, long long seq,
                      long long vocab, long long ldl, int shift, cudaStream_t s) {
  const long long rows = batch * seq;
  if (rows <= 0 || shift < 0 || shift > 1) return cudaErrorInvalidValue;
  if (dtype == VB_BF16) {
    const __nv_bfloat16* lg = reinterpret_cast<const __nv_bfloat16*>(logits);
    launch_pdl(ce_row_kernel<__nv_bfloat16>, dim3(static_cast<unsigned>(rows)), dim3(256), 0, s, lg, labels, row_lse, seq, vocab, ldl, shift);
    launch_pdl(ce_finalize_kernel<__nv_bfloat16>, dim3(1), dim3(256), 0, s, lg, labels, row_lse, loss, n_valid, rows, seq, vocab, ldl, shift);
  } else if (dtype == VB_F32) {
    const float* lg = reinterpret_cast<const float*>(logits);
    launch_pdl(ce_row_kernel<float>, dim3(static_cast<unsigned>(rows)), dim3(256), 0, s, lg, labels, row_lse, seq, vocab, ldl, shift);
    launch_pdl(ce_finalize_kernel<float>, dim3(1), dim3(256), 0, s, lg, labels, row_lse, loss, n_valid, rows, seq, vocab, ldl, shift);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t ce_bwd_launch(const void* logits, int dtype, const long long* labels,
                          const float* row_lse, const int* n_valid, const float* grad_scale,
                          void* dlogits, long long batch, long long seq, long long vocab,
                          long long ldl, long long ldd, int shift, cudaStream_t s) {
  const long long rows = batch * seq;
  if (rows <= 0 || shift < 0 || shift > 1) return cudaErrorInvalidValue;
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dlogits);
  if (dtype == VB_BF16)
    launch_pdl(ce_bwd_kernel<__nv_bfloat16>, dim3(static_cast<unsigned>(rows)), dim3(256), 0, s, 
        reinterpret_cast<const __nv_bfloat16*>(logits), labels, row_lse, n_valid, grad_scale, d, seq, vocab, ldl, ldd, shift);
  else if (dtype == VB_F32)
    launch_pdl(ce_bwd_kernel<float>, dim3(static_cast<unsigned>(rows)), dim3(256), 0, s, 
        reinterpret_cast<const float*>(logits), labels, row_lse, n_valid, grad_scale, d, seq, vocab, ldl, ldd, shift);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ------------------------------------------------------------------ transpose / convert
__global__ void __launch_bounds__(256)
transpose_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                 long long rows, long long cols, long long ld_in, long long ld_out) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  __shared__ __nv_bfloat16 tile[32][34];
  const long long c0 = static_cast<long long>(blockIdx.x) * 32;
  const long long r0 = static_cast<long long>(blockIdx.y) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const long long r = r0 + ty + i, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + i][tx] = in[r * ld_in + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const long long c = c0 + ty + i, r = r0 + tx;
    if (r < rows && c < cols) out[c * ld_out + r] = tile[tx][ty + i];
  }
}

cudaError_t transpose_launch(const void* in, void* out, long long rows, long long cols,
                             long long ld_in, long long ld_out, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  dim3 grid(static_cast<unsigned>((cols + 31) / 32), static_cast<unsigned>((rows + 31) / 32));
  if (grid.y > 65535) return cudaErrorInvalidValue;
  launch_pdl(transpose_kernel, dim3(grid), dim3(256), 0, s, reinterpret_cast<const __nv_bfloat16*>(in),
                                        reinterpret_cast<__nv_bfloat16*>(out), rows, cols, ld_in,
                                        ld_out);
  return cudaGetLastError();
}

template <typename S, typename D>
__global__ void __launch_bounds__(256) convert_kernel(const S* src, D* dst, long long n) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = from_f32<D>(to_f32<S>(src[i]));
}

template <typename S>
static cudaError_t convert_from(const S* src, void* dst, int dd, long long n, cudaStream_t s) {
  const unsigned grid = static_cast<unsigned>(n / 256 + 1 < 148 * 16 ? n / 256 + 1 : 148 * 16);
  if (dd == VB_BF16) launch_pdl(convert_kernel<S, __nv_bfloat16>, dim3(grid), dim3(256), 0, s, src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  else if (dd == VB_F32) launch_pdl(convert_kernel<S, float>, dim3(grid), dim3(256), 0, s, src, reinterpret_cast<float*>(dst), n);
  else if (dd == VB_F16) launch_pdl(convert_kernel<S, __half>, dim3(grid), dim3(256), 0, s, src, reinterpret_cast<__half*>(dst), n);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t convert_launch(const void* src, int sd, void* dst, int dd, long long n, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (sd == VB_BF16) return convert_from(reinterpret_cast<const __nv_bfloat16*>(src), dst, dd, n, s);
  if (sd == VB_F32) return convert_from(reinterpret_cast<const float*>(src), dst, dd, n, s);
  if (sd == VB_F16) return convert_from(reinterpret_cast<const __half*>(src), dst, dd, n, s);
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------ activation bwd, sums
__global__ void __launch_bounds__(256)
act_bwd_kernel(const __nv_bfloat16* dy, const __nv_bfloat16* saved, __nv_bfloat16* dx, int epi,
               long long n) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float g = __bfloat162float(dy[i]);
    const float sv = __bfloat162float(saved[i]);
    float d;
    if (epi == VB_EPI_GELU) d = g * gelu_erf_grad(sv);
    else if (epi == VB_EPI_RELU) d = sv > 0.0f ? g : 0.0f;
    else d = g;
    dx[i] = __float2bfloat16(d);
  }
}

cudaError_t act_bwd_launch(const void* dy, const void* saved, void* dx, int epi, long long n,
                           cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>(n / 256 + 1 < 148 * 16 ? n / 256 + 1 : 148 * 16);
  launch_pdl(act_bwd_kernel, dim3(grid), dim3(256), 0, s, reinterpret_cast<const __nv_bfloat16*>(dy),
                                      reinterpret_cast<const __nv_bfloat16*>(saved),
                                      reinterpret_cast<__nv_bfloat16*>(dx), epi, n);
  return cudaGetLastError();
}

// Block = 32 columns x 32 row lanes (1024 threads, four loads in flight per thread); grid.y splits very tall
// inputs, atomics only across grid.y (one row block up to 16 384 rows: a fixed summation order).
__global__ void __launch_bounds__(1024)
colsum_kernel(const __nv_bfloat16* x, float* out, long long rows, long long cols, long long ldx) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  __shared__ float sm[32][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const long long c = static_cast<long long>(blockIdx.x) * 32 + cx;
  const long long rows_per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r_lo = static_cast<long long>(blockIdx.y) * rows_per;
  const long long r_hi = r_lo + rows_per < rows ? r_lo + rows_per : rows;
  float acc = 0.0f;
  if (c < cols) {
#pragma unroll 4
    for (long long r = r_lo + ry; r < r_hi; r += 32) acc += __bfloat162float(x[r * ldx + c]);
  }
  sm[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += sm[i][cx];
    if (gridDim.y == 1) out[c] += t;
    else atomicAdd(&out[c], t);
  }
}

cudaError_t colsum_launch(const void* x, float* out, long long rows, long long cols, long long ldx,
                          int accumulate, cudaStream_t s) {
  if (cols <= 0) return cudaSuccess;
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * cols, s);
    if (e != cudaSuccess) return e;
  }
  if (rows <= 0) return cudaSuccess;
  unsigned gy = static_cast<unsigned>(rows / 16384 + 1);
  if (gy > 64) gy = 64;
  dim3 grid(static_cast<unsigned>((cols + 31) / 32), gy);
  launch_pdl(colsum_kernel, dim3(grid), dim3(1024), 0, s, reinterpret_cast<const __nv_bfloat16*>(x), out, rows, cols, ldx);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
add_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* y, long long n) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = __float2bfloat16(__bfloat162float(a[i]) + __bfloat162float(b[i]));
}

cudaError_t add_launch(const void* a, const void* b, void* y, long long n, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>(n / 256 + 1 < 148 * 16 ? n / 256 + 1 : 148 * 16);
  launch_pdl(add_kernel, dim3(grid), dim3(256), 0, s, reinterpret_cast<const __nv_bfloat16*>(a),
                                  reinterpret_cast<const __nv_bfloat16*>(b),
                                  reinterpret_cast<__nv_bfloat16*>(y), n);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
dropout_kernel(const __nv_bfloat16* x, __nv_bfloat16* y, long long rows, long long cols, long long ldx,
               long long ldy, unsigned int thresh, float scale, const unsigned long long* seed_ptr,
               unsigned long long salt) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const uint64_t seed = *seed_ptr + salt;
  const long long total = rows * cols;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / cols, c = i % cols;
    const float v = __bfloat162float(x[r * ldx + c]);
    y[r * ldy + c] = __float2bfloat16(dropout_keep(seed, static_cast<uint64_t>(i), thresh) ? v * scale : 0.0f);
  }
}

cudaError_t dropout_launch(const void* x, void* y, long long rows, long long cols, long long ldx,
                           long long ldy, float p, const unsigned long long* seed, unsigned long long salt,
                           cudaStream_t s) {
  const long long n = rows * cols;
  if (n <= 0) return cudaSuccess;
  if (seed == nullptr || !(p >= 0.0f) || p >= 1.0f) return cudaErrorInvalidValue;
  const unsigned grid = static_cast<unsigned>(n / 256 + 1 < 148 * 16 ? n / 256 + 1 : 148 * 16);
  launch_pdl(dropout_kernel, dim3(grid), dim3(256), 0, s, reinterpret_cast<const __nv_bfloat16*>(x),
                                      reinterpret_cast<__nv_bfloat16*>(y), rows, cols, ldx, ldy,
                                      dropout_threshold(p), 1.0f / (1.0f - p), seed, salt);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ optimizer
__global__ void __launch_bounds__(256)
adamw_kernel(float* p, const float* g, float* m, float* v, long long n, float lr, float b1,
             float b2, float eps, float wd, float bc1, float bc2_sqrt, const float* grad_scale) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const float gs = grad_scale != nullptr ? *grad_scale : 1.0f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float grad = g[i] * gs;
    float param = p[i];
    param -= lr * wd * param;  // decoupled weight decay (torch.optim.AdamW)
    const float mi = b1 * m[i] + (1.0f - b1) * grad;
    const float vi = b2 * v[i] + (1.0f - b2) * grad * grad;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    param -= (lr / bc1) * (mi / denom);
    p[i] = param;
  }
}

cudaError_t adamw_launch(float* p, const float* g, float* m, float* v, long long n, float lr,
                         float b1, float b2, float eps, float wd, long long step,
                         const float* grad_scale, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const float bc1 = 1.0f - powf(b1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(b2, static_cast<float>(step));
  const unsigned grid = static_cast<unsigned>(n / 256 + 1 < 148 * 8 ? n / 256 + 1 : 148 * 8);
  launch_pdl(adamw_kernel, dim3(grid), dim3(256), 0, s, p, g, m, v, n, lr, b1, b2, eps, wd, bc1, sqrtf(bc2), grad_scale);
  return cudaGetLastError();
}

// Sum of squares of the flat gradient buffer (global-norm clipping).  DETERMINISTIC: every data-parallel rank must
// derive the same clip coefficient from the same all-reduced gradient, or the replicas drift apart by an ulp per
// step (an fp32 atomicAdd per block made the result depend on block arrival order).  Blocks write their partial sums
// to a fixed slot; the last block to arrive (ticket) adds them up in slot order.  One launch at a time per device
// (the optimizer step's only user).
constexpr int kSumsqMaxBlocks = 148 * 4;
__device__ float g_sumsq_part[kSumsqMaxBlocks];
__device__ unsigned int g_sumsq_ticket;

__global__ void __launch_bounds__(256) sumsq_kernel(const float* x, long long n, float* out) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  __shared__ float red[8];
  __shared__ int s_last;
  float acc = 0.0f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    acc += x[i] * x[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    g_sumsq_part[blockIdx.x] = t;
    __threadfence();
    const unsigned int ticket = atomicAdd(&g_sumsq_ticket, 1u);
    s_last = (ticket == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last != 0) {
    __threadfence();
    // fixed order: thread t adds slots t, t + 256, ...; then the same warp / block tree as above
    float v = 0.0f;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += 256) v += __ldcg(&g_sumsq_part[i]);
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i];
      *out += t;            // the caller zeroes `out`; single writer
      g_sumsq_ticket = 0;   // ready for the next launch / graph replay
    }
  }
}

cudaError_t sumsq_launch(const float* x, long long n, float* out, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>(n / 256 + 1 < kSumsqMaxBlocks ? n / 256 + 1 : kSumsqMaxBlocks);
  launch_pdl(sumsq_kernel, dim3(grid), dim3(256), 0, s, x, n, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ classify helpers
// Merges two softmax-attention partial results over disjoint key sets (the shared prompt
// context and the per-class continuation): out = (e^{l1} o1 + e^{l2} o2) / (e^{l1} + e^{l2}).
// Rows are (H*D)-wide and contiguous; partial i keeps its log-sum-exp as (rows/S_i, H, S_i).
__global__ void __launch_bounds__(256)
attn_merge_kernel(const __nv_bfloat16* o1, const float* lse1, long long s1, const __nv_bfloat16* o2,
                  const float* lse2, long long s2, __nv_bfloat16* out, long long rows, int heads, int d) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long hd = static_cast<long long>(heads) * d;
  const long long total = rows * hd;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / hd;
    const int h = static_cast<int>((i % hd) / d);
    const float l1 = lse1[((r / s1) * heads + h) * s1 + r % s1];
    const float l2 = lse2[((r / s2) * heads + h) * s2 + r % s2];
    const float m = fmaxf(l1, l2);
    float w1 = 0.0f, w2 = 0.0f;
    if (m != -INFINITY) {
      w1 = __expf(l1 - m);
      w2 = __expf(l2 - m);
    }
    const float den = w1 + w2;
    const float inv = den > 0.0f ? 1.0f / den : 0.0f;
    out[i] = __float2bfloat16((w1 * __bfloat162float(o1[i]) + w2 * __bfloat162float(o2[i])) * inv);
  }
}

cudaError_t attn_merge_launch(const void* o1, const float* lse1, long long s1, const void* o2,
                              const float* lse2, long long s2, void* out, long long rows,
                              long long heads, long long d, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  if (s1 <= 0 || s2 <= 0 || rows % s1 != 0 || rows % s2 != 0) return cudaErrorInvalidValue;
  const long long total = rows * heads * d;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(attn_merge_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s, 
      reinterpret_cast<const __nv_bfloat16*>(o1), lse1, s1, reinterpret_cast<const __nv_bfloat16*>(o2),
      lse2, s2, reinterpret_cast<__nv_bfloat16*>(out), rows, static_cast<int>(heads), static_cast<int>(d));
  return cudaGetLastError();
}

// out[i] = log softmax(logits[row_i, :])[target_i]  (0 when target_i is outside [0, vocab):
// the ignore_index rows of nn.CrossEntropyLoss(reduction="none")); row_i = row_index[i] or i.
template <typename T>
__global__ void __launch_bounds__(256)
token_logprob_kernel(const T* logits, const long long* row_index, const long long* targets, float* out,
                     long long vocab, long long ldl) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long i = blockIdx.x;
  const long long target = targets[i];
  if (target < 0 || target >= vocab) {
    if (threadIdx.x == 0) out[i] = 0.0f;
    return;
  }
  const T* row = logits + (row_index != nullptr ? row_index[i] : i) * ldl;
  __shared__ float red[8];
  float mx = -INFINITY;
  for (long long c = threadIdx.x; c < vocab; c += 256) mx = fmaxf(mx, to_f32<T>(row[c]));
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int j = 1; j < 8; ++j) mx = fmaxf(mx, red[j]);
  __syncthreads();
  float sum = 0.0f;
  for (long long c = threadIdx.x; c < vocab; c += 256) sum += __expf(to_f32<T>(row[c]) - mx);
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) tot += red[j];
    out[i] = to_f32<T>(row[target]) - (mx + logf(tot));
  }
}

cudaError_t token_logprob_launch(const void* logits, int dtype, const long long* row_index,
                                 const long long* targets, float* out, long long n, long long vocab,
                                 long long ldl, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (dtype == VB_BF16)
    launch_pdl(token_logprob_kernel<__nv_bfloat16>, dim3(static_cast<unsigned>(n)), dim3(256), 0, s, 
        reinterpret_cast<const __nv_bfloat16*>(logits), row_index, targets, out, vocab, ldl);
  else if (dtype == VB_F32)
    launch_pdl(token_logprob_kernel<float>, dim3(static_cast<unsigned>(n)), dim3(256), 0, s, 
        reinterpret_cast<const float*>(logits), row_index, targets, out, vocab, ldl);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ------------------------------------------------------------------ splice launchers
cudaError_t embed_splice_launch(const long long* ids, const long long* attn, const long long* vmask,
                                const void* embed, const void* feats, const void* pos_table,
                                long long pos_offset, void* inputs_embeds, void* hidden,
                                int* slot_index, int* pos_ids, int* status, long long batch,
                                long long seq, long long dim, long long vocab, long long n_features,
                                cudaStream_t s) {
  const long long n = batch * seq;
  if (n <= 0) return cudaErrorInvalidValue;
  launch_pdl(splice_index_kernel, dim3(1), dim3(1024), 0, s, attn, vmask, slot_index, pos_ids, status, batch, seq,
                                         pos_offset, n_features);
  launch_pdl(splice_gather_kernel, dim3(static_cast<unsigned>(n)), dim3(128), 0, s, 
      ids, slot_index, pos_ids, reinterpret_cast<const __nv_bfloat16*>(embed),
      reinterpret_cast<const __nv_bfloat16*>(feats), reinterpret_cast<const __nv_bfloat16*>(pos_table),
      reinterpret_cast<__nv_bfloat16*>(inputs_embeds), reinterpret_cast<__nv_bfloat16*>(hidden), dim,
      vocab, n_features);
  return cudaGetLastError();
}

cudaError_t splice_bwd_launch(const void* d_embeds, const int* slot_index, void* d_feats,
                              long long positions, long long dim, long long n_features,
                              cudaStream_t s) {
  if (positions <= 0) return cudaSuccess;
  launch_pdl(splice_bwd_kernel, dim3(static_cast<unsigned>(positions)), dim3(128), 0, s, 
      reinterpret_cast<const __nv_bfloat16*>(d_embeds), slot_index,
      reinterpret_cast<__nv_bfloat16*>(d_feats), dim, n_features);
  return cudaGetLastError();
}

}  // namespace vb
