// Token-by-token generation kernels (HBM-bound): a weight-streaming GEMV family for
// M <= 16 rows and single-query attention over a paged KV cache.
//   bytes per decode step ~ all LM weights (5.29 GB for OPT-2.7B in bf16) + the KV pages
//   touched, so every weight element is read exactly once with 16-byte loads and the
//   cache is laid out (page, slot, head*dim) for contiguous per-token rows.
#include "common.cuh"
#include "internal.h"

namespace vb {

constexpr int kGemvMaxM = 16;

struct GemvParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;
  const float* bias;
  const __nv_bfloat16* residual;
  void* y;
  long long m, n, k, ldx, ldw, ldy, ldr;
  float alpha;
  long long alpha_cols;
  int epilogue, out_f32, vec;
};

// One warp per output feature n: streams W[n, :] once, dots it with all M rows of x.
template <int M>
__global__ void __launch_bounds__(128) gemv_kernel(const GemvParams p) {
  const int lane = threadIdx.x & 31;
  const long long n = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (n >= p.n) return;
  float acc[M];
#pragma unroll
  for (int i = 0; i < M; ++i) acc[i] = 0.0f;
  const __nv_bfloat16* wr = p.w + n * p.ldw;
  if (p.vec) {
    for (long long k0 = lane * 8; k0 < p.k; k0 += 256) {
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wr + k0));
      const float2 w0 = unpack_bf16x2(wv.x), w1 = unpack_bf16x2(wv.y), w2 = unpack_bf16x2(wv.z),
                   w3 = unpack_bf16x2(wv.w);
#pragma unroll
      for (int i = 0; i < M; ++i) {
        if (i < p.m) {
          const uint4 xv = *reinterpret_cast<const uint4*>(p.x + i * p.ldx + k0);
          const float2 x0 = unpack_bf16x2(xv.x), x1 = unpack_bf16x2(xv.y), x2 = unpack_bf16x2(xv.z),
                       x3 = unpack_bf16x2(xv.w);
          acc[i] += w0.x * x0.x + w0.y * x0.y + w1.x * x1.x + w1.y * x1.y + w2.x * x2.x +
                    w2.y * x2.y + w3.x * x3.x + w3.y * x3.y;
        }
      }
    }
  } else {
    for (long long k0 = lane; k0 < p.k; k0 += 32) {
      const float wv = __bfloat162float(wr[k0]);
#pragma unroll
      for (int i = 0; i < M; ++i)
        if (i < p.m) acc[i] += wv * __bfloat162float(p.x[i * p.ldx + k0]);
    }
  }
#pragma unroll
  for (int i = 0; i < M; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
    const long long ac = p.alpha_cols <= 0 ? p.n : p.alpha_cols;
    for (int i = 0; i < M; ++i) {
      if (i >= p.m) break;
      float v = acc[i];
      if (p.bias != nullptr) v += p.bias[n];
      if (n < ac) v *= p.alpha;
      if (p.epilogue == VB_EPI_GELU) v = gelu_erf(v);
      else if (p.epilogue == VB_EPI_RELU) v = fmaxf(v, 0.0f);
      if (p.residual != nullptr) v += __bfloat162float(p.residual[i * p.ldr + n]);
      if (p.out_f32) reinterpret_cast<float*>(p.y)[i * p.ldy + n] = v;
      else reinterpret_cast<__nv_bfloat16*>(p.y)[i * p.ldy + n] = __float2bfloat16(v);
    }
  }
}

cudaError_t gemv_launch(const void* x, const void* w, const float* bias, const void* residual,
                        void* y, long long m, long long n, long long k, long long ldx,
                        long long ldw, long long ldy, long long ldr, float alpha,
                        long long alpha_cols, int epilogue, int out_dtype, cudaStream_t s) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  if (m > kGemvMaxM || k <= 0) return cudaErrorInvalidValue;
  GemvParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.w = reinterpret_cast<const __nv_bfloat16*>(w);
  p.bias = bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.y = y;
  p.m = m; p.n = n; p.k = k; p.ldx = ldx; p.ldw = ldw; p.ldy = ldy; p.ldr = ldr;
  p.alpha = alpha; p.alpha_cols = alpha_cols; p.epilogue = epilogue;
  p.out_f32 = out_dtype == VB_F32 ? 1 : 0;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  p.vec = (k % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0 && al(x) && al(w)) ? 1 : 0;
  const unsigned grid = static_cast<unsigned>((n + 3) / 4);
  if (m <= 1) gemv_kernel<1><<<grid, 128, 0, s>>>(p);
  else if (m <= 2) gemv_kernel<2><<<grid, 128, 0, s>>>(p);
  else if (m <= 4) gemv_kernel<4><<<grid, 128, 0, s>>>(p);
  else if (m <= 8) gemv_kernel<8><<<grid, 128, 0, s>>>(p);
  else gemv_kernel<16><<<grid, 128, 0, s>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ paged KV cache
__global__ void __launch_bounds__(128)
paged_kv_write_kernel(const __nv_bfloat16* k, const __nv_bfloat16* v, long long ld,
                      __nv_bfloat16* kc, __nv_bfloat16* vc, const int* page_table, long long seq,
                      long long hd, long long page_size, long long max_pages) {
  const long long tok = blockIdx.x;  // b*seq + l
  const long long b = tok / seq, l = tok % seq;
  const long long page = page_table[b * max_pages + l / page_size];
  const long long dst = (page * page_size + l % page_size) * hd;
  for (long long c = threadIdx.x; c < hd; c += blockDim.x) {
    kc[dst + c] = k[tok * ld + c];
    vc[dst + c] = v[tok * ld + c];
  }
}

cudaError_t paged_kv_write_launch(const void* k, const void* v, long long ld, void* k_cache,
                                  void* v_cache, const int* page_table, long long batch,
                                  long long seq, long long hd, long long page_size,
                                  long long max_pages, cudaStream_t s) {
  if (batch * seq <= 0) return cudaSuccess;
  paged_kv_write_kernel<<<static_cast<unsigned>(batch * seq), 128, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(k), reinterpret_cast<const __nv_bfloat16*>(v), ld,
      reinterpret_cast<__nv_bfloat16*>(k_cache), reinterpret_cast<__nv_bfloat16*>(v_cache),
      page_table, seq, hd, page_size, max_pages);
  return cudaGetLastError();
}

// One CTA per (head, sequence): append this step's k/v, then softmax(q.K^T).V over the
// cached context.  Dynamic smem: D floats (q) + ctx floats (scores).
__global__ void __launch_bounds__(128)
paged_decode_attn_kernel(const __nv_bfloat16* qkv, __nv_bfloat16* kc, __nv_bfloat16* vc,
                         const int* page_table, const int* ctx_len, const int* first_valid,
                         __nv_bfloat16* out, int heads, int D, int page_size, int max_pages,
                         float scale) {
  extern __shared__ float sm[];
  __shared__ float red[4];
  const int h = blockIdx.x, b = blockIdx.y;
  const int hd = heads * D;
  const int ctx = ctx_len[b];
  const int fv = first_valid != nullptr ? first_valid[b] : 0;
  float* sq = sm;
  float* sc = sm + D;
  const int* pt = page_table + static_cast<long long>(b) * max_pages;
  const __nv_bfloat16* row = qkv + static_cast<long long>(b) * 3 * hd;
  // append k, v of the new token (position ctx-1)
  {
    const int l = ctx - 1;
    const long long dst = (static_cast<long long>(pt[l / page_size]) * page_size + l % page_size) * hd + h * D;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      kc[dst + c] = row[hd + h * D + c];
      vc[dst + c] = row[2 * hd + h * D + c];
      sq[c] = __bfloat162float(row[h * D + c]);
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int l = warp; l < ctx; l += 4) {
    float s = -INFINITY;
    if (l >= fv) {
      const __nv_bfloat16* kr = kc + (static_cast<long long>(pt[l / page_size]) * page_size + l % page_size) * hd + h * D;
      float acc = 0.0f;
      for (int c = lane; c < D; c += 32) acc += sq[c] * __bfloat162float(kr[c]);
      s = warp_sum(acc) * scale;
    }
    if (lane == 0) sc[l] = s;
    mx = fmaxf(mx, s);
  }
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.0f;
  for (int l = threadIdx.x; l < ctx; l += blockDim.x) {
    const float pr = (sc[l] == -INFINITY) ? 0.0f : __expf(sc[l] - mx);
    sc[l] = pr;
    sum += pr;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  const float tot = red[0] + red[1] + red[2] + red[3];
  const float inv = tot > 0.0f ? 1.0f / tot : 0.0f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.0f;
    for (int l = fv; l < ctx; ++l) {
      const __nv_bfloat16* vr = vc + (static_cast<long long>(pt[l / page_size]) * page_size + l % page_size) * hd + h * D;
      acc += sc[l] * __bfloat162float(vr[c]);
    }
    out[static_cast<long long>(b) * hd + h * D + c] = __float2bfloat16(acc * inv);
  }
}

cudaError_t paged_decode_attention_launch(const void* qkv, void* k_cache, void* v_cache,
                                          const int* page_table, const int* ctx_len,
                                          const int* first_valid, void* out, long long batch,
                                          long long heads, long long d, long long page_size,
                                          long long max_pages, float scale, cudaStream_t s) {
  if (batch <= 0 || heads <= 0) return cudaSuccess;
  const long long max_ctx = page_size * max_pages;
  const size_t smem = sizeof(float) * static_cast<size_t>(d + max_ctx);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(paged_decode_attn_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  dim3 grid(static_cast<unsigned>(heads), static_cast<unsigned>(batch));
  paged_decode_attn_kernel<<<grid, 128, smem, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(k_cache),
      reinterpret_cast<__nv_bfloat16*>(v_cache), page_table, ctx_len, first_valid,
      reinterpret_cast<__nv_bfloat16*>(out), static_cast<int>(heads), static_cast<int>(d),
      static_cast<int>(page_size), static_cast<int>(max_pages), scale);
  return cudaGetLastError();
}

}  // namespace vb
