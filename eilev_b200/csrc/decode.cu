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

// v2: the M activation rows are staged once per CTA in shared memory (optionally
// LayerNorm-ed on the way in, which removes the separate LN launch from the decode step);
// every warp then streams TWO weight rows with several 16-byte loads in flight per lane.
template <int M>
__global__ void __launch_bounds__(128) gemv_smem_kernel(const GemvParams p, const float* ln_g,
                                                        const float* ln_b, float ln_eps) {
  extern __shared__ __align__(16) uint8_t gsm[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(gsm);  // [M][K]
  __shared__ float red[2][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = static_cast<int>(p.k);
  const long long n0 = (static_cast<long long>(blockIdx.x) * 4 + warp) * 2;
  const bool active = n0 < p.n;
  const bool two = n0 + 1 < p.n;
  const __nv_bfloat16* w0 = p.w + (active ? n0 : 0) * p.ldw;
  const __nv_bfloat16* w1 = p.w + (two ? n0 + 1 : (active ? n0 : 0)) * p.ldw;
  // Issue the first weight loads before touching x: their DRAM latency hides the staging /
  // LayerNorm prologue below.  (K % 256 == 0 is guaranteed by the launcher.)
  constexpr int PF = 4;  // 256-element chunks prefetched per row
  uint4 pa[PF], pb[PF];
#pragma unroll
  for (int c = 0; c < PF; ++c) {
    const int k0 = lane * 8 + c * 256;
    if (k0 < K) {
      pa[c] = __ldg(reinterpret_cast<const uint4*>(w0 + k0));
      pb[c] = __ldg(reinterpret_cast<const uint4*>(w1 + k0));
    } else {
      pa[c] = make_uint4(0, 0, 0, 0);
      pb[c] = make_uint4(0, 0, 0, 0);
    }
  }
  for (int m = 0; m < M; ++m) {
    if (m >= p.m) break;
    const __nv_bfloat16* xr = p.x + m * p.ldx;
    if (ln_g == nullptr) {
      for (int c = threadIdx.x * 8; c < K; c += 128 * 8)
        *reinterpret_cast<uint4*>(xs + m * K + c) = *reinterpret_cast<const uint4*>(xr + c);
    } else {
      float s1 = 0.0f;
      for (int c = threadIdx.x; c < K; c += 128) s1 += __bfloat162float(xr[c]);
      s1 = warp_sum(s1);
      if (lane == 0) red[0][warp] = s1;
      __syncthreads();
      const float mean = (red[0][0] + red[0][1] + red[0][2] + red[0][3]) / static_cast<float>(K);
      float s2 = 0.0f;
      for (int c = threadIdx.x; c < K; c += 128) {
        const float d = __bfloat162float(xr[c]) - mean;
        s2 += d * d;
      }
      s2 = warp_sum(s2);
      if (lane == 0) red[1][warp] = s2;
      __syncthreads();
      const float rstd = rsqrtf((red[1][0] + red[1][1] + red[1][2] + red[1][3]) / static_cast<float>(K) + ln_eps);
      for (int c = threadIdx.x; c < K; c += 128)
        xs[m * K + c] = __float2bfloat16((__bfloat162float(xr[c]) - mean) * rstd * ln_g[c] + ln_b[c]);
      __syncthreads();
    }
  }
  __syncthreads();
  if (!active) return;
  float acc0[M], acc1[M];
#pragma unroll
  for (int i = 0; i < M; ++i) { acc0[i] = 0.0f; acc1[i] = 0.0f; }
  auto fma8 = [&](const uint4& wv, const uint4& xv, float& acc) {
    const float2 a0 = unpack_bf16x2(wv.x), a1 = unpack_bf16x2(wv.y), a2 = unpack_bf16x2(wv.z), a3 = unpack_bf16x2(wv.w);
    const float2 b0 = unpack_bf16x2(xv.x), b1 = unpack_bf16x2(xv.y), b2 = unpack_bf16x2(xv.z), b3 = unpack_bf16x2(xv.w);
    acc += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
  };
  // software pipeline: consume the PF chunks in registers while the next PF are in flight
  for (int kb = 0; kb < K; kb += PF * 256) {
    uint4 na[PF], nb[PF];
#pragma unroll
    for (int c = 0; c < PF; ++c) {
      const int k1 = kb + PF * 256 + lane * 8 + c * 256;
      if (k1 < K) {
        na[c] = __ldg(reinterpret_cast<const uint4*>(w0 + k1));
        nb[c] = __ldg(reinterpret_cast<const uint4*>(w1 + k1));
      } else {
        na[c] = make_uint4(0, 0, 0, 0);
        nb[c] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int c = 0; c < PF; ++c) {
      const int k0 = kb + lane * 8 + c * 256;
      if (k0 < K) {
#pragma unroll
        for (int i = 0; i < M; ++i) {
          if (i < p.m) {
            const uint4 x0 = *reinterpret_cast<const uint4*>(xs + i * K + k0);
            fma8(pa[c], x0, acc0[i]);
            fma8(pb[c], x0, acc1[i]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < PF; ++c) { pa[c] = na[c]; pb[c] = nb[c]; }
  }
#pragma unroll
  for (int i = 0; i < M; ++i) { acc0[i] = warp_sum(acc0[i]); acc1[i] = warp_sum(acc1[i]); }
  if (lane == 0) {
    const long long ac = p.alpha_cols <= 0 ? p.n : p.alpha_cols;
    for (int r = 0; r < (two ? 2 : 1); ++r) {
      const long long n = n0 + r;
      for (int i = 0; i < M; ++i) {
        if (i >= p.m) break;
        float v = r == 0 ? acc0[i] : acc1[i];
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
}

template <int M>
static cudaError_t launch_gemv_smem(const GemvParams& p, const float* g, const float* b, float eps,
                                    cudaStream_t s) {
  const size_t smem = static_cast<size_t>(M) * p.k * 2;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(gemv_smem_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return e;
    attr = 200 * 1024;
  }
  const unsigned grid = static_cast<unsigned>((p.n + 7) / 8);
  gemv_smem_kernel<M><<<grid, 128, smem, s>>>(p, g, b, eps);
  return cudaGetLastError();
}

cudaError_t gemv_launch(const void* x, const void* w, const float* bias, const void* residual,
                        void* y, long long m, long long n, long long k, long long ldx,
                        long long ldw, long long ldy, long long ldr, float alpha,
                        long long alpha_cols, int epilogue, int out_dtype, const float* ln_gamma,
                        const float* ln_beta, float ln_eps, cudaStream_t s) {
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
  const int mt = m <= 1 ? 1 : (m <= 2 ? 2 : (m <= 4 ? 4 : (m <= 8 ? 8 : 16)));
  const bool smem_ok = p.vec && k % 256 == 0 && static_cast<size_t>(mt) * k * 2 <= 160 * 1024;
  if (smem_ok) {
    switch (mt) {
      case 1: return launch_gemv_smem<1>(p, ln_gamma, ln_beta, ln_eps, s);
      case 2: return launch_gemv_smem<2>(p, ln_gamma, ln_beta, ln_eps, s);
      case 4: return launch_gemv_smem<4>(p, ln_gamma, ln_beta, ln_eps, s);
      case 8: return launch_gemv_smem<8>(p, ln_gamma, ln_beta, ln_eps, s);
      default: return launch_gemv_smem<16>(p, ln_gamma, ln_beta, ln_eps, s);
    }
  }
  if (ln_gamma != nullptr) return cudaErrorInvalidValue;  // LN fusion needs the staged path
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

// Flash-decoding: grid (head, sequence, split).  Every CTA appends nothing but its own range
// of the cached context ([lo, hi) of ctx tokens; the CTA owning the last position first
// appends this step's k/v), computes a partial softmax(q.K^T).V relative to its local max
// and parks (max, sum, acc[D]) in the workspace; the last CTA to finish a (head, sequence)
// merges the splits.  Phase 1: one thread per cached token (16-byte loads of its K row);
// phase 2: thread = (token group, 8-wide d vector).
__global__ void __launch_bounds__(128)
paged_decode_attn_kernel(const __nv_bfloat16* qkv, __nv_bfloat16* kc, __nv_bfloat16* vc,
                         const int* page_table, const int* ctx_len, const int* first_valid,
                         __nv_bfloat16* out, float* ws, int* counters, int heads, int D,
                         int page_size, int max_pages, float scale, int splits, int chunk_cap) {
  extern __shared__ float sm[];
  __shared__ float red[4];
  __shared__ int s_last;
  const int h = blockIdx.x, b = blockIdx.y, sp = blockIdx.z;
  const int hd = heads * D;
  const int ctx = ctx_len[b];
  const int fv = first_valid != nullptr ? first_valid[b] : 0;
  const int chunk = (ctx + splits - 1) / splits;
  const int lo = sp * chunk;
  const int hi = lo + chunk < ctx ? lo + chunk : ctx;
  float* sq = sm;                 // D
  float* sc = sm + D;             // chunk_cap scores
  float* part = sc + chunk_cap;   // groups * D
  const int* pt = page_table + static_cast<long long>(b) * max_pages;
  const __nv_bfloat16* row = qkv + static_cast<long long>(b) * 3 * hd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto tok_off = [&](int l) {
    return (static_cast<long long>(pt[l / page_size]) * page_size + l % page_size) * hd + h * D;
  };
  const bool owns_new = (ctx - 1 >= lo && ctx - 1 < hi);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    if (owns_new) {
      const long long dst = tok_off(ctx - 1);
      kc[dst + c] = row[hd + h * D + c];
      vc[dst + c] = row[2 * hd + h * D + c];
    }
    sq[c] = __bfloat162float(row[h * D + c]) * scale;
  }
  __syncthreads();
  const bool vec = (D % 8 == 0) && (hd % 8 == 0);
  float mx = -INFINITY;
  for (int l = lo + threadIdx.x; l < hi; l += blockDim.x) {
    float s = -INFINITY;
    if (l >= fv) {
      const __nv_bfloat16* kr = kc + tok_off(l);
      float acc = 0.0f;
      if (vec) {
#pragma unroll 5
        for (int c = 0; c < D; c += 8) {
          const uint4 u = *reinterpret_cast<const uint4*>(kr + c);
          const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
          acc += sq[c] * a0.x + sq[c + 1] * a0.y + sq[c + 2] * a1.x + sq[c + 3] * a1.y + sq[c + 4] * a2.x +
                 sq[c + 5] * a2.y + sq[c + 6] * a3.x + sq[c + 7] * a3.y;
        }
      } else {
        for (int c = 0; c < D; ++c) acc += sq[c] * __bfloat162float(kr[c]);
      }
      s = acc;
    }
    sc[l - lo] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.0f;
  for (int l = lo + threadIdx.x; l < hi; l += blockDim.x) {
    const float pr = (sc[l - lo] == -INFINITY) ? 0.0f : __expf(sc[l - lo] - mx);
    sc[l - lo] = pr;
    sum += pr;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  const float tot = red[0] + red[1] + red[2] + red[3];
  // phase 2
  const int nvec = (D + 7) / 8;
  int groups = static_cast<int>(blockDim.x) / nvec;
  if (groups > 16) groups = 16;
  const int gidx = threadIdx.x / nvec, vi = threadIdx.x % nvec;
  if (gidx < groups) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int start = lo > fv ? lo : fv;
    for (int l = start + gidx; l < hi; l += groups) {
      const float pr = sc[l - lo];
      const __nv_bfloat16* vr = vc + tok_off(l) + vi * 8;
      if (vec) {
        const uint4 u = *reinterpret_cast<const uint4*>(vr);
        const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
        acc[0] += pr * a0.x; acc[1] += pr * a0.y; acc[2] += pr * a1.x; acc[3] += pr * a1.y;
        acc[4] += pr * a2.x; acc[5] += pr * a2.y; acc[6] += pr * a3.x; acc[7] += pr * a3.y;
      } else {
        for (int j = 0; j < 8; ++j)
          if (vi * 8 + j < D) acc[j] += pr * __bfloat162float(vr[j]);
      }
    }
    for (int j = 0; j < 8; ++j)
      if (vi * 8 + j < D) part[gidx * D + vi * 8 + j] = acc[j];
  }
  __syncthreads();
  float* my = ws + ((static_cast<long long>(b) * heads + h) * splits + sp) * (D + 2);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.0f;
    for (int gi = 0; gi < groups; ++gi) a += part[gi * D + c];
    my[2 + c] = a;
  }
  if (threadIdx.x == 0) {
    my[0] = mx;
    my[1] = tot;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(&counters[b * heads + h], 1);
    s_last = (ticket == splits - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* base = ws + (static_cast<long long>(b) * heads + h) * splits * (D + 2);
  float gm = -INFINITY;
  for (int i = 0; i < splits; ++i) gm = fmaxf(gm, base[i * (D + 2)]);
  float denom = 0.0f;
  for (int i = 0; i < splits; ++i) {
    const float mi = base[i * (D + 2)];
    denom += (mi == -INFINITY) ? 0.0f : __expf(mi - gm) * base[i * (D + 2) + 1];
  }
  const float inv = denom > 0.0f ? 1.0f / denom : 0.0f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.0f;
    for (int i = 0; i < splits; ++i) {
      const float mi = base[i * (D + 2)];
      if (mi != -INFINITY) a += __expf(mi - gm) * base[i * (D + 2) + 2 + c];
    }
    out[static_cast<long long>(b) * hd + h * D + c] = __float2bfloat16(a * inv);
  }
  if (threadIdx.x == 0) counters[b * heads + h] = 0;  // ready for the next step / graph replay
}

cudaError_t paged_decode_attention_launch(const void* qkv, void* k_cache, void* v_cache,
                                          const int* page_table, const int* ctx_len,
                                          const int* first_valid, void* out, float* workspace,
                                          int* counters, long long splits, long long batch,
                                          long long heads, long long d, long long page_size,
                                          long long max_pages, float scale, cudaStream_t s) {
  if (batch <= 0 || heads <= 0) return cudaSuccess;
  if (splits <= 0 || splits > 64 || workspace == nullptr || counters == nullptr) return cudaErrorInvalidValue;
  const long long max_ctx = page_size * max_pages;
  const long long chunk_cap = (max_ctx + splits - 1) / splits;
  const size_t smem = sizeof(float) * static_cast<size_t>(d + chunk_cap + 16 * d);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(paged_decode_attn_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr = 200 * 1024;
  }
  dim3 grid(static_cast<unsigned>(heads), static_cast<unsigned>(batch), static_cast<unsigned>(splits));
  paged_decode_attn_kernel<<<grid, 128, smem, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(k_cache),
      reinterpret_cast<__nv_bfloat16*>(v_cache), page_table, ctx_len, first_valid,
      reinterpret_cast<__nv_bfloat16*>(out), workspace, counters, static_cast<int>(heads),
      static_cast<int>(d), static_cast<int>(page_size), static_cast<int>(max_pages), scale,
      static_cast<int>(splits), static_cast<int>(chunk_cap));
  return cudaGetLastError();
}

}  // namespace vb
