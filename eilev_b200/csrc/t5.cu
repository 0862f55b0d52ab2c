// flan-T5 language-model kernels (eilev/model/v2.py:228-238 -> HF:t5/modeling_t5.py): RMSNorm
// (T5LayerNorm) forward / backward, the gated tanh-GELU of T5DenseGatedActDense forward /
// backward, and the decoder-side embedding gather.  All HBM-bound: one read + one write per
// element, fp32 statistics, bf16 storage.  The relative-position bias lives in attention.cu.
#include "common.cuh"
#include "internal.h"

namespace vb {

// One warp per row; the row is read twice (statistics, then normalise) — the second read hits L1/L2.
__global__ void __launch_bounds__(256)
rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                   __nv_bfloat16* __restrict__ y, float* __restrict__ rstd_out, long long rows, int cols,
                   long long ldx, long long ldy, float eps, int vec) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __nv_bfloat16* xr = x + row * ldx;
  __nv_bfloat16* yr = y + row * ldy;
  float ss = 0.0f;
  if (vec) {
    for (int c = lane * 8; c < cols; c += 256) {
      const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
      const float2 a0 = unpack_bf16x2(v.x), a1 = unpack_bf16x2(v.y), a2 = unpack_bf16x2(v.z), a3 = unpack_bf16x2(v.w);
      ss += a0.x * a0.x + a0.y * a0.y + a1.x * a1.x + a1.y * a1.y + a2.x * a2.x + a2.y * a2.y + a3.x * a3.x + a3.y * a3.y;
    }
  } else {
    for (int c = lane; c < cols; c += 32) {
      const float v = __bfloat162float(xr[c]);
      ss += v * v;
    }
  }
  const float r = rsqrtf(warp_sum(ss) / static_cast<float>(cols) + eps);
  if (lane == 0 && rstd_out != nullptr) rstd_out[row] = r;
  if (vec) {
    for (int c = lane * 8; c < cols; c += 256) {
      const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
      const float2 a0 = unpack_bf16x2(v.x), a1 = unpack_bf16x2(v.y), a2 = unpack_bf16x2(v.z), a3 = unpack_bf16x2(v.w);
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
      uint4 o;
      o.x = pack_bf16x2(a0.x * r * g0.x, a0.y * r * g0.y);
      o.y = pack_bf16x2(a1.x * r * g0.z, a1.y * r * g0.w);
      o.z = pack_bf16x2(a2.x * r * g1.x, a2.y * r * g1.y);
      o.w = pack_bf16x2(a3.x * r * g1.z, a3.y * r * g1.w);
      *reinterpret_cast<uint4*>(yr + c) = o;
    }
  } else {
    for (int c = lane; c < cols; c += 32) yr[c] = __float2bfloat16(__bfloat162float(xr[c]) * r * gamma[c]);
  }
}

cudaError_t rmsnorm_fwd_launch(const void* x, const float* gamma, void* y, float* rstd, long long rows,
                               long long cols, long long ldx, long long ldy, float eps, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  if (x == nullptr || gamma == nullptr || y == nullptr) return cudaErrorInvalidValue;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const int vec = (cols % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && al(x) && al(y) && al(gamma)) ? 1 : 0;
  launch_pdl(rmsnorm_fwd_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, s, 
      reinterpret_cast<const __nv_bfloat16*>(x), gamma, reinterpret_cast<__nv_bfloat16*>(y), rstd, rows,
      static_cast<int>(cols), ldx, ldy, eps, vec);
  return cudaGetLastError();
}

// dx_j = r * (g_j - x_j * r^2 * mean_i(g_i x_i)),  g = gamma * dy   (+ dx_add).  One warp per
// row, 16-byte accesses when the row length is a multiple of 8.
__global__ void __launch_bounds__(256)
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                   const float* __restrict__ gamma, const float* __restrict__ rstd,
                   const __nv_bfloat16* __restrict__ dx_add, __nv_bfloat16* __restrict__ dx, long long rows,
                   int cols, int vec) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __nv_bfloat16* dyr = dy + row * cols;
  const __nv_bfloat16* xr = x + row * cols;
  float dot = 0.0f;
  if (vec) {
    for (int c = lane * 8; c < cols; c += 256) {
      const uint4 d = *reinterpret_cast<const uint4*>(dyr + c);
      const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
      const float2 d0 = unpack_bf16x2(d.x), d1 = unpack_bf16x2(d.y), d2 = unpack_bf16x2(d.z), d3 = unpack_bf16x2(d.w);
      const float2 x0 = unpack_bf16x2(v.x), x1 = unpack_bf16x2(v.y), x2 = unpack_bf16x2(v.z), x3 = unpack_bf16x2(v.w);
      dot += g0.x * d0.x * x0.x + g0.y * d0.y * x0.y + g0.z * d1.x * x1.x + g0.w * d1.y * x1.y +
             g1.x * d2.x * x2.x + g1.y * d2.y * x2.y + g1.z * d3.x * x3.x + g1.w * d3.y * x3.y;
    }
  } else {
    for (int c = lane; c < cols; c += 32)
      dot += gamma[c] * __bfloat162float(dyr[c]) * __bfloat162float(xr[c]);
  }
  const float r = rstd[row];
  const float k = warp_sum(dot) / static_cast<float>(cols) * r * r;
  if (vec) {
    for (int c = lane * 8; c < cols; c += 256) {
      const uint4 d = *reinterpret_cast<const uint4*>(dyr + c);
      const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
      const float2 d0 = unpack_bf16x2(d.x), d1 = unpack_bf16x2(d.y), d2 = unpack_bf16x2(d.z), d3 = unpack_bf16x2(d.w);
      const float2 x0 = unpack_bf16x2(v.x), x1 = unpack_bf16x2(v.y), x2 = unpack_bf16x2(v.z), x3 = unpack_bf16x2(v.w);
      float o[8] = {r * (g0.x * d0.x - x0.x * k), r * (g0.y * d0.y - x0.y * k), r * (g0.z * d1.x - x1.x * k),
                    r * (g0.w * d1.y - x1.y * k), r * (g1.x * d2.x - x2.x * k), r * (g1.y * d2.y - x2.y * k),
                    r * (g1.z * d3.x - x3.x * k), r * (g1.w * d3.y - x3.y * k)};
      if (dx_add != nullptr) {
        const uint4 a = *reinterpret_cast<const uint4*>(dx_add + row * cols + c);
        const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
        o[0] += a0.x; o[1] += a0.y; o[2] += a1.x; o[3] += a1.y; o[4] += a2.x; o[5] += a2.y; o[6] += a3.x; o[7] += a3.y;
      }
      uint4 w;
      w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]); w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(dx + row * cols + c) = w;
    }
  } else {
    for (int c = lane; c < cols; c += 32) {
      float v = r * (gamma[c] * __bfloat162float(dyr[c]) - __bfloat162float(xr[c]) * k);
      if (dx_add != nullptr) v += __bfloat162float(dx_add[row * cols + c]);
      dx[row * cols + c] = __float2bfloat16(v);
    }
  }
}

cudaError_t rmsnorm_bwd_launch(const void* dy, const void* x, const float* gamma, const float* rstd,
                               const void* dx_add, void* dx, long long rows, long long cols, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  if (dy == nullptr || x == nullptr || gamma == nullptr || rstd == nullptr || dx == nullptr)
    return cudaErrorInvalidValue;
  auto al = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const int vec = (cols % 8 == 0 && al(dy) && al(x) && al(gamma) && al(dx_add) && al(dx)) ? 1 : 0;
  launch_pdl(rmsnorm_bwd_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, s, 
      reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x), gamma, rstd,
      reinterpret_cast<const __nv_bfloat16*>(dx_add), reinterpret_cast<__nv_bfloat16*>(dx), rows,
      static_cast<int>(cols), vec);
  return cudaGetLastError();
}

// gelu_new(x) = 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))   (HF:activations.py NewGELUActivation)
VB_DEVICE float gelu_tanh(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}
VB_DEVICE float gelu_tanh_grad(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  const float th = tanhf(u);
  const float du = 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * x * x);
  return 0.5f * (1.0f + th) + 0.5f * x * (1.0f - th * th) * du;
}

// vec: dff % 8 == 0 and 16-byte aligned buffers -> one thread handles 8 consecutive columns
__global__ void __launch_bounds__(256)
gated_gelu_fwd_kernel(const __nv_bfloat16* __restrict__ h01, __nv_bfloat16* __restrict__ out, long long rows,
                      long long dff, int vec) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long step = vec ? 8 : 1;
  const long long total = rows * dff / step;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = i * step;
    const long long r = e / dff, c = e - r * dff;
    if (vec) {
      const uint4 a = *reinterpret_cast<const uint4*>(h01 + r * 2 * dff + c);
      const uint4 b = *reinterpret_cast<const uint4*>(h01 + r * 2 * dff + dff + c);
      const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
      const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
      uint4 o;
      o.x = pack_bf16x2(gelu_tanh(a0.x) * b0.x, gelu_tanh(a0.y) * b0.y);
      o.y = pack_bf16x2(gelu_tanh(a1.x) * b1.x, gelu_tanh(a1.y) * b1.y);
      o.z = pack_bf16x2(gelu_tanh(a2.x) * b2.x, gelu_tanh(a2.y) * b2.y);
      o.w = pack_bf16x2(gelu_tanh(a3.x) * b3.x, gelu_tanh(a3.y) * b3.y);
      *reinterpret_cast<uint4*>(out + e) = o;
    } else {
      const float h0 = __bfloat162float(h01[r * 2 * dff + c]);
      const float h1 = __bfloat162float(h01[r * 2 * dff + dff + c]);
      out[e] = __float2bfloat16(gelu_tanh(h0) * h1);
    }
  }
}

__global__ void __launch_bounds__(256)
gated_gelu_bwd_kernel(const __nv_bfloat16* __restrict__ d_out, const __nv_bfloat16* __restrict__ h01,
                      __nv_bfloat16* __restrict__ d_h01, long long rows, long long dff, int vec) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long step = vec ? 8 : 1;
  const long long total = rows * dff / step;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = i * step;
    const long long r = e / dff, c = e - r * dff;
    if (vec) {
      const uint4 a = *reinterpret_cast<const uint4*>(h01 + r * 2 * dff + c);
      const uint4 b = *reinterpret_cast<const uint4*>(h01 + r * 2 * dff + dff + c);
      const uint4 gq = *reinterpret_cast<const uint4*>(d_out + e);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, gw[4] = {gq.x, gq.y, gq.z, gq.w};
      uint32_t o0[4], o1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 h0 = unpack_bf16x2(aw[j]), h1 = unpack_bf16x2(bw[j]), g = unpack_bf16x2(gw[j]);
        o0[j] = pack_bf16x2(g.x * h1.x * gelu_tanh_grad(h0.x), g.y * h1.y * gelu_tanh_grad(h0.y));
        o1[j] = pack_bf16x2(g.x * gelu_tanh(h0.x), g.y * gelu_tanh(h0.y));
      }
      *reinterpret_cast<uint4*>(d_h01 + r * 2 * dff + c) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
      *reinterpret_cast<uint4*>(d_h01 + r * 2 * dff + dff + c) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    } else {
      const float h0 = __bfloat162float(h01[r * 2 * dff + c]);
      const float h1 = __bfloat162float(h01[r * 2 * dff + dff + c]);
      const float g = __bfloat162float(d_out[e]);
      d_h01[r * 2 * dff + c] = __float2bfloat16(g * h1 * gelu_tanh_grad(h0));
      d_h01[r * 2 * dff + dff + c] = __float2bfloat16(g * gelu_tanh(h0));
    }
  }
}

static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }
static unsigned ew_grid(long long total) {
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  return static_cast<unsigned>(blocks);
}

cudaError_t gated_gelu_fwd_launch(const void* h01, void* out, long long rows, long long dff, cudaStream_t s) {
  if (rows * dff <= 0) return cudaSuccess;
  const int vec = (dff % 8 == 0 && al16(h01) && al16(out)) ? 1 : 0;
  launch_pdl(gated_gelu_fwd_kernel, dim3(ew_grid(rows * dff / (vec ? 8 : 1))), dim3(256), 0, s, 
      reinterpret_cast<const __nv_bfloat16*>(h01), reinterpret_cast<__nv_bfloat16*>(out), rows, dff, vec);
  return cudaGetLastError();
}

cudaError_t gated_gelu_bwd_launch(const void* d_out, const void* h01, void* d_h01, long long rows,
                                  long long dff, cudaStream_t s) {
  if (rows * dff <= 0) return cudaSuccess;
  const int vec = (dff % 8 == 0 && al16(d_out) && al16(h01) && al16(d_h01)) ? 1 : 0;
  launch_pdl(gated_gelu_bwd_kernel, dim3(ew_grid(rows * dff / (vec ? 8 : 1))), dim3(256), 0, s, 
      reinterpret_cast<const __nv_bfloat16*>(d_out), reinterpret_cast<const __nv_bfloat16*>(h01),
      reinterpret_cast<__nv_bfloat16*>(d_h01), rows, dff, vec);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(128)
embedding_kernel(const long long* ids, const __nv_bfloat16* __restrict__ table, __nv_bfloat16* __restrict__ out,
                 long long dim, long long vocab) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long i = blockIdx.x;
  long long id = ids[i];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  for (long long c = threadIdx.x; c < dim; c += blockDim.x) out[i * dim + c] = table[id * dim + c];
}

cudaError_t embedding_launch(const long long* ids, const void* table, void* out, long long n, long long dim,
                             long long vocab, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  launch_pdl(embedding_kernel, dim3(static_cast<unsigned>(n)), dim3(128), 0, s, ids, reinterpret_cast<const __nv_bfloat16*>(table),
                                                            reinterpret_cast<__nv_bfloat16*>(out), dim, vocab);
  return cudaGetLastError();
}

}  // namespace vb
