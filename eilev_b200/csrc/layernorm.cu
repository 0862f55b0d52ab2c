// LayerNorm forward / backward, one warp per row, bf16 I/O with fp32 statistics.
// The row is held in registers between the statistics and the normalisation pass
// (16-byte vector loads) so HBM sees one read and one write per element; a strided
// multi-pass variant covers shapes that break the vector path (tiny test configs).
#include <cstdlib>

#include "common.cuh"
#include "internal.h"

namespace vb {

constexpr int kLnWarps = 4;

struct LnParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* res;
  const float* gamma;
  const float* beta;
  __nv_bfloat16* y;
  float* mean;
  float* rstd;
  long long rows, cols, ldx, ldr, ldy;
  float eps;
};

// Sum over the threads that share a row: one warp (WPR = 1) or the whole 4-warp block (WPR = 4: wide rows with few
// of them — OPT's 976 x 2560 — leave one-warp-per-row at 6 warps per SM, each holding the row in 160 registers).
template <int WPR>
VB_DEVICE float row_sum(float v, float* red) {
  v = warp_sum(v);
  if (WPR == 1) return v;
  __syncthreads();  // red[] may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
#pragma unroll
  for (int i = 0; i < WPR; ++i) t += red[i];
  return t;
}

// VPL = 16-byte vectors per thread; cols <= VPL * 256 * WPR, cols % 8 == 0.
template <int VPL, int WPR>
__global__ void __launch_bounds__(kLnWarps * 32) ln_fwd_vec_kernel(const LnParams p) {
  static_assert(WPR == 1 || WPR == kLnWarps, "a row belongs to one warp or to the whole block");
  __shared__ float red[kLnWarps];
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const int lane = WPR == 1 ? (threadIdx.x & 31) : threadIdx.x;   // index of this thread within its row
  constexpr int kStride = 32 * WPR;
  const long long row = WPR == 1 ? static_cast<long long>(blockIdx.x) * kLnWarps + (threadIdx.x >> 5) : blockIdx.x;
  if (row >= p.rows) return;   // (WPR = 4: grid == rows, never taken)
  const int nvec = static_cast<int>(p.cols / 8);
  float v[VPL][8];
  float sum = 0.0f;
  const __nv_bfloat16* xr = p.x + row * p.ldx;
  const __nv_bfloat16* rr = p.res != nullptr ? p.res + row * p.ldr : nullptr;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * kStride;
    if (vi < nvec) {
      uint4 u = *reinterpret_cast<const uint4*>(xr + vi * 8);
      float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
             d = unpack_bf16x2(u.w);
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = b.x; v[i][3] = b.y;
      v[i][4] = c.x; v[i][5] = c.y; v[i][6] = d.x; v[i][7] = d.y;
      if (rr != nullptr) {
        uint4 w = *reinterpret_cast<const uint4*>(rr + vi * 8);
        float2 a2 = unpack_bf16x2(w.x), b2 = unpack_bf16x2(w.y), c2 = unpack_bf16x2(w.z),
               d2 = unpack_bf16x2(w.w);
        v[i][0] += a2.x; v[i][1] += a2.y; v[i][2] += b2.x; v[i][3] += b2.y;
        v[i][4] += c2.x; v[i][5] += c2.y; v[i][6] += d2.x; v[i][7] += d2.y;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.0f;
    }
  }
  const float mean = row_sum<WPR>(sum, red) / static_cast<float>(p.cols);
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (lane + i * kStride < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(row_sum<WPR>(sq, red) / static_cast<float>(p.cols) + p.eps);
  if (lane == 0) {
    if (p.mean != nullptr) p.mean[row] = mean;
    if (p.rstd != nullptr) p.rstd[row] = rstd;
  }
  __nv_bfloat16* yr = p.y + row * p.ldy;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * kStride;
    if (vi < nvec) {
      float o[8];
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + vi * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + vi * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.beta + vi * 8 + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * gg[j] + bb[j];
      uint4 u;
      u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
      u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(yr + vi * 8) = u;
    }
  }
}

__global__ void __launch_bounds__(kLnWarps * 32) ln_fwd_generic_kernel(const LnParams p) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kLnWarps + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const __nv_bfloat16* xr = p.x + row * p.ldx;
  const __nv_bfloat16* rr = p.res != nullptr ? p.res + row * p.ldr : nullptr;
  auto at = [&](long long c) {
    float v = __bfloat162float(xr[c]);
    if (rr != nullptr) v += __bfloat162float(rr[c]);
    return v;
  };
  float sum = 0.0f;
  for (long long c = lane; c < p.cols; c += 32) sum += at(c);
  const float mean = warp_sum(sum) / static_cast<float>(p.cols);
  float sq = 0.0f;
  for (long long c = lane; c < p.cols; c += 32) {
    const float d = at(c) - mean;
    sq += d * d;
  }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(p.cols) + p.eps);
  if (lane == 0) {
    if (p.mean != nullptr) p.mean[row] = mean;
    if (p.rstd != nullptr) p.rstd[row] = rstd;
  }
  for (long long c = lane; c < p.cols; c += 32)
    p.y[row * p.ldy + c] = __float2bfloat16((at(c) - mean) * rstd * p.gamma[c] + p.beta[c]);
}

template <int VPL>
static void launch_vec(const LnParams& p, cudaStream_t s) {
  const unsigned grid = static_cast<unsigned>((p.rows + kLnWarps - 1) / kLnWarps);
  launch_pdl(ln_fwd_vec_kernel<VPL, 1>, dim3(grid), dim3(kLnWarps * 32), 0, s, p);
}
template <int VPL>
static void launch_vec_block(const LnParams& p, cudaStream_t s) {   // one block per row
  launch_pdl(ln_fwd_vec_kernel<VPL, kLnWarps>, dim3(static_cast<unsigned>(p.rows)), dim3(kLnWarps * 32), 0, s, p);
}
// rows of >= 2048 columns (OPT 2560, flan-T5 2048) are split over the block's four warps
static bool ln_block_per_row(long long rows, long long cols) {
  static const bool on = [] {
    const char* e = std::getenv("VB_LN_BLOCK");   // A/B measurements: 0 = one warp per row for every shape
    return e == nullptr || e[0] != '0';
  }();
  return on && cols >= 2048 && cols <= 3 * 8 * 32 * kLnWarps && rows < (1ll << 31);
}

cudaError_t layernorm_launch(const LnParams& p, cudaStream_t s) {
  if (p.rows <= 0 || p.cols <= 0) return cudaSuccess;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool vec = p.cols % 8 == 0 && p.cols <= 12 * 256 && al(p.x) && al(p.y) && al(p.gamma) &&
                   al(p.beta) && p.ldx % 8 == 0 && p.ldy % 8 == 0 &&
                   (p.res == nullptr || (al(p.res) && p.ldr % 8 == 0));
  if (vec && ln_block_per_row(p.rows, p.cols)) {
    if (p.cols <= 2 * 8 * 32 * kLnWarps) launch_vec_block<2>(p, s);
    else launch_vec_block<3>(p, s);
  } else if (vec) {
    const int vpl = static_cast<int>((p.cols / 8 + 31) / 32);
    if (vpl <= 1) launch_vec<1>(p, s);
    else if (vpl <= 2) launch_vec<2>(p, s);
    else if (vpl <= 3) launch_vec<3>(p, s);
    else if (vpl <= 4) launch_vec<4>(p, s);
    else if (vpl <= 6) launch_vec<6>(p, s);
    else if (vpl <= 8) launch_vec<8>(p, s);
    else if (vpl <= 10) launch_vec<10>(p, s);
    else launch_vec<12>(p, s);
  } else {
    const unsigned grid = static_cast<unsigned>((p.rows + kLnWarps - 1) / kLnWarps);
    launch_pdl(ln_fwd_generic_kernel, dim3(grid), dim3(kLnWarps * 32), 0, s, p);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------ backward
struct LnBwdParams {
  const __nv_bfloat16* dy;
  const __nv_bfloat16* xin;
  const float* gamma;
  const float* mean;
  const float* rstd;
  const __nv_bfloat16* dx_add;
  __nv_bfloat16* dx;
  float* dgamma;
  float* dbeta;
  long long rows, cols;
  // optional second output: dropout(dx) with the counter-hash mask of (seed + salt, row * cols + col) — the
  // gradient entering the dgrad GEMM of the linear layer in front of a residual dropout (OPT backward)
  __nv_bfloat16* dx_drop;
  const unsigned long long* drop_seed;
  unsigned long long drop_salt;
  unsigned int drop_thresh;
  float drop_scale;
};

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma.  One warp per row.
__global__ void __launch_bounds__(kLnWarps * 32) ln_bwd_dx_kernel(const LnBwdParams p) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kLnWarps + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const __nv_bfloat16* dyr = p.dy + row * p.cols;
  const __nv_bfloat16* xr = p.xin + row * p.cols;
  const float mean = p.mean[row], rstd = p.rstd[row];
  float s1 = 0.0f, s2 = 0.0f;
  for (long long c = lane; c < p.cols; c += 32) {
    const float g = __bfloat162float(dyr[c]) * p.gamma[c];
    const float xh = (__bfloat162float(xr[c]) - mean) * rstd;
    s1 += g;
    s2 += g * xh;
  }
  s1 = warp_sum(s1) / static_cast<float>(p.cols);
  s2 = warp_sum(s2) / static_cast<float>(p.cols);
  for (long long c = lane; c < p.cols; c += 32) {
    const float g = __bfloat162float(dyr[c]) * p.gamma[c];
    const float xh = (__bfloat162float(xr[c]) - mean) * rstd;
    float d = rstd * (g - s1 - xh * s2);
    if (p.dx_add != nullptr) d += __bfloat162float(p.dx_add[row * p.cols + c]);
    const __nv_bfloat16 db = __float2bfloat16(d);
    p.dx[row * p.cols + c] = db;
    if (p.dx_drop != nullptr) {
      const bool keep = dropout_keep(*p.drop_seed + p.drop_salt, static_cast<uint64_t>(row * p.cols + c), p.drop_thresh);
      p.dx_drop[row * p.cols + c] = __float2bfloat16(keep ? __bfloat162float(db) * p.drop_scale : 0.0f);
    }
  }
}

// Vectorised variant: the row (g = dy*gamma and xhat) is held in registers between the two
// reductions and the store, 16-byte loads/stores (cols % 8 == 0, cols <= VPL*256).
template <int VPL, int WPR>
__global__ void __launch_bounds__(kLnWarps * 32) ln_bwd_dx_vec_kernel(const LnBwdParams p) {
  static_assert(WPR == 1 || WPR == kLnWarps, "a row belongs to one warp or to the whole block");
  __shared__ float red[kLnWarps];
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const int lane = WPR == 1 ? (threadIdx.x & 31) : threadIdx.x;   // index of this thread within its row
  constexpr int kStride = 32 * WPR;
  const long long row = WPR == 1 ? static_cast<long long>(blockIdx.x) * kLnWarps + (threadIdx.x >> 5) : blockIdx.x;
  if (row >= p.rows) return;   // (WPR = 4: grid == rows, never taken)
  const int nvec = static_cast<int>(p.cols / 8);
  const __nv_bfloat16* dyr = p.dy + row * p.cols;
  const __nv_bfloat16* xr = p.xin + row * p.cols;
  const float mean = p.mean[row], rstd = p.rstd[row];
  float g[VPL][8], xh[VPL][8];
  float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * kStride;
    if (vi < nvec) {
      const uint4 ud = *reinterpret_cast<const uint4*>(dyr + vi * 8);
      const uint4 ux = *reinterpret_cast<const uint4*>(xr + vi * 8);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + vi * 8 + 4));
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const uint32_t dw[4] = {ud.x, ud.y, ud.z, ud.w};
      const uint32_t xw[4] = {ux.x, ux.y, ux.z, ux.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 d2 = unpack_bf16x2(dw[j]);
        const float2 x2 = unpack_bf16x2(xw[j]);
        g[i][2 * j] = d2.x * gm[2 * j];
        g[i][2 * j + 1] = d2.y * gm[2 * j + 1];
        xh[i][2 * j] = (x2.x - mean) * rstd;
        xh[i][2 * j + 1] = (x2.y - mean) * rstd;
        s1 += g[i][2 * j] + g[i][2 * j + 1];
        s2 += g[i][2 * j] * xh[i][2 * j] + g[i][2 * j + 1] * xh[i][2 * j + 1];
      }
    }
  }
  s1 = row_sum<WPR>(s1, red) / static_cast<float>(p.cols);
  s2 = row_sum<WPR>(s2, red) / static_cast<float>(p.cols);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * kStride;
    if (vi < nvec) {
      float d[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = rstd * (g[i][j] - s1 - xh[i][j] * s2);
      if (p.dx_add != nullptr) {
        const uint4 ua = *reinterpret_cast<const uint4*>(p.dx_add + row * p.cols + vi * 8);
        const uint32_t aw[4] = {ua.x, ua.y, ua.z, ua.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a2 = unpack_bf16x2(aw[j]);
          d[2 * j] += a2.x;
          d[2 * j + 1] += a2.y;
        }
      }
      uint4 u;
      u.x = pack_bf16x2(d[0], d[1]); u.y = pack_bf16x2(d[2], d[3]);
      u.z = pack_bf16x2(d[4], d[5]); u.w = pack_bf16x2(d[6], d[7]);
      *reinterpret_cast<uint4*>(p.dx + row * p.cols + vi * 8) = u;
      if (p.dx_drop != nullptr) {  // the mask applies to the stored (bf16-rounded) gradient, as vb_dropout would
        const uint64_t seed = *p.drop_seed + p.drop_salt;
        const uint64_t base = static_cast<uint64_t>(row * p.cols + vi * 8);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(w[j]);
          o[j] = pack_bf16x2(dropout_keep(seed, base + 2 * j, p.drop_thresh) ? f.x * p.drop_scale : 0.0f,
                             dropout_keep(seed, base + 2 * j + 1, p.drop_thresh) ? f.y * p.drop_scale : 0.0f);
        }
        *reinterpret_cast<uint4*>(p.dx_drop + row * p.cols + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// dgamma[c] += sum_r dy*xhat ; dbeta[c] += sum_r dy.  Block = 32 columns x 32 row lanes;
// every column is owned by exactly one block, so the accumulation is deterministic.
__global__ void __launch_bounds__(1024) ln_bwd_param_kernel(const LnBwdParams p) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  __shared__ float sg[32][33], sb[32][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;   // 32 columns x 32 row lanes
  const long long c = static_cast<long long>(blockIdx.x) * 32 + cx;
  float ag = 0.0f, ab = 0.0f;
  if (c < p.cols) {
#pragma unroll 4
    for (long long r = ry; r < p.rows; r += 32) {
      const float dyv = __bfloat162float(p.dy[r * p.cols + c]);
      const float xh = (__bfloat162float(p.xin[r * p.cols + c]) - p.mean[r]) * p.rstd[r];
      ag += dyv * xh;
      ab += dyv;
    }
  }
  sg[ry][cx] = ag;
  sb[ry][cx] = ab;
  __syncthreads();
  if (ry == 0 && c < p.cols) {
    float tg = 0.0f, tb = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { tg += sg[i][cx]; tb += sb[i][cx]; }
    p.dgamma[c] += tg;
    p.dbeta[c] += tb;
  }
}

cudaError_t layernorm_bwd_launch(const LnBwdParams& p, cudaStream_t s) {
  if (p.rows <= 0 || p.cols <= 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((p.rows + kLnWarps - 1) / kLnWarps);
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool vec = p.cols % 8 == 0 && p.cols <= 12 * 256 && al(p.dy) && al(p.xin) && al(p.dx) &&
                   al(p.gamma) && (p.dx_add == nullptr || al(p.dx_add)) && (p.dx_drop == nullptr || al(p.dx_drop));
  const int vpl = static_cast<int>((p.cols / 8 + 31) / 32);
  const dim3 row_grid(static_cast<unsigned>(p.rows));
  if (!vec) launch_pdl(ln_bwd_dx_kernel, dim3(grid), dim3(kLnWarps * 32), 0, s, p);
  else if (ln_block_per_row(p.rows, p.cols) && p.cols <= 2 * 8 * 32 * kLnWarps)
    launch_pdl(ln_bwd_dx_vec_kernel<2, kLnWarps>, row_grid, dim3(kLnWarps * 32), 0, s, p);
  else if (ln_block_per_row(p.rows, p.cols))
    launch_pdl(ln_bwd_dx_vec_kernel<3, kLnWarps>, row_grid, dim3(kLnWarps * 32), 0, s, p);
  else if (vpl <= 1) launch_pdl(ln_bwd_dx_vec_kernel<1, 1>, dim3(grid), dim3(kLnWarps * 32), 0, s, p);
  else if (vpl <= 3) launch_pdl(ln_bwd_dx_vec_kernel<3, 1>, dim3(grid), dim3(kLnWarps * 32), 0, s, p);
  else if (vpl <= 6) launch_pdl(ln_bwd_dx_vec_kernel<6, 1>, dim3(grid), dim3(kLnWarps * 32), 0, s, p);
  else if (vpl <= 10) launch_pdl(ln_bwd_dx_vec_kernel<10, 1>, dim3(grid), dim3(kLnWarps * 32), 0, s, p);
  else launch_pdl(ln_bwd_dx_vec_kernel<12, 1>, dim3(grid), dim3(kLnWarps * 32), 0, s, p);
  if (p.dgamma != nullptr && p.dbeta != nullptr) {
    launch_pdl(ln_bwd_param_kernel, dim3(static_cast<unsigned>((p.cols + 31) / 32)), dim3(1024), 0, s, p);
  }
  return cudaGetLastError();
}


cudaError_t layernorm_fwd(const void* x, const void* residual, const float* gamma, const float* beta,
                          void* y, float* mean, float* rstd, long long rows, long long cols,
                          long long ldx, long long ldr, long long ldy, float eps, cudaStream_t s) {
  LnParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.res = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.gamma = gamma; p.beta = beta;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.mean = mean; p.rstd = rstd;
  p.rows = rows; p.cols = cols; p.ldx = ldx; p.ldr = ldr; p.ldy = ldy; p.eps = eps;
  return layernorm_launch(p, s);
}

cudaError_t layernorm_bwd(const void* dy, const void* xin, const float* gamma, const float* mean,
                          const float* rstd, const void* dx_add, void* dx, float* dgamma,
                          float* dbeta, long long rows, long long cols, void* dx_drop, float drop_p,
                          const unsigned long long* drop_seed, unsigned long long drop_salt, cudaStream_t s) {
  LnBwdParams p;
  const bool drop = dx_drop != nullptr && drop_seed != nullptr && drop_p > 0.0f;
  p.dx_drop = drop ? reinterpret_cast<__nv_bfloat16*>(dx_drop) : nullptr;
  p.drop_seed = drop_seed;
  p.drop_salt = drop_salt;
  p.drop_thresh = drop ? dropout_threshold(drop_p) : 0u;
  p.drop_scale = drop ? 1.0f / (1.0f - drop_p) : 1.0f;
  p.dy = reinterpret_cast<const __nv_bfloat16*>(dy);
  p.xin = reinterpret_cast<const __nv_bfloat16*>(xin);
  p.gamma = gamma; p.mean = mean; p.rstd = rstd;
  p.dx_add = reinterpret_cast<const __nv_bfloat16*>(dx_add);
  p.dx = reinterpret_cast<__nv_bfloat16*>(dx);
  p.dgamma = dgamma; p.dbeta = dbeta; p.rows = rows; p.cols = cols;
  return layernorm_bwd_launch(p, s);
}


// ------------------------------------------------------------------ row statistics
// stats[row] = [sum_c x[row, c], sum_c x[row, c]^2] (f32): the input of a LayerNorm folded into
// the consuming GEMM (vb_gemm_args.ln_stats) when the producer of x is not a GEMM epilogue.
// One warp per row; read-only pass (2 B per element).
__global__ void __launch_bounds__(kLnWarps * 32)
row_stats_kernel(const __nv_bfloat16* x, double* stats, long long rows, long long cols, long long ldx) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const __nv_bfloat16* xr = x + row * ldx;
  float s = 0.0f, q = 0.0f;
  if (cols % 8 == 0 && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
    for (long long v = lane; v < cols / 8; v += 32) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + v * 8));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        s += f.x + f.y;
        q = fmaf(f.x, f.x, fmaf(f.y, f.y, q));
      }
    }
  } else {
    for (long long c = lane; c < cols; c += 32) {
      const float f = __bfloat162float(xr[c]);
      s += f;
      q = fmaf(f, f, q);
    }
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane == 0) {
    stats[2 * row] = static_cast<double>(s);
    stats[2 * row + 1] = static_cast<double>(q);
  }
}

cudaError_t row_stats_launch(const void* x, double* stats, long long rows, long long cols, long long ldx,
                             cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((rows + kLnWarps - 1) / kLnWarps);
  launch_pdl(row_stats_kernel, dim3(grid), dim3(kLnWarps * 32), 0, s, reinterpret_cast<const __nv_bfloat16*>(x), stats, rows, cols, ldx);
  return cudaGetLastError();
}

}  // namespace vb
