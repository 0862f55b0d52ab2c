// Shape-agnostic CUDA-core GEMM with the same epilogue contract as the tcgen05 kernel.
// It exists because the drop-in module surface must also run the reference's tiny test
// configurations (hidden_size 8, head_dim 2, K = 192 ...) whose operands violate the
// TMA / UMMA alignment rules, and it doubles as the on-device cross-check of the
// tensor-core path.  No CPU fallback exists anywhere.
#include "common.cuh"
#include "gemm.h"

namespace vb {

constexpr int GT = 64;   // tile edge
constexpr int GK = 16;   // k step

struct GenParams {
  const __nv_bfloat16* a;
  const __nv_bfloat16* b;
  void* c;
  const float* bias;
  const __nv_bfloat16* residual;
  long long m, n, k, lda, ldb, ldc, ldr;
  float alpha, beta;
  long long alpha_cols, row_group;
  int epilogue, out_f32;
  const unsigned long long* drop_seed;
  unsigned long long drop_salt;
  unsigned int drop_thresh;
  float drop_scale;
};

__global__ void __launch_bounds__(256) gemm_generic_kernel(const GenParams p) {
  __shared__ float sa[GK][GT + 1];
  __shared__ float sb[GK][GT + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const long long m0 = static_cast<long long>(blockIdx.x) * GT;
  const long long n0 = static_cast<long long>(blockIdx.y) * GT;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (long long k0 = 0; k0 < p.k; k0 += GK) {
    // 64 rows x 16 k per operand, 256 threads -> 4 elements each
    for (int e = threadIdx.x; e < GT * GK; e += 256) {
      const int r = e / GK, kk = e % GK;
      const long long gk = k0 + kk;
      float va = 0.0f, vb_ = 0.0f;
      if (gk < p.k) {
        if (m0 + r < p.m) va = __bfloat162float(p.a[(m0 + r) * p.lda + gk]);
        if (n0 + r < p.n) vb_ = __bfloat162float(p.b[(n0 + r) * p.ldb + gk]);
      }
      sa[kk][r] = va;
      sb[kk][r] = vb_;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float ra[4], rb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) ra[i] = sa[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = sb[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ra[i], rb[j], acc[i][j]);
    }
    __syncthreads();
  }

  const long long ac = p.alpha_cols <= 0 ? p.n : p.alpha_cols;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long row = m0 + ty * 4 + i;
    if (row >= p.m) continue;
    long long out_row = row, res_row = row;
    if (p.row_group > 0) {
      out_row = row + row / p.row_group + 1;
      res_row = 1 + row % p.row_group;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long col = n0 + tx * 4 + j;
      if (col >= p.n) continue;
      float v = acc[i][j];
      if (p.bias != nullptr) v += p.bias[col];
      if (col < ac) v *= p.alpha;
      if (p.epilogue == VB_EPI_GELU) v = gelu_erf(v);
      else if (p.epilogue == VB_EPI_RELU) v = fmaxf(v, 0.0f);
      if (p.drop_thresh != 0u)
        v = dropout_keep(*p.drop_seed + p.drop_salt, static_cast<uint64_t>(row) * p.n + col, p.drop_thresh)
                ? v * p.drop_scale : 0.0f;
      if (p.residual != nullptr) {
        const float sv = __bfloat162float(p.residual[res_row * p.ldr + col]);
        if (p.epilogue == VB_EPI_RELU_BWD) v = sv > 0.0f ? v : 0.0f;
        else if (p.epilogue == VB_EPI_GELU_BWD) v *= gelu_erf_grad(sv);
        else v += sv;
      }
      if (p.out_f32) {
        float* c = reinterpret_cast<float*>(p.c) + out_row * p.ldc + col;
        if (p.beta != 0.0f) v += p.beta * *c;
        *c = v;
      } else {
        __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.c) + out_row * p.ldc + col;
        if (p.beta != 0.0f) v += p.beta * __bfloat162float(*c);
        *c = __float2bfloat16(v);
      }
    }
  }
}

cudaError_t gemm_generic_launch(const vb_gemm_args& a, cudaStream_t stream) {
  if (a.m <= 0 || a.n <= 0) return cudaSuccess;
  GenParams p;
  p.a = reinterpret_cast<const __nv_bfloat16*>(a.a);
  p.b = reinterpret_cast<const __nv_bfloat16*>(a.b);
  p.c = a.c;
  p.bias = a.bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a.residual);
  p.m = a.m; p.n = a.n; p.k = a.k; p.lda = a.lda; p.ldb = a.ldb; p.ldc = a.ldc; p.ldr = a.ldr;
  p.alpha = a.alpha; p.beta = a.beta; p.alpha_cols = a.alpha_cols; p.row_group = a.row_group;
  p.epilogue = a.epilogue; p.out_f32 = (a.out_dtype == VB_F32) ? 1 : 0;
  const bool drop = a.dropout_p > 0.0f && a.dropout_seed != nullptr;
  p.drop_seed = reinterpret_cast<const unsigned long long*>(a.dropout_seed);
  p.drop_salt = a.dropout_salt;
  p.drop_thresh = drop ? dropout_threshold(a.dropout_p) : 0u;
  p.drop_scale = drop ? 1.0f / (1.0f - a.dropout_p) : 1.0f;
  dim3 grid(static_cast<unsigned>((a.m + GT - 1) / GT), static_cast<unsigned>((a.n + GT - 1) / GT));
  if (grid.y > 65535) return cudaErrorInvalidValue;
  gemm_generic_kernel<<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace vb
