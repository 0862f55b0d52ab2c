// Epilogue shared by the 1-CTA and 2-CTA tcgen05 GEMM kernels: one thread finishes 16
// consecutive fp32 accumulator columns of one output row (bias, alpha, GELU/ReLU, residual,
// beta*C, optional patch-embedding row remap) and stores bf16 / f32 with 16-byte accesses.
#pragma once
#include "common.cuh"

namespace vb {

struct EpiParams {
  void* c;
  const float* bias;
  const __nv_bfloat16* residual;
  long long m, n;
  long long ldc, ldr;
  float alpha, beta;
  long long alpha_cols;
  long long row_group;
  const unsigned long long* drop_seed;
  unsigned long long drop_salt;
  unsigned int drop_thresh;  // 0 = no dropout
  float drop_scale;
  int tma_store;             // 1: bf16 tile staged in smem and written with one TMA store per slab
  int epilogue;
  int out_f32;
  // LayerNorm folded into this GEMM (vb_gemm_args.ln_stats / ln_colsum): per-row [sum, sum of squares]
  // of A's rows over K columns; out = rstd * acc - rstd * mean * colsum[n] + bias[n]
  // The statistics are f64: a row's sum is assembled from up to N / 16 f32 partials by atomics in no fixed order,
  // and f64 addition of a few dozen f32 values is exact (their bits fit the 53-bit mantissa), so the result does
  // not depend on that order -- the step stays reproducible bit for bit (f32 atomics were not:
  // scripts/micro/vision_repeat.py).
  const double* ln_stats;
  const float* ln_colsum;
  float ln_inv_k, ln_eps;
  double* stats_out;         // per stored row [sum, sum of squares] of the bf16 output, f64 atomics
  double* stats_zero;        // (M, 2) buffer cleared by the tiles of the first column block
};

// acc (+) the `residual` operand: a residual add, or — for the activation-backward epilogues — the product
// with the activation's derivative at the saved forward value
VB_DEVICE float epi_combine(int epilogue, bool act_bwd, float acc, float saved) {
  if (!act_bwd) return acc + saved;
  if (epilogue == VB_EPI_RELU_BWD) return saved > 0.0f ? acc : 0.0f;
  return acc * gelu_erf_grad(saved);
}

// (rstd, -rstd * mean) of row `row` from the [sum, sum of squares] pair
VB_DEVICE float2 ln_fold_coeffs(const EpiParams& p, long long row) {
  const double2 st = __ldg(reinterpret_cast<const double2*>(p.ln_stats) + row);
  const double inv_k = static_cast<double>(p.ln_inv_k);
  const double mean_d = st.x * inv_k;
  const float mean = static_cast<float>(mean_d);
  const float var = fmaxf(static_cast<float>(fma(-mean_d, mean_d, st.y * inv_k)), 0.0f);
  const float rstd = rsqrtf(var + p.ln_eps);
  return make_float2(rstd, -rstd * mean);
}
// acc[j] <- rstd * acc[j] - rstd * mean * colsum[col0 + j];  c = ln_fold_coeffs(row)
VB_DEVICE void ln_fold_apply(const EpiParams& p, const float2 c, long long col0, float (&v)[16]) {
  if (col0 + 16 <= p.n) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 cs = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + col0 + j));
      v[j] = fmaf(v[j], c.x, c.y * cs.x);
      v[j + 1] = fmaf(v[j + 1], c.x, c.y * cs.y);
      v[j + 2] = fmaf(v[j + 2], c.x, c.y * cs.z);
      v[j + 3] = fmaf(v[j + 3], c.x, c.y * cs.w);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (col0 + j < p.n) v[j] = fmaf(v[j], c.x, c.y * __ldg(p.ln_colsum + col0 + j));
  }
}
// sum / sum of squares of the two bf16 values packed in `u` (the values as they are stored)
VB_DEVICE void stats_accum(uint32_t u, float& s, float& q) {
  const float2 f = unpack_bf16x2(u);
  s += f.x + f.y;
  q = fmaf(f.x, f.x, fmaf(f.y, f.y, q));
}

// One thread finishes 16 consecutive columns of one row.
VB_DEVICE void epilogue_row16(const EpiParams& p, long long row, long long col0,
                              const uint32_t (&acc)[16]) {
  if (row >= p.m || col0 >= p.n) return;
  long long out_row = row, res_row = row;
  if (p.row_group > 0) {
    out_row = row + row / p.row_group + 1;
    res_row = 1 + row % p.row_group;
  }
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
  const bool full = (col0 + 16 <= p.n);
  if (p.stats_zero != nullptr && col0 == 0) *reinterpret_cast<double2*>(p.stats_zero + 2 * row) = make_double2(0.0, 0.0);
  if (p.ln_stats != nullptr) ln_fold_apply(p, ln_fold_coeffs(p, row), col0, v);
  if (p.bias != nullptr) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
      for (int j = 0; j < 16; ++j)
        if (col0 + j < p.n) v[j] += __ldg(p.bias + col0 + j);
    }
  }
  if (p.alpha != 1.0f) {
    const long long ac = p.alpha_cols <= 0 ? p.n : p.alpha_cols;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (col0 + j < ac) v[j] *= p.alpha;
  }
  if (p.epilogue == VB_EPI_GELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = gelu_fast(v[j]);
  } else if (p.epilogue == VB_EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (p.drop_thresh != 0u) {
    const uint64_t seed = *p.drop_seed + p.drop_salt;
    const uint64_t base = static_cast<uint64_t>(row) * static_cast<uint64_t>(p.n) + static_cast<uint64_t>(col0);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = dropout_keep(seed, base + j, p.drop_thresh) ? v[j] * p.drop_scale : 0.0f;
  }
  if (p.residual != nullptr) {
    const __nv_bfloat16* r = p.residual + res_row * p.ldr + col0;
    const bool act_bwd = p.epilogue == VB_EPI_GELU_BWD || p.epilogue == VB_EPI_RELU_BWD;
    if (full) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(r + 8 * h));
        float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z),
               f3 = unpack_bf16x2(u.w);
        const float f[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * h + j] = epi_combine(p.epilogue, act_bwd, v[8 * h + j], f[j]);
      }
    } else {
      for (int j = 0; j < 16; ++j)
        if (col0 + j < p.n) v[j] = epi_combine(p.epilogue, act_bwd, v[j], __bfloat162float(r[j]));
    }
  }
  if (p.out_f32) {
    float* c = reinterpret_cast<float*>(p.c) + out_row * p.ldc + col0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        if (p.beta != 0.0f) {
          float4 old = *reinterpret_cast<const float4*>(c + j);
          o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
        }
        *reinterpret_cast<float4*>(c + j) = o;
      }
    } else {
      for (int j = 0; j < 16; ++j)
        if (col0 + j < p.n) c[j] = v[j] + (p.beta != 0.0f ? p.beta * c[j] : 0.0f);
    }
  } else {
    __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.c) + out_row * p.ldc + col0;
    float st_s = 0.0f, st_q = 0.0f;
    if (full) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (p.beta != 0.0f) {
          uint4 u = *reinterpret_cast<const uint4*>(c + 8 * h);
          float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z),
                 f3 = unpack_bf16x2(u.w);
          v[8 * h + 0] += p.beta * f0.x; v[8 * h + 1] += p.beta * f0.y;
          v[8 * h + 2] += p.beta * f1.x; v[8 * h + 3] += p.beta * f1.y;
          v[8 * h + 4] += p.beta * f2.x; v[8 * h + 5] += p.beta * f2.y;
          v[8 * h + 6] += p.beta * f3.x; v[8 * h + 7] += p.beta * f3.y;
        }
        uint4 o;
        o.x = pack_bf16x2(v[8 * h + 0], v[8 * h + 1]);
        o.y = pack_bf16x2(v[8 * h + 2], v[8 * h + 3]);
        o.z = pack_bf16x2(v[8 * h + 4], v[8 * h + 5]);
        o.w = pack_bf16x2(v[8 * h + 6], v[8 * h + 7]);
        *reinterpret_cast<uint4*>(c + 8 * h) = o;
        if (p.stats_out != nullptr) {
          stats_accum(o.x, st_s, st_q); stats_accum(o.y, st_s, st_q);
          stats_accum(o.z, st_s, st_q); stats_accum(o.w, st_s, st_q);
        }
      }
    } else {
      for (int j = 0; j < 16; ++j)
        if (col0 + j < p.n) {
          float o = v[j] + (p.beta != 0.0f ? p.beta * __bfloat162float(c[j]) : 0.0f);
          const __nv_bfloat16 ob = __float2bfloat16(o);
          c[j] = ob;
          const float of = __bfloat162float(ob);
          st_s += of;
          st_q = fmaf(of, of, st_q);
        }
    }
    if (p.stats_out != nullptr) {
      atomicAdd(p.stats_out + 2 * out_row, static_cast<double>(st_s));
      atomicAdd(p.stats_out + 2 * out_row + 1, static_cast<double>(st_q));
    }
  }
}



// Staged epilogue (CTA-pair kernel): 16 accumulator columns of one row -> bias / GELU|ReLU /
// dropout / residual (already sitting in the swizzled slab) -> bf16 back into the slab.
// `slab_row` points at this thread's 128-byte row of a [128 rows][64 cols] SWIZZLE_128B slab,
// `row7` = row & 7 (swizzle phase), `c16` = 16-column chunk inside the slab (0..3).
VB_DEVICE void epilogue_row16_staged(const EpiParams& p, long long row, long long col0, const uint32_t (&acc)[16],
                                     uint8_t* slab_row, int row7, int c16, bool has_res, const float2 ln_c,
                                     float& st_s, float& st_q) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
  if (p.ln_stats != nullptr) ln_fold_apply(p, ln_c, col0, v);  // ln_c: per-row coefficients, once per tile
  if (p.bias != nullptr) {
    if (col0 + 16 <= p.n) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (col0 + j < p.n) v[j] += __ldg(p.bias + col0 + j);
    }
  }
  if (p.alpha != 1.0f) {
    const long long ac = p.alpha_cols <= 0 ? p.n : p.alpha_cols;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (col0 + j < ac) v[j] *= p.alpha;
  }
  if (p.epilogue == VB_EPI_GELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = gelu_fast(v[j]);
  } else if (p.epilogue == VB_EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (p.drop_thresh != 0u) {
    const uint64_t seed = *p.drop_seed + p.drop_salt;
    const uint64_t base = static_cast<uint64_t>(row) * static_cast<uint64_t>(p.n) + static_cast<uint64_t>(col0);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = dropout_keep(seed, base + j, p.drop_thresh) ? v[j] * p.drop_scale : 0.0f;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint4* cell = reinterpret_cast<uint4*>(slab_row + (((2 * c16 + h) ^ row7) << 4));
    if (has_res) {
      const uint4 u = *cell;
      const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      const float f[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
      const bool act_bwd = p.epilogue == VB_EPI_GELU_BWD || p.epilogue == VB_EPI_RELU_BWD;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[8 * h + j] = epi_combine(p.epilogue, act_bwd, v[8 * h + j], f[j]);
    }
    uint4 o;
    o.x = pack_bf16x2(v[8 * h + 0], v[8 * h + 1]);
    o.y = pack_bf16x2(v[8 * h + 2], v[8 * h + 3]);
    o.z = pack_bf16x2(v[8 * h + 4], v[8 * h + 5]);
    o.w = pack_bf16x2(v[8 * h + 6], v[8 * h + 7]);
    *cell = o;
    if (p.stats_out != nullptr && col0 + 8 * h + 8 <= p.n) {  // n % 8 == 0: whole 8-column groups
      stats_accum(o.x, st_s, st_q); stats_accum(o.y, st_s, st_q);
      stats_accum(o.z, st_s, st_q); stats_accum(o.w, st_s, st_q);
    }
  }
}

}  // namespace vb
