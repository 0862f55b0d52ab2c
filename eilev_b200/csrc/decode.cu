// Token-by-token generation kernels (HBM-bound).
//   bytes per decode step ~ all LM weights (5.29 GB for OPT-2.7B in bf16) + the KV pages
//   touched, so the whole design is about keeping the weight stream running:
//   * a tensor-core GEMV (mma.sync m16n8k16; weights are the 16-row A operand read straight
//     from global memory with 16-byte loads, the <= 16 rows of x are the B operand) whose
//     instruction cost per KB of weights is ~20x below a CUDA-core dot product, so a batch
//     of 8 sequences costs what a single one costs;
//   * single-query attention over a paged KV cache (flash-decoding splits);
//   * ONE persistent cooperative kernel per generated token (`decode_step_kernel`): it
//     walks a device-resident op list (embed, 32 x [qkv, attention, out, fc1, fc2], head)
//     with grid barriers in between, and every warp issues the first weight loads of the
//     NEXT projection before it enters the barrier — the HBM pipe stays busy across op
//     boundaries instead of draining at ~160 kernel boundaries per token.
#include <cstdlib>

#include "common.cuh"
#include "internal.h"

namespace vb {

constexpr int kGemvMaxM = 16;

// ------------------------------------------------------------------ op accessors
// A projection / attention / embed step is described by a vb_decode_op (C ABI, see the
// header): the same record drives the stand-alone launches and the persistent kernel.
struct GemvView {
  const vb_decode_op& o;
  VB_DEVICE explicit GemvView(const vb_decode_op& op) : o(op) {}
  VB_DEVICE const __nv_bfloat16* w() const { return reinterpret_cast<const __nv_bfloat16*>(o.ptr[0]); }
  VB_DEVICE const float* bias() const { return reinterpret_cast<const float*>(o.ptr[1]); }
  VB_DEVICE const __nv_bfloat16* residual() const { return reinterpret_cast<const __nv_bfloat16*>(o.ptr[2]); }
  VB_DEVICE const __nv_bfloat16* x() const { return reinterpret_cast<const __nv_bfloat16*>(o.ptr[3]); }
  VB_DEVICE void* y() const { return const_cast<void*>(o.ptr[4]); }
  VB_DEVICE const float* ln_g() const { return reinterpret_cast<const float*>(o.ptr[5]); }
  VB_DEVICE const float* ln_b() const { return reinterpret_cast<const float*>(o.ptr[6]); }
  VB_DEVICE long long n() const { return o.i64[0]; }
  VB_DEVICE int k() const { return static_cast<int>(o.i64[1]); }
  VB_DEVICE long long ldw() const { return o.i64[2]; }
  VB_DEVICE long long ldx() const { return o.i64[3]; }
  VB_DEVICE long long ldy() const { return o.i64[4]; }
  VB_DEVICE long long ldr() const { return o.i64[5]; }
  VB_DEVICE long long alpha_cols() const { return o.i64[6]; }
  VB_DEVICE float alpha() const { return o.f32[0]; }
  VB_DEVICE float ln_eps() const { return o.f32[1]; }
  VB_DEVICE int epilogue() const { return o.i32[0]; }
  VB_DEVICE int out_f32() const { return o.i32[1]; }
};

// bf16 load that bypasses L1 (the value was written by another SM a moment ago)
VB_DEVICE float ldcg_bf16(const __nv_bfloat16* p) {
  return __bfloat162float(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}

VB_DEVICE void fence_gpu() { asm volatile("fence.acq_rel.gpu;\n" ::: "memory"); }
VB_DEVICE unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ------------------------------------------------------------------ legacy GEMV
// One warp per output feature: any shape / alignment (the reference's tiny test configs).
template <int M>
__global__ void __launch_bounds__(128) gemv_kernel(const vb_decode_op op, int m, int vec) {
  const GemvView p(op);
  const int lane = threadIdx.x & 31;
  const long long n = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (n >= p.n()) return;
  const int K = p.k();
  float acc[M];
#pragma unroll
  for (int i = 0; i < M; ++i) acc[i] = 0.0f;
  const __nv_bfloat16* wr = p.w() + n * p.ldw();
  if (vec) {
    for (int k0 = lane * 8; k0 < K; k0 += 256) {
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wr + k0));
      const float2 w0 = unpack_bf16x2(wv.x), w1 = unpack_bf16x2(wv.y), w2 = unpack_bf16x2(wv.z),
                   w3 = unpack_bf16x2(wv.w);
#pragma unroll
      for (int i = 0; i < M; ++i) {
        if (i < m) {
          const uint4 xv = *reinterpret_cast<const uint4*>(p.x() + i * p.ldx() + k0);
          const float2 x0 = unpack_bf16x2(xv.x), x1 = unpack_bf16x2(xv.y), x2 = unpack_bf16x2(xv.z),
                       x3 = unpack_bf16x2(xv.w);
          acc[i] += w0.x * x0.x + w0.y * x0.y + w1.x * x1.x + w1.y * x1.y + w2.x * x2.x +
                    w2.y * x2.y + w3.x * x3.x + w3.y * x3.y;
        }
      }
    }
  } else {
    for (int k0 = lane; k0 < K; k0 += 32) {
      const float wv = __bfloat162float(wr[k0]);
#pragma unroll
      for (int i = 0; i < M; ++i)
        if (i < m) acc[i] += wv * __bfloat162float(p.x()[i * p.ldx() + k0]);
    }
  }
#pragma unroll
  for (int i = 0; i < M; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
    const long long ac = p.alpha_cols() <= 0 ? p.n() : p.alpha_cols();
    for (int i = 0; i < M; ++i) {
      if (i >= m) break;
      float v = acc[i];
      if (p.bias() != nullptr) v += p.bias()[n];
      if (n < ac) v *= p.alpha();
      if (p.epilogue() == VB_EPI_GELU) v = gelu_erf(v);
      else if (p.epilogue() == VB_EPI_RELU) v = fmaxf(v, 0.0f);
      if (p.residual() != nullptr) v += __bfloat162float(p.residual()[i * p.ldr() + n]);
      if (p.out_f32()) reinterpret_cast<float*>(p.y())[i * p.ldy() + n] = v;
      else reinterpret_cast<__nv_bfloat16*>(p.y())[i * p.ldy() + n] = __float2bfloat16(v);
    }
  }
}

// ------------------------------------------------------------------ tensor-core GEMV
//   * A fragments come straight from global memory: the contraction index is order-free, so
//     thread (g, t) of a warp takes two 16-byte pieces (columns 8t.. and 32+8t.. of the
//     64-column step) of weight rows g and g+8 and feeds the element pairs (4j, 4j+1 | 4j+2,
//     4j+3) of its 16 values to the k-slots (2t, 2t+1 | 2t+8, 2t+9) of MMA j = 0..3; the x
//     fragments use the same permutation (thread (g, t) reads x[g][same 16 columns] from
//     shared memory), so no shuffles and no ldmatrix, and every load instruction covers 64
//     contiguous bytes per row.
//   * one k-step = 16 rows x 64 columns = 2 KB per warp; kMD steps are in flight per warp.
//   * a CTA owns a balanced contiguous range of 16-row blocks; its (block, k-step) units are
//     dealt round-robin to the NW warps (see GemvGeom), so N = 2560 still keeps every warp of
//     every SM loading; the per-warp partial 16 x 8 tiles meet in shared memory and are summed
//     in warp order (deterministic).
constexpr int kXPad = 32;  // x rows are K + 32 elements apart in shared memory (bank spread)
constexpr int kMD = 4;  // k-steps in flight per warp, stand-alone kernel (2 CTAs / SM)
constexpr int kPD = 8;  // same, persistent kernel (1 CTA / SM, 255 registers available)

VB_DEVICE void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                              uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct WFrag {
  uint4 lo0, lo1, hi0, hi1;  // 32 B of row g, 32 B of row g+8
};

// Work decomposition.  Global k-step index s = row_block * spr + k_step (spr = K / 64 steps
// per 16-row block).  A CTA owns the contiguous range [s0, s0 + U); locally q = h0 + u walks
// it (h0 = steps of the first touched block that belong to predecessor CTAs), block = q / spr.
// Warp w takes u = w, w + NW, w + 2 NW, ... — the NW warps of a CTA walk the SAME 16 weight
// rows side by side, so at any moment the CTA reads NW * 128 contiguous bytes of each row.
//   * stand-alone kernel: ranges are whole row blocks (h0 = 0, U a multiple of spr);
//   * persistent kernel: ranges are S / gridDim steps wherever they fall ("stream-K"), so
//     N = 2560 (160 blocks on 148 SMs) is balanced to one k-step.  A block cut by a range
//     boundary is finalised by the CTA holding its LAST step; the other contributors export
//     their partial 16 x 8 tile through a global slot + flag, summed in CTA order.
struct GemvGeom {
  long long s0;   // first global k-step of the CTA
  long long rbf;  // first row block touched (= s0 / spr)
  int U;          // k-steps of the CTA
  int spr;        // k-steps per row block
  int h0;         // s0 - rbf * spr
  int u0;         // this warp's first local step (= warp index)
};

template <int NW>
VB_DEVICE GemvGeom gemv_geom_blocks(const GemvView& p, int cta, int ncta, int warp) {
  GemvGeom G;
  const long long nrb = (p.n() + 15) / 16;
  G.spr = p.k() >> 6;
  G.rbf = nrb * cta / ncta;
  G.s0 = G.rbf * G.spr;
  G.U = static_cast<int>(nrb * (cta + 1) / ncta - G.rbf) * G.spr;
  G.h0 = 0;
  G.u0 = warp;
  return G;
}

template <int NW>
VB_DEVICE GemvGeom gemv_geom_steps(const GemvView& p, int cta, int ncta, int warp) {
  GemvGeom G;
  const long long nrb = (p.n() + 15) / 16;
  G.spr = p.k() >> 6;
  const long long S = nrb * G.spr;
  G.s0 = S * cta / ncta;
  G.U = static_cast<int>(S * (cta + 1) / ncta - G.s0);
  G.rbf = G.s0 / G.spr;
  G.h0 = static_cast<int>(G.s0 - G.rbf * G.spr);
  G.u0 = warp;
  return G;
}

// load cursor of the rolling weight prefetch
struct WCursor {
  const __nv_bfloat16 *pa, *pb;
  int lr, lc;
};

VB_DEVICE void wc_set_rows(const GemvView& p, const GemvGeom& G, WCursor& c, int rbl, int g, int t) {
  long long ra = (G.rbf + rbl) * 16 + g, rh = ra + 8;  // (rbl may run one past the range: clamped)
  const long long last = p.n() - 1;
  ra = ra < last ? ra : last;  // rows past N re-read the last row; masked at the store
  rh = rh < last ? rh : last;
  c.pa = p.w() + ra * p.ldw() + t * 8;
  c.pb = p.w() + rh * p.ldw() + t * 8;
}

template <int NW>
VB_DEVICE void wc_load(const GemvView& p, const GemvGeom& G, WCursor& c, WFrag& f, int u, int g, int t) {
  if (u < G.U) {
    const uint4* qa = reinterpret_cast<const uint4*>(c.pa + c.lc * 64);
    const uint4* qb = reinterpret_cast<const uint4*>(c.pb + c.lc * 64);
    f.lo0 = __ldcs(qa);  // one instruction = 64 contiguous bytes (two full sectors) per row
    f.lo1 = __ldcs(qa + 4);
    f.hi0 = __ldcs(qb);
    f.hi1 = __ldcs(qb + 4);
    c.lc += NW;
    if (c.lc >= G.spr) {
      do {
        c.lc -= G.spr;
        ++c.lr;
      } while (c.lc >= G.spr);
      wc_set_rows(p, G, c, c.lr, g, t);
    }
  }
}

// first MD loads of this warp's steps (weights are constant on the stream: may be issued
// before the producer of x has finished)
template <int NW, int MD>
VB_DEVICE void gemv_prime(const GemvView& p, const GemvGeom& G, WCursor& c, WFrag (&buf)[MD], int g, int t,
                          int depth = MD) {
  const int q0 = G.h0 + G.u0;
  c.lr = q0 / G.spr;
  c.lc = q0 - c.lr * G.spr;
  wc_set_rows(p, G, c, c.lr, g, t);
#pragma unroll
  for (int d = 0; d < MD; ++d) {
    // every slot is (re)defined here, so nothing of the previous op stays live across the
    // code between two projections
    buf[d].lo0 = buf[d].lo1 = buf[d].hi0 = buf[d].hi1 = make_uint4(0, 0, 0, 0);
    if (d < depth) wc_load<NW>(p, G, c, buf[d], G.u0 + d * NW, g, t);
  }
}

// tops a partially primed buffer up to MD steps (call once the latency-critical loads of
// the op have been issued)
template <int NW, int MD>
VB_DEVICE void gemv_prime_rest(const GemvView& p, const GemvGeom& G, WCursor& c, WFrag (&buf)[MD], int g, int t,
                               int depth) {
#pragma unroll
  for (int d = 0; d < MD; ++d)
    if (d >= depth) wc_load<NW>(p, G, c, buf[d], G.u0 + d * NW, g, t);
}

// x rows -> shared memory (bf16, row stride K + kXPad), LayerNorm-ed on the way in by the whole
// CTA: every thread owns the same 16-byte column groups of every row in all three passes
// (load + sum, centred sum of squares, normalise), so only the two statistics reductions
// synchronise.  Loads bypass L1 (the rows were written by another SM a moment ago).
// `red`: 2 * NW * MR floats of scratch.  Ends with a CTA barrier.
struct NoHook {
  VB_DEVICE void operator()() const {}
};
template <int NW, int MR, typename Hook = NoHook>
VB_DEVICE void gemv_stage_x(const GemvView& p, int m, __nv_bfloat16* xs, float* red, int warp, int lane,
                            Hook after_loads = Hook(), const float* ln_g_s = nullptr,
                            const float* ln_b_s = nullptr) {
  const int K = p.k();
  const int xstride = K + kXPad;
  // ln_*_s: shared-memory copies requested with cp.async before the bulk weight loads
  const float* ln_g = p.ln_g() == nullptr ? nullptr : (ln_g_s != nullptr ? ln_g_s : p.ln_g());
  const float* ln_b = ln_b_s != nullptr ? ln_b_s : p.ln_b();
  const bool rms = ln_b == nullptr;  // gamma without beta: T5LayerNorm (no mean subtraction, no bias)
  const int tid = warp * 32 + lane;
  float* red2 = red + NW * MR;
  // pass 1: copy + row sums; the loads of up to four rows are issued back to back
  bool hooked = false;
  for (int r0 = 0; r0 < m; r0 += 4) {
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int c = tid * 8; c < K; c += NW * 256) {
      uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (r0 + j < m) v[j] = __ldcg(reinterpret_cast<const uint4*>(p.x() + (r0 + j) * p.ldx() + c));
      if (!hooked) {
        after_loads();  // bulk loads queue behind the latency-critical x loads
        hooked = true;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (r0 + j < m) {
          *reinterpret_cast<uint4*>(xs + static_cast<size_t>(r0 + j) * xstride + c) = v[j];
          const float2 a0 = unpack_bf16x2(v[j].x), a1 = unpack_bf16x2(v[j].y), a2 = unpack_bf16x2(v[j].z),
                       a3 = unpack_bf16x2(v[j].w);
          acc[j] += a0.x + a0.y + a1.x + a1.y + a2.x + a2.y + a3.x + a3.y;
        }
      }
    }
    if (ln_g != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (r0 + j < m) {
          const float v = warp_sum(acc[j]);
          if (lane == 0) red[warp * MR + r0 + j] = v;
        }
      }
    }
  }
  if (!hooked) after_loads();
  asm volatile("cp.async.wait_all;\n" ::: "memory");  // (no-op without pending copies)
  __syncthreads();
  if (ln_g == nullptr) return;
  auto total = [&](const float* scratch, int r) {  // fixed order: identical in every thread
    float v = 0.0f;
#pragma unroll
    for (int w = 0; w < NW; ++w) v += scratch[w * MR + r];
    return v;
  };
  for (int r = 0; r < m; ++r) {  // pass 2: centred sums of squares (own columns only)
    const __nv_bfloat16* xd = xs + static_cast<size_t>(r) * xstride;
    const float mu = rms ? 0.0f : total(red, r) / static_cast<float>(K);
    float acc = 0.0f;
    for (int c = tid * 8; c < K; c += NW * 256) {
      const uint4 v = *reinterpret_cast<const uint4*>(xd + c);
      const float2 a0 = unpack_bf16x2(v.x), a1 = unpack_bf16x2(v.y), a2 = unpack_bf16x2(v.z), a3 = unpack_bf16x2(v.w);
      const float d0 = a0.x - mu, d1 = a0.y - mu, d2 = a1.x - mu, d3 = a1.y - mu, d4 = a2.x - mu, d5 = a2.y - mu,
                  d6 = a3.x - mu, d7 = a3.y - mu;
      acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3 + d4 * d4 + d5 * d5 + d6 * d6 + d7 * d7;
    }
    acc = warp_sum(acc);
    if (lane == 0) red2[warp * MR + r] = acc;
  }
  __syncthreads();
  for (int r = 0; r < m; ++r) {  // pass 3: normalise in place
    __nv_bfloat16* xd = xs + static_cast<size_t>(r) * xstride;
    const float mu = rms ? 0.0f : total(red, r) / static_cast<float>(K);
    const float rstd = rsqrtf(total(red2, r) / static_cast<float>(K) + p.ln_eps());
    for (int c = tid * 8; c < K; c += NW * 256) {
      const uint4 v = *reinterpret_cast<const uint4*>(xd + c);
      const float2 a0 = unpack_bf16x2(v.x), a1 = unpack_bf16x2(v.y), a2 = unpack_bf16x2(v.z), a3 = unpack_bf16x2(v.w);
      const float4 g0 = *reinterpret_cast<const float4*>(ln_g + c), g1 = *reinterpret_cast<const float4*>(ln_g + c + 4);
      const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      const float4 b0 = rms ? zero4 : *reinterpret_cast<const float4*>(ln_b + c);
      const float4 b1 = rms ? zero4 : *reinterpret_cast<const float4*>(ln_b + c + 4);
      uint4 o;
      o.x = pack_bf16x2((a0.x - mu) * rstd * g0.x + b0.x, (a0.y - mu) * rstd * g0.y + b0.y);
      o.y = pack_bf16x2((a1.x - mu) * rstd * g0.z + b0.z, (a1.y - mu) * rstd * g0.w + b0.w);
      o.z = pack_bf16x2((a2.x - mu) * rstd * g1.x + b1.x, (a2.y - mu) * rstd * g1.y + b1.y);
      o.w = pack_bf16x2((a3.x - mu) * rstd * g1.z + b1.z, (a3.y - mu) * rstd * g1.w + b1.w);
      *reinterpret_cast<uint4*>(xd + c) = o;
    }
  }
  __syncthreads();
}

// consumes this warp's steps (buf was primed by gemv_prime) and parks its partial tiles:
// psum[local block][warp] = 16 x (NT*8) floats laid out [m][row]
template <int NT, int NW, int MD>
VB_DEVICE void gemv_main(const GemvView& p, int m, const GemvGeom& G, WCursor& c, WFrag (&buf)[MD],
                         const __nv_bfloat16* xs, float* psum, int warp, int g, int t) {
  const int xstride = p.k() + kXPad;
  float acc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[n][i] = 0.0f;
  const int q0 = G.h0 + G.u0;
  int cr = q0 / G.spr, cc = q0 - cr * G.spr;
  auto flush = [&]() {
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      float* ps = psum + (static_cast<size_t>(cr) * NW + warp) * (NT * 8 * 16) + (n * 8 + 2 * t) * 16;
      ps[g] = acc[n][0];
      ps[16 + g] = acc[n][1];
      ps[g + 8] = acc[n][2];
      ps[16 + g + 8] = acc[n][3];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[n][i] = 0.0f;
    }
  };
  for (int ub = G.u0; ub < G.U; ub += MD * NW) {
#pragma unroll
    for (int d = 0; d < MD; ++d) {
      const int u = ub + d * NW;
      if (u < G.U) {
        {
          const WFrag& f = buf[d];
          const uint32_t ra[8] = {f.lo0.x, f.lo0.y, f.lo0.z, f.lo0.w, f.lo1.x, f.lo1.y, f.lo1.z, f.lo1.w};
          const uint32_t rh[8] = {f.hi0.x, f.hi0.y, f.hi0.z, f.hi0.w, f.hi1.x, f.hi1.y, f.hi1.z, f.hi1.w};
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            uint4 x0 = make_uint4(0, 0, 0, 0), x1 = x0;  // columns beyond m multiply zeros
            if (n * 8 + g < m) {
              const __nv_bfloat16* xk = xs + static_cast<size_t>(n * 8 + g) * xstride + cc * 64 + t * 8;
              x0 = *reinterpret_cast<const uint4*>(xk);
              x1 = *reinterpret_cast<const uint4*>(xk + 32);
            }
            const uint32_t xb[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
              mma_bf16_16816(acc[n], ra[2 * j], rh[2 * j], ra[2 * j + 1], rh[2 * j + 1], xb[2 * j], xb[2 * j + 1]);
          }
        }
        // refill the slot only after its fragments were consumed: the loads land directly in
        // the slot's registers (no staging copy that would wait for them)
        wc_load<NW>(p, G, c, buf[d], u + MD * NW, g, t);
        cc += NW;
        if (cc >= G.spr || u + NW >= G.U) {  // this warp's next step is in another row block
          flush();
          do {
            cc -= G.spr;
            ++cr;
          } while (cc >= G.spr);
        }
      }
    }
  }
}

// sum over the warps that touched local block rbl (fixed order), element (m i, row r)
template <int NT, int NW>
VB_DEVICE float gemv_block_sum(const GemvGeom& G, const float* psum, int rbl, int i, int r) {
  int lo = rbl * G.spr - G.h0, hi = lo + G.spr;  // local steps of this block held by the CTA
  lo = lo > 0 ? lo : 0;
  hi = hi < G.U ? hi : G.U;
  float v = 0.0f;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    const int first = lo + (((w - lo) % NW) + NW) % NW;  // first step >= lo owned by warp w
    if (first < hi) v += psum[(static_cast<size_t>(rbl) * NW + w) * (NT * 8 * 16) + i * 16 + r];
  }
  return v;
}

// Stream-K hand-over (persistent kernel): `flags` / `slots` are indexed by CTA; a CTA whose
// range ends inside a row block exports its partial tile of that block.
struct GemvXfer {
  unsigned* flags;   // [gridDim.x], zeroed by the launcher
  float* slots;      // [gridDim.x][8*16]
  unsigned epoch;    // op ordinal (> 0): flags[c] == epoch <=> CTA c exported for this op
};

template <int NW>
VB_DEVICE void gemv_export_tail(const GemvGeom& G, const float* psum, const GemvXfer& x, int tid) {
  const int end = G.h0 + G.U;
  if (G.U == 0 || end % G.spr == 0) return;  // CTA-uniform
  const int rbl = end / G.spr;               // the block the range ends in
  if (tid < 8 * 16) x.slots[static_cast<size_t>(blockIdx.x) * (8 * 16) + tid] =
      gemv_block_sum<1, NW>(G, psum, rbl, tid >> 4, tid & 15);
  fence_gpu();
  __syncthreads();
  if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(x.flags + blockIdx.x), "r"(x.epoch) : "memory");
}

// fixed-order reduction + epilogue of the blocks this CTA owns (those whose last step it
// holds).  xfer != nullptr: the first block may have predecessor partials to add.
template <int NT, int NW>
VB_DEVICE void gemv_finalize(const GemvView& p, int m, const GemvGeom& G, const float* psum, int tid,
                             const GemvXfer* xfer, const float* bias_s = nullptr) {
  const long long ac = p.alpha_cols() <= 0 ? p.n() : p.alpha_cols();
  const int owned = (G.h0 + G.U) / G.spr;  // local blocks 0 .. owned-1 end inside the range
  int c_lo = blockIdx.x;
  if (xfer != nullptr && G.h0 > 0 && owned > 0) {
    // contributors: earlier CTAs whose range reaches into my first block
    const long long nrb = (p.n() + 15) / 16;
    const long long S = nrb * G.spr, B0 = G.rbf * G.spr;
    while (c_lo > 0 && S * c_lo / gridDim.x > B0) --c_lo;
    if (tid == 0) {
      for (int c = c_lo; c < static_cast<int>(blockIdx.x); ++c) {
        if (S * (c + 1) / gridDim.x == S * c / gridDim.x) continue;  // empty range
        if (ld_acquire_u32(xfer->flags + c) != xfer->epoch) {
          const long long t0 = clock64();
          while (ld_acquire_u32(xfer->flags + c) != xfer->epoch)
            if (clock64() - t0 > 4000000000LL) __trap();
        }
      }
    }
    __syncthreads();
  }
  for (int it = tid; it < owned * m * 16; it += NW * 32) {
    const int r = it & 15, i = (it >> 4) % m, rbl = (it >> 4) / m;
    const long long row = (G.rbf + rbl) * 16 + r;
    if (row >= p.n()) continue;
    float v = 0.0f;
    if (rbl == 0 && c_lo < static_cast<int>(blockIdx.x)) {
      const long long nrb = (p.n() + 15) / 16;
      const long long S = nrb * G.spr;
      for (int c = c_lo; c < static_cast<int>(blockIdx.x); ++c)
        if (S * (c + 1) / gridDim.x != S * c / gridDim.x)
          v += __ldcg(xfer->slots + static_cast<size_t>(c) * (8 * 16) + i * 16 + r);
    }
    v += gemv_block_sum<NT, NW>(G, psum, rbl, i, r);
    if (p.bias() != nullptr) v += bias_s != nullptr ? bias_s[rbl * 16 + r] : p.bias()[row];
    if (row < ac) v *= p.alpha();
    if (p.epilogue() == VB_EPI_GELU) v = gelu_erf(v);
    else if (p.epilogue() == VB_EPI_RELU) v = fmaxf(v, 0.0f);
    if (p.residual() != nullptr) v += ldcg_bf16(p.residual() + i * p.ldr() + row);
    if (p.out_f32()) reinterpret_cast<float*>(p.y())[i * p.ldy() + row] = v;
    else reinterpret_cast<__nv_bfloat16*>(p.y())[i * p.ldy() + row] = __float2bfloat16(v);
  }
}

constexpr int kGW = 8;  // warps of the stand-alone GEMV

template <int NT>
__global__ void __launch_bounds__(kGW * 32, 2) gemv_mma_kernel(const vb_decode_op op, int m, int rbmax) {
  extern __shared__ __align__(128) uint8_t gsm[];
  const GemvView p(op);
  const int K = p.k();
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(gsm);                               // [m][K+pad]
  float* psum = reinterpret_cast<float*>(gsm + static_cast<size_t>(m) * (K + kXPad) * 2);   // [rbmax][NW] tiles
  float* bias_s = psum + static_cast<size_t>(rbmax) * kGW * (NT * 8 * 16);                  // [rbmax*16]
  float* lng_s = bias_s + rbmax * 16;                                                       // [K] (LN only)
  float* lnb_s = lng_s + K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const GemvGeom G = gemv_geom_blocks<kGW>(p, blockIdx.x, gridDim.x, warp);
  // The small constants of this op (LN affine, bias rows) are requested FIRST: the per-SM load
  // path is in order, so behind 128 KB of weight prefetch they would wait microseconds.
  const bool ln = p.ln_g() != nullptr, ln_bias = p.ln_b() != nullptr;
  if (ln) {
    for (int c = threadIdx.x * 4; c < K; c += kGW * 128) {
      cp_async_16(lng_s + c, p.ln_g() + c);
      if (ln_bias) cp_async_16(lnb_s + c, p.ln_b() + c);
    }
  }
  const long long row0 = G.rbf * 16;
  const int nb = static_cast<int>((p.n() - row0) < (G.U / G.spr) * 16 ? (p.n() - row0) : (G.U / G.spr) * 16);
  const bool bias_al = p.bias() != nullptr && (reinterpret_cast<uintptr_t>(p.bias()) & 15u) == 0;
  if (p.bias() != nullptr) {
    for (int c = threadIdx.x * 4; c < nb; c += kGW * 128) {
      if (bias_al && c + 4 <= nb) cp_async_16(bias_s + c, p.bias() + row0 + c);
      else
        for (int j = c; j < nb && j < c + 4; ++j) bias_s[j] = p.bias()[row0 + j];
    }
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  WFrag buf[kMD];
  WCursor cur;
  gemv_prime<kGW, kMD>(p, G, cur, buf, g, t);
  pdl_trigger();
  pdl_wait();  // x / residual come from the previous kernel of the stream
  gemv_stage_x<kGW, NT * 8>(p, m, xs, psum, warp, lane, NoHook(), ln ? lng_s : nullptr,
                            (ln && ln_bias) ? lnb_s : nullptr);
  gemv_main<NT, kGW, kMD>(p, m, G, cur, buf, xs, psum, warp, g, t);
  __syncthreads();
  gemv_finalize<NT, kGW>(p, m, G, psum, threadIdx.x, nullptr, p.bias() != nullptr ? bias_s : nullptr);
}

static int sm_count() { return device_sm_count(); }

// Grid of the stand-alone tensor-core GEMV: two CTAs per SM when x + the partial-tile table
// fit, more waves of smaller ranges otherwise; 0 = shape not taken.
static long long gemv_mma_grid(int nt, long long m, long long n, long long k, bool ln, size_t* smem_out,
                               int* rbmax_out) {
  if (k % 64 != 0 || k > (1 << 20)) return 0;
  const long long nrb = (n + 15) / 16;
  const long long xb = m * (k + kXPad) * 2;
  for (int cps = 2; cps >= 1; --cps) {
    const long long limit = (cps == 1 ? 200 : 110) * 1024;
    for (long long waves = 1; waves <= 32; ++waves) {
      long long grid = static_cast<long long>(sm_count()) * cps * waves;
      if (grid > nrb) grid = nrb;
      const long long rbmax = (nrb + grid - 1) / grid;
      // x + [blocks][warps] partial tiles + bias rows + LN affine
      const long long smem = xb + rbmax * kGW * nt * 8 * 16 * 4 + rbmax * 16 * 4 + (ln ? 2 * k * 4 : 0);
      if (smem <= limit) {
        *smem_out = static_cast<size_t>(smem);
        *rbmax_out = static_cast<int>(rbmax);
        return grid;
      }
      if (grid == nrb) break;
    }
  }
  return 0;
}

template <int NT>
static cudaError_t launch_gemv_mma(const vb_decode_op& op, int m, long long grid, size_t smem, int rbmax,
                                   cudaStream_t s) {
  static DeviceOnce attr_once;
  bool& attr = attr_once();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemv_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  return launch_pdl(gemv_mma_kernel<NT>, dim3(static_cast<unsigned>(grid)), dim3(kGW * 32), smem, s, op, m, rbmax);
}

static vb_decode_op make_gemv_op(const void* x, const void* w, const float* bias, const void* residual, void* y,
                                 long long n, long long k, long long ldx, long long ldw, long long ldy,
                                 long long ldr, float alpha, long long alpha_cols, int epilogue, int out_dtype,
                                 const float* ln_gamma, const float* ln_beta, float ln_eps) {
  vb_decode_op op = {};
  op.type = VB_OP_GEMV;
  op.ptr[0] = w; op.ptr[1] = bias; op.ptr[2] = residual; op.ptr[3] = x; op.ptr[4] = y;
  op.ptr[5] = ln_gamma; op.ptr[6] = ln_beta;
  op.i64[0] = n; op.i64[1] = k; op.i64[2] = ldw; op.i64[3] = ldx; op.i64[4] = ldy; op.i64[5] = ldr;
  op.i64[6] = alpha_cols;
  op.f32[0] = alpha; op.f32[1] = ln_eps;
  op.i32[0] = epilogue; op.i32[1] = out_dtype == VB_F32 ? 1 : 0;
  return op;
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

cudaError_t gemv_launch(const void* x, const void* w, const float* bias, const void* residual,
                        void* y, long long m, long long n, long long k, long long ldx,
                        long long ldw, long long ldy, long long ldr, float alpha,
                        long long alpha_cols, int epilogue, int out_dtype, const float* ln_gamma,
                        const float* ln_beta, float ln_eps, cudaStream_t s) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  if (m > kGemvMaxM || k <= 0) return cudaErrorInvalidValue;
  const bool vec = (k % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0 && aligned16(x) && aligned16(w));
  const bool ln_al = ln_gamma == nullptr || (aligned16(ln_gamma) && (ln_beta == nullptr || aligned16(ln_beta)));
  if (vec && ln_al && k % 64 == 0) {
    const int nt = m <= 8 ? 1 : 2;
    const bool ln = ln_gamma != nullptr;
    size_t smem = 0;
    int rbmax = 0;
    long long grid = gemv_mma_grid(nt, m, n, k, ln, &smem, &rbmax);
    if (grid > 0) {
      const vb_decode_op op = make_gemv_op(x, w, bias, residual, y, n, k, ldx, ldw, ldy, ldr, alpha, alpha_cols,
                                           epilogue, out_dtype, ln_gamma, ln_beta, ln_eps);
      return nt == 1 ? launch_gemv_mma<1>(op, static_cast<int>(m), grid, smem, rbmax, s)
                     : launch_gemv_mma<2>(op, static_cast<int>(m), grid, smem, rbmax, s);
    }
    // more rows than one n8 tile and not enough shared memory for all of x: passes of 8 rows
    if (nt == 2 && gemv_mma_grid(1, 8, n, k, ln, &smem, &rbmax) > 0) {
      const size_t esz = out_dtype == VB_F32 ? 4 : 2;
      for (long long m0 = 0; m0 < m; m0 += 8) {
        const long long mm = m - m0 < 8 ? m - m0 : 8;
        size_t sm2 = 0;
        int rb2 = 0;
        const long long g2 = gemv_mma_grid(1, mm, n, k, ln, &sm2, &rb2);
        const vb_decode_op op = make_gemv_op(
            reinterpret_cast<const __nv_bfloat16*>(x) + m0 * ldx, w, bias,
            residual ? reinterpret_cast<const __nv_bfloat16*>(residual) + m0 * ldr : nullptr,
            reinterpret_cast<uint8_t*>(y) + static_cast<size_t>(m0) * ldy * esz, n, k, ldx, ldw, ldy, ldr, alpha,
            alpha_cols, epilogue, out_dtype, ln_gamma, ln_beta, ln_eps);
        cudaError_t e = launch_gemv_mma<1>(op, static_cast<int>(mm), g2, sm2, rb2, s);
        if (e != cudaSuccess) return e;
      }
      return cudaSuccess;
    }
  }
  if (ln_gamma != nullptr) return cudaErrorInvalidValue;  // LN fusion needs the staged path
  const vb_decode_op op = make_gemv_op(x, w, bias, residual, y, n, k, ldx, ldw, ldy, ldr, alpha, alpha_cols,
                                       epilogue, out_dtype, nullptr, nullptr, 0.0f);
  const unsigned grid = static_cast<unsigned>((n + 3) / 4);
  const int mi = static_cast<int>(m), v = vec ? 1 : 0;
  if (m <= 1) gemv_kernel<1><<<grid, 128, 0, s>>>(op, mi, v);
  else if (m <= 2) gemv_kernel<2><<<grid, 128, 0, s>>>(op, mi, v);
  else if (m <= 4) gemv_kernel<4><<<grid, 128, 0, s>>>(op, mi, v);
  else if (m <= 8) gemv_kernel<8><<<grid, 128, 0, s>>>(op, mi, v);
  else gemv_kernel<16><<<grid, 128, 0, s>>>(op, mi, v);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ decode-step prologue
// x[b, :] = embed[tokens[b], :] + pos_table[n_valid[b] + pos_offset, :]; then the per-sequence
// counters advance (n_valid, ctx_len += 1).  `nthr` threads cooperate on sequence b.
// HF:opt/modeling_opt.py:45-70 (offset 2), :350-354 (positions from the mask cumsum).
VB_DEVICE void decode_embed_row(const long long* tokens, const __nv_bfloat16* embed, const __nv_bfloat16* pos_table,
                                const int* n_valid, __nv_bfloat16* x, long long dim, long long vocab,
                                long long pos_rows, long long pos_offset, int b, int tid, int nthr) {
  long long tok = tokens[b];
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
  long long pos = static_cast<long long>(n_valid[b]) + pos_offset;
  pos = pos >= pos_rows ? pos_rows - 1 : pos;
  const __nv_bfloat16* e = embed + tok * dim;
  const __nv_bfloat16* q = pos_table + pos * dim;
  for (long long c = tid; c < dim; c += nthr)
    x[b * dim + c] = __float2bfloat16(__bfloat162float(e[c]) + __bfloat162float(q[c]));
}

__global__ void __launch_bounds__(256)
decode_embed_kernel(const long long* tokens, const __nv_bfloat16* embed, const __nv_bfloat16* pos_table,
                    int* n_valid, int* ctx_len, __nv_bfloat16* x, long long dim, long long vocab,
                    long long pos_rows, long long pos_offset) {
  // No early trigger here: the attention kernel two launches later reads ctx_len before ITS
  // dependency wait, which is only safe once this kernel has completed.
  pdl_wait();
  const int b = blockIdx.x;
  decode_embed_row(tokens, embed, pos_table, n_valid, x, dim, vocab, pos_rows, pos_offset, b, threadIdx.x, 256);
  __syncthreads();
  if (threadIdx.x == 0) {
    n_valid[b] += 1;
    ctx_len[b] += 1;
  }
}

cudaError_t decode_embed_launch(const long long* tokens, const void* embed, const void* pos_table,
                                int* n_valid, int* ctx_len, void* x, long long batch, long long dim,
                                long long vocab, long long pos_rows, long long pos_offset,
                                cudaStream_t s) {
  if (batch <= 0) return cudaSuccess;
  return launch_pdl(decode_embed_kernel, dim3(static_cast<unsigned>(batch)), dim3(256), 0, s, tokens,
                    reinterpret_cast<const __nv_bfloat16*>(embed),
                    reinterpret_cast<const __nv_bfloat16*>(pos_table), n_valid, ctx_len,
                    reinterpret_cast<__nv_bfloat16*>(x), dim, vocab, pos_rows, pos_offset);
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

// ------------------------------------------------------------------ single-query attention
// Flash-decoding: one unit = (head, sequence, split), worked by 128 threads.  A unit covers
// [lo, hi) of the ctx cached tokens (the unit owning the last position first appends this
// step's k/v), computes a partial softmax(q.K^T).V relative to its local max and parks
// (max, sum, acc[D]) in the workspace; the last unit to finish a (head, sequence) merges the
// splits.  Phase 1: one thread per cached token (its whole K row in flight as 16-byte
// loads); phase 2: thread = (token group, 8-wide d vector), 8 V-row loads in flight.
struct AttnArgs {
  const __nv_bfloat16* qkv;
  __nv_bfloat16 *kc, *vc;
  const int *page_table, *ctx_len, *first_valid;
  __nv_bfloat16* out;
  float* ws;
  int* counters;
  int heads, D, page_size, max_pages, splits, chunk_cap;
  float scale;
  // T5: additive relative-position bias of key l seen from the newest position ctx-1:
  // rel_bias[h * rel_stride + rel_center + (l - (ctx - 1))]; nullptr = none
  const float* rel_bias;
  int rel_stride, rel_center;
  // generalisations for the one-query cross-attention of an encoder-decoder LM: elements
  // between consecutive cached tokens (0 = heads*D), between the q rows of consecutive
  // sequences (0 = 3*heads*D), and whether this step's k/v row is appended (self-attention)
  long long kv_stride, q_stride;
  int append;
};

__host__ __device__ inline int attn_unit_smem_floats(int D, int chunk_cap) { return D + chunk_cap + 16 * D + 8; }

// sm: attn_unit_smem_floats() floats private to these 128 threads; bar_id: their named barrier.
// WAIT: the stand-alone kernel — everything that does not depend on this step's qkv row (page
// lookups and the K / V rows of all previously cached tokens) is requested BEFORE
// griddepcontrol.wait, i.e. while the qkv projection is still streaming its weights; after the
// wait only q, the new token and the arithmetic remain.
template <bool WAIT>
VB_DEVICE void attn_unit(const AttnArgs& a, int h, int b, int sp, int tid, float* sm, uint32_t bar_id) {
  const int D = a.D, heads = a.heads, splits = a.splits, page_size = a.page_size;
  const int hd = heads * D;
  const int ctx = a.ctx_len[b];
  const int fv = a.first_valid != nullptr ? a.first_valid[b] : 0;
  const int chunk = (ctx + splits - 1) / splits;
  const int lo = sp * chunk;
  const int hi = lo + chunk < ctx ? lo + chunk : ctx;
  float* sq = sm;                    // D
  float* sc = sm + D;                // chunk_cap scores
  float* part = sc + a.chunk_cap;    // groups * D
  float* red = part + 16 * D;        // 4
  int* s_last = reinterpret_cast<int*>(red + 4);
  const int* pt = a.page_table + static_cast<long long>(b) * a.max_pages;
  const long long kvs = a.kv_stride > 0 ? a.kv_stride : hd;
  const __nv_bfloat16* row = a.qkv + static_cast<long long>(b) * (a.q_stride > 0 ? a.q_stride : 3 * hd);
  const int warp = tid >> 5, lane = tid & 31;
  __nv_bfloat16* kc = a.kc;
  __nv_bfloat16* vc = a.vc;
  auto tok_off = [&](int l) {
    return (static_cast<long long>(pt[l / page_size]) * page_size + l % page_size) * kvs + h * D;
  };
  const bool vec = (D % 8 == 0) && (hd % 8 == 0) && (kvs % 8 == 0);
  const bool fast = vec && D <= 128;  // whole K row (<= 16 x 16 B) and 8 V pieces live in registers
  // this step's token: its k/v come from the qkv row (self-attention); none for cross-attention
  const int newest = a.append ? ctx - 1 : -1;
  // phase-2 mapping: thread = (token group, 8-wide d vector)
  const int nvec = (D + 7) / 8;
  int groups = 128 / nvec;
  if (groups > 16) groups = 16;
  const int gidx = tid / nvec, vi = tid % nvec;
  const int start = lo > fv ? lo : fv;
  constexpr int U = 8;  // independent V-row loads in flight per thread

  // ---- early requests (cached tokens only)
  uint4 kreg[16];
  uint4 vreg[U];
  const int l1 = lo + tid;  // this thread's first phase-1 token
  if (fast) {
    if (l1 < hi && l1 >= fv && l1 != newest) {
      const __nv_bfloat16* kr = kc + tok_off(l1);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i * 8 < D) kreg[i] = *reinterpret_cast<const uint4*>(kr + i * 8);
    }
    if (gidx < groups) {
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int ll = start + gidx + q * groups;
        vreg[q] = make_uint4(0, 0, 0, 0);
        if (ll < hi && ll != newest) vreg[q] = *reinterpret_cast<const uint4*>(vc + tok_off(ll) + vi * 8);
      }
    }
  }
  if (WAIT) pdl_wait();

  // ---- q (pre-scaled) and the new token's k / v
  const bool owns_new = a.append && (newest >= lo && newest < hi);
  for (int c = tid; c < D; c += 128) {
    if (owns_new) {
      const long long dst = tok_off(newest);
      kc[dst + c] = __float2bfloat16(ldcg_bf16(row + hd + h * D + c));
      vc[dst + c] = __float2bfloat16(ldcg_bf16(row + 2 * hd + h * D + c));
    }
    sq[c] = ldcg_bf16(row + h * D + c) * a.scale;
  }
  if (fast) {
    if (l1 == newest && l1 < hi) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i * 8 < D) kreg[i] = __ldcg(reinterpret_cast<const uint4*>(row + hd + h * D + i * 8));
    }
    if (gidx < groups) {
#pragma unroll
      for (int q = 0; q < U; ++q)
        if (start + gidx + q * groups == newest && newest < hi)
          vreg[q] = __ldcg(reinterpret_cast<const uint4*>(row + 2 * hd + h * D + vi * 8));
    }
  }
  named_bar_sync(bar_id, 128);

  // ---- phase 1: scores, one thread per token
  float mx = -INFINITY;
  for (int l = l1; l < hi; l += 128) {
    float s = -INFINITY;
    if (l >= fv) {
      float acc = 0.0f;
      if (fast) {
        if (l != l1) {  // contexts longer than 128 tokens per split: later rows are loaded here
          const __nv_bfloat16* kr = kc + tok_off(l);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i * 8 < D) kreg[i] = *reinterpret_cast<const uint4*>(kr + i * 8);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (i * 8 < D) {
            const int c = i * 8;
            const float2 a0 = unpack_bf16x2(kreg[i].x), a1 = unpack_bf16x2(kreg[i].y), a2 = unpack_bf16x2(kreg[i].z),
                         a3 = unpack_bf16x2(kreg[i].w);
            acc += sq[c] * a0.x + sq[c + 1] * a0.y + sq[c + 2] * a1.x + sq[c + 3] * a1.y + sq[c + 4] * a2.x +
                   sq[c + 5] * a2.y + sq[c + 6] * a3.x + sq[c + 7] * a3.y;
          }
        }
      } else {
        const __nv_bfloat16* kr = kc + tok_off(l);
        if (vec) {
#pragma unroll 4
          for (int c = 0; c < D; c += 8) {
            const uint4 u = *reinterpret_cast<const uint4*>(kr + c);
            const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
            acc += sq[c] * a0.x + sq[c + 1] * a0.y + sq[c + 2] * a1.x + sq[c + 3] * a1.y + sq[c + 4] * a2.x +
                   sq[c + 5] * a2.y + sq[c + 6] * a3.x + sq[c + 7] * a3.y;
          }
        } else {
          for (int c = 0; c < D; ++c) acc += sq[c] * __bfloat162float(kr[c]);
        }
      }
      s = acc;
      if (a.rel_bias != nullptr) s += a.rel_bias[h * a.rel_stride + a.rel_center + (l - newest)];
    }
    sc[l - lo] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  named_bar_sync(bar_id, 128);
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  named_bar_sync(bar_id, 128);
  float sum = 0.0f;
  for (int l = l1; l < hi; l += 128) {
    const float pr = (sc[l - lo] == -INFINITY) ? 0.0f : __expf(sc[l - lo] - mx);
    sc[l - lo] = pr;
    sum += pr;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  named_bar_sync(bar_id, 128);
  const float tot = red[0] + red[1] + red[2] + red[3];

  // ---- phase 2: P.V
  if (gidx < groups) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (vec) {
      for (int l = start + gidx; l < hi; l += groups * U) {
        float pr[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const int ll = l + q * groups;
          pr[q] = ll < hi ? sc[ll - lo] : 0.0f;
          if (!(fast && l == start + gidx)) {  // first batch is already in registers
            vreg[q] = make_uint4(0, 0, 0, 0);
            if (ll < hi) vreg[q] = *reinterpret_cast<const uint4*>(vc + tok_off(ll) + vi * 8);
          }
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const float2 a0 = unpack_bf16x2(vreg[q].x), a1 = unpack_bf16x2(vreg[q].y), a2 = unpack_bf16x2(vreg[q].z),
                       a3 = unpack_bf16x2(vreg[q].w);
          acc[0] += pr[q] * a0.x; acc[1] += pr[q] * a0.y; acc[2] += pr[q] * a1.x; acc[3] += pr[q] * a1.y;
          acc[4] += pr[q] * a2.x; acc[5] += pr[q] * a2.y; acc[6] += pr[q] * a3.x; acc[7] += pr[q] * a3.y;
        }
      }
    } else {
      for (int l = start + gidx; l < hi; l += groups) {
        const float pr = sc[l - lo];
        const __nv_bfloat16* vr = vc + tok_off(l) + vi * 8;
        for (int j = 0; j < 8; ++j)
          if (vi * 8 + j < D) acc[j] += pr * __bfloat162float(vr[j]);
      }
    }
    for (int j = 0; j < 8; ++j)
      if (vi * 8 + j < D) part[gidx * D + vi * 8 + j] = acc[j];
  }
  named_bar_sync(bar_id, 128);
  float* my = a.ws + ((static_cast<long long>(b) * heads + h) * splits + sp) * (D + 2);
  for (int c = tid; c < D; c += 128) {
    float v = 0.0f;
    for (int gi = 0; gi < groups; ++gi) v += part[gi * D + c];
    my[2 + c] = v;
  }
  if (tid == 0) {
    my[0] = mx;
    my[1] = tot;
  }
  fence_gpu();
  named_bar_sync(bar_id, 128);
  if (tid == 0) {
    const int ticket = atomicAdd(&a.counters[b * heads + h], 1);
    *s_last = (ticket == splits - 1) ? 1 : 0;
  }
  named_bar_sync(bar_id, 128);
  if (*s_last) {
    fence_gpu();
    const float* base = a.ws + (static_cast<long long>(b) * heads + h) * splits * (D + 2);
    float gm = -INFINITY;
    for (int i = 0; i < splits; ++i) gm = fmaxf(gm, __ldcg(base + i * (D + 2)));
    float denom = 0.0f;
    for (int i = 0; i < splits; ++i) {
      const float mi = __ldcg(base + i * (D + 2));
      denom += (mi == -INFINITY) ? 0.0f : __expf(mi - gm) * __ldcg(base + i * (D + 2) + 1);
    }
    const float inv = denom > 0.0f ? 1.0f / denom : 0.0f;
    for (int c = tid; c < D; c += 128) {
      float v = 0.0f;
      for (int i = 0; i < splits; ++i) {
        const float mi = __ldcg(base + i * (D + 2));
        if (mi != -INFINITY) v += __expf(mi - gm) * __ldcg(base + i * (D + 2) + 2 + c);
      }
      a.out[static_cast<long long>(b) * hd + h * D + c] = __float2bfloat16(v * inv);
    }
    if (tid == 0) a.counters[b * heads + h] = 0;  // ready for the next step / graph replay
  }
  named_bar_sync(bar_id, 128);  // sm is reused by the next unit
}

__global__ void __launch_bounds__(128) paged_decode_attn_kernel(const AttnArgs a) {
  extern __shared__ float attn_sm[];
  pdl_trigger();  // lets the out-projection GEMV prefetch its weights under this kernel
  attn_unit<true>(a, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, attn_sm, 1);
}

cudaError_t paged_decode_attention_launch(const void* qkv, void* k_cache, void* v_cache,
                                          const int* page_table, const int* ctx_len,
                                          const int* first_valid, void* out, float* workspace,
                                          int* counters, long long splits, long long batch,
                                          long long heads, long long d, long long page_size,
                                          long long max_pages, float scale, const float* rel_bias,
                                          long long rel_stride, long long rel_center, cudaStream_t s) {
  if (batch <= 0 || heads <= 0) return cudaSuccess;
  if (splits <= 0 || splits > 64 || workspace == nullptr || counters == nullptr) return cudaErrorInvalidValue;
  const long long max_ctx = page_size * max_pages;
  const long long chunk_cap = (max_ctx + splits - 1) / splits;
  const size_t smem = sizeof(float) * static_cast<size_t>(attn_unit_smem_floats(static_cast<int>(d), static_cast<int>(chunk_cap)));
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(paged_decode_attn_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr = 200 * 1024;
  }
  AttnArgs a;
  a.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  a.kc = reinterpret_cast<__nv_bfloat16*>(k_cache);
  a.vc = reinterpret_cast<__nv_bfloat16*>(v_cache);
  a.page_table = page_table; a.ctx_len = ctx_len; a.first_valid = first_valid;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.ws = workspace; a.counters = counters;
  a.heads = static_cast<int>(heads); a.D = static_cast<int>(d); a.page_size = static_cast<int>(page_size);
  a.max_pages = static_cast<int>(max_pages); a.splits = static_cast<int>(splits);
  a.chunk_cap = static_cast<int>(chunk_cap); a.scale = scale;
  a.rel_bias = rel_bias; a.rel_stride = static_cast<int>(rel_stride); a.rel_center = static_cast<int>(rel_center);
  a.kv_stride = 0; a.q_stride = 0; a.append = 1;
  dim3 grid(static_cast<unsigned>(heads), static_cast<unsigned>(batch), static_cast<unsigned>(splits));
  return launch_pdl(paged_decode_attn_kernel, grid, dim3(128), smem, s, a);
}

// One query per sequence over a dense (B, L, stride) key / value buffer — the cross-attention of
// a decoder step over the encoder's projected K|V (no append, no paging: page b = sequence b).
cudaError_t decode_cross_attention_launch(const void* q, long long q_stride, const void* k, const void* v,
                                          long long kv_stride, const int* seq_ids, const int* ctx_len,
                                          const int* first_valid, void* out, float* workspace, int* counters,
                                          long long splits, long long batch, long long heads, long long d,
                                          long long max_ctx, float scale, cudaStream_t s) {
  if (batch <= 0 || heads <= 0) return cudaSuccess;
  if (splits <= 0 || splits > 64 || workspace == nullptr || counters == nullptr || seq_ids == nullptr ||
      out == nullptr || max_ctx <= 0)
    return cudaErrorInvalidValue;
  const long long chunk_cap = (max_ctx + splits - 1) / splits;
  const size_t smem = sizeof(float) * static_cast<size_t>(attn_unit_smem_floats(static_cast<int>(d), static_cast<int>(chunk_cap)));
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(paged_decode_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return e;
  }
  AttnArgs a;
  a.qkv = reinterpret_cast<const __nv_bfloat16*>(q);
  a.kc = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(k));
  a.vc = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(v));
  a.page_table = seq_ids; a.ctx_len = ctx_len; a.first_valid = first_valid;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.ws = workspace; a.counters = counters;
  a.heads = static_cast<int>(heads); a.D = static_cast<int>(d); a.page_size = static_cast<int>(max_ctx);
  a.max_pages = 1; a.splits = static_cast<int>(splits); a.chunk_cap = static_cast<int>(chunk_cap);
  a.scale = scale;
  a.rel_bias = nullptr; a.rel_stride = 0; a.rel_center = 0;
  a.kv_stride = kv_stride; a.q_stride = q_stride; a.append = 0;
  dim3 grid(static_cast<unsigned>(heads), static_cast<unsigned>(batch), static_cast<unsigned>(splits));
  return launch_pdl(paged_decode_attn_kernel, grid, dim3(128), smem, s, a);
}

// ------------------------------------------------------------------ persistent decode step
// One cooperative launch per generated token.  Every CTA (one per SM, 8 warps) walks the
// op list; ops are separated by a grid barrier, and before entering it every warp has
// already issued the first kPD weight loads of the next projection (128 regs per lane = 128 KB
// per SM in flight), so HBM keeps streaming while the tiny activation vector makes its way
// through the barrier.
constexpr int kPW = 8;  // warps per CTA of the persistent kernel

// All CTAs are co-resident (cooperative launch).  `target` = barrier ordinal * gridDim.x; the
// counter is zeroed by the launcher.  Bounded spin: a mis-programmed op list traps instead of
// hanging the GPU.
VB_DEVICE void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    // release: everything this CTA wrote (ordered before by the CTA barrier) is visible to
    // whoever acquires the count
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
    if (ld_acquire_u32(counter) < target) {
      const long long t0 = clock64();
      while (ld_acquire_u32(counter) < target) {
        if (clock64() - t0 > 4000000000LL) __trap();
      }
    }
    fence_gpu();
  }
  __syncthreads();
}

// Workspace of a decode-step program (uint32 words): [0] barrier counter, [16 .. 16+grid)
// stream-K flags, byte 4096 onwards the per-CTA partial-tile slots (512 B each).  The first
// 4096 bytes are zeroed by the launcher.
constexpr int kWsFlagWord = 16;
constexpr int kWsSlotByte = 4096;
constexpr int kWsMaxCtas = (kWsSlotByte / 4) - kWsFlagWord;

__global__ void __launch_bounds__(kPW * 32, 1)
decode_step_kernel(const vb_decode_op* __restrict__ ops, int n_ops, int m, unsigned* ws,
                   unsigned long long* trace, int early) {
  extern __shared__ __align__(128) uint8_t gsm[];
  // optional phase timeline (VB_DECODE_TRACE tooling): 6 globaltimer stamps per (op, CTA)
  auto stamp = [&](int op_i, int k) {
    if (trace != nullptr && threadIdx.x == 0) {
      unsigned long long tns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
      trace[(static_cast<size_t>(op_i) * gridDim.x + blockIdx.x) * 6 + k] = tns;
    }
  };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  WFrag buf[kPD];
  WCursor cur;
  GemvGeom G = {};
  unsigned epoch = 0;
  GemvXfer xfer;
  xfer.flags = ws + kWsFlagWord;
  xfer.slots = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + kWsSlotByte);
  xfer.epoch = 0;
  // Invariant: when op i is a projection, its first kPD weight loads were issued at the end
  // of op i-1 (before the barrier); `prime_next` (re)defines every slot of buf either way.
  auto prime_next = [&](int nxt) {
    if (nxt < n_ops && ops[nxt].type == VB_OP_GEMV) {
      const GemvView pn(ops[nxt]);
      G = gemv_geom_steps<kPW>(pn, blockIdx.x, gridDim.x, warp);
      gemv_prime<kPW, kPD>(pn, G, cur, buf, g, t, early);
    } else {
#pragma unroll
      for (int d = 0; d < kPD; ++d) buf[d].lo0 = buf[d].lo1 = buf[d].hi0 = buf[d].hi1 = make_uint4(0, 0, 0, 0);
    }
  };
  prime_next(0);
  for (int i = 0; i < n_ops; ++i) {
    const vb_decode_op& op = ops[i];
    const int type = op.type;
    stamp(i, 0);
    if (type == VB_OP_GEMV) {
      const GemvView p(op);
      __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(gsm);
      float* psum = reinterpret_cast<float*>(gsm + static_cast<size_t>(m) * (p.k() + kXPad) * 2);
      gemv_stage_x<kPW, 8>(p, m, xs, psum, warp, lane, [&]() {
        gemv_prime_rest<kPW, kPD>(p, G, cur, buf, g, t, early);  // bulk loads go out behind the x loads
      });
      stamp(i, 1);
      gemv_main<1, kPW, kPD>(p, m, G, cur, buf, xs, psum, warp, g, t);
      stamp(i, 2);
      __syncthreads();
      stamp(i, 3);
      // latency-critical tail first (bias / residual / neighbour partials), then the first
      // `early` steps of the next projection, which stay in flight across the barrier
      xfer.epoch = static_cast<unsigned>(i) + 1u;
      gemv_export_tail<kPW>(G, psum, xfer, threadIdx.x);
      gemv_finalize<1, kPW>(p, m, G, psum, threadIdx.x, &xfer);
      prime_next(i + 1);
      stamp(i, 4);
    } else {
      if (type == VB_OP_ATTN) {
        AttnArgs a;
        a.qkv = reinterpret_cast<const __nv_bfloat16*>(op.ptr[0]);
        a.kc = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(op.ptr[1]));
        a.vc = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(op.ptr[2]));
        a.page_table = reinterpret_cast<const int*>(op.ptr[3]);
        a.ctx_len = reinterpret_cast<const int*>(op.ptr[4]);
        a.first_valid = reinterpret_cast<const int*>(op.ptr[5]);
        a.out = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(op.ptr[6]));
        a.ws = reinterpret_cast<float*>(const_cast<void*>(op.ptr[7]));
        a.counters = reinterpret_cast<int*>(const_cast<void*>(op.ptr[8]));
        a.heads = op.i32[0]; a.D = op.i32[1]; a.page_size = op.i32[2]; a.max_pages = op.i32[3];
        a.splits = op.i32[4]; a.chunk_cap = op.i32[5]; a.scale = op.f32[0];
        a.rel_bias = nullptr; a.rel_stride = 0; a.rel_center = 0;
        a.kv_stride = 0; a.q_stride = 0; a.append = 1;
        const int sub = threadIdx.x >> 7, tid = threadIdx.x & 127;
        float* sm = reinterpret_cast<float*>(gsm) + static_cast<size_t>(sub) * attn_unit_smem_floats(a.D, a.chunk_cap);
        const int units = a.heads * m * a.splits;
        for (int u = blockIdx.x * (kPW / 4) + sub; u < units; u += gridDim.x * (kPW / 4)) {
          const int h = u % a.heads, rest = u / a.heads;
          attn_unit<false>(a, h, rest % m, rest / m, tid, sm, 1 + sub);
        }
      } else if (type == VB_OP_EMBED) {
        const long long* tokens = reinterpret_cast<const long long*>(op.ptr[0]);
        int* n_valid = reinterpret_cast<int*>(const_cast<void*>(op.ptr[3]));
        int* ctx_len = reinterpret_cast<int*>(const_cast<void*>(op.ptr[4]));
        for (int b = blockIdx.x; b < m; b += gridDim.x) {
          decode_embed_row(tokens, reinterpret_cast<const __nv_bfloat16*>(op.ptr[1]),
                           reinterpret_cast<const __nv_bfloat16*>(op.ptr[2]), n_valid,
                           reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(op.ptr[5])), op.i64[0], op.i64[1],
                           op.i64[2], op.i64[3], b, threadIdx.x, kPW * 32);
          __syncthreads();
          if (threadIdx.x == 0) {
            n_valid[b] += 1;
            ctx_len[b] += 1;
          }
        }
      }
      prime_next(i + 1);
    }
    if (type != VB_OP_GEMV) stamp(i, 4);
    if (i + 1 < n_ops) grid_barrier(ws, ++epoch * gridDim.x);
    stamp(i, 5);
  }
}

// Shared memory the persistent kernel needs for an op list (host copy of the ops); -1 = an
// op the kernel does not take.
static long long decode_step_smem(const vb_decode_op* ops, int n_ops, int m, int grid) {
  long long need = 0;
  for (int i = 0; i < n_ops; ++i) {
    const vb_decode_op& o = ops[i];
    long long s = 0;
    if (o.type == VB_OP_GEMV) {
      const long long n = o.i64[0], k = o.i64[1];
      if (n <= 0 || k <= 0 || k % 64 != 0 || o.i64[2] % 8 != 0 || o.i64[3] % 8 != 0 || !aligned16(o.ptr[0]) ||
          !aligned16(o.ptr[3]))
        return -1;
      const long long spr = k / 64, S = (n + 15) / 16 * spr;
      const long long umax = (S + grid - 1) / grid;
      const long long blocks = (umax + spr - 1) / spr + 1;  // touched by one CTA range
      s = static_cast<long long>(m) * (k + kXPad) * 2 + blocks * kPW * 8 * 16 * 4;
    } else if (o.type == VB_OP_ATTN) {
      if (o.i32[4] <= 0 || o.i32[4] > 64) return -1;
      s = 4LL * (kPW / 4) * attn_unit_smem_floats(o.i32[1], o.i32[5]);
    } else if (o.type != VB_OP_EMBED) {
      return -1;
    }
    need = s > need ? s : need;
  }
  return need;
}

static unsigned long long* g_decode_trace = nullptr;
void decode_step_set_trace(void* buffer) { g_decode_trace = reinterpret_cast<unsigned long long*>(buffer); }

cudaError_t decode_step_launch(const vb_decode_op* ops_host, const vb_decode_op* ops_dev, int n_ops, int m,
                               unsigned* workspace, cudaStream_t s) {
  if (n_ops <= 0) return cudaSuccess;
  if (m <= 0 || m > 8 || ops_host == nullptr || ops_dev == nullptr || workspace == nullptr)
    return cudaErrorInvalidValue;
  const int grid = sm_count();
  if (grid > kWsMaxCtas) return cudaErrorInvalidValue;
  const long long smem = decode_step_smem(ops_host, n_ops, m, grid);
  if (smem < 0 || smem > 200 * 1024) return cudaErrorInvalidValue;
  static DeviceOnce attr_once;
  bool& attr = attr_once();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  cudaError_t e = cudaMemsetAsync(workspace, 0, kWsSlotByte, s);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kPW * 32);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  static const int early = [] {
    const char* e = std::getenv("VB_DECODE_EARLY");
    const int v = e ? std::atoi(e) : 2;
    return v < 0 ? 0 : (v > kPD ? kPD : v);
  }();
  return cudaLaunchKernelEx(&cfg, decode_step_kernel, ops_dev, n_ops, m, workspace, g_decode_trace, early);
}

}  // namespace vb
