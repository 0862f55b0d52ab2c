// Fused softmax attention (forward + backward), FlashAttention-style tiling:
// 64 queries x 64 keys per step, online softmax in fp32, bf16 tensor-core MMAs
// (mma.sync m16n8k16) with operands staged in padded shared memory via cp.async.
// Head dim is a template parameter DP (multiple of 16, zero padded): 96 serves the
// ViT's d=88, 80 OPT's d=80, 64 the Q-Former, 16 the reference's tiny test configs.
//
// Addressing: element (b, s, h, d) at base + b*bs + s*rs + h*D + d, so q/k/v may alias
// one fused-QKV activation buffer and o may be written straight into the (tokens, H*D)
// layout the following projection GEMM consumes — no transposes anywhere.
#include <cstdlib>

#include "common.cuh"
#include "internal.h"

namespace vb {

constexpr int kAM = 64;  // queries per CTA
constexpr int kAN = 64;  // keys per step
constexpr int kAttnThreads = 128;

VB_DEVICE void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
VB_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
VB_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

VB_DEVICE void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
VB_DEVICE void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
VB_DEVICE void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Copy `rows_valid` rows of D bf16 (zero-padded to DP columns and to 64 rows) into a
// [64][DP+8] shared tile.  vec: 16-byte cp.async path (D%8==0, aligned), else scalar.
template <int DP>
VB_DEVICE void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, long long row_stride,
                         int rows_valid, int D, bool vec) {
  constexpr int LDS = DP + 8;
  if (vec) {
    constexpr int VPR = DP / 8;
    for (int e = threadIdx.x; e < 64 * VPR; e += kAttnThreads) {
      const int r = e / VPR, c = (e % VPR) * 8;
      __nv_bfloat16* dst = s + r * LDS + c;
      if (r < rows_valid && c < D) {
        cp_async16(dst, g + r * row_stride + c);
      } else {
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
      }
    }
  } else {
    for (int e = threadIdx.x; e < 64 * DP; e += kAttnThreads) {
      const int r = e / DP, c = e % DP;
      __nv_bfloat16 v = __float2bfloat16(0.0f);
      if (r < rows_valid && c < D) v = g[r * row_stride + c];
      s[r * LDS + c] = v;
    }
  }
}

struct AttnParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* o;
  float* lse;
  const uint8_t* key_mask;
  int sq, skv, d, heads;
  long long q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;
  float scale_log2;  // scale * log2(e)
  int causal;
  int vec;
  int o_vec2;
  const unsigned long long* drop_seed;
  unsigned long long drop_salt;
  unsigned int drop_thresh;  // 0 = no dropout on the probabilities
  float drop_scale;
  const float* rel_bias;      // T5 relative-position bias table [heads][sq + skv - 1] or nullptr
  long long rel_bias_stride;
};

template <int DP>
__global__ void __launch_bounds__(kAttnThreads) attn_fwd_kernel(const AttnParams p) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  constexpr int LDS = DP + 8;
  constexpr int KS = DP / 16;  // k-steps of Q.K^T
  constexpr int NB = DP / 8;   // n-blocks of the output
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sk = sq + 64 * LDS;      // 2 stages
  __nv_bfloat16* sv = sk + 2 * 64 * LDS;  // 2 stages

  // grid = (heads, q tiles, batch).  Causal tiles are issued heaviest first (the last q tile
  // sees every key tile) across ALL heads, so the light tiles fill the tail of the launch.
  const int mt = p.causal ? static_cast<int>(gridDim.y) - 1 - static_cast<int>(blockIdx.y) : static_cast<int>(blockIdx.y);
  const int m0 = mt * kAM;
  const int h = blockIdx.x, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int D = p.d;
  const bool vec = p.vec != 0;

  const __nv_bfloat16* qg = p.q + b * p.q_bs + static_cast<long long>(m0) * p.q_rs + h * D;
  const __nv_bfloat16* kg = p.k + b * p.k_bs + h * D;
  const __nv_bfloat16* vg = p.v + b * p.v_bs + h * D;
  const int causal_off = p.skv - p.sq;

  int kv_end = p.skv;
  if (p.causal) {
    const int last = m0 + kAM - 1 + causal_off;  // last visible key for the last row
    kv_end = last + 1 < p.skv ? last + 1 : p.skv;
    if (kv_end < 0) kv_end = 0;
  }
  const int n_tiles = (kv_end + kAN - 1) / kAN;

  const int q_rows = p.sq - m0 < kAM ? p.sq - m0 : kAM;
  load_tile<DP>(sq, qg, p.q_rs, q_rows, D, vec);
  if (n_tiles > 0) {
    const int rows = p.skv < kAN ? p.skv : kAN;
    load_tile<DP>(sk, kg, p.k_rs, rows, D, vec);
    load_tile<DP>(sv, vg, p.v_rs, rows, D, vec);
  }
  cp_async_commit();

  float o_acc[NB][4];
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.0f;
  float row_m[2] = {-INFINITY, -INFINITY};
  float row_l[2] = {0.0f, 0.0f};
  uint32_t qf[KS][4];
  const bool warp_active = (m0 + warp * 16) < p.sq;

  for (int it = 0; it < n_tiles; ++it) {
    const int st = it & 1;
    const int n0 = it * kAN;
    // prefetch next K/V tile into the other stage
    if (it + 1 < n_tiles) {
      const int nn0 = n0 + kAN;
      const int rows = p.skv - nn0 < kAN ? p.skv - nn0 : kAN;
      load_tile<DP>(sk + (st ^ 1) * 64 * LDS, kg + static_cast<long long>(nn0) * p.k_rs, p.k_rs,
                    rows, D, vec);
      load_tile<DP>(sv + (st ^ 1) * 64 * LDS, vg + static_cast<long long>(nn0) * p.v_rs, p.v_rs,
                    rows, D, vec);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    if (it == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        ldsm_x4(qf[ks], sq + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
    }

    const __nv_bfloat16* ks_ = sk + st * 64 * LDS;
    const __nv_bfloat16* vs_ = sv + st * 64 * LDS;
    // 16-key groups of this tile that hold at least one in-range key (S=257: the 5th tile
    // has one) and whether this warp owns any in-range query row (the 5th query tile: one)
    const int np_cnt = (((kv_end - n0 < kAN) ? kv_end - n0 : kAN) + 15) >> 4;
    if (warp_active) {
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key blocks
        if (np >= np_cnt) continue;
        uint32_t kf[4];
        ldsm_x4(kf, ks_ + (np * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 +
                        ((lane >> 3) & 1) * 8);
        mma_bf16(s[2 * np], qf[ks], kf[0], kf[1]);
        mma_bf16(s[2 * np + 1], qf[ks], kf[2], kf[3]);
      }
    }

    // scale + mask (log2 domain)
    const int qi0 = m0 + warp * 16 + g;
    const uint8_t* km = p.key_mask != nullptr ? p.key_mask + static_cast<long long>(b) * p.skv : nullptr;
    constexpr float kLog2e = 1.4426950408889634f;
    // rb[key - qi]: entry (key - qi) + (sq - 1) of this head's relative-position table
    const float* rb = p.rel_bias != nullptr ? p.rel_bias + h * p.rel_bias_stride + (p.sq - 1) : nullptr;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = n0 + nb * 8 + 2 * t + (j & 1);
        const int qi = qi0 + (j >> 1) * 8;
        bool ok = key < p.skv;
        if (p.causal) ok = ok && (key <= qi + causal_off);
        if (km != nullptr && ok) ok = km[key] != 0;
        float sv = s[nb][j] * p.scale_log2;
        if (rb != nullptr && ok) sv += rb[key - qi] * kLog2e;
        s[nb][j] = ok ? sv : -INFINITY;
      }
    }
    // online softmax for the two rows this thread owns (g, g+8)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float mx = -INFINITY;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) mx = fmaxf(mx, fmaxf(s[nb][2 * r], s[nb][2 * r + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(row_m[r], mx);
      const float m_safe = (m_new == -INFINITY) ? 0.0f : m_new;
      const float corr = exp2f(row_m[r] - m_safe);  // row_m=-inf -> 0
      float sum = 0.0f;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const float p0 = exp2f(s[nb][2 * r] - m_safe);
        const float p1 = exp2f(s[nb][2 * r + 1] - m_safe);
        s[nb][2 * r] = p0;
        s[nb][2 * r + 1] = p1;
        sum += p0 + p1;
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      row_l[r] = row_l[r] * corr + sum;
      row_m[r] = m_new;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        o_acc[nb][2 * r] *= corr;
        o_acc[nb][2 * r + 1] *= corr;
      }
    }
    if (p.drop_thresh != 0u) {  // dropout on the probabilities; the normaliser stays undropped
      const uint64_t seed = *p.drop_seed + p.drop_salt;
      const uint64_t bh = static_cast<uint64_t>(b) * p.heads + h;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t key = n0 + nb * 8 + 2 * t + (j & 1);
          const uint64_t qi = qi0 + (j >> 1) * 8;
          const uint64_t idx = (bh * p.sq + qi) * p.skv + key;
          s[nb][j] = dropout_keep(seed, idx, p.drop_thresh) ? s[nb][j] * p.drop_scale : 0.0f;
        }
      }
    }
    // O += P . V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 16-key steps
      if (kk >= np_cnt) continue;
      uint32_t pf[4];
      pf[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < NB / 2; ++dp) {  // pairs of 8-wide d blocks
        uint32_t vf[4];
        ldsm_x4_t(vf, vs_ + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dp * 16 +
                          (lane >> 4) * 8);
        mma_bf16(o_acc[2 * dp], pf, vf[0], vf[1]);
        mma_bf16(o_acc[2 * dp + 1], pf, vf[2], vf[3]);
      }
    }
    }  // warp_active
    __syncthreads();  // everyone done with stage st before it is refilled
  }
  cp_async_wait<0>();

  // finalize
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qi = m0 + warp * 16 + g + r * 8;
    if (qi >= p.sq) continue;
    const float inv = row_l[r] > 0.0f ? 1.0f / row_l[r] : 0.0f;
    __nv_bfloat16* og = p.o + b * p.o_bs + static_cast<long long>(qi) * p.o_rs + h * D;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int c = nb * 8 + 2 * t;
      const float v0 = o_acc[nb][2 * r] * inv, v1 = o_acc[nb][2 * r + 1] * inv;
      if (c + 1 < D && p.o_vec2) {
        *reinterpret_cast<uint32_t*>(og + c) = pack_bf16x2(v0, v1);
      } else {
        if (c < D) og[c] = __float2bfloat16(v0);
        if (c + 1 < D) og[c + 1] = __float2bfloat16(v1);
      }
    }
    if (p.lse != nullptr && t == 0) {
      const float l2 = row_l[r] > 0.0f ? row_m[r] + log2f(row_l[r]) : -INFINITY;
      p.lse[(static_cast<long long>(b) * p.heads + h) * p.sq + qi] = l2 * 0.69314718055994531f;
    }
  }
}

template <int DP>
static cudaError_t launch_fwd(const AttnParams& p, int batch, cudaStream_t stream) {
  constexpr int smem = 5 * 64 * (DP + 8) * 2;
  static DeviceOnce attr_once;
  bool& attr = attr_once();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<DP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  dim3 grid(p.heads, (p.sq + kAM - 1) / kAM, batch);
  launch_pdl(attn_fwd_kernel<DP>, dim3(grid), dim3(kAttnThreads), smem, stream, p);
  return cudaGetLastError();
}

static bool attn_vec_ok(const vb_attn_args& a) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return a.d % 8 == 0 && al(a.q) && al(a.k) && al(a.v) && a.q_rs % 8 == 0 && a.k_rs % 8 == 0 &&
         a.v_rs % 8 == 0 && a.q_bs % 8 == 0 && a.k_bs % 8 == 0 && a.v_bs % 8 == 0;
}

cudaError_t attention_fwd_launch(const vb_attn_args& a, cudaStream_t stream) {
  if (a.batch <= 0 || a.heads <= 0 || a.sq <= 0) return cudaSuccess;
  if (a.d <= 0 || a.d > 128 || a.skv <= 0) return cudaErrorInvalidValue;
  if (attention_tcgen05_eligible(a)) return attention_tcgen05_launch(a, stream);
  if (attention_flash_tcgen05_eligible(a)) return attention_flash_tcgen05_launch(a, stream);
  if (a.heads > 65535 || a.batch > 65535) return cudaErrorInvalidValue;
  AttnParams p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(a.q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(a.k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(a.v);
  p.o = reinterpret_cast<__nv_bfloat16*>(a.o);
  p.lse = a.lse;
  p.key_mask = a.key_mask;
  p.sq = static_cast<int>(a.sq); p.skv = static_cast<int>(a.skv);
  p.d = static_cast<int>(a.d); p.heads = static_cast<int>(a.heads);
  p.q_bs = a.q_bs; p.q_rs = a.q_rs; p.k_bs = a.k_bs; p.k_rs = a.k_rs;
  p.v_bs = a.v_bs; p.v_rs = a.v_rs; p.o_bs = a.o_bs; p.o_rs = a.o_rs;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.causal = a.causal;
  {
    const bool drop = a.dropout_p > 0.0f && a.dropout_seed != nullptr;
    p.drop_seed = reinterpret_cast<const unsigned long long*>(a.dropout_seed);
    p.drop_salt = a.dropout_salt;
    p.drop_thresh = drop ? dropout_threshold(a.dropout_p) : 0u;
    p.drop_scale = drop ? 1.0f / (1.0f - a.dropout_p) : 1.0f;
  }
  p.rel_bias = a.rel_bias;
  p.rel_bias_stride = a.rel_bias_stride;
  p.vec = attn_vec_ok(a) ? 1 : 0;
  // o is written with 4-byte stores when every (row, head) start is 4-byte aligned
  p.o_vec2 = (a.d % 2 == 0 && a.o_rs % 2 == 0 && a.o_bs % 2 == 0 &&
              (reinterpret_cast<uintptr_t>(a.o) & 3u) == 0) ? 1 : 0;
  const int dp = static_cast<int>((a.d + 15) / 16 * 16);
  const int batch = static_cast<int>(a.batch);
  switch (dp) {
    case 16: return launch_fwd<16>(p, batch, stream);
    case 32: return launch_fwd<32>(p, batch, stream);
    case 48: return launch_fwd<48>(p, batch, stream);
    case 64: return launch_fwd<64>(p, batch, stream);
    case 80: return launch_fwd<80>(p, batch, stream);
    case 96: return launch_fwd<96>(p, batch, stream);
    case 112: return launch_fwd<112>(p, batch, stream);
    case 128: return launch_fwd<128>(p, batch, stream);
    default: return cudaErrorInvalidValue;
  }
}


// =================================================================== backward
// One CTA owns 64 keys of one (batch, head): K_j, V_j stay in shared memory while the
// CTA walks the query tiles that can see them.  Per query tile:
//   S = Q K^T, P = exp(S*scale - lse), dP = dO V^T, dS = P o (dP - delta)
//   dV += P^T dO, dK += dS^T Q (register accumulators, keys x d)
//   dQ += dS K   (fp32 atomics into dq_acc; converted to bf16 afterwards)
struct AttnBwdParams {
  AttnParams f;
  const __nv_bfloat16* d_o;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  long long dk_bs, dk_rs, dv_bs, dv_rs;
  const float* delta;
  float* dq_acc;
  float scale;
  int kt_per_cta;   // key tiles walked by one CTA (> 1 only with a single, non-causal query tile)
};

__global__ void __launch_bounds__(128)
attn_delta_kernel(const __nv_bfloat16* o, const __nv_bfloat16* d_o, float* delta, int sq, int heads,
                  int D, long long o_bs, long long o_rs, long long total) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= total) return;
  const int qi = static_cast<int>(wid % sq);
  const int h = static_cast<int>((wid / sq) % heads);
  const long long b = wid / (static_cast<long long>(sq) * heads);
  const long long off = b * o_bs + static_cast<long long>(qi) * o_rs + h * D;
  float acc = 0.0f;
  for (int c = lane; c < D; c += 32)
    acc += __bfloat162float(o[off + c]) * __bfloat162float(d_o[off + c]);
  acc = warp_sum(acc);
  if (lane == 0) delta[wid] = acc;
}

// KEEP_DQ: the CTA walks several key tiles and keeps dQ in registers (single query tile, not causal); a
// template parameter so that the general instantiation pays no registers for it.
template <int DP, bool KEEP_DQ>
__global__ void __launch_bounds__(kAttnThreads) attn_bwd_kernel(const AttnBwdParams bp) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const AttnParams& p = bp.f;
  constexpr int LDS = DP + 8;
  constexpr int LDP = 72;
  constexpr int KS = DP / 16;
  constexpr int NB = DP / 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sK0 = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sV0 = sK0 + 64 * LDS;
  __nv_bfloat16* sQ = sV0 + 64 * LDS;
  __nv_bfloat16* sdO = sQ + 64 * LDS;
  __nv_bfloat16* sP = sdO + 64 * LDS;
  __nv_bfloat16* sdS = sP + 64 * LDP;
  // KEEP_DQ: second K / V buffers, the next key tile is fetched while the current one is processed
  __nv_bfloat16* sK1 = sdS + 64 * LDP;
  __nv_bfloat16* sV1 = sK1 + 64 * LDS;

  // grid = (heads, key-tile groups, batch): with a causal mask key tile 0 is the heaviest, and it
  // is issued first for every head.  A group is ONE key tile in general; when all queries fit one tile
  // (the Q-Former's 32 queries against 2 056 image tokens) a CTA walks `kt_per_cta` key tiles and keeps dQ
  // in registers across them: one round of f32 atomics per group instead of one per key tile (33 CTAs used
  // to hit every dQ element).
  const int h = blockIdx.x, b = blockIdx.z;
  const int key_tiles = (p.skv + kAN - 1) / kAN;
  const int kt_begin = blockIdx.y * bp.kt_per_cta;
  const int kt_end = kt_begin + bp.kt_per_cta < key_tiles ? kt_begin + bp.kt_per_cta : key_tiles;
  // KEEP_DQ: the host guarantees a single query tile, not causal
  float dq_sum[KEEP_DQ ? NB : 1][4];
#pragma unroll
  for (int i = 0; i < (KEEP_DQ ? NB : 1); ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq_sum[i][j] = 0.0f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int D = p.d;
  const bool vec = p.vec != 0;
  const int causal_off = p.skv - p.sq;
  constexpr float kLog2e = 1.4426950408889634f;
  auto load_kv = [&](int tile, __nv_bfloat16* dk_s, __nv_bfloat16* dv_s) {
    const int n_first = tile * kAN;
    const int rows_kv = p.skv - n_first < kAN ? p.skv - n_first : kAN;
    load_tile<DP>(dk_s, p.k + b * p.k_bs + static_cast<long long>(n_first) * p.k_rs + h * D, p.k_rs, rows_kv, D, vec);
    load_tile<DP>(dv_s, p.v + b * p.v_bs + static_cast<long long>(n_first) * p.v_rs + h * D, p.v_rs, rows_kv, D, vec);
  };
  if constexpr (KEEP_DQ) {  // the single Q / dO tile and the first key tile, once
    const int q_rows0 = p.sq < kAM ? p.sq : kAM;
    load_kv(kt_begin, sK0, sV0);
    load_tile<DP>(sQ, p.q + b * p.q_bs + h * D, p.q_rs, q_rows0, D, vec);
    load_tile<DP>(sdO, bp.d_o + b * p.o_bs + h * D, p.o_rs, q_rows0, D, vec);
    cp_async_commit();
  }
  for (int kt = kt_begin; kt < kt_end; ++kt) {
  const int n0 = kt * kAN;
  const bool odd = KEEP_DQ && (((kt - kt_begin) & 1) != 0);
  __nv_bfloat16* sK = odd ? sK1 : sK0;
  __nv_bfloat16* sV = odd ? sV1 : sV0;
  if constexpr (KEEP_DQ) {
    cp_async_wait<0>();
    __syncthreads();  // this tile has landed; the other buffer's readers finished at the end of the last tile
    if (kt + 1 < kt_end) {
      load_kv(kt + 1, odd ? sK0 : sK1, odd ? sV0 : sV1);
      cp_async_commit();
    }
  } else {
    load_kv(kt, sK, sV);
    cp_async_commit();
  }

  float dk_acc[NB][4], dv_acc[NB][4];
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { dk_acc[i][j] = 0.0f; dv_acc[i][j] = 0.0f; }

  int m_start = 0;
  if (p.causal) {
    int first_q = n0 - causal_off;
    if (first_q < 0) first_q = 0;
    m_start = first_q / kAM;
  }
  const int m_tiles = (p.sq + kAM - 1) / kAM;
  const uint8_t* km = p.key_mask != nullptr ? p.key_mask + static_cast<long long>(b) * p.skv : nullptr;
  const float* rb = p.rel_bias != nullptr ? p.rel_bias + h * p.rel_bias_stride + (p.sq - 1) : nullptr;
  const float* lse_bh = p.lse + (static_cast<long long>(b) * p.heads + h) * p.sq;
  const float* delta_bh = bp.delta + (static_cast<long long>(b) * p.heads + h) * p.sq;

  for (int mt = m_start; mt < m_tiles; ++mt) {
    const int m0 = mt * kAM;
    const int q_rows = p.sq - m0 < kAM ? p.sq - m0 : kAM;
    if constexpr (!KEEP_DQ) {
      load_tile<DP>(sQ, p.q + b * p.q_bs + static_cast<long long>(m0) * p.q_rs + h * D, p.q_rs,
                    q_rows, D, vec);
      load_tile<DP>(sdO, bp.d_o + b * p.o_bs + static_cast<long long>(m0) * p.o_rs + h * D, p.o_rs,
                    q_rows, D, vec);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
    }

    // ---- S and dP for this warp's 16 query rows
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[i][j] = 0.0f; dp[i][j] = 0.0f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t qf[4], of[4];
      ldsm_x4(qf, sQ + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
      ldsm_x4(of, sdO + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t kf[4], vf[4];
        const int off = (np * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(kf, sK + off);
        ldsm_x4(vf, sV + off);
        mma_bf16(s[2 * np], qf, kf[0], kf[1]);
        mma_bf16(s[2 * np + 1], qf, kf[2], kf[3]);
        mma_bf16(dp[2 * np], of, vf[0], vf[1]);
        mma_bf16(dp[2 * np + 1], of, vf[2], vf[3]);
      }
    }
    // ---- P, dS
    float lse2[2], dl[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qi = m0 + warp * 16 + g + r * 8;
      if (qi < p.sq) {
        lse2[r] = lse_bh[qi] * kLog2e;
        dl[r] = delta_bh[qi];
      } else {
        lse2[r] = -INFINITY;
        dl[r] = 0.0f;
      }
    }
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = j >> 1;
        const int key = n0 + nb * 8 + 2 * t + (j & 1);
        const int qi = m0 + warp * 16 + g + r * 8;
        bool ok = key < p.skv && qi < p.sq && lse2[r] != -INFINITY;
        if (p.causal) ok = ok && (key <= qi + causal_off);
        if (km != nullptr && ok) ok = km[key] != 0;
        float sv = s[nb][j] * p.scale_log2;
        if (rb != nullptr && ok) sv += rb[key - qi] * kLog2e;
        const float pr = ok ? exp2f(sv - lse2[r]) : 0.0f;
        float p_used = pr, dpv = dp[nb][j];
        if (p.drop_thresh != 0u) {  // regenerate the forward mask
          const uint64_t idx = ((static_cast<uint64_t>(b) * p.heads + h) * p.sq + qi) * p.skv + key;
          const bool keep = dropout_keep(*p.drop_seed + p.drop_salt, idx, p.drop_thresh);
          p_used = keep ? pr * p.drop_scale : 0.0f;
          dpv = keep ? dpv * p.drop_scale : 0.0f;
        }
        s[nb][j] = p_used;                 // P actually multiplied with V in the forward -> dV
        dp[nb][j] = pr * (dpv - dl[r]);    // dS
      }
    }
    // stash P and dS (bf16) for the transposed products
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = warp * 16 + g + r * 8;
        const int col = nb * 8 + 2 * t;
        *reinterpret_cast<uint32_t*>(sP + row * LDP + col) = pack_bf16x2(s[nb][2 * r], s[nb][2 * r + 1]);
        *reinterpret_cast<uint32_t*>(sdS + row * LDP + col) = pack_bf16x2(dp[nb][2 * r], dp[nb][2 * r + 1]);
      }
    }
    // ---- dQ = dS . K  (this warp's 16 rows), fp32 atomics
    {
      float dq[NB][4];
#pragma unroll
      for (int i = 0; i < NB; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dq[i][j] = 0.0f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t af[4];
        af[0] = pack_bf16x2(dp[2 * kk][0], dp[2 * kk][1]);
        af[1] = pack_bf16x2(dp[2 * kk][2], dp[2 * kk][3]);
        af[2] = pack_bf16x2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
        af[3] = pack_bf16x2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
        for (int dpi = 0; dpi < NB / 2; ++dpi) {
          uint32_t kf[4];
          ldsm_x4_t(kf, sK + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dpi * 16 +
                            (lane >> 4) * 8);
          mma_bf16(dq[2 * dpi], af, kf[0], kf[1]);
          mma_bf16(dq[2 * dpi + 1], af, kf[2], kf[3]);
        }
      }
      if constexpr (KEEP_DQ) {
#pragma unroll
        for (int i = 0; i < NB; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dq_sum[i][j] += dq[i][j];
      } else {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int qi = m0 + warp * 16 + g + r * 8;
          if (qi >= p.sq) continue;
          float* dst = bp.dq_acc + (static_cast<long long>(b) * p.sq + qi) * (static_cast<long long>(p.heads) * D) + h * D;
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const int c = nb * 8 + 2 * t;
            if (c < D) atomicAdd(dst + c, dq[nb][2 * r] * bp.scale);
            if (c + 1 < D) atomicAdd(dst + c + 1, dq[nb][2 * r + 1] * bp.scale);
          }
        }
      }
    }
    __syncthreads();
    // ---- dV += P^T dO ; dK += dS^T Q   (this warp's 16 keys)
#pragma unroll
    for (int qk = 0; qk < 4; ++qk) {  // 16-query steps
      uint32_t pf[4], sf[4];
      const int aoff = (qk * 16 + (lane & 7) + (lane >> 4) * 8) * LDP + warp * 16 + ((lane >> 3) & 1) * 8;
      ldsm_x4_t(pf, sP + aoff);
      ldsm_x4_t(sf, sdS + aoff);
#pragma unroll
      for (int dpi = 0; dpi < NB / 2; ++dpi) {
        uint32_t of[4], qf[4];
        const int boff = (qk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dpi * 16 + (lane >> 4) * 8;
        ldsm_x4_t(of, sdO + boff);
        ldsm_x4_t(qf, sQ + boff);
        mma_bf16(dv_acc[2 * dpi], pf, of[0], of[1]);
        mma_bf16(dv_acc[2 * dpi + 1], pf, of[2], of[3]);
        mma_bf16(dk_acc[2 * dpi], sf, qf[0], qf[1]);
        mma_bf16(dk_acc[2 * dpi + 1], sf, qf[2], qf[3]);
      }
    }
    __syncthreads();
  }
  if constexpr (!KEEP_DQ) cp_async_wait<0>();

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = n0 + warp * 16 + g + r * 8;
    if (key >= p.skv) continue;
    __nv_bfloat16* dkg = bp.dk + b * bp.dk_bs + static_cast<long long>(key) * bp.dk_rs + h * D;
    __nv_bfloat16* dvg = bp.dv + b * bp.dv_bs + static_cast<long long>(key) * bp.dv_rs + h * D;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int c = nb * 8 + 2 * t;
      if (c < D) {
        dkg[c] = __float2bfloat16(dk_acc[nb][2 * r] * bp.scale);
        dvg[c] = __float2bfloat16(dv_acc[nb][2 * r]);
      }
      if (c + 1 < D) {
        dkg[c + 1] = __float2bfloat16(dk_acc[nb][2 * r + 1] * bp.scale);
        dvg[c + 1] = __float2bfloat16(dv_acc[nb][2 * r + 1]);
      }
    }
  }
  __syncthreads();  // the K / V tiles are reloaded by the next key tile of this CTA
  }  // key tiles of this CTA
  if constexpr (KEEP_DQ) {  // single query tile (m0 = 0): the group's dQ in one round of atomics
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qi = warp * 16 + g + r * 8;
      if (qi >= p.sq) continue;
      float* dst = bp.dq_acc + (static_cast<long long>(b) * p.sq + qi) * (static_cast<long long>(p.heads) * D) + h * D;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int c = nb * 8 + 2 * t;
        if (c < D) atomicAdd(dst + c, dq_sum[nb][2 * r] * bp.scale);
        if (c + 1 < D) atomicAdd(dst + c + 1, dq_sum[nb][2 * r + 1] * bp.scale);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
attn_dq_convert_kernel(const float* dq_acc, __nv_bfloat16* dq, int sq, int hd, long long dq_bs,
                       long long dq_rs, float mul, long long total) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long c = i % hd;
    const long long srow = (i / hd) % sq;
    const long long b = i / (static_cast<long long>(hd) * sq);
    dq[b * dq_bs + srow * dq_rs + c] = __float2bfloat16(dq_acc[i] * mul);
  }
}

template <int DP>
static cudaError_t launch_bwd(const AttnBwdParams& bp, int batch, cudaStream_t stream) {
  constexpr int smem = (4 * 64 * (DP + 8) + 2 * 64 * 72) * 2;
  constexpr int smem_keep = smem + 2 * 64 * (DP + 8) * 2;  // + the second K / V buffers
  static DeviceOnce attr_once;
  bool& attr = attr_once();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<DP, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_kernel<DP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_keep);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int key_tiles = (bp.f.skv + kAN - 1) / kAN;
  dim3 grid(bp.f.heads, (key_tiles + bp.kt_per_cta - 1) / bp.kt_per_cta, batch);
  if (bp.kt_per_cta > 1) launch_pdl(attn_bwd_kernel<DP, true>, dim3(grid), dim3(kAttnThreads), smem_keep, stream, bp);
  else launch_pdl(attn_bwd_kernel<DP, false>, dim3(grid), dim3(kAttnThreads), smem, stream, bp);
  return cudaGetLastError();
}

cudaError_t attention_bwd_launch(const vb_attn_bwd_args& a, cudaStream_t stream) {
  const vb_attn_args& f = a.fwd;
  if (f.batch <= 0 || f.heads <= 0 || f.sq <= 0 || f.skv <= 0) return cudaSuccess;
  if (f.d <= 0 || f.d > 128) return cudaErrorInvalidValue;
  if (f.lse == nullptr || a.delta == nullptr || a.dq_acc == nullptr) return cudaErrorInvalidValue;
  if (f.heads > 65535 || f.batch > 65535) return cudaErrorInvalidValue;
  AttnBwdParams bp;
  AttnParams& p = bp.f;
  p.q = reinterpret_cast<const __nv_bfloat16*>(f.q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(f.k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(f.v);
  p.o = reinterpret_cast<__nv_bfloat16*>(f.o);
  p.lse = f.lse;
  p.key_mask = f.key_mask;
  p.sq = static_cast<int>(f.sq); p.skv = static_cast<int>(f.skv);
  p.d = static_cast<int>(f.d); p.heads = static_cast<int>(f.heads);
  p.q_bs = f.q_bs; p.q_rs = f.q_rs; p.k_bs = f.k_bs; p.k_rs = f.k_rs;
  p.v_bs = f.v_bs; p.v_rs = f.v_rs; p.o_bs = f.o_bs; p.o_rs = f.o_rs;
  p.scale_log2 = f.scale * 1.4426950408889634f;
  p.causal = f.causal;
  {
    const bool drop = f.dropout_p > 0.0f && f.dropout_seed != nullptr;
    p.drop_seed = reinterpret_cast<const unsigned long long*>(f.dropout_seed);
    p.drop_salt = f.dropout_salt;
    p.drop_thresh = drop ? dropout_threshold(f.dropout_p) : 0u;
    p.drop_scale = drop ? 1.0f / (1.0f - f.dropout_p) : 1.0f;
  }
  p.rel_bias = f.rel_bias;
  p.rel_bias_stride = f.rel_bias_stride;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  p.vec = (attn_vec_ok(f) && al(a.d_o) && f.o_rs % 8 == 0 && f.o_bs % 8 == 0) ? 1 : 0;
  p.o_vec2 = 0;
  bp.d_o = reinterpret_cast<const __nv_bfloat16*>(a.d_o);
  bp.dk = reinterpret_cast<__nv_bfloat16*>(a.dk);
  bp.dv = reinterpret_cast<__nv_bfloat16*>(a.dv);
  bp.dk_bs = a.dk_bs; bp.dk_rs = a.dk_rs; bp.dv_bs = a.dv_bs; bp.dv_rs = a.dv_rs;
  bp.delta = a.delta;
  bp.dq_acc = a.dq_acc;
  bp.scale = f.scale;
  // cross-attention of a few queries over a long memory (Q-Former: 32 x 2 056): one CTA per 8 key tiles,
  // dQ accumulated in registers; VB_ATTN_BWD_KT overrides (1 = one key tile per CTA, the general scheme)
  bp.kt_per_cta = 1;
  if (f.sq <= kAM && !f.causal && f.skv >= 8 * kAN) {
    static const int kt = [] {
      const char* e = std::getenv("VB_ATTN_BWD_KT");
      const int v = e != nullptr ? std::atoi(e) : 8;
      return v < 1 ? 1 : v;
    }();
    bp.kt_per_cta = kt;
  }

  // tcgen05 path (attention_flash_tcgen05.cu): dQ and dK / dV by two passes with the reduction index on the TMEM
  // columns; no dq_acc round trip, delta computed inside the dQ pass
  if (attention_bwd_tcgen05_eligible(a)) return attention_bwd_tcgen05_launch(a, stream);
  const long long rows = f.batch * f.heads * f.sq;
  launch_pdl(attn_delta_kernel, dim3(static_cast<unsigned>((rows * 32 + 127) / 128)), dim3(128), 0, stream, 
      p.o, bp.d_o, a.delta, p.sq, p.heads, p.d, p.o_bs, p.o_rs, rows);
  const long long hd = f.heads * f.d;
  const long long total = f.batch * f.sq * hd;
  cudaError_t e = cudaMemsetAsync(a.dq_acc, 0, sizeof(float) * total, stream);
  if (e != cudaSuccess) return e;
  const int dp = static_cast<int>((f.d + 15) / 16 * 16);
  const int batch = static_cast<int>(f.batch);
  switch (dp) {
    case 16: e = launch_bwd<16>(bp, batch, stream); break;
    case 32: e = launch_bwd<32>(bp, batch, stream); break;
    case 48: e = launch_bwd<48>(bp, batch, stream); break;
    case 64: e = launch_bwd<64>(bp, batch, stream); break;
    case 80: e = launch_bwd<80>(bp, batch, stream); break;
    case 96: e = launch_bwd<96>(bp, batch, stream); break;
    case 112: e = launch_bwd<112>(bp, batch, stream); break;
    case 128: e = launch_bwd<128>(bp, batch, stream); break;
    default: return cudaErrorInvalidValue;
  }
  if (e != cudaSuccess) return e;
  const unsigned grid = static_cast<unsigned>(total / 256 + 1 < 148 * 8 ? total / 256 + 1 : 148 * 8);
  launch_pdl(attn_dq_convert_kernel, dim3(grid), dim3(256), 0, stream, a.dq_acc, reinterpret_cast<__nv_bfloat16*>(a.dq),
                                                   p.sq, static_cast<int>(hd), a.dq_bs, a.dq_rs,
                                                   a.dq_scale == 0.0f ? 1.0f : a.dq_scale, total);
  return cudaGetLastError();
}


// ------------------------------------------------------------------ attention maps (diagnostic output)
// probs[b, h, i, :] = softmax_j(scale * q_i . k_j (+ causal / key-padding mask)) written out in full — the
// `attentions` the reference returns with output_attentions=True (eilev/model/v2.py:87-95;
// HF:blip_2/modeling_blip_2.py:319-353).  The fused kernels never materialise these maps; this plain CUDA-core
// kernel (one CTA per query row, keys dealt to the threads) exists only for that optional output and is not on
// the training / generation path.
struct ProbsParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const uint8_t* key_mask;
  int sq, skv, d, heads;
  long long q_bs, q_rs, k_bs, k_rs;
  float scale;
  int causal;
};

__global__ void __launch_bounds__(128) attn_probs_kernel(const ProbsParams p, void* probs, int out_bf16) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sh[];            // [d] q row, then [skv] scores
  float* sq = sh;
  float* sc = sh + p.d;
  __shared__ float red[4];
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const __nv_bfloat16* q = p.q + static_cast<long long>(b) * p.q_bs + static_cast<long long>(i) * p.q_rs + h * p.d;
  for (int c = tid; c < p.d; c += 128) sq[c] = __bfloat162float(q[c]);
  __syncthreads();
  const int limit = p.causal ? (i + (p.skv - p.sq) + 1) : p.skv;  // keys j < limit are visible
  float mx = -INFINITY;
  for (int j = tid; j < p.skv; j += 128) {
    float s = -INFINITY;
    const bool ok = j < limit && (p.key_mask == nullptr || p.key_mask[static_cast<long long>(b) * p.skv + j] != 0);
    if (ok) {
      const __nv_bfloat16* k = p.k + static_cast<long long>(b) * p.k_bs + static_cast<long long>(j) * p.k_rs + h * p.d;
      float acc = 0.0f;
      for (int c = 0; c < p.d; ++c) acc = fmaf(sq[c], __bfloat162float(k[c]), acc);
      s = acc * p.scale;
    }
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.0f;
  for (int j = tid; j < p.skv; j += 128) {
    const float e = sc[j] == -INFINITY ? 0.0f : __expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  sum = (red[0] + red[1]) + (red[2] + red[3]);
  const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
  const long long base = ((static_cast<long long>(b) * p.heads + h) * p.sq + i) * p.skv;
  for (int j = tid; j < p.skv; j += 128) {
    const float v = sc[j] * inv;
    if (out_bf16) reinterpret_cast<__nv_bfloat16*>(probs)[base + j] = __float2bfloat16(v);
    else reinterpret_cast<float*>(probs)[base + j] = v;
  }
}

cudaError_t attention_probs_launch(const vb_attn_args& a, void* probs, int out_bf16, cudaStream_t stream) {
  if (a.batch <= 0 || a.sq <= 0 || a.skv <= 0) return cudaSuccess;
  if (a.heads > 65535 || a.batch > 65535) return cudaErrorInvalidValue;
  ProbsParams p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(a.q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(a.k);
  p.key_mask = a.key_mask;
  p.sq = static_cast<int>(a.sq); p.skv = static_cast<int>(a.skv);
  p.d = static_cast<int>(a.d); p.heads = static_cast<int>(a.heads);
  p.q_bs = a.q_bs; p.q_rs = a.q_rs; p.k_bs = a.k_bs; p.k_rs = a.k_rs;
  p.scale = a.scale;
  p.causal = a.causal;
  const size_t smem = static_cast<size_t>(a.d + a.skv) * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  const dim3 grid(static_cast<unsigned>(a.sq), static_cast<unsigned>(a.heads), static_cast<unsigned>(a.batch));
  return launch_pdl(attn_probs_kernel, grid, dim3(128), smem, stream, p, probs, out_bf16);
}

}  // namespace vb
