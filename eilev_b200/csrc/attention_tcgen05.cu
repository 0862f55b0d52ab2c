// Non-causal self-attention for the ViT tower on the 5th-gen tensor cores.
//
// Shape class: one (frame, head) has S <= 272 keys (ViT-g: 257) and head dim d <= 128
// (ViT-g: 88), so the whole score row fits TMEM and the softmax is single-pass — no online
// rescaling.  One persistent CTA per SM walks (frame, head) items:
//
//   warp 0   TMA producer: Q tile (128 rows), K, V via 3-D tensor maps (d, head, token);
//            d is the innermost dim with extent 88, so the 64-wide 128B-swizzled boxes are
//            zero-filled beyond the head (no padding pass, no transposes).
//   warp 1   MMA issuer:  S = Q K^T   (tcgen05.mma SS, M=128, N=256 (+16), K-major operands)
//                         O = P V     (tcgen05.mma TS: P is read from TMEM where it aliases S,
//                                      V is the MN-major B operand straight from its row layout)
//   warps 4-11 softmax + epilogue: thread = query row (TMEM lane), two warps per lane quarter
//            split the columns: tcgen05.ld the row, max, exp2, sum (combined through smem),
//            bf16 P back into TMEM (tcgen05.st), later O * 1/sum -> global.
//
// TMEM columns: [0,272) S (fp32) / [0,136) P (bf16x2, in place), [288,416) O.
#include <cstdlib>

#include "common.cuh"
#include "internal.h"
#include "tc_attention.cuh"

namespace vb {

constexpr int kTaThreads = 384;
constexpr int kTaQRows = 128;
constexpr int kTaKRows = 272;            // keys padded to a multiple of 16
constexpr int kTaHalf = 136;             // K/V are loaded as two boxes of 136 rows
constexpr int kTaChunkBytesQ = kTaQRows * 128;   // one 64-wide d chunk of a Q tile
constexpr int kTaChunkBytesK = kTaKRows * 128;   // one 64-wide d chunk of K or V
constexpr int kTaSmem = 2 * 2 * kTaChunkBytesQ + 2 * 2 * kTaChunkBytesK + 1024 + 128 + 1024 + 64;
constexpr uint32_t kTaColO = 288;

struct TaParams {
  __nv_bfloat16* o;
  long long o_rs;      // row stride of o (elements); rows are batch*S contiguous
  int items;           // batch * heads
  int heads, s, d;
  float scale_log2;
  int pv_n;            // N of the P.V instruction (128, or d rounded up to 16)
};

__global__ void __launch_bounds__(kTaThreads, 1)
attn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const TaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                               // [2 buffers][2 chunks][128 rows][128 B]
  uint8_t* sK = sQ + 4 * kTaChunkBytesQ;            // [2 chunks][272 rows][128 B]
  uint8_t* sV = sK + 2 * kTaChunkBytesK;            // [2 chunks][272 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * kTaChunkBytesK);
  uint64_t* q_full = bars;        // [2]
  uint64_t* q_empty = bars + 2;   // [2]
  uint64_t* k_full = bars + 4;
  uint64_t* k_empty = bars + 5;
  uint64_t* v_full = bars + 6;
  uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 8;
  uint64_t* p_ready = bars + 9;
  uint64_t* o_full = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.s + kTaQRows - 1) / kTaQRows;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 8);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // programmatic dependent launch: q / k / v of the preceding GEMM are visible from here on
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {  // (not lane == 0: see gemm_tcgen05_2cta.cu)
      uint32_t k_ph = 0, v_ph = 0, q_ph[2] = {0, 0};
      int qt = 0;  // running Q tile counter -> buffer = qt & 1
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int b = item / p.heads, h = item % p.heads;
        const int row0 = b * p.s;
        auto load_q = [&](int t) {
          const int buf = qt & 1;
          mbar_wait(&q_empty[buf], q_ph[buf] ^ 1u);
          q_ph[buf] ^= 1u;
          mbar_expect_tx(&q_full[buf], 2 * kTaChunkBytesQ);
          uint8_t* dst = sQ + buf * 2 * kTaChunkBytesQ;
          tma_load_3d(dst, &tmap_q, &q_full[buf], 0, h, row0 + t * kTaQRows);
          tma_load_3d(dst + kTaChunkBytesQ, &tmap_q, &q_full[buf], 64, h, row0 + t * kTaQRows);
          ++qt;
        };
        mbar_wait(k_empty, k_ph ^ 1u);
        k_ph ^= 1u;
        mbar_expect_tx(k_full, 2 * kTaChunkBytesK);
        for (int c = 0; c < 2; ++c)
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(sK + c * kTaChunkBytesK + hf * kTaHalf * 128, &tmap_k, k_full, c * 64, h,
                        row0 + hf * kTaHalf);
        load_q(0);
        if (m_tiles > 1) load_q(1);
        mbar_wait(v_empty, v_ph ^ 1u);
        v_ph ^= 1u;
        mbar_expect_tx(v_full, 2 * kTaChunkBytesK);
        for (int c = 0; c < 2; ++c)
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(sV + c * kTaChunkBytesK + hf * kTaHalf * 128, &tmap_v, v_full, c * 64, h,
                        row0 + hf * kTaHalf);
        for (int t = 2; t < m_tiles; ++t) load_q(t);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s256 = umma_idesc_bf16(128, 256);
      const uint32_t idesc_s16 = umma_idesc_bf16(128, 16);
      // B is MN-major; pv_n = 128, or round_up(d, 16) (partial 64-element swizzle atoms) when p.pv_n is set
      const uint32_t idesc_o = umma_idesc_bf16(128, static_cast<uint32_t>(p.pv_n)) | (1u << 16);
      const int k_steps = (p.d + 15) / 16;            // 16-wide steps along d (6 for d = 88)
      const int kv_steps = (p.s + 15) / 16;           // 16-key steps of P.V (17 for S = 257)
      const bool tail16 = p.s > 256;
      uint32_t k_ph = 0, v_ph = 0, q_ph[2] = {0, 0}, p_ph = 0;
      int qt = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        mbar_wait(k_full, k_ph);
        k_ph ^= 1u;
        for (int t = 0; t < m_tiles; ++t) {
          const int buf = qt & 1;
          ++qt;
          mbar_wait(&q_full[buf], q_ph[buf]);
          q_ph[buf] ^= 1u;
          tc_fence_after();
          // ---- S = Q K^T
          for (int ks = 0; ks < k_steps; ++ks) {
            const int c = ks >> 2, kk = ks & 3;
            const uint64_t a_desc =
                umma_desc_k_sw128(smem_u32(sQ + buf * 2 * kTaChunkBytesQ + c * kTaChunkBytesQ)) + 2 * kk;
            const uint64_t b_desc = umma_desc_k_sw128(smem_u32(sK + c * kTaChunkBytesK)) + 2 * kk;
            umma_bf16(tmem_base, a_desc, b_desc, idesc_s256, ks != 0 ? 1u : 0u);
            if (tail16) {
              const uint64_t b16 = umma_desc_k_sw128(smem_u32(sK + c * kTaChunkBytesK + 256 * 128)) + 2 * kk;
              umma_bf16(tmem_base + 256, a_desc, b16, idesc_s16, ks != 0 ? 1u : 0u);
            }
          }
          umma_commit(&q_empty[buf]);
          if (t == m_tiles - 1) umma_commit(k_empty);
          umma_commit(s_full);
          // ---- O = P V  (P written by the softmax warps into TMEM columns [0, 136))
          if (t == 0) {
            mbar_wait(v_full, v_ph);
            v_ph ^= 1u;
          }
          mbar_wait(p_ready, p_ph);
          p_ph ^= 1u;
          tc_fence_after();
          for (int js = 0; js < kv_steps; ++js) {
            const uint64_t b_desc = umma_desc_mn_sw128(smem_u32(sV + js * 16 * 128), kTaChunkBytesK);
            umma_bf16_ts(tmem_base + kTaColO, tmem_base + js * 8, b_desc, idesc_o, js != 0 ? 1u : 0u);
          }
          if (t == m_tiles - 1) umma_commit(v_empty);
          umma_commit(o_full);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax + epilogue
    // Two warps share every TMEM lane quarter and split the score columns: `half` 0 owns the
    // 32-column chunks [0, split), half 1 owns [split, n_chunks).  Row max / row sum are
    // combined through shared memory with a 64-thread named barrier per quarter.
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r_in_tile = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t s_ph = 0, o_ph = 0;
    const int n_chunks = (p.s + 31) / 32;  // 32-column chunks holding in-range keys (9 for 257)
    const int split = (n_chunks + 1) / 2;
    const int c_lo = half == 0 ? 0 : split;
    const int c_hi = half == 0 ? split : n_chunks;
    const int my_chunks = c_hi - c_lo;     // <= 5
    float* xch = reinterpret_cast<float*>(bars + 16);  // [2 halves][128 rows] exchange buffer
    const uint32_t bar_id = 1 + quarter;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int b = item / p.heads, h = item % p.heads;
      for (int t = 0; t < m_tiles; ++t) {
        const int qi = t * kTaQRows + r_in_tile;
        const bool warp_has_rows = (t * kTaQRows + quarter * 32) < p.s;
        mbar_wait(s_full, s_ph);
        s_ph ^= 1u;
        tc_fence_after();
        float inv_sum = 0.0f;
        if (warp_has_rows) {
          // ---- pass 1: partial row maximum over this warp's columns
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
          for (int ch = c_lo; ch < c_hi; ++ch) {
            uint32_t r[32];
            tmem_ld_32(t_lane + ch * 32, r);
            tmem_ld_wait();
            if (ch * 32 + 32 <= p.s) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                m0 = fmaxf(m0, __uint_as_float(r[j]));
                m1 = fmaxf(m1, __uint_as_float(r[j + 1]));
                m2 = fmaxf(m2, __uint_as_float(r[j + 2]));
                m3 = fmaxf(m3, __uint_as_float(r[j + 3]));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (ch * 32 + j < p.s) m0 = fmaxf(m0, __uint_as_float(r[j]));
            }
          }
          float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          xch[half * 128 + r_in_tile] = mx;
          asm volatile("bar.sync %0, 64;\n" ::"r"(bar_id) : "memory");
          mx = fmaxf(mx, xch[(half ^ 1) * 128 + r_in_tile]);
          const float mxs = mx * p.scale_log2;
          // ---- pass 2: p = 2^(s*c - max*c) packed to bf16 pairs (P column = S column / 2)
          uint32_t pk[5][16];
          float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
          for (int ci = 0; ci < 5; ++ci) {
            if (ci < my_chunks) {
              const int ch = c_lo + ci;
              uint32_t r[32];
              tmem_ld_32(t_lane + ch * 32, r);
              tmem_ld_wait();
              if (ch * 32 + 32 <= p.s) {
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                  const float p0 = exp2f(fmaf(__uint_as_float(r[2 * j]), p.scale_log2, -mxs));
                  const float p1 = exp2f(fmaf(__uint_as_float(r[2 * j + 1]), p.scale_log2, -mxs));
                  const float p2 = exp2f(fmaf(__uint_as_float(r[2 * j + 2]), p.scale_log2, -mxs));
                  const float p3 = exp2f(fmaf(__uint_as_float(r[2 * j + 3]), p.scale_log2, -mxs));
                  s0 += p0; s1 += p1; s2 += p2; s3 += p3;
                  pk[ci][j] = pack_bf16x2(p0, p1);
                  pk[ci][j + 1] = pack_bf16x2(p2, p3);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int k0 = ch * 32 + 2 * j;
                  const float p0 = k0 < p.s ? exp2f(fmaf(__uint_as_float(r[2 * j]), p.scale_log2, -mxs)) : 0.0f;
                  const float p1 = k0 + 1 < p.s ? exp2f(fmaf(__uint_as_float(r[2 * j + 1]), p.scale_log2, -mxs)) : 0.0f;
                  s0 += p0; s1 += p1;
                  pk[ci][j] = pack_bf16x2(p0, p1);
                }
              }
            }
          }
          float sum = (s0 + s1) + (s2 + s3);
          asm volatile("bar.sync %0, 64;\n" ::"r"(bar_id) : "memory");  // max exchange slot free
          xch[half * 128 + r_in_tile] = sum;
          // both warps of the quarter have now consumed all of S: P may overwrite it in place
          asm volatile("bar.sync %0, 64;\n" ::"r"(bar_id) : "memory");
          sum += xch[(half ^ 1) * 128 + r_in_tile];
          inv_sum = 1.0f / sum;
#pragma unroll
          for (int ci = 0; ci < 5; ++ci) {
            if (ci < my_chunks) {
              const int ch = c_lo + ci;
              if (ch * 32 + 32 <= kTaKRows) {
                tmem_st_16(t_lane + ch * 16, pk[ci]);
              } else {  // last chunk of S = 257: only 16 key columns (256..271) exist
                uint32_t pk8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) pk8[j] = pk[ci][j];
                tmem_st_8(t_lane + ch * 16, pk8);
              }
            }
          }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        // ---- epilogue: O * 1/sum -> global (each half stores its share of the d columns)
        mbar_wait(o_full, o_ph);
        o_ph ^= 1u;
        tc_fence_after();
        if (warp_has_rows) {
          __nv_bfloat16* orow = p.o + (static_cast<long long>(b) * p.s + qi) * p.o_rs + h * p.d;
          const int n16 = (p.d + 15) / 16;             // 16-column groups of O (6 for d = 88)
          const int g_lo = half == 0 ? 0 : (n16 + 1) / 2;
          const int g_hi = half == 0 ? (n16 + 1) / 2 : n16;
          for (int gi = g_lo; gi < g_hi; ++gi) {
            uint32_t r[16];
            tmem_ld_16(t_lane + kTaColO + gi * 16, r);
            tmem_ld_wait();
            if (qi < p.s) {
#pragma unroll
              for (int j = 0; j < 16; j += 8) {
                const int c0 = gi * 16 + j;
                if (c0 < p.d) {  // d % 8 == 0: whole 16-byte groups
                  uint4 u;
                  u.x = pack_bf16x2(__uint_as_float(r[j]) * inv_sum, __uint_as_float(r[j + 1]) * inv_sum);
                  u.y = pack_bf16x2(__uint_as_float(r[j + 2]) * inv_sum, __uint_as_float(r[j + 3]) * inv_sum);
                  u.z = pack_bf16x2(__uint_as_float(r[j + 4]) * inv_sum, __uint_as_float(r[j + 5]) * inv_sum);
                  u.w = pack_bf16x2(__uint_as_float(r[j + 6]) * inv_sum, __uint_as_float(r[j + 7]) * inv_sum);
                  *reinterpret_cast<uint4*>(orow + c0) = u;
                }
              }
            }
          }
        }
        tc_fence_before();  // O reads retire before the next P.V overwrites the accumulator
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================
// Ping-pong variant (round 2): two Q tiles in flight per CTA, S / P / O of each tile in its own
// 256-column TMEM slot, so the tensor pipe works on one tile while the other tile's softmax runs.
//
//   slot s (columns [256 s, 256 s + 256)):  S = Q K^T over keys 0..255 (fp32, 256 columns)
//       -> P (bf16 pairs) written IN PLACE into [0, 128) by the row's own thread while it streams S,
//       -> O = P V accumulates into [128, 128 + dpad) (dead S columns), dpad = d rounded up to 16 <= 96,
//       -> the 257th key (ViT-g: 256 patches + CLS) does not fit the slot: its score q.k_256 is
//          computed by the softmax warp itself with mma.sync straight from the swizzled Q / K tiles
//          in shared memory, and its probability goes into the 8 spare columns [224, 232) as one
//          more 16-key block of P for the last P.V instruction.
//   warp 0      TMA producer (Q tiles double-buffered; K and V single-buffered per (frame, head))
//   warp 1      MMA issuer: QK^T of tile g, then P.V of tile g-1 (the other slot)
//   warps 4-7   softmax + epilogue of the tiles in slot 0; warps 8-11: slot 1.  One thread owns one
//               query row (TMEM lane): pass 1 = row maximum (64 columns in flight, FMNMX3), pass 2 = exp2 /
//               sum / pack / store P (a rolled loop of two 32-column chunks, the next chunk's TMEM load
//               under the current chunk's exponentials), epilogue = O * (1 / sum) -> a dense staging tile
//               shared by both slots (taken in tile order) -> ONE clipped 4-D bulk tensor store per tile.
//               No cross-warp exchange.  setmaxnreg: control warps 88 registers, softmax warps 208.
// Shape class: non-causal, unmasked, 64 <= S <= 257, d <= 96 (the old single-slot kernel above keeps
// S <= 272 / d <= 128).
constexpr int kPpColO = 128;
constexpr int kPpColTail = 224;
// alignment slack + Q (2 buffers) + K + V + 2 KB of barriers / row-256 scratch; the O staging tile follows
constexpr int kPpSmemBase = 1024 + 2 * 2 * kTaChunkBytesQ + 2 * 2 * kTaChunkBytesK + 2048;

// -DVB_PP_TRACE (scripts/micro/pp_trace.sh): clock64 stamps of CTA 0's MMA issuer (who 0) and of the first warp of
// each softmax group (who 1, 2), read back through vb_debug_pp_trace.  Off in the product build.
#ifdef VB_PP_TRACE
__device__ long long g_pp_trace[3 * 1024];
__device__ int g_pp_trace_n[3];
#define PP_TRACE_DECL int tr_i_ = 0;
#define PP_TRACE(who, tag) do { if (blockIdx.x == 0 && tr_i_ < 511) { \
  g_pp_trace[(who) * 1024 + 2 * tr_i_] = (tag); g_pp_trace[(who) * 1024 + 2 * tr_i_ + 1] = clock64(); ++tr_i_; } } while (0)
#define PP_TRACE_END(who) do { if (blockIdx.x == 0) g_pp_trace_n[who] = tr_i_; } while (0)
#else
#define PP_TRACE_DECL
#define PP_TRACE(who, tag)
#define PP_TRACE_END(who)
#endif

struct PpParams {
  __nv_bfloat16* o;
  long long o_rs;
  const __nv_bfloat16* q;   // raw q rows (row stride q_rs): the 257th query row is read straight from global memory
  long long q_rs;
  int row256;               // 1: the last query row of S = 257 is computed by warps 2-3 (no third Q tile)
  int stagger;              // cycles the first tile of slot 1 is held back (phase offset between the warp groups)
  int flags;                // bit 0: software-pipelined softmax passes (S = 256 + tail), bit 1: O through a staged TMA store,
                            // bit 2: L2 prefetch of the next item's K / V / Q
  int items, heads, s, d;
  int dpad;            // d rounded up to 16: N of the P.V instruction
  float scale_log2;
};

VB_DEVICE void ldmatrix_x4(uint32_t (&r)[4], uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_addr));
}
VB_DEVICE void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

VB_DEVICE float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
VB_DEVICE float max3(float a, float b, float c) {
  float y;
  asm("max.ftz.f32 %0, %1, %2, %3;\n" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
// L2 prefetch of one (d, head, token) box: the later TMA load of the same box hits L2
VB_DEVICE void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// 4-D tiled store smem -> global (bulk async group): (d, head, token, frame) coordinates; clipped at the extents
VB_DEVICE void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__global__ void __launch_bounds__(kTaThreads, 1)
attn_tcgen05_pp_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                       const PpParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                               // [2 buffers][2 chunks][128 rows][128 B]
  uint8_t* sK = sQ + 4 * kTaChunkBytesQ;            // [2 chunks][272 rows][128 B]
  uint8_t* sV = sK + 2 * kTaChunkBytesK;            // [2 chunks][272 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * kTaChunkBytesK);
  uint64_t* q_full = bars;          // [2]  TMA bytes of a Q tile
  uint64_t* q_empty = bars + 2;     // [2]  QK^T retired (1) + the slot's 4 softmax warps read the tile (4)
  uint64_t* k_full = bars + 4;
  uint64_t* k_empty = bars + 5;     // last QK^T of the item retired (1) + 4 warps per tile read row 256
  uint64_t* v_full = bars + 6;
  uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 8;      // [2]
  uint64_t* p_ready = bars + 10;    // [2]  4 softmax warps
  uint64_t* o_full = bars + 12;     // [2]
  uint64_t* slot_free = bars + 14;  // [2]  4 softmax warps finished reading O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  uint64_t* stage_free = bars + 17;  // the O staging tile was read by its TMA store (one completion per tile, in tile order)
  uint8_t* sO = reinterpret_cast<uint8_t*>(bars) + 2048;   // [128 rows][d] bf16, dense: the source box of the O store

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // S = 257 = 2 x 128 + 1: a third 128-row tile for ONE query row would cost a full QK^T / softmax / P.V
  // round (a third of the kernel).  That row is computed instead by the two otherwise idle warps 2-3 with
  // mma.sync straight from the K / V tiles already in shared memory (below), and only two tiles go through
  // the TMEM pipeline.
  const bool row256 = p.row256 != 0 && p.s == 257;
  const int m_tiles = row256 ? 2 : (p.s + kTaQRows - 1) / kTaQRows;
  const bool has_tail = p.s > 256;           // exactly one key (index 256) beyond the slot
  const int s_main = has_tail ? 256 : p.s;   // keys whose scores live in TMEM

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
    prefetch_tmap(&tmap_o);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 5);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&slot_free[i], 4);
    }
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1 + 4 * m_tiles + (row256 ? 2 : 0));
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1 + (row256 ? 2 : 0));
    mbar_init(stage_free, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // programmatic dependent launch: q / k / v of the preceding GEMM are visible from here on
  pdl_trigger();

  // Register split (the kernel is launched at 168 per thread = 64 512): the control warp group hands 80 per
  // thread back, the two softmax groups take 40 more each.  At 168 ptxas serialised the softmax passes (TMEM load
  // -> wait -> exponentials on the SAME registers) to stay under the cap; with 208 the next chunk's load is in
  // flight under the current chunk's exponentials.
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 88;\n" ::: "memory");
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {  // (not lane == 0: see gemm_tcgen05_2cta.cu)
      uint32_t k_ph = 0, v_ph = 0, q_ph[2] = {0, 0};
      int g = 0;  // running Q tile counter -> buffer = slot = g & 1
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int b = item / p.heads, h = item % p.heads;
        const int row0 = b * p.s;
        auto load_q = [&](int t) {
          const int buf = g & 1;
          mbar_wait(&q_empty[buf], q_ph[buf] ^ 1u);
          q_ph[buf] ^= 1u;
          mbar_expect_tx(&q_full[buf], 2 * kTaChunkBytesQ);
          uint8_t* dst = sQ + buf * 2 * kTaChunkBytesQ;
          tma_load_3d(dst, &tmap_q, &q_full[buf], 0, h, row0 + t * kTaQRows);
          tma_load_3d(dst + kTaChunkBytesQ, &tmap_q, &q_full[buf], 64, h, row0 + t * kTaQRows);
          ++g;
        };
        // K and V are single-buffered: the load of the next item's K can only be issued when this item's last
        // Q K^T has retired, and its ~4 000 clk of DRAM latency then sat on the critical path (k_full wait of the
        // MMA warp in profiles/r02_attn_pp_trace.txt).  The next item's boxes are pulled into L2 one item ahead.
        if ((p.flags & 4) != 0 && item + static_cast<int>(gridDim.x) < p.items) {
          const int nb = (item + gridDim.x) / p.heads, nh = (item + gridDim.x) % p.heads;
          const int nrow0 = nb * p.s;
          for (int c = 0; c < 2; ++c) {
            if (c * 64 < p.d) {
              for (int hf = 0; hf < 2; ++hf) {
                tma_prefetch_3d(&tmap_k, c * 64, nh, nrow0 + hf * kTaHalf);
                tma_prefetch_3d(&tmap_v, c * 64, nh, nrow0 + hf * kTaHalf);
              }
              for (int t = 0; t < m_tiles; ++t) tma_prefetch_3d(&tmap_q, c * 64, nh, nrow0 + t * kTaQRows);
            }
          }
        }
        mbar_wait(k_empty, k_ph ^ 1u);
        k_ph ^= 1u;
        mbar_expect_tx(k_full, 2 * kTaChunkBytesK);
        for (int c = 0; c < 2; ++c)
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(sK + c * kTaChunkBytesK + hf * kTaHalf * 128, &tmap_k, k_full, c * 64, h,
                        row0 + hf * kTaHalf);
        load_q(0);
        if (m_tiles > 1) load_q(1);
        mbar_wait(v_empty, v_ph ^ 1u);
        v_ph ^= 1u;
        mbar_expect_tx(v_full, 2 * kTaChunkBytesK);
        for (int c = 0; c < 2; ++c)
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(sV + c * kTaChunkBytesK + hf * kTaHalf * 128, &tmap_v, v_full, c * 64, h,
                        row0 + hf * kTaHalf);
        for (int t = 2; t < m_tiles; ++t) load_q(t);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      PP_TRACE_DECL
      const uint32_t idesc_s = umma_idesc_bf16(128, 256);
      const uint32_t idesc_o = umma_idesc_bf16(128, static_cast<uint32_t>(p.dpad)) | (1u << 16);  // B MN-major
      const int k_steps = (p.d + 15) / 16;
      const int kv_main = (s_main + 15) / 16;          // 16-key blocks of P in columns [0, 128)
      uint32_t k_ph = 0, v_ph = 0, q_ph[2] = {0, 0}, free_ph[2] = {0, 0}, p_ph[2] = {0, 0};
      bool have_prev = false, prev_first = false, prev_last = false;
      int prev_slot = 0;
      auto issue_pv = [&]() {
        const int slot = prev_slot;
        if (prev_first) {
          mbar_wait(v_full, v_ph);
          v_ph ^= 1u;
        }
        PP_TRACE(0, 3);
        mbar_wait(&p_ready[slot], p_ph[slot]);
        p_ph[slot] ^= 1u;
        tc_fence_after();
        PP_TRACE(0, 4);
        const uint32_t t0 = tmem_base + slot * 256;
        for (int js = 0; js < kv_main; ++js) {
          const uint64_t b_desc = umma_desc_mn_sw128(smem_u32(sV + js * 16 * 128), kTaChunkBytesK);
          umma_bf16_ts(t0 + kPpColO, t0 + js * 8, b_desc, idesc_o, js != 0 ? 1u : 0u);
        }
        if (has_tail) {  // keys 256..271: only key 256 has a non-zero probability
          const uint64_t b_desc = umma_desc_mn_sw128(smem_u32(sV + 256 * 128), kTaChunkBytesK);
          umma_bf16_ts(t0 + kPpColO, t0 + kPpColTail, b_desc, idesc_o, 1u);
        }
        if (prev_last) umma_commit(v_empty);
        umma_commit(&o_full[slot]);
      };
      int g = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        mbar_wait(k_full, k_ph);
        k_ph ^= 1u;
        for (int t = 0; t < m_tiles; ++t) {
          const int slot = g & 1;
          PP_TRACE(0, 1);
          mbar_wait(&q_full[slot], q_ph[slot]);
          q_ph[slot] ^= 1u;
          mbar_wait(&slot_free[slot], free_ph[slot] ^ 1u);  // epilogue of tile g-2 has drained O
          free_ph[slot] ^= 1u;
          PP_TRACE(0, 2);
          if (g == 1 && p.stagger > 0) {
            // With two tiles per item both warp groups would start every item together and stay in lockstep:
            // softmax of both tiles at the same time (contending for the MUFU pipe) with the tensor pipe idle,
            // then both P.V back to back with the softmax warps idle.  Each slot's chain (QK^T -> softmax ->
            // P.V -> epilogue) is self-timed, so holding back the very first tile of slot 1 by about half a
            // period keeps the two groups out of phase for the rest of the kernel.
            const long long t0 = clock64();
            while (clock64() - t0 < p.stagger) {
            }
          }
          tc_fence_after();
          for (int ks = 0; ks < k_steps; ++ks) {
            const int c = ks >> 2, kk = ks & 3;
            const uint64_t a_desc =
                umma_desc_k_sw128(smem_u32(sQ + slot * 2 * kTaChunkBytesQ + c * kTaChunkBytesQ)) + 2 * kk;
            const uint64_t b_desc = umma_desc_k_sw128(smem_u32(sK + c * kTaChunkBytesK)) + 2 * kk;
            umma_bf16(tmem_base + slot * 256, a_desc, b_desc, idesc_s, ks != 0 ? 1u : 0u);
          }
          umma_commit(&q_empty[slot]);
          if (t == m_tiles - 1) umma_commit(k_empty);
          umma_commit(&s_full[slot]);
          if (have_prev) issue_pv();
          have_prev = true;
          prev_slot = slot;
          prev_first = (t == 0);
          prev_last = (t == m_tiles - 1);
          ++g;
        }
      }
      if (have_prev) issue_pv();
      PP_TRACE_END(0);
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------ warps 2-3: the 257th query row
    if (row256) {
      const int w = warp - 2;
      const int k_steps = (p.d + 15) / 16;
      const int n_mb = p.dpad / 16;                      // 16-wide blocks of the head dim (6 for d = 88)
      const int d_lo = w == 0 ? 0 : (n_mb + 1) / 2, d_hi = w == 0 ? (n_mb + 1) / 2 : n_mb;
      const int kb_lo = w == 0 ? 0 : 9, kb_hi = w == 0 ? 9 : 17;   // 16-key blocks of the 272 staged keys
      __nv_bfloat16* pbuf = reinterpret_cast<__nv_bfloat16*>(bars + 20);   // [272] probabilities (bf16)
      float* xch = reinterpret_cast<float*>(bars + 20) + 140;              // [4] max / sum of the two warps
      int n_item = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n_item) {
        const int b = item / p.heads, h = item % p.heads;
        const long long row = static_cast<long long>(b) * p.s + 256;
        // q_256 as the n = 0 column of the B fragments (lanes 0-3), 16 head-dim elements per k-step
        uint32_t qb[8][2];
        {
          const __nv_bfloat16* qrow = p.q + row * p.q_rs + h * p.d;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const int d0 = ks * 16 + lane * 2;
            qb[ks][0] = (lane < 4 && ks < k_steps && d0 < p.d) ? *reinterpret_cast<const uint32_t*>(qrow + d0) : 0u;
            qb[ks][1] = (lane < 4 && ks < k_steps && d0 + 8 < p.d) ? *reinterpret_cast<const uint32_t*>(qrow + d0 + 8) : 0u;
          }
        }
        mbar_wait(k_full, static_cast<uint32_t>(n_item & 1));
        // scores s_j = q_256 . k_j: A = 16 keys x 16 dims of the swizzled K tile, D column 0 holds the scores
        float sc[9][2];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const int kb = kb_lo + i;
          sc[i][0] = sc[i][1] = -INFINITY;
          if (kb < kb_hi) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              if (ks < k_steps) {
                const int c = ks >> 2, kk = ks & 3;
                const int krow = kb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int unit = 2 * kk + (lane >> 4);
                uint32_t a[4];
                ldmatrix_x4(a, smem_u32(sK + c * kTaChunkBytesK + krow * 128 + ((unit ^ (krow & 7)) << 4)));
                mma_bf16_16816(acc, a, qb[ks][0], qb[ks][1]);
              }
            }
            const int key0 = kb * 16 + (lane >> 2);
            if ((lane & 3) == 0) {
              if (key0 < p.s) sc[i][0] = acc[0];
              if (key0 + 8 < p.s) sc[i][1] = acc[2];
            }
            mx = fmaxf(mx, fmaxf(sc[i][0], sc[i][1]));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(k_empty);   // this warp is done with K
        mx = warp_max(mx);
        if (lane == 0) xch[w] = mx;
        named_bar_sync(5, 64);
        mx = fmaxf(xch[0], xch[1]);
        const float mxs = mx * p.scale_log2;
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const int kb = kb_lo + i;
          if (kb < kb_hi && (lane & 3) == 0) {
            const int key0 = kb * 16 + (lane >> 2);
            const float p0 = sc[i][0] == -INFINITY ? 0.0f : exp2f(fmaf(sc[i][0], p.scale_log2, -mxs));
            const float p1 = sc[i][1] == -INFINITY ? 0.0f : exp2f(fmaf(sc[i][1], p.scale_log2, -mxs));
            sum += p0 + p1;
            pbuf[key0] = __float2bfloat16(p0);
            pbuf[key0 + 8] = __float2bfloat16(p1);
          }
        }
        sum = warp_sum(sum);
        if (lane == 0) xch[2 + w] = sum;
        named_bar_sync(5, 64);                 // probabilities of both warps are in pbuf
        const float inv_sum = 1.0f / (xch[2] + xch[3]);
        // o[d] = sum_j p_j V[j][d]: A = V^T (16 dims x 16 keys, transposed on load), B column 0 = p
        mbar_wait(v_full, static_cast<uint32_t>(n_item & 1));
        __nv_bfloat16* orow = p.o + row * p.o_rs + h * p.d;
        for (int mb = d_lo; mb < d_hi; ++mb) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          for (int ks = 0; ks < 17; ++ks) {
            const int key = ks * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
            const int ug = 2 * mb + ((lane >> 3) & 1);
            uint32_t a[4];
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                         : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                         : "r"(smem_u32(sV + (ug >> 3) * kTaChunkBytesK + key * 128 + (((ug & 7) ^ (key & 7)) << 4))));
            uint32_t b0 = 0u, b1 = 0u;
            if (lane < 4) {
              b0 = *reinterpret_cast<const uint32_t*>(pbuf + ks * 16 + lane * 2);
              b1 = *reinterpret_cast<const uint32_t*>(pbuf + ks * 16 + lane * 2 + 8);
            }
            mma_bf16_16816(acc, a, b0, b1);
          }
          if ((lane & 3) == 0) {
            const int d0 = mb * 16 + (lane >> 2);
            if (d0 < p.d) orow[d0] = __float2bfloat16(acc[0] * inv_sum);
            if (d0 + 8 < p.d) orow[d0 + 8] = __float2bfloat16(acc[2] * inv_sum);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(v_empty);   // this warp is done with V
        named_bar_sync(5, 64);                 // pbuf / xch are free for the next item
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;\n" ::: "memory");
    // ------------------------------------------------------------ softmax + epilogue (one slot per warp group)
    const int wg = (warp - 4) >> 2;  // == slot == Q buffer
    const int quarter = warp & 3;
    const int r_in_tile = quarter * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + wg * 256;
    const int n_chunks = (s_main + 31) / 32;   // 32-key chunks of S (8 for ViT-g)
    const int n16 = p.dpad / 16;
    uint32_t q_ph = 0, s_ph = 0, o_ph = 0;
    int g = 0, n_item = 0;
    PP_TRACE_DECL
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n_item) {
      const int b = item / p.heads, h = item % p.heads;
      for (int t = 0; t < m_tiles; ++t, ++g) {
        if ((g & 1) != wg) continue;
        const int qi = t * kTaQRows + r_in_tile;
        const bool warp_has_rows = (t * kTaQRows + quarter * 32) < p.s;
        // ---- score of key 256 for this warp's 32 rows: mma.sync from the swizzled Q / K tiles
        if (quarter == 0 && lane == 0) PP_TRACE(1 + wg, 1);
        mbar_wait(&q_full[wg], q_ph);
        q_ph ^= 1u;
        if (quarter == 0 && lane == 0) PP_TRACE(1 + wg, 2);
        float tail = -INFINITY;
        if (has_tail) {
          mbar_wait(k_full, static_cast<uint32_t>(n_item & 1));
          if (warp_has_rows) {
            // every load first, then four independent mma.sync chains (two row blocks x even / odd k-steps): the
            // rolled loop of round 1 (load -> wait -> mma per k-step) took ~1 200 clk on the tile's critical path
            const uint32_t q_base = smem_u32(sQ + wg * 2 * kTaChunkBytesQ);
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            float acc_odd[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            static_assert(kPpColO / 16 >= 6, "k-steps of the tail score are unrolled for d <= 96");
            uint32_t kf[6][2];
            uint32_t af[6][2][4];
            // (k-steps beyond d multiply zero-filled columns: the 64-wide boxes are zero beyond the head dim)
#pragma unroll
            for (int ks = 0; ks < 6; ++ks) {
              const int c = ks >> 2, kk = ks & 3;
              const uint8_t* kr = sK + c * kTaChunkBytesK + 256 * 128 + kk * 32 + (lane & 3) * 4;
              // B fragment column n = 0 <-> key 256 (row 256: swizzle phase 0); the other columns are unused
              kf[ks][0] = *reinterpret_cast<const uint32_t*>(kr);
              kf[ks][1] = *reinterpret_cast<const uint32_t*>(kr + 16);
#pragma unroll
              for (int rb = 0; rb < 2; ++rb) {
                const int row = quarter * 32 + rb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int unit = 2 * kk + (lane >> 4);
                ldmatrix_x4(af[ks][rb], q_base + c * kTaChunkBytesQ + row * 128 + ((unit ^ (row & 7)) << 4));
              }
            }
#pragma unroll
            for (int ks = 0; ks < 6; ks += 2) {
              mma_bf16_16816(acc[0], af[ks][0], kf[ks][0], kf[ks][1]);
              mma_bf16_16816(acc[1], af[ks][1], kf[ks][0], kf[ks][1]);
              mma_bf16_16816(acc_odd[0], af[ks + 1][0], kf[ks + 1][0], kf[ks + 1][1]);
              mma_bf16_16816(acc_odd[1], af[ks + 1][1], kf[ks + 1][0], kf[ks + 1][1]);
            }
            acc[0][0] += acc_odd[0][0];
            acc[0][2] += acc_odd[0][2];
            acc[1][0] += acc_odd[1][0];
            acc[1][2] += acc_odd[1][2];
            // column 0 of the C fragments sits in lanes 0, 4, ..., 28 (c0: row lane/4, c2: row lane/4 + 8)
            const int src = (lane & 7) * 4;
            const float v00 = __shfl_sync(0xffffffffu, acc[0][0], src);
            const float v02 = __shfl_sync(0xffffffffu, acc[0][2], src);
            const float v10 = __shfl_sync(0xffffffffu, acc[1][0], src);
            const float v12 = __shfl_sync(0xffffffffu, acc[1][2], src);
            tail = lane < 16 ? ((lane & 8) ? v02 : v00) : ((lane & 8) ? v12 : v10);
          }
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&q_empty[wg]);
          mbar_arrive(k_empty);
        }
        if (quarter == 0 && lane == 0) PP_TRACE(1 + wg, 3);
        mbar_wait(&s_full[wg], s_ph);
        s_ph ^= 1u;
        tc_fence_after();
        if (quarter == 0 && lane == 0) PP_TRACE(1 + wg, 4);
        float inv_sum = 0.0f;
        if (warp_has_rows) {
          float mxs;
          float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
          if ((p.flags & 1) != 0 && s_main == 256) {
            // ---- S = 256 (+ tail), software-pipelined.  One warp per scheduler is in this code at a time, so
            // latencies are hidden by instruction-level parallelism or not at all (the round-1 form issued a MUFU
            // and its dependent FADD two instructions apart: 21 clk per element against 8 at the MUFU rate).
            // pass 1: row maximum, 64 columns in flight while 64 are reduced with 3-input FMNMX
            float m0 = tail, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
            {
              uint32_t a0[32], a1[32], b0[32], b1[32];
              auto max64 = [&](const uint32_t (&x)[32], const uint32_t (&y)[32]) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  m0 = max3(m0, __uint_as_float(x[j]), __uint_as_float(y[j]));
                  m1 = max3(m1, __uint_as_float(x[j + 1]), __uint_as_float(y[j + 1]));
                  m2 = max3(m2, __uint_as_float(x[j + 2]), __uint_as_float(y[j + 2]));
                  m3 = max3(m3, __uint_as_float(x[j + 3]), __uint_as_float(y[j + 3]));
                }
              };
              tmem_ld_32(t_row, a0);
              tmem_ld_32(t_row + 32, a1);
              tmem_ld_wait();
              tmem_ld_32(t_row + 64, b0);
              tmem_ld_32(t_row + 96, b1);
              max64(a0, a1);
              tmem_ld_wait();
              tmem_ld_32(t_row + 128, a0);
              tmem_ld_32(t_row + 160, a1);
              max64(b0, b1);
              tmem_ld_wait();
              tmem_ld_32(t_row + 192, b0);
              tmem_ld_32(t_row + 224, b1);
              max64(a0, a1);
              tmem_ld_wait();
              max64(b0, b1);
            }
            mxs = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * p.scale_log2;
            if (quarter == 0 && lane == 0) PP_TRACE(1 + wg, 5);
            // pass 2: p = 2^(s c - max c), row sums, bf16 P over the consumed S columns.  A ROLLED loop of two
            // 32-column chunks (the fully unrolled form spent a third of its cycles waiting for instruction fetch,
            // stall_no_inst in profiles/r02_ncu_attn_pp_source.txt: two softmax warps per scheduler stream through
            // different code); the TMEM load of the next chunk is in flight under the current chunk's exponentials.
            {
              uint32_t ra[32], rb[32];
              const float c = p.scale_log2;
              auto chunk32 = [&](const uint32_t (&r)[32], uint32_t col) {
                float e[32];
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 32; ++j) e[j] = ex2_ftz(fmaf(__uint_as_float(r[j]), c, -mxs));
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  s0 += e[j]; s1 += e[j + 1]; s2 += e[j + 2]; s3 += e[j + 3];
                  pk[j / 2] = pack_bf16x2(e[j], e[j + 1]);
                  pk[j / 2 + 1] = pack_bf16x2(e[j + 2], e[j + 3]);
                }
                tmem_st_16(t_row + col, pk);  // P columns alias S columns of an earlier chunk: consumed
              };
              tmem_ld_32(t_row, ra);
              tmem_ld_wait();
#pragma unroll 1
              for (int i = 0; i < 4; ++i) {
                tmem_ld_32(t_row + (2 * i + 1) * 32, rb);
                chunk32(ra, (2 * i) * 16);
                tmem_ld_wait();
                if (i < 3) tmem_ld_32(t_row + (2 * i + 2) * 32, ra);
                chunk32(rb, (2 * i + 1) * 16);
                tmem_ld_wait();
              }
            }
          } else {
            // ---- pass 1: row maximum (two 32-column loads in flight)
            float m0 = tail, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
            for (int ch = 0; ch < n_chunks; ch += 2) {
              uint32_t r0[32], r1[32];
              const bool two = ch + 1 < n_chunks;
              tmem_ld_32(t_row + ch * 32, r0);
              if (two) tmem_ld_32(t_row + (ch + 1) * 32, r1);
              tmem_ld_wait();
              if (ch * 32 + 32 <= s_main) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  m0 = fmaxf(m0, __uint_as_float(r0[j]));
                  m1 = fmaxf(m1, __uint_as_float(r0[j + 1]));
                  m2 = fmaxf(m2, __uint_as_float(r0[j + 2]));
                  m3 = fmaxf(m3, __uint_as_float(r0[j + 3]));
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (ch * 32 + j < s_main) m0 = fmaxf(m0, __uint_as_float(r0[j]));
              }
              if (two) {
                if (ch * 32 + 64 <= s_main) {
#pragma unroll
                  for (int j = 0; j < 32; j += 4) {
                    m0 = fmaxf(m0, __uint_as_float(r1[j]));
                    m1 = fmaxf(m1, __uint_as_float(r1[j + 1]));
                    m2 = fmaxf(m2, __uint_as_float(r1[j + 2]));
                    m3 = fmaxf(m3, __uint_as_float(r1[j + 3]));
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (ch * 32 + 32 + j < s_main) m0 = fmaxf(m0, __uint_as_float(r1[j]));
                }
              }
            }
            mxs = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * p.scale_log2;
            if (quarter == 0 && lane == 0) PP_TRACE(1 + wg, 5);
            // ---- pass 2: p = 2^(s*c - max*c), row sum, bf16 P over the consumed S columns; the load of
            // chunk ch+1 is in flight while chunk ch is exponentiated
            auto chunk = [&](const uint32_t (&r)[32], int ch) {
              uint32_t pk[16];
              if (ch * 32 + 32 <= s_main) {
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                  const float p0 = exp2f(fmaf(__uint_as_float(r[2 * j]), p.scale_log2, -mxs));
                  const float p1 = exp2f(fmaf(__uint_as_float(r[2 * j + 1]), p.scale_log2, -mxs));
                  const float p2 = exp2f(fmaf(__uint_as_float(r[2 * j + 2]), p.scale_log2, -mxs));
                  const float p3 = exp2f(fmaf(__uint_as_float(r[2 * j + 3]), p.scale_log2, -mxs));
                  s0 += p0; s1 += p1; s2 += p2; s3 += p3;
                  pk[j] = pack_bf16x2(p0, p1);
                  pk[j + 1] = pack_bf16x2(p2, p3);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int k0 = ch * 32 + 2 * j;
                  const float p0 = k0 < s_main ? exp2f(fmaf(__uint_as_float(r[2 * j]), p.scale_log2, -mxs)) : 0.0f;
                  const float p1 = k0 + 1 < s_main ? exp2f(fmaf(__uint_as_float(r[2 * j + 1]), p.scale_log2, -mxs)) : 0.0f;
                  s0 += p0; s1 += p1;
                  pk[j] = pack_bf16x2(p0, p1);
                }
              }
              // P columns [16 ch, 16 ch + 16) alias S columns of chunk ch / 2 <= ch: already consumed
              tmem_st_16(t_row + ch * 16, pk);
            };
            {
              uint32_t ra[32], rb[32];
              tmem_ld_32(t_row, ra);
              for (int ch = 0; ch < n_chunks; ch += 2) {
                tmem_ld_wait();
                if (ch + 1 < n_chunks) tmem_ld_32(t_row + (ch + 1) * 32, rb);
                chunk(ra, ch);
                if (ch + 1 < n_chunks) {
                  tmem_ld_wait();
                  if (ch + 2 < n_chunks) tmem_ld_32(t_row + (ch + 2) * 32, ra);
                  chunk(rb, ch + 1);
                }
              }
            }
          }
          if (has_tail) {
            const float pt = exp2f(fmaf(tail, p.scale_log2, -mxs));
            s0 += pt;
            uint32_t pk8[8] = {pack_bf16x2(pt, 0.0f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            tmem_st_8(t_row + kPpColTail, pk8);
          }
          inv_sum = 1.0f / ((s0 + s1) + (s2 + s3));
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[wg]);
        if (quarter == 0 && lane == 0) PP_TRACE(1 + wg, 6);
        // ---- epilogue: O * 1/sum -> global
        mbar_wait(&o_full[wg], o_ph);
        o_ph ^= 1u;
        tc_fence_after();
        if (quarter == 0 && lane == 0) PP_TRACE(1 + wg, 7);
        if ((p.flags & 2) != 0) {
          // ---- O through shared memory and ONE bulk tensor store per tile.  (Row-per-thread 16-byte global stores
          // touch 32 lines per instruction: 1 408 LSU wavefronts per tile, ~3 800 clk of epilogue in the round-1
          // trace.)  All of O is pulled into registers first so that the TMEM slot is released before the stores.
          // the staging tile is shared by both groups; tiles take it in tile order (g), one completion each
          if (g > 0) mbar_wait(stage_free, static_cast<uint32_t>((g - 1) & 1));
          uint8_t* srow = sO + r_in_tile * (p.d * 2);
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            uint32_t ro[3][16];
#pragma unroll
            for (int gi = 0; gi < 3; ++gi)
              tmem_ld_16(t_row + kPpColO + (half * 3 + gi) * 16, ro[gi]);  // (columns beyond dpad: unused)
            tmem_ld_wait();
            if (half == 1) {
              tc_fence_before();  // O reads retire before the next QK^T overwrites the slot
              __syncwarp();
              if (lane == 0) mbar_arrive(&slot_free[wg]);
            }
#pragma unroll
            for (int gi = 0; gi < 3; ++gi) {
#pragma unroll
              for (int j = 0; j < 16; j += 8) {
                const int c0 = (half * 3 + gi) * 16 + j;
                if (c0 < p.d) {  // d % 8 == 0: whole 16-byte groups
                  uint4 u;
                  u.x = pack_bf16x2(__uint_as_float(ro[gi][j]) * inv_sum, __uint_as_float(ro[gi][j + 1]) * inv_sum);
                  u.y = pack_bf16x2(__uint_as_float(ro[gi][j + 2]) * inv_sum, __uint_as_float(ro[gi][j + 3]) * inv_sum);
                  u.z = pack_bf16x2(__uint_as_float(ro[gi][j + 4]) * inv_sum, __uint_as_float(ro[gi][j + 5]) * inv_sum);
                  u.w = pack_bf16x2(__uint_as_float(ro[gi][j + 6]) * inv_sum, __uint_as_float(ro[gi][j + 7]) * inv_sum);
                  *reinterpret_cast<uint4*>(srow + c0 * 2) = u;
                }
              }
            }
          }
          fence_proxy_async();
          named_bar_sync(6 + wg, 128);
          if (quarter == 0 && lane == 0) {
            tma_store_4d(&tmap_o, sO, 0, h, t * kTaQRows, b);  // rows beyond S are clipped
            bulk_commit();
            bulk_wait_read<0>();
            mbar_arrive(stage_free);
          }
          __syncwarp();
        } else {
          if (warp_has_rows) {
            __nv_bfloat16* orow = p.o + (static_cast<long long>(b) * p.s + qi) * p.o_rs + h * p.d;
            for (int gi = 0; gi < n16; gi += 2) {
              uint32_t r0[16], r1[16];
              const bool two = gi + 1 < n16;
              tmem_ld_16(t_row + kPpColO + gi * 16, r0);
              if (two) tmem_ld_16(t_row + kPpColO + (gi + 1) * 16, r1);
              tmem_ld_wait();
              if (qi < p.s) {
#pragma unroll
                for (int j = 0; j < 16; j += 8) {
                  const int c0 = gi * 16 + j;
                  if (c0 < p.d) {  // d % 8 == 0: whole 16-byte groups
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(r0[j]) * inv_sum, __uint_as_float(r0[j + 1]) * inv_sum);
                    u.y = pack_bf16x2(__uint_as_float(r0[j + 2]) * inv_sum, __uint_as_float(r0[j + 3]) * inv_sum);
                    u.z = pack_bf16x2(__uint_as_float(r0[j + 4]) * inv_sum, __uint_as_float(r0[j + 5]) * inv_sum);
                    u.w = pack_bf16x2(__uint_as_float(r0[j + 6]) * inv_sum, __uint_as_float(r0[j + 7]) * inv_sum);
                    *reinterpret_cast<uint4*>(orow + c0) = u;
                  }
                }
                if (two) {
#pragma unroll
                  for (int j = 0; j < 16; j += 8) {
                    const int c0 = (gi + 1) * 16 + j;
                    if (c0 < p.d) {
                      uint4 u;
                      u.x = pack_bf16x2(__uint_as_float(r1[j]) * inv_sum, __uint_as_float(r1[j + 1]) * inv_sum);
                      u.y = pack_bf16x2(__uint_as_float(r1[j + 2]) * inv_sum, __uint_as_float(r1[j + 3]) * inv_sum);
                      u.z = pack_bf16x2(__uint_as_float(r1[j + 4]) * inv_sum, __uint_as_float(r1[j + 5]) * inv_sum);
                      u.w = pack_bf16x2(__uint_as_float(r1[j + 6]) * inv_sum, __uint_as_float(r1[j + 7]) * inv_sum);
                      *reinterpret_cast<uint4*>(orow + c0) = u;
                    }
                  }
                }
              }
            }
          }
          tc_fence_before();  // O reads retire before the next QK^T overwrites the slot
          __syncwarp();
          if (lane == 0) mbar_arrive(&slot_free[wg]);
        }
        if (quarter == 0 && lane == 0) PP_TRACE(1 + wg, 8);
      }
    }
    if (quarter == 0 && lane == 0) PP_TRACE_END(1 + wg);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn3 encode_fn3() {
  static EncodeTiledFn3 fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn3>(sym);
  }
  return fn;
}

// (rows, heads, d) view of a (rows, >= heads*d) bf16 buffer; box = 64 x 1 x box_rows.
bool make_tmap_heads(CUtensorMap* map, const void* ptr, long long rows, long long heads,
                            long long d, long long row_stride, int box_rows) {
  const TmapKey key = {{3ull, reinterpret_cast<unsigned long long>(ptr), static_cast<unsigned long long>(rows),
                        static_cast<unsigned long long>(heads), static_cast<unsigned long long>(d),
                        static_cast<unsigned long long>(row_stride), static_cast<unsigned long long>(box_rows), 0ull}};
  if (tmap_cache_lookup(key, map)) return true;
  EncodeTiledFn3 fn = encode_fn3();
  if (fn == nullptr) return false;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(heads),
                        static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(d) * 2, static_cast<cuuint64_t>(row_stride) * 2};
  cuuint32_t box[3] = {64, 1, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[3] = {1, 1, 1};
  const bool ok = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  if (ok) tmap_cache_store(key, *map);
  return ok;
}

// (frames, tokens, heads, d) view of the output for the staged store: box = d x 1 x 128 x 1 out of a dense
// [128][d] shared-memory tile, no swizzle; rows beyond a frame's S tokens are clipped.
static bool make_tmap_o(CUtensorMap* map, const void* ptr, long long frames, long long s, long long heads, long long d,
                        long long row_stride) {
  const TmapKey key = {{4ull, reinterpret_cast<unsigned long long>(ptr), static_cast<unsigned long long>(frames),
                        static_cast<unsigned long long>(s), static_cast<unsigned long long>(heads),
                        static_cast<unsigned long long>(d), static_cast<unsigned long long>(row_stride), 0ull}};
  if (tmap_cache_lookup(key, map)) return true;
  EncodeTiledFn3 fn = encode_fn3();
  if (fn == nullptr) return false;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(heads), static_cast<cuuint64_t>(s),
                        static_cast<cuuint64_t>(frames)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(d) * 2, static_cast<cuuint64_t>(row_stride) * 2,
                        static_cast<cuuint64_t>(s * row_stride) * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(d), 1, static_cast<cuuint32_t>(kTaQRows), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const bool ok = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  if (ok) tmap_cache_store(key, *map);
  return ok;
}

bool attention_tcgen05_eligible(const vb_attn_args& a) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if (a.causal || a.key_mask != nullptr || a.lse != nullptr || a.rel_bias != nullptr) return false;
  if (a.dropout_p > 0.0f && a.dropout_seed != nullptr) return false;
  if (a.sq != a.skv || a.sq > kTaKRows || a.sq < 64) return false;
  if (a.d % 8 != 0 || a.d > 128 || a.d < 16) return false;
  if (!al(a.q) || !al(a.k) || !al(a.v) || !al(a.o)) return false;
  if (a.q_rs % 8 || a.k_rs % 8 || a.v_rs % 8 || a.o_rs % 8) return false;
  // frames must be back to back: row of (b, s) = b*S + s
  if (a.q_bs != a.sq * a.q_rs || a.k_bs != a.skv * a.k_rs || a.v_bs != a.skv * a.v_rs ||
      a.o_bs != a.sq * a.o_rs)
    return false;
  if ((a.d * 2) % 16 != 0) return false;
  if (a.batch * a.heads > (1ll << 30)) return false;
  return true;
}

// The two-slot ping-pong kernel takes d <= 96 (O fits 96 columns) and S <= 257 (at most one key beyond the
// 256-column S slot).  VB_ATTN_PP=0 keeps the single-slot kernel (A/B measurements).
static bool attention_pp_eligible(const vb_attn_args& a) {
  static const bool on = [] {
    const char* e = std::getenv("VB_ATTN_PP");
    return e == nullptr || e[0] != '0';
  }();
  return on && a.d <= 96 && a.sq <= 257;
}

cudaError_t attention_tcgen05_launch(const vb_attn_args& a, cudaStream_t stream) {
  CUtensorMap tq, tk, tv;
  const long long rows = a.batch * a.sq;
  if (!make_tmap_heads(&tq, a.q, rows, a.heads, a.d, a.q_rs, kTaQRows)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&tk, a.k, rows, a.heads, a.d, a.k_rs, kTaHalf)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&tv, a.v, rows, a.heads, a.d, a.v_rs, kTaHalf)) return cudaErrorInvalidValue;
  static DeviceOnce attr_once;
  bool& attr = attr_once();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTaSmem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int sms = device_sm_count();
  if (attention_pp_eligible(a)) {
    static DeviceOnce attr_pp_once;
  bool& attr_pp = attr_pp_once();
    if (!attr_pp) {
      cudaError_t e = cudaFuncSetAttribute(attn_tcgen05_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kPpSmemBase + kTaQRows * 96 * 2);  // = 232 448, the 227 KB limit
      if (e != cudaSuccess) return e;
      attr_pp = true;
    }
    PpParams pp;
    pp.o = reinterpret_cast<__nv_bfloat16*>(a.o);
    pp.o_rs = a.o_rs;
    pp.items = static_cast<int>(a.batch * a.heads);
    pp.heads = static_cast<int>(a.heads);
    pp.s = static_cast<int>(a.sq);
    pp.d = static_cast<int>(a.d);
    pp.dpad = (pp.d + 15) / 16 * 16;
    pp.scale_log2 = a.scale * 1.4426950408889634f;
    pp.q = reinterpret_cast<const __nv_bfloat16*>(a.q);
    pp.q_rs = a.q_rs;
    static const int row256 = [] {
      const char* e = std::getenv("VB_ATTN_ROW256");
      return (e != nullptr && e[0] == '0') ? 0 : 1;
    }();
    pp.row256 = row256;
    static const int stagger = [] {
      const char* e = std::getenv("VB_ATTN_STAGGER");
      return e != nullptr ? std::atoi(e) : 0;  // measured: no effect (the shared K tile re-aligns the groups)
    }();
    pp.stagger = stagger;
    static const int flags = [] {
      const char* e = std::getenv("VB_ATTN_PP_FLAGS");   // A/B measurements: 0 = the round-1 softmax / epilogue
      return e != nullptr ? std::atoi(e) : 3;
    }();
    pp.flags = flags;
    CUtensorMap to;
    if (!make_tmap_o(&to, a.o, a.batch, a.sq, a.heads, a.d, a.o_rs)) return cudaErrorInvalidValue;
    const int smem_pp = kPpSmemBase + kTaQRows * pp.d * 2;
    const int grid_pp = pp.items < sms ? pp.items : sms;
    return launch_pdl(attn_tcgen05_pp_kernel, dim3(static_cast<unsigned>(grid_pp)), dim3(kTaThreads), smem_pp, stream,
                      tq, tk, tv, to, pp);
  }
  TaParams p;
  p.o = reinterpret_cast<__nv_bfloat16*>(a.o);
  p.o_rs = a.o_rs;
  p.items = static_cast<int>(a.batch * a.heads);
  p.heads = static_cast<int>(a.heads);
  p.s = static_cast<int>(a.sq);
  p.d = static_cast<int>(a.d);
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.pv_n = 128;
  if (const char* e = std::getenv("VB_ATTN_PV_NPAD"); e != nullptr && e[0] == '1') p.pv_n = (p.d + 15) / 16 * 16;
  const int grid = p.items < sms ? p.items : sms;
  return launch_pdl(attn_tcgen05_kernel, dim3(static_cast<unsigned>(grid)), dim3(kTaThreads), kTaSmem, stream, tq, tk,
                    tv, p);
}

}  // namespace vb

#ifdef VB_PP_TRACE
extern "C" int vb_debug_pp_trace(long long* host_out, int* host_n, int reset) {
  int e = 0;
  if (host_out != nullptr) {
    e |= static_cast<int>(cudaMemcpyFromSymbol(host_out, vb::g_pp_trace, sizeof(vb::g_pp_trace)));
    e |= static_cast<int>(cudaMemcpyFromSymbol(host_n, vb::g_pp_trace_n, sizeof(vb::g_pp_trace_n)));
  }
  if (reset != 0) {
    int z[3] = {0, 0, 0};
    e |= static_cast<int>(cudaMemcpyToSymbol(vb::g_pp_trace_n, z, sizeof(z)));
  }
  return e;
}
#endif
