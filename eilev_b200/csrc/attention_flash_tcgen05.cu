// Flash attention on the 5th-gen tensor cores (tcgen05 / TMEM / TMA) for the sequence-length-general shape
// class: the OPT causal self-attention (d = 80, L <= 2048), the Q-Former self / cross attention and the T5
// attentions (d = 64) — forward and backward.  (The ViT's S = 257 class keeps its single-pass kernel,
// attention_tcgen05.cu.)
//
// One kernel body, three instantiations.  A CTA owns a 128-row tile of the "row side" (TMEM lanes) of one
// (batch, head) and sweeps 128-wide tiles of the "column side":
//
//   kFwd     rows = queries (R1 = Q_i),             columns = keys    (C1 = K_j, C2 = V_j)    -> O, lse
//   kBwdQ    rows = queries (R1 = Q_i, R2 = dO_i),  columns = keys    (C1 = K_j, C2 = V_j)    -> dQ
//   kBwdKV   rows = keys    (R1 = K_j, R2 = V_j),   columns = queries (C1 = Q_i, C2 = dO_i)   -> dV, dK
//
//   T1 = R1 . C1^T   (the scores, or their transpose)     tcgen05.mma SS, M = 128, N = 64, K = d
//   T2 = R2 . C2^T   (dP = dO . V^T, or its transpose; backward only)
//   P  = exp2(T1 * scale * log2 e - lse),  dS = P o (T2 - delta)          one thread per TMEM lane
//   acc1 += P  . C2  (O or dV)                              tcgen05.mma TS: P / dS are read from TMEM where they
//   acc2 += dS . C1  (dK or dQ)                             overwrite T1 / T2 as bf16 pairs; C1 / C2 = MN-major B
//
// Every product has its reduction index on the TMEM columns, so nothing is transposed through shared memory and
// dQ needs no fp32 atomics (the mma.sync kernel's memset + atomics + convert pass): the backward computes the
// scores twice instead, once per orientation.  The forward sweeps the keys twice — sweep 0 reduces the row
// maximum / sum (the lse), sweep 1 multiplies P = exp2(s - max) with V and the epilogue divides by the sum —
// which keeps O accumulating in TMEM without the rescaling of an online softmax; the tensor work is a few
// percent of the kernel, the exp2 of the extra sweep is what it costs.
//
// The column side is walked in 64-column blocks ("items"), even items by warps 4-7 in TMEM slot 0, odd items by
// warps 8-11 in slot 1:
//   warp 0      TMA producer: R tiles once per unit, C blocks through a 4-deep ring, 3-D maps (d, head, token)
//   warp 1      MMA issuer: accumulate(item k - 2) then scores(item k) into the slot that frees, so one group's
//               tensor work runs under the other group's exponentials
//   warps 4-11  elementwise + epilogue (lane quarter = warp & 3, group = (warp - 4) / 4)
// TMEM columns: slot s at [128 s, 128 s + 128): T1 / P in [0,64), T2 / dS in [64,128); acc1 at 256, acc2 at
// 256 + dpad (dpad = d rounded up to 16 <= 128).
//
// Masks are folded into the exponent: exponent = T1 * c - (row term + column term) where an invalid key (beyond
// Skv, key-padding mask) or an invalid / fully masked query (beyond Sq, lse = -inf) contributes +inf, so P = 0
// exactly.  Causal-diagonal items, dropout on the probabilities and the T5 relative bias are compile-time variants
// of the same straight-line loop (fa_chunk_plain<.., DROP, BIAS, CAUSAL>): per-element branches made the ptxas
// output three times slower.
// Reference ops: HF OPTAttention / Blip2QFormerMultiHeadAttention / T5Attention forward and autograd backward of
// softmax(Q K^T * scale + mask) V, reached from eilev/model/v2.py:132-252 through language_model / qformer.
#include <cstdlib>

#include "common.cuh"
#include "internal.h"
#include "tc_attention.cuh"

namespace vb {

constexpr int kFaThreads = 384;
constexpr int kFaEdge = 128;                    // tile edge: rows (lanes) and columns
constexpr int kFaSub = 64;                      // columns per item / TMEM slot
constexpr int kFaChunk = kFaEdge * 128;         // one 64-wide d chunk of a 128-row R tile: 16 KB
constexpr int kFaTile = 2 * kFaChunk;           // d <= 128
constexpr int kFaBlkChunk = kFaSub * 128;       // one 64-wide d chunk of a 64-row C block: 8 KB
constexpr int kFaBlk = 2 * kFaBlkChunk;         // one operand of a C block
constexpr int kFaRing = 4;                      // C blocks (C1 + C2 each) in flight
constexpr int kFaColBytes = 2 * 2 * kFaSub * 8; // per-column terms: [2 groups][2 buffers][64] float2
constexpr int kFaStatBytes = 2 * kFaEdge * 8;   // forward: [2 groups][128] (max, sum)
constexpr int kFaSmem = 2 * kFaTile + kFaRing * 2 * kFaBlk + kFaColBytes + kFaStatBytes + 256 + 1024;
constexpr float kFaLog2e = 1.4426950408889634f;

// -DVB_FA_TRACE (scripts/micro/attn_trace.sh): clock64 stamps of CTA 0's MMA issuer and of one warp per group,
// read back through vb_debug_attn_trace.  Off in the product build.
#ifdef VB_FA_TRACE
__device__ long long g_fa_trace[3 * 512];
__device__ int g_fa_trace_n[3];
#define FA_TRACE(who, tag) do { if (blockIdx.x == 0) { int i_ = g_fa_trace_n[who]; if (i_ < 255) { \
  g_fa_trace[(who) * 512 + 2 * i_] = (tag); g_fa_trace[(who) * 512 + 2 * i_ + 1] = clock64(); g_fa_trace_n[who] = i_ + 1; } } } while (0)
#else
#define FA_TRACE(who, tag)
#endif

enum FaMode : int { kFwd = 0, kBwdQ = 1, kBwdKV = 2 };

struct FaParams {
  const float* lse;         // backward: (B, H, Sq) natural log
  float* lse_out;           // forward: optional
  const float* delta;       // backward, kBwdKV: (B, H, Sq) rowsum(dO o O), written by the kBwdQ pass
  float* delta_out;         // kBwdQ: computes delta from o / d_o (below) and stores it here for the kBwdKV pass
  const __nv_bfloat16* o;   // kBwdQ: forward output and its gradient, (b, s, h, d) through o_bs / o_rs
  const __nv_bfloat16* d_o;
  long long o_bs, o_rs;
  const uint8_t* key_mask;  // (B, Skv) or nullptr
  __nv_bfloat16* out1;      // kFwd: O, kBwdKV: dV
  __nv_bfloat16* out2;      // kBwdKV: dK, kBwdQ: dQ
  long long out1_bs, out1_rs, out2_bs, out2_rs;
  float out2_mul;           // softmax scale (dK) / scale * dq_scale (dQ)
  int batch, heads, sq, skv, d, dpad;
  int causal;
  float scale_log2;
  const unsigned long long* drop_seed;
  unsigned long long drop_salt;
  unsigned int drop_thresh;
  float drop_scale;
  const float* rel_bias;
  long long rel_bias_stride;
  int row_tiles, col_blocks;  // 128-row tiles on the lane side, 64-column blocks on the column side
};

// The n-th unit of CTA c: passes run forwards and backwards over the heavy-first unit list so that every CTA
// gets a heavy and a light unit (causal: the number of column tiles falls / rises with the row tile).
VB_DEVICE int fa_unit(int n) {
  const int g = static_cast<int>(gridDim.x), c = static_cast<int>(blockIdx.x);
  return (n & 1) ? (n + 1) * g - 1 - c : n * g + c;
}

template <int MODE>
VB_DEVICE void fa_unit_coords(const FaParams& p, int unit, int& b, int& h, int& rt, int& c_begin, int& c_end) {
  const int bh = p.batch * p.heads;
  const int slot = unit / bh;  // heavy first
  const int rem = unit % bh;
  b = rem / p.heads;
  h = rem % p.heads;
  const int off = p.skv - p.sq;
  c_begin = 0;
  c_end = p.col_blocks;
  if (MODE == kBwdKV) {
    rt = slot;  // key tile: low tiles are seen by the most queries
    if (p.causal) {
      const int first_q = rt * kFaEdge - off;  // first query that sees the tile's first key
      c_begin = first_q <= 0 ? 0 : first_q / kFaSub;
      if (c_begin > c_end) c_begin = c_end;
    }
  } else {
    rt = p.causal ? p.row_tiles - 1 - slot : slot;  // query tile: high tiles see the most keys
    if (p.causal) {
      const int last_key = rt * kFaEdge + kFaEdge - 1 + off;
      const int e = last_key < 0 ? 0 : last_key / kFaSub + 1;
      if (e < c_end) c_end = e;
    }
  }
}

// Named barrier over `threads` threads that also ANDs a predicate across them
VB_DEVICE bool named_bar_and(uint32_t id, uint32_t threads, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %3, 0;\n\t"
      "barrier.cta.red.and.pred p, %1, %2, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(r)
      : "r"(id), "r"(threads), "r"(static_cast<uint32_t>(pred))
      : "memory");
  return r != 0;
}

// (neg, delta) of two neighbouring columns from the per-column table in shared memory
VB_DEVICE float4 fa_col_pair(uint32_t smem_addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_addr));
  return v;
}

// Dropout on the probabilities / T5 relative bias of one 32-column chunk of one row: column j hashes index
// idx0 + j * idx_step and reads bias rb[clamp(rb0 + j * rb_step)] (the bias depends on key - query only)
struct FaExtra {
  uint64_t seed, idx0;
  long long idx_step;
  uint32_t thresh;
  float scale;
  const float* rb;
  int rb0, rb_step, rb_lo, rb_hi;
  int vis_lo, vis_hi;   // causal items: column j of the chunk is visible iff vis_lo <= j <= vis_hi
};

// 32 columns of one row, no per-element masks: P = exp2(T1 * c - neg), dS = P * (T2 - delta), packed to bf16
// pairs.  neg / delta = the row's term plus (COL_TERMS) the column's, or (!ROW_TERMS) the column's alone.
// Straight-line code: the variants are chosen once per item, outside the unrolled loop.
template <bool HAS_P, bool HAS_DS, bool COL_TERMS, bool ROW_TERMS, bool DROP = false, bool BIAS = false,
          bool CAUSAL = false>
VB_DEVICE void fa_chunk_plain(const uint32_t (&t1)[32], const uint32_t (&t2)[32], uint32_t (&u1)[16], uint32_t (&u2)[16],
                              float c, float neg_row, float d_row, uint32_t col_addr, const FaExtra& x = FaExtra()) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float na = neg_row, nb = neg_row, da = d_row, db = d_row;
    if (COL_TERMS) {
      const float4 cc = fa_col_pair(col_addr + 8 * j);
      if (ROW_TERMS) {
        na += cc.x; nb += cc.z; da += cc.y; db += cc.w;
      } else {
        na = cc.x; nb = cc.z; da = cc.y; db = cc.w;
      }
    }
    if (BIAS) {  // exponent = T1 * c + bias * log2(e) - neg; an infinite neg still wins
      const int ia = min(max(x.rb0 + j * x.rb_step, x.rb_lo), x.rb_hi);
      const int ib = min(max(x.rb0 + (j + 1) * x.rb_step, x.rb_lo), x.rb_hi);
      na = fmaf(-__ldg(x.rb + ia), kFaLog2e, na);
      nb = fmaf(-__ldg(x.rb + ib), kFaLog2e, nb);
    }
    if (CAUSAL) {  // an invisible pair gets an infinite neg: P = 0, dS = 0
      na = (j >= x.vis_lo && j <= x.vis_hi) ? na : INFINITY;
      nb = (j + 1 >= x.vis_lo && j + 1 <= x.vis_hi) ? nb : INFINITY;
    }
    const float pa = exp2f(fmaf(__uint_as_float(t1[j]), c, -na));
    const float pb = exp2f(fmaf(__uint_as_float(t1[j + 1]), c, -nb));
    float ua = pa, ub = pb;                                                   // P that multiplies V / dO
    float ga = HAS_DS ? __uint_as_float(t2[j]) : 0.0f, gb = HAS_DS ? __uint_as_float(t2[j + 1]) : 0.0f;  // dP
    if (DROP) {  // the forward's mask, regenerated: dropped probabilities pass no value and no gradient
      const bool ka = dropout_keep(x.seed, x.idx0 + static_cast<uint64_t>(j * x.idx_step), x.thresh);
      const bool kb = dropout_keep(x.seed, x.idx0 + static_cast<uint64_t>((j + 1) * x.idx_step), x.thresh);
      ua = ka ? pa * x.scale : 0.0f;
      ub = kb ? pb * x.scale : 0.0f;
      ga = ka ? ga * x.scale : 0.0f;
      gb = kb ? gb * x.scale : 0.0f;
    }
    if (HAS_P) u1[j >> 1] = pack_bf16x2(ua, ub);
    if (HAS_DS) u2[j >> 1] = pack_bf16x2(pa * (ga - da), pb * (gb - db));
  }
}

// All variants with the column and row terms on: chosen by three run-time flags of the item
template <bool HAS_P, bool HAS_DS>
VB_DEVICE void fa_chunk_extras(bool causal, bool drop, bool bias, const uint32_t (&t1)[32], const uint32_t (&t2)[32],
                               uint32_t (&u1)[16], uint32_t (&u2)[16], float c, float neg_row, float d_row,
                               uint32_t col_addr, const FaExtra& x) {
#define VB_FA_CALL(D, B, C) fa_chunk_plain<HAS_P, HAS_DS, true, true, D, B, C>(t1, t2, u1, u2, c, neg_row, d_row, col_addr, x)
  if (causal) {
    if (drop) { if (bias) VB_FA_CALL(true, true, true); else VB_FA_CALL(true, false, true); }
    else { if (bias) VB_FA_CALL(false, true, true); else VB_FA_CALL(false, false, true); }
  } else {
    if (drop) { if (bias) VB_FA_CALL(true, true, false); else VB_FA_CALL(true, false, false); }
    else { if (bias) VB_FA_CALL(false, true, false); else VB_FA_CALL(false, false, false); }
  }
#undef VB_FA_CALL
}

// Forward statistics of 32 unmasked columns: running maximum (log2 units) and sum, four independent chains
VB_DEVICE void fa_stats_plain(const uint32_t (&t1)[32], float c, float& run_m, float& run_l) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    m0 = fmaxf(m0, __uint_as_float(t1[j]) * c);
    m1 = fmaxf(m1, __uint_as_float(t1[j + 1]) * c);
    m2 = fmaxf(m2, __uint_as_float(t1[j + 2]) * c);
    m3 = fmaxf(m3, __uint_as_float(t1[j + 3]) * c);
  }
  const float m_new = fmaxf(run_m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));  // finite: the scores are
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    s0 += exp2f(fmaf(__uint_as_float(t1[j]), c, -m_new));
    s1 += exp2f(fmaf(__uint_as_float(t1[j + 1]), c, -m_new));
    s2 += exp2f(fmaf(__uint_as_float(t1[j + 2]), c, -m_new));
    s3 += exp2f(fmaf(__uint_as_float(t1[j + 3]), c, -m_new));
  }
  run_l = run_l * exp2f(run_m - m_new) + ((s0 + s1) + (s2 + s3));
  run_m = m_new;
}

// The same with a per-column term (+inf for a masked / out-of-range key: the column drops out of both reductions)
template <bool BIAS = false, bool CAUSAL = false>
VB_DEVICE void fa_stats_cols(const uint32_t (&t1)[32], float c, uint32_t col_addr, float& run_m, float& run_l,
                             const FaExtra& e = FaExtra()) {
  float x[32];
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const float4 cc = fa_col_pair(col_addr + 8 * j);
    float na = cc.x, nb = cc.z;
    if (BIAS) {
      const int ia = min(max(e.rb0 + j * e.rb_step, e.rb_lo), e.rb_hi);
      const int ib = min(max(e.rb0 + (j + 1) * e.rb_step, e.rb_lo), e.rb_hi);
      na = fmaf(-__ldg(e.rb + ia), kFaLog2e, na);
      nb = fmaf(-__ldg(e.rb + ib), kFaLog2e, nb);
    }
    if (CAUSAL) {
      na = (j >= e.vis_lo && j <= e.vis_hi) ? na : INFINITY;
      nb = (j + 1 >= e.vis_lo && j + 1 <= e.vis_hi) ? nb : INFINITY;
    }
    x[j] = fmaf(__uint_as_float(t1[j]), c, -na);
    x[j + 1] = fmaf(__uint_as_float(t1[j + 1]), c, -nb);
    m0 = fmaxf(m0, x[j]);
    m1 = fmaxf(m1, x[j + 1]);
  }
  const float m_new = fmaxf(run_m, fmaxf(m0, m1));
  if (m_new == -INFINITY) return;  // nothing visible so far
  float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    s0 += exp2f(x[j] - m_new);
    s1 += exp2f(x[j + 1] - m_new);
  }
  run_l = run_l * exp2f(run_m - m_new) + (s0 + s1);
  run_m = m_new;
}

template <int MODE>
__global__ void __launch_bounds__(kFaThreads, 1)
attn_flash_tc_kernel(const __grid_constant__ CUtensorMap tmap_r1, const __grid_constant__ CUtensorMap tmap_r2,
                     const __grid_constant__ CUtensorMap tmap_c1, const __grid_constant__ CUtensorMap tmap_c2,
                     const FaParams p) {
  constexpr bool kHasT2 = MODE != kFwd;           // second score-shaped product (dP)
  constexpr bool kHasAcc1 = MODE != kBwdQ;        // P . C2
  constexpr bool kHasAcc2 = MODE != kFwd;         // dS . C1
  constexpr bool kRowsAreKeys = MODE == kBwdKV;
  constexpr int kSweeps = MODE == kFwd ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sR1 = smem;
  uint8_t* sR2 = smem + kFaTile;
  uint8_t* sC = smem + 2 * kFaTile;            // [kFaRing][C1 block, C2 block]
  constexpr int kTiles = 2 * kFaTile + kFaRing * 2 * kFaBlk;
  float2* sCol = reinterpret_cast<float2*>(smem + kTiles);                  // [group][buffer][64]
  float2* sStat = reinterpret_cast<float2*>(smem + kTiles + kFaColBytes);   // [group][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTiles + kFaColBytes + kFaStatBytes);
  uint64_t* r_full = bars;
  uint64_t* r_empty = bars + 1;
  uint64_t* c_full = bars + 2;    // [kFaRing]
  uint64_t* c_empty = bars + 6;   // [kFaRing]
  uint64_t* t_full = bars + 10;   // [2]  scores of one item in the group's TMEM slot
  uint64_t* u_ready = bars + 12;  // [2]  4 warps: the slot is consumed / holds P, dS
  uint64_t* acc_full = bars + 14;
  uint64_t* acc_free = bars + 15; // 8 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int units = p.batch * p.heads * p.row_tiles;
  const int n_chunks = (p.d + 63) / 64;              // 64-wide d chunks that hold data
  const int k_steps = (p.d + 15) / 16;
  const int rows_total = kRowsAreKeys ? p.skv : p.sq;  // per batch
  const int cols_total = kRowsAreKeys ? p.sq : p.skv;
  const uint32_t col_acc1 = 256, col_acc2 = 256 + (kHasAcc1 ? static_cast<uint32_t>(p.dpad) : 0u);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_r1);
    prefetch_tmap(&tmap_r2);
    prefetch_tmap(&tmap_c1);
    prefetch_tmap(&tmap_c2);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(r_full, 1);
    mbar_init(r_empty, 1);
    for (int i = 0; i < kFaRing; ++i) {
      mbar_init(&c_full[i], 1);
      mbar_init(&c_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&u_ready[i], 4);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 8);
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
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t r_ph = 0;
      int ct = 0;  // running item counter: ring slot = ct % kFaRing, phase = (ct / kFaRing) & 1
      for (int n = 0; n * static_cast<int>(gridDim.x) < units; ++n) {
        const int unit = fa_unit(n);
        if (unit >= units) continue;
        int b, h, rt, c_begin, c_end;
        fa_unit_coords<MODE>(p, unit, b, h, rt, c_begin, c_end);
        if (c_begin >= c_end) continue;
        mbar_wait(r_empty, r_ph ^ 1u);
        r_ph ^= 1u;
        mbar_expect_tx(r_full, (kHasT2 ? 2 : 1) * n_chunks * kFaChunk);
        const int r_row = b * rows_total + rt * kFaEdge;
        for (int c = 0; c < n_chunks; ++c) {
          tma_load_3d(sR1 + c * kFaChunk, &tmap_r1, r_full, c * 64, h, r_row);
          if (kHasT2) tma_load_3d(sR2 + c * kFaChunk, &tmap_r2, r_full, c * 64, h, r_row);
        }
        for (int sweep = 0; sweep < kSweeps; ++sweep) {
          const bool need_c2 = MODE != kFwd || sweep == 1;  // the forward's statistics sweep reads K only
          for (int t = c_begin; t < c_end; ++t, ++ct) {
            const int buf = ct % kFaRing;
            mbar_wait(&c_empty[buf], ((ct / kFaRing) & 1) ^ 1u);
            mbar_expect_tx(&c_full[buf], (need_c2 ? 2 : 1) * n_chunks * kFaBlkChunk);
            const int c_row = b * cols_total + t * kFaSub;
            uint8_t* dst = sC + buf * 2 * kFaBlk;
            for (int c = 0; c < n_chunks; ++c) {
              tma_load_3d(dst + c * kFaBlkChunk, &tmap_c1, &c_full[buf], c * 64, h, c_row);
              if (need_c2) tma_load_3d(dst + kFaBlk + c * kFaBlkChunk, &tmap_c2, &c_full[buf], c * 64, h, c_row);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // (elect_one, not lane == 0: ptxas keeps the operands of a region guarded by elect.sync in uniform registers;
    // under a lane test every tcgen05.mma was wrapped in a waterfall loop, ~180 clk per instruction in
    // profiles/r02_attn_flash_trace.txt, and the issuer paced the whole kernel)
    if (elect_one()) {
      const uint32_t idesc_acc = umma_idesc_bf16(128, static_cast<uint32_t>(p.dpad)) | (1u << 16);  // B MN-major
      uint32_t r_ph = 0, u_ph[2] = {0, 0}, free_ph = 0;
      int ct = 0;
      // columns of 64-column block blk that hold data
      auto blk_valid = [&](int blk) {
        const int v = cols_total - blk * kFaSub;
        return v > kFaSub ? kFaSub : v;
      };
      for (int n = 0; n * static_cast<int>(gridDim.x) < units; ++n) {
        const int unit = fa_unit(n);
        if (unit >= units) continue;
        int b, h, rt, c_begin, c_end;
        fa_unit_coords<MODE>(p, unit, b, h, rt, c_begin, c_end);
        if (c_begin >= c_end) continue;
        mbar_wait(r_full, r_ph);
        r_ph ^= 1u;
        const int n_blk = c_end - c_begin;
        const int n_items = kSweeps * n_blk;
        const int ct0 = ct;
        bool acc_started = false;
        // accumulate instructions of item k (its P / dS are in its group's TMEM slot once u_ready has fired);
        // releases the item's C block
        auto issue_acc = [&](int k) {
          const int s = k & 1;
          const int buf = (ct0 + k) % kFaRing;
          FA_TRACE(0, 100 + k);
          mbar_wait(&u_ready[s], u_ph[s]);
          u_ph[s] ^= 1u;
          FA_TRACE(0, 200 + k);
          if (!(MODE == kFwd && k < n_blk)) {  // (the forward's statistics sweep has nothing to accumulate)
            if (!acc_started) {  // the previous unit's accumulators have been read out
              mbar_wait(acc_free, free_ph ^ 1u);
              free_ph ^= 1u;
            }
            tc_fence_after();
            const uint8_t* c1 = sC + buf * 2 * kFaBlk;
            const uint8_t* c2 = c1 + kFaBlk;
            const uint32_t slot = tmem_base + static_cast<uint32_t>(128 * s);
            const int u_steps = (blk_valid(c_begin + k % n_blk) + 15) / 16;
            for (int js = 0; js < u_steps; ++js) {
              const uint32_t accum = (acc_started || js != 0) ? 1u : 0u;
              const int row_b = js * 16 * 128;  // 16 column-side elements per step
              if (kHasAcc1)
                umma_bf16_ts(tmem_base + col_acc1, slot + 8 * js,
                             umma_desc_mn_sw128(smem_u32(c2 + row_b), kFaBlkChunk), idesc_acc, accum);
              if (kHasAcc2)
                umma_bf16_ts(tmem_base + col_acc2, slot + 64 + 8 * js,
                             umma_desc_mn_sw128(smem_u32(c1 + row_b), kFaBlkChunk), idesc_acc, accum);
            }
            acc_started = true;
          }
          FA_TRACE(0, 300 + k);
          umma_commit(&c_empty[buf]);
        };
        auto issue_scores = [&](int k) {
          const int s = k & 1;
          const int buf = (ct0 + k) % kFaRing;
          FA_TRACE(0, 400 + k);
          mbar_wait(&c_full[buf], ((ct0 + k) / kFaRing) & 1);
          FA_TRACE(0, 500 + k);
          tc_fence_after();
          const int valid = blk_valid(c_begin + k % n_blk);
          const uint32_t idesc_t = umma_idesc_bf16(128, static_cast<uint32_t>((valid + 15) / 16 * 16));
          const uint8_t* c1 = sC + buf * 2 * kFaBlk;
          const uint8_t* c2 = c1 + kFaBlk;
          const uint32_t slot = tmem_base + static_cast<uint32_t>(128 * s);
          for (int ks = 0; ks < k_steps; ++ks) {
            const int c = ks >> 2, kk = ks & 3;
            umma_bf16(slot, umma_desc_k_sw128(smem_u32(sR1 + c * kFaChunk)) + 2 * kk,
                      umma_desc_k_sw128(smem_u32(c1 + c * kFaBlkChunk)) + 2 * kk, idesc_t, ks != 0 ? 1u : 0u);
          }
          if (kHasT2) {
            for (int ks = 0; ks < k_steps; ++ks) {
              const int c = ks >> 2, kk = ks & 3;
              umma_bf16(slot + 64, umma_desc_k_sw128(smem_u32(sR2 + c * kFaChunk)) + 2 * kk,
                        umma_desc_k_sw128(smem_u32(c2 + c * kFaBlkChunk)) + 2 * kk, idesc_t, ks != 0 ? 1u : 0u);
            }
          }
          umma_commit(&t_full[s]);
          FA_TRACE(0, 600 + k);
        };
        for (int k = 0; k < n_items; ++k) {
          if (k >= 2) issue_acc(k - 2);  // frees the group's TMEM slot (in issue order)
          issue_scores(k);
        }
        umma_commit(r_empty);            // the R tiles are read by the scores only: the next unit's may load
        for (int k = n_items >= 2 ? n_items - 2 : 0; k < n_items; ++k) issue_acc(k);
        ct = ct0 + n_items;
        umma_commit(acc_full);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ elementwise + epilogue
    const int quarter = warp & 3;
    const int grp = (warp - 4) >> 2;                             // warp group = TMEM slot = item parity
    const int r_in = quarter * 32 + lane;                        // row (TMEM lane) inside the tile
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_slot = t_row + static_cast<uint32_t>(128 * grp);
    const int off = p.skv - p.sq;
    const bool has_drop = p.drop_thresh != 0u, has_bias = p.rel_bias != nullptr;
    const bool extras = has_drop || has_bias;  // handled by the branch-free variants, with the column terms on
    uint32_t t_ph = 0, acc_ph = 0;
    int tt = 0;  // items this group has worked on (per-column buffer = tt & 1)
    for (int n = 0; n * static_cast<int>(gridDim.x) < units; ++n) {
      const int unit = fa_unit(n);
      if (unit >= units) continue;
      int b, h, rt, c_begin, c_end;
      fa_unit_coords<MODE>(p, unit, b, h, rt, c_begin, c_end);
      const int row_g = rt * kFaEdge + r_in;   // key (kBwdKV) or query index of this lane
      const long long bh_off = (static_cast<long long>(b) * p.heads + h) * p.sq;
      const uint8_t* km = p.key_mask != nullptr ? p.key_mask + static_cast<long long>(b) * p.skv : nullptr;
      const float* rb = p.rel_bias != nullptr ? p.rel_bias + h * p.rel_bias_stride + (p.sq - 1) : nullptr;
      // per-row terms: +inf in neg_row makes P = 0 for the whole row
      float neg_row = INFINITY, d_row = 0.0f;
      if (MODE == kBwdKV) {
        if (row_g < p.skv && (km == nullptr || km[row_g] != 0)) neg_row = 0.0f;
      } else if (MODE == kBwdQ) {
        if (row_g < p.sq) {
          const float l = p.lse[bh_off + row_g];
          if (l != -INFINITY) neg_row = l * kFaLog2e;
          // delta = rowsum(dO o O) of this query row, computed here (one pass over two d-element rows) instead of
          // by a kernel of its own; the dK / dV pass, launched after this one, reads it back
          const long long ro = b * p.o_bs + static_cast<long long>(row_g) * p.o_rs + h * p.d;
          float acc = 0.0f;
#pragma unroll
          for (int c0 = 0; c0 < 128; c0 += 64) {  // eight 16-byte loads of each row in flight at a time
            uint4 uo[8], ug[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int c = c0 + 8 * i;
              uo[i] = ug[i] = make_uint4(0u, 0u, 0u, 0u);
              if (c < p.d) {
                uo[i] = __ldg(reinterpret_cast<const uint4*>(p.o + ro + c));
                ug[i] = __ldg(reinterpret_cast<const uint4*>(p.d_o + ro + c));
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t wo[4] = {uo[i].x, uo[i].y, uo[i].z, uo[i].w}, wg[4] = {ug[i].x, ug[i].y, ug[i].z, ug[i].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 fo = unpack_bf16x2(wo[j]), fg = unpack_bf16x2(wg[j]);
                acc = fmaf(fo.x, fg.x, fmaf(fo.y, fg.y, acc));
              }
            }
          }
          d_row = acc;
          if (grp == 0) p.delta_out[bh_off + row_g] = acc;
        }
      }
      // do the rows of this unit carry a term at all (other than the lse of a query row)?
      const bool rows_plain = MODE == kBwdKV ? (km == nullptr && rt * kFaEdge + kFaEdge <= p.skv) : false;
      float run_m = -INFINITY, run_l = 0.0f;  // forward, sweep 0: running maximum (log2 units) and sum
      float fwd_scale = 0.0f;                 // forward: 1 / row sum, applied to O in the epilogue
      const int n_blk = c_end - c_begin;
      for (int sweep = 0; sweep < kSweeps; ++sweep) {
        for (int kb = 0; kb < n_blk; ++kb) {
          if (((sweep * n_blk + kb) & 1) != grp) continue;  // the other group's item
          const int blk = c_begin + kb;
          int valid = cols_total - blk * kFaSub;
          if (valid > kFaSub) valid = kFaSub;
          const int col_g0 = blk * kFaSub;
          // ---- per-column terms -> shared memory (the group's first 64 threads, one column each).  Key columns
          // carry a term only when some key of the tile is masked or out of range.
          // causal: does the diagonal cut through this item?  (then the branch-free causal variants run, with the
          // column terms on)
          bool diag = false;
          if (p.causal) {
            const int key_max = kRowsAreKeys ? rt * kFaEdge + kFaEdge - 1 : col_g0 + kFaSub - 1;
            const int q_min = kRowsAreKeys ? col_g0 : rt * kFaEdge;
            diag = key_max > q_min + off;
          }
          const bool item_extras = extras || diag;
          bool cols_plain = (kRowsAreKeys || item_extras) ? false : (km == nullptr && valid == kFaSub);
          if (quarter == 0 && lane == 0) FA_TRACE(1 + grp, 100 + sweep * n_blk + kb);
          float2* col = sCol + (grp * 2 + (tt & 1)) * kFaSub;
          if (!cols_plain) {
            bool all_valid = true;
            if (r_in < kFaSub) {
              const int cg = col_g0 + r_in;
              float2 v = make_float2(INFINITY, 0.0f);
              if (kRowsAreKeys) {
                if (cg < p.sq) {
                  const float l = p.lse[bh_off + cg];
                  if (l != -INFINITY) v.x = l * kFaLog2e;
                  v.y = p.delta[bh_off + cg];
                }
              } else {
                if (cg < p.skv && (km == nullptr || km[cg] != 0)) v.x = 0.0f;
              }
              col[r_in] = v;
              all_valid = v.x == 0.0f;
            }
            // the barrier also tells whether any key of the block is masked at all: a key-padding mask that is all
            // ones over this block (the common case away from the padded end) takes the term-free variants
            all_valid = named_bar_and(1 + grp, 128, all_valid);
            if (!kRowsAreKeys && !item_extras && all_valid && valid == kFaSub) cols_plain = true;
          }
          ++tt;
          // causal: is every (row, column) pair of this half visible?
          const int n_used = (valid + 15) / 16 * 16;  // columns the instructions computed / will read
          if (quarter == 0 && lane == 0) FA_TRACE(1 + grp, 200 + sweep * n_blk + kb);
          mbar_wait(&t_full[grp], t_ph);
          t_ph ^= 1u;
          if (quarter == 0 && lane == 0) FA_TRACE(1 + grp, 300 + sweep * n_blk + kb);
          tc_fence_after();
#pragma unroll 1
          for (int c2 = 0; c2 < 2; ++c2) {
            const int col0 = 32 * c2;
            if (col0 >= n_used) continue;
            uint32_t t1[32], t2[32];
            tmem_ld_32(t_slot + col0, t1);
            if constexpr (kHasT2) tmem_ld_32(t_slot + 64 + col0, t2);
            tmem_ld_wait();
            // dropout index / bias index of this chunk's first column (see FaExtra)
            FaExtra ex;
            if (item_extras) {
              const uint64_t bh = static_cast<uint64_t>(b) * p.heads + h;
              const int c_first = col_g0 + col0;
              ex.seed = has_drop ? *p.drop_seed + p.drop_salt : 0ull;
              ex.thresh = p.drop_thresh;
              ex.scale = p.drop_scale;
              if (kRowsAreKeys) {  // row = key, columns = queries
                ex.idx0 = (bh * p.sq + c_first) * p.skv + row_g;
                ex.idx_step = p.skv;
                ex.rb0 = row_g - c_first;
                ex.rb_step = -1;
                ex.vis_lo = row_g - off - c_first;   // key <= query + off  <=>  j >= key - off - first query
                ex.vis_hi = 1 << 30;
              } else {             // row = query, columns = keys
                ex.idx0 = (bh * p.sq + row_g) * p.skv + c_first;
                ex.idx_step = 1;
                ex.rb0 = c_first - row_g;
                ex.rb_step = 1;
                ex.vis_lo = -(1 << 30);
                ex.vis_hi = row_g + off - c_first;   // key <= query + off  <=>  j <= query + off - first key
              }
              ex.rb = rb;
              ex.rb_lo = -(p.sq - 1);
              ex.rb_hi = p.skv - 1;
            }
            if (MODE == kFwd && sweep == 0) {
              // ---- statistics sweep: running row maximum / sum of this half's columns
              if (valid == kFaSub) {  // (a ragged item reads TMEM columns no instruction wrote: per-element path)
                if (cols_plain) fa_stats_plain(t1, p.scale_log2, run_m, run_l);
                else if (diag && has_bias) fa_stats_cols<true, true>(t1, p.scale_log2, smem_u32(col + col0), run_m, run_l, ex);
                else if (diag) fa_stats_cols<false, true>(t1, p.scale_log2, smem_u32(col + col0), run_m, run_l, ex);
                else if (has_bias) fa_stats_cols<true, false>(t1, p.scale_log2, smem_u32(col + col0), run_m, run_l, ex);
                else fa_stats_cols<false, false>(t1, p.scale_log2, smem_u32(col + col0), run_m, run_l);
                continue;
              }
              float cmax = -INFINITY;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float v = __uint_as_float(t1[j]) * p.scale_log2;
                {
                  const int key = col_g0 + col0 + j;
                  bool ok = col0 + j < valid;
                  if (!cols_plain && ok) ok = col[col0 + j].x == 0.0f;
                  if (p.causal) ok = ok && key <= row_g + off;
                  if (rb != nullptr && ok) v += rb[key - row_g] * kFaLog2e;
                  if (!ok) v = -INFINITY;
                }
                t1[j] = __float_as_uint(v);
                cmax = fmaxf(cmax, v);
              }
              const float m_new = fmaxf(run_m, cmax);
              if (m_new != -INFINITY) {
                float sum = 0.0f;
#pragma unroll
                for (int j = 0; j < 32; ++j) sum += exp2f(__uint_as_float(t1[j]) - m_new);
                run_l = run_l * exp2f(run_m - m_new) + sum;
                run_m = m_new;
              }
              continue;
            }
            uint32_t u1[16], u2[16];
            {
              const uint32_t col_addr = smem_u32(col + col0);
              if (item_extras)
                fa_chunk_extras<kHasAcc1, kHasAcc2>(diag, has_drop, has_bias, t1, t2, u1, u2, p.scale_log2, neg_row, d_row,
                                                    col_addr, ex);
              else if (cols_plain)
                fa_chunk_plain<kHasAcc1, kHasAcc2, false, true>(t1, t2, u1, u2, p.scale_log2, neg_row, d_row, col_addr);
              else if (rows_plain)
                fa_chunk_plain<kHasAcc1, kHasAcc2, true, false>(t1, t2, u1, u2, p.scale_log2, neg_row, d_row, col_addr);
              else
                fa_chunk_plain<kHasAcc1, kHasAcc2, true, true>(t1, t2, u1, u2, p.scale_log2, neg_row, d_row, col_addr);
            }
            // bf16 pairs of columns [col0, col0 + 32) -> the first half of the columns just consumed
            if constexpr (kHasAcc1) tmem_st_16(t_slot + 16 * c2, u1);
            if constexpr (kHasAcc2) tmem_st_16(t_slot + 64 + 16 * c2, u2);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&u_ready[grp]);
          if (quarter == 0 && lane == 0) FA_TRACE(1 + grp, 400 + sweep * n_blk + kb);
        }
        if (MODE == kFwd && sweep == 0) {
          // ---- the two column halves of a row meet: lse = max + log2(sum)
          sStat[grp * kFaEdge + r_in] = make_float2(run_m, run_l);
          named_bar_sync(3, 256);
          const float2 o = sStat[(grp ^ 1) * kFaEdge + r_in];
          const float m = fmaxf(run_m, o.x);
          // sweep 1 forms P = exp2(s - max) (the row's largest term is exactly 1, as in an online softmax) and
          // the epilogue divides by the sum
          neg_row = INFINITY;
          fwd_scale = 0.0f;
          float lse = -INFINITY;
          if (m != -INFINITY && row_g < p.sq) {
            const float l = run_l * exp2f(run_m - m) + o.y * exp2f(o.x - m);
            neg_row = m;
            fwd_scale = 1.0f / l;
            lse = (m + log2f(l)) * 0.69314718055994531f;
          }
          if (grp == 0 && row_g < p.sq && p.lse_out != nullptr) p.lse_out[bh_off + row_g] = lse;
          named_bar_sync(3, 256);  // sStat is free for the next unit
        }
      }
      // ---- epilogue: accumulators -> global
      const bool have_acc = c_begin < c_end;
      if (have_acc) {
        mbar_wait(acc_full, acc_ph);
        acc_ph ^= 1u;
        tc_fence_after();
      } else if (MODE == kFwd && grp == 0 && row_g < p.sq && p.lse_out != nullptr) {
        p.lse_out[bh_off + row_g] = -INFINITY;  // a query tile that sees no key at all
      }
      const int n16 = p.dpad / 16;
      // kBwdKV: group 0 writes dV (acc1), group 1 writes dK (acc2); otherwise the groups split the 16-column blocks
      __nv_bfloat16* out = (MODE == kFwd || (MODE == kBwdKV && grp == 0)) ? p.out1 : p.out2;
      const bool first = MODE == kFwd || (MODE == kBwdKV && grp == 0);
      const long long bs = first ? p.out1_bs : p.out2_bs, rs = first ? p.out1_rs : p.out2_rs;
      const float mul = MODE == kFwd ? fwd_scale : (first ? 1.0f : p.out2_mul);
      const uint32_t acc_col = (MODE == kFwd || (MODE == kBwdKV && grp == 0)) ? col_acc1 : col_acc2;
      const int g_lo = MODE == kBwdKV ? 0 : (grp == 0 ? 0 : (n16 + 1) / 2);
      const int g_hi = MODE == kBwdKV ? n16 : (grp == 0 ? (n16 + 1) / 2 : n16);
      __nv_bfloat16* orow = out + b * bs + static_cast<long long>(row_g) * rs + h * p.d;
      for (int gi = g_lo; gi < g_hi; ++gi) {
        uint32_t r[16];
        if (have_acc) {
          tmem_ld_16(t_row + acc_col + gi * 16, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = 0u;
        }
        if (row_g < rows_total) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            const int c0 = gi * 16 + j;
            if (c0 < p.d) {  // d % 8 == 0
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(r[j]) * mul, __uint_as_float(r[j + 1]) * mul);
              u.y = pack_bf16x2(__uint_as_float(r[j + 2]) * mul, __uint_as_float(r[j + 3]) * mul);
              u.z = pack_bf16x2(__uint_as_float(r[j + 4]) * mul, __uint_as_float(r[j + 5]) * mul);
              u.w = pack_bf16x2(__uint_as_float(r[j + 6]) * mul, __uint_as_float(r[j + 7]) * mul);
              *reinterpret_cast<uint4*>(orow + c0) = u;
            }
          }
        }
      }
      if (have_acc) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_free);
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

// ------------------------------------------------------------------ host side
static bool fa_enabled(const char* env) {
  const char* e = std::getenv(env);  // "0": keep the mma.sync kernels (A/B measurements)
  return e == nullptr || e[0] != '0';
}

static bool fa_layout_ok(const vb_attn_args& f) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  // Dropout on the probabilities and the T5 relative bias: branch-free variants of the elementwise code
  // (fa_chunk_plain<.., DROP, BIAS>).  VB_ATTN_TC_SLOW=0 sends those calls to the mma.sync kernels (A/B).
  static const bool slow_ok = [] {
    const char* e = std::getenv("VB_ATTN_TC_SLOW");
    return e == nullptr || e[0] != '0';
  }();
  if (!slow_ok && (f.rel_bias != nullptr || (f.dropout_p > 0.0f && f.dropout_seed != nullptr))) return false;
  if (f.d % 16 != 0 || f.d < 16 || f.d > 128) return false;
  if (!al(f.q) || !al(f.k) || !al(f.v) || !al(f.o)) return false;
  if (f.q_rs % 8 || f.k_rs % 8 || f.v_rs % 8 || f.o_rs % 8 || f.o_bs % 8) return false;
  // batches back to back: row of (b, s) = b * S + s (the 3-D tensor maps index tokens by one coordinate)
  if (f.batch > 1 && (f.q_bs != f.sq * f.q_rs || f.k_bs != f.skv * f.k_rs || f.v_bs != f.skv * f.v_rs ||
                      f.o_bs != f.sq * f.o_rs))
    return false;
  if (f.batch * f.heads * ((f.skv + f.sq) / kFaEdge + 2) > (1ll << 30)) return false;
  if (f.batch * f.sq > (1ll << 30) || f.batch * f.skv > (1ll << 30)) return false;
  return true;
}

bool attention_flash_tcgen05_eligible(const vb_attn_args& f) {
  static const bool on = fa_enabled("VB_ATTN_FWD_TC");
  // The forward has queries on the 128 TMEM lanes: with fewer than one full row tile (the Q-Former's 32 queries)
  // most lanes idle and the mma.sync kernel's 64-row tiles are faster (profiles/r02_attn_flash.txt).  The
  // backward keeps the tcgen05 path there: its dK / dV pass has the keys on the lanes.
  return on && f.sq >= kFaEdge && fa_layout_ok(f);
}

bool attention_bwd_tcgen05_eligible(const vb_attn_bwd_args& a) {
  static const bool on = fa_enabled("VB_ATTN_BWD_TC");
  if (!on) return false;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if (!fa_layout_ok(a.fwd)) return false;
  if (a.fwd.lse == nullptr || a.delta == nullptr) return false;
  if (!al(a.d_o) || !al(a.dq) || !al(a.dk) || !al(a.dv)) return false;
  if (a.dq_rs % 8 || a.dk_rs % 8 || a.dv_rs % 8 || a.dq_bs % 8 || a.dk_bs % 8 || a.dv_bs % 8) return false;
  return true;
}

static int fa_sm_count() { return device_sm_count(); }

static cudaError_t fa_set_attrs() {
  static DeviceOnce attr_once;
  bool& attr = attr_once();
  if (attr) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(attn_flash_tc_kernel<kFwd>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(attn_flash_tc_kernel<kBwdQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(attn_flash_tc_kernel<kBwdKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmem);
  if (e == cudaSuccess) attr = true;
  return e;
}

static void fa_common(FaParams& p, const vb_attn_args& f) {
  p.lse = nullptr; p.lse_out = nullptr; p.delta = nullptr; p.delta_out = nullptr;
  p.o = nullptr; p.d_o = nullptr; p.o_bs = p.o_rs = 0;
  p.key_mask = f.key_mask;
  p.out1 = nullptr; p.out2 = nullptr;
  p.out1_bs = p.out1_rs = p.out2_bs = p.out2_rs = 0;
  p.out2_mul = 1.0f;
  p.batch = static_cast<int>(f.batch); p.heads = static_cast<int>(f.heads);
  p.sq = static_cast<int>(f.sq); p.skv = static_cast<int>(f.skv);
  p.d = static_cast<int>(f.d); p.dpad = (p.d + 15) / 16 * 16;
  p.causal = f.causal;
  p.scale_log2 = f.scale * kFaLog2e;
  const bool drop = f.dropout_p > 0.0f && f.dropout_seed != nullptr;
  p.drop_seed = reinterpret_cast<const unsigned long long*>(f.dropout_seed);
  p.drop_salt = f.dropout_salt;
  p.drop_thresh = drop ? dropout_threshold(f.dropout_p) : 0u;
  p.drop_scale = drop ? 1.0f / (1.0f - f.dropout_p) : 1.0f;
  p.rel_bias = f.rel_bias;
  p.rel_bias_stride = f.rel_bias_stride;
}

cudaError_t attention_flash_tcgen05_launch(const vb_attn_args& f, cudaStream_t stream) {
  // row-side operands travel as 128-row tiles, column-side operands as 64-row blocks
  CUtensorMap tq, tk, tv;
  if (!make_tmap_heads(&tq, f.q, f.batch * f.sq, f.heads, f.d, f.q_rs, kFaEdge)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&tk, f.k, f.batch * f.skv, f.heads, f.d, f.k_rs, kFaSub)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&tv, f.v, f.batch * f.skv, f.heads, f.d, f.v_rs, kFaSub)) return cudaErrorInvalidValue;
  cudaError_t e = fa_set_attrs();
  if (e != cudaSuccess) return e;
  FaParams p;
  fa_common(p, f);
  p.lse_out = f.lse;
  p.out1 = reinterpret_cast<__nv_bfloat16*>(f.o); p.out1_bs = f.o_bs; p.out1_rs = f.o_rs;
  p.row_tiles = (p.sq + kFaEdge - 1) / kFaEdge;
  p.col_blocks = (p.skv + kFaSub - 1) / kFaSub;
  const long long units = static_cast<long long>(p.batch) * p.heads * p.row_tiles;
  const int sms = fa_sm_count();
  return launch_pdl(attn_flash_tc_kernel<kFwd>, dim3(static_cast<unsigned>(units < sms ? units : sms)),
                    dim3(kFaThreads), kFaSmem, stream, tq, tq, tk, tv, p);
}

cudaError_t attention_bwd_tcgen05_launch(const vb_attn_bwd_args& a, cudaStream_t stream) {
  const vb_attn_args& f = a.fwd;
  // row-side operands travel as 128-row tiles (t*), column-side operands as 64-row blocks (b*)
  CUtensorMap tq, tk, tv, tdo, bq, bk, bv, bdo;
  if (!make_tmap_heads(&tq, f.q, f.batch * f.sq, f.heads, f.d, f.q_rs, kFaEdge)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&tk, f.k, f.batch * f.skv, f.heads, f.d, f.k_rs, kFaEdge)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&tv, f.v, f.batch * f.skv, f.heads, f.d, f.v_rs, kFaEdge)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&tdo, a.d_o, f.batch * f.sq, f.heads, f.d, f.o_rs, kFaEdge)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&bq, f.q, f.batch * f.sq, f.heads, f.d, f.q_rs, kFaSub)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&bk, f.k, f.batch * f.skv, f.heads, f.d, f.k_rs, kFaSub)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&bv, f.v, f.batch * f.skv, f.heads, f.d, f.v_rs, kFaSub)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&bdo, a.d_o, f.batch * f.sq, f.heads, f.d, f.o_rs, kFaSub)) return cudaErrorInvalidValue;
  cudaError_t e = fa_set_attrs();
  if (e != cudaSuccess) return e;
  const int sms = fa_sm_count();
  FaParams p;
  fa_common(p, f);
  p.lse = f.lse;
  const int q_tiles = (p.sq + kFaEdge - 1) / kFaEdge, k_tiles = (p.skv + kFaEdge - 1) / kFaEdge;
  const float dq_scale = a.dq_scale == 0.0f ? 1.0f : a.dq_scale;

  // dQ first: rows = queries; it also produces delta for the second pass
  p.out1 = nullptr; p.out1_bs = 0; p.out1_rs = 0;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a.dq); p.out2_bs = a.dq_bs; p.out2_rs = a.dq_rs;
  p.out2_mul = f.scale * dq_scale;
  p.delta = nullptr;
  p.delta_out = a.delta;
  p.o = reinterpret_cast<const __nv_bfloat16*>(f.o);
  p.d_o = reinterpret_cast<const __nv_bfloat16*>(a.d_o);
  p.o_bs = f.o_bs; p.o_rs = f.o_rs;
  p.row_tiles = q_tiles; p.col_blocks = (p.skv + kFaSub - 1) / kFaSub;
  long long units = static_cast<long long>(p.batch) * p.heads * p.row_tiles;
  e = launch_pdl(attn_flash_tc_kernel<kBwdQ>, dim3(static_cast<unsigned>(units < sms ? units : sms)),
                 dim3(kFaThreads), kFaSmem, stream, tq, tdo, bk, bv, p);
  if (e != cudaSuccess) return e;
  // dK, dV: rows = keys
  p.out1 = reinterpret_cast<__nv_bfloat16*>(a.dv); p.out1_bs = a.dv_bs; p.out1_rs = a.dv_rs;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a.dk); p.out2_bs = a.dk_bs; p.out2_rs = a.dk_rs;
  p.out2_mul = f.scale;
  p.delta = a.delta;
  p.delta_out = nullptr;
  p.row_tiles = k_tiles; p.col_blocks = (p.sq + kFaSub - 1) / kFaSub;
  units = static_cast<long long>(p.batch) * p.heads * p.row_tiles;
  return launch_pdl(attn_flash_tc_kernel<kBwdKV>, dim3(static_cast<unsigned>(units < sms ? units : sms)),
                    dim3(kFaThreads), kFaSmem, stream, tk, tv, bq, bdo, p);
}

#ifdef VB_FA_TRACE
extern "C" int vb_debug_attn_trace(long long* host_out, int* host_n, int reset) {
  int e = 0;
  if (host_out != nullptr) {
    e |= static_cast<int>(cudaMemcpyFromSymbol(host_out, g_fa_trace, sizeof(g_fa_trace)));
    e |= static_cast<int>(cudaMemcpyFromSymbol(host_n, g_fa_trace_n, sizeof(g_fa_trace_n)));
  }
  if (reset) {
    const int z[3] = {0, 0, 0};
    e |= static_cast<int>(cudaMemcpyToSymbol(g_fa_trace_n, z, sizeof(z)));
  }
  return e;
}
#endif

}  // namespace vb
