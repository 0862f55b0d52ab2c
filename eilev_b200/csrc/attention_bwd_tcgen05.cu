// Attention backward on the 5th-gen tensor cores (tcgen05 / TMEM / TMA): the OPT causal self-attention
// (d = 80, L <= 2048), the Q-Former self / cross attention and the T5 attentions (d = 64).
//
// One kernel body, instantiated twice.  A CTA owns a 128-row tile of the "row side" (TMEM lanes) of one
// (batch, head) and walks 128-wide tiles of the "column side":
//
//   TRANSPOSED = true   rows = keys   (R1 = K_j, R2 = V_j),   columns = queries (C1 = Q_i, C2 = dO_i)  -> dV, dK
//   TRANSPOSED = false  rows = queries (R1 = Q_i, R2 = dO_i), columns = keys    (C1 = K_j, C2 = V_j)   -> dQ
//
//   T1 = R1 . C1^T   (the scores, or their transpose)     tcgen05.mma SS, M = 128, N = 128, K = d
//   T2 = R2 . C2^T   (dP = dO . V^T, or its transpose)
//   P  = exp2(T1 * scale * log2 e - lse),  dS = P o (T2 - delta)          one thread per TMEM lane
//   acc1 += P  . C2  (dV; TRANSPOSED only)                  tcgen05.mma TS: P / dS are read from TMEM where they
//   acc2 += dS . C1  (dK, or dQ)                            overwrite T1 / T2 as bf16 pairs; C1 / C2 = MN-major B
//
// Because the dQ pass has queries on the lanes and the dK / dV pass has keys on the lanes, every product has its
// reduction index on the TMEM columns: no shared-memory transposes, no fp32 atomics, no dq_acc round trip (the
// mma.sync kernel's memset + atomics + convert), at the price of computing the scores and exponentials twice.
//
//   warp 0      TMA producer: R tiles once per unit, C tiles double-buffered, 3-D maps (d, head, token)
//   warp 1      MMA issuer
//   warps 4-11  elementwise + epilogue: lane quarter = warp & 3, column half = (warp - 4) / 4
//
// TMEM columns: [0,128) T1 / P, [128,256) T2 / dS, [256, 256 + dpad) acc1, [256 + dpad, 256 + 2 dpad) acc2.
// Masks are folded into the exponent: exponent = T1 * c - (neg_row + neg_col[column]) where an invalid key
// (beyond Skv, key-padding mask) or an invalid / fully masked query (beyond Sq, lse = -inf) contributes +inf, so
// P = 0 exactly; only causal-diagonal tiles, dropout and the T5 relative bias take the per-element path.
// Reference op: HF OPTAttention / Blip2QFormerMultiHeadAttention / T5Attention backward (autograd of
// softmax(QK^T * scale + mask) V), eilev/model/v2.py:132-252 calls them through language_model / qformer.
#include <cstdlib>

#include "common.cuh"
#include "internal.h"
#include "tc_attention.cuh"

namespace vb {

constexpr int kBtThreads = 384;
constexpr int kBtEdge = 128;                    // tile edge: rows (lanes) and columns
constexpr int kBtChunk = kBtEdge * 128;         // one 64-wide d chunk of a 128-row tile: 16 KB
constexpr int kBtTile = 2 * kBtChunk;           // d <= 128
constexpr int kBtSmem = 6 * kBtTile + 2 * kBtEdge * 8 + 256 + 1024;  // R1 R2, 2 x (C1 C2), column data, barriers

struct BtParams {
  const float* lse;        // (B, H, Sq) natural log
  const float* delta;      // (B, H, Sq) rowsum(dO o O)
  const uint8_t* key_mask; // (B, Skv) or nullptr
  __nv_bfloat16* out1;     // TRANSPOSED: dV
  __nv_bfloat16* out2;     // TRANSPOSED: dK, else dQ
  long long out1_bs, out1_rs, out2_bs, out2_rs;
  float out2_mul;          // softmax scale (dK) / scale * dq_scale (dQ)
  int batch, heads, sq, skv, d, dpad;
  int causal;
  float scale_log2;
  const unsigned long long* drop_seed;
  unsigned long long drop_salt;
  unsigned int drop_thresh;
  float drop_scale;
  const float* rel_bias;
  long long rel_bias_stride;
  int row_tiles, col_tiles;
};

// The n-th unit of CTA c: passes run forwards and backwards over the heavy-first unit list so that every CTA
// gets a heavy and a light unit (causal: the number of column tiles falls / rises with the row tile).
VB_DEVICE int bt_unit(int n, int units) {
  const int g = static_cast<int>(gridDim.x), c = static_cast<int>(blockIdx.x);
  return (n & 1) ? (n + 1) * g - 1 - c : n * g + c;
}

template <bool TRANSPOSED>
VB_DEVICE void bt_unit_coords(const BtParams& p, int unit, int& b, int& h, int& rt, int& c_begin, int& c_end) {
  const int bh = p.batch * p.heads;
  const int slot = unit / bh;  // heavy first
  const int rem = unit % bh;
  b = rem / p.heads;
  h = rem % p.heads;
  const int off = p.skv - p.sq;
  if (TRANSPOSED) {
    rt = slot;  // key tile: low tiles are seen by the most queries
    c_begin = 0;
    c_end = p.col_tiles;
    if (p.causal) {
      const int first_q = rt * kBtEdge - off;  // first query that sees the tile's first key
      c_begin = first_q <= 0 ? 0 : first_q / kBtEdge;
      if (c_begin > c_end) c_begin = c_end;
    }
  } else {
    rt = p.causal ? p.row_tiles - 1 - slot : slot;  // query tile: high tiles see the most keys
    c_begin = 0;
    c_end = p.col_tiles;
    if (p.causal) {
      const int last_key = rt * kBtEdge + kBtEdge - 1 + off;
      const int e = last_key < 0 ? 0 : last_key / kBtEdge + 1;
      if (e < c_end) c_end = e;
    }
  }
}

template <bool TRANSPOSED>
__global__ void __launch_bounds__(kBtThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_r1, const __grid_constant__ CUtensorMap tmap_r2,
                   const __grid_constant__ CUtensorMap tmap_c1, const __grid_constant__ CUtensorMap tmap_c2,
                   const BtParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sR1 = smem;
  uint8_t* sR2 = smem + kBtTile;
  uint8_t* sC = smem + 2 * kBtTile;            // [2 buffers][C1, C2]
  float2* sCol = reinterpret_cast<float2*>(smem + 6 * kBtTile);   // [2][128] (neg_col, delta_col)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kBtTile + 2 * kBtEdge * 8);
  uint64_t* r_full = bars;
  uint64_t* r_empty = bars + 1;
  uint64_t* c_full = bars + 2;    // [2]
  uint64_t* c_empty = bars + 4;   // [2]
  uint64_t* t_full = bars + 6;
  uint64_t* u_ready = bars + 7;   // 8 warps
  uint64_t* acc_full = bars + 8;
  uint64_t* acc_free = bars + 9;  // 8 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int units = p.batch * p.heads * p.row_tiles;
  const int n_chunks = (p.d + 63) / 64;              // 64-wide d chunks that hold data
  const int k_steps = (p.d + 15) / 16;
  const int rows_total = TRANSPOSED ? p.skv : p.sq;  // per batch
  const int cols_total = TRANSPOSED ? p.sq : p.skv;
  const uint32_t col_acc1 = 256, col_acc2 = 256 + static_cast<uint32_t>(p.dpad);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_r1);
    prefetch_tmap(&tmap_r2);
    prefetch_tmap(&tmap_c1);
    prefetch_tmap(&tmap_c2);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(r_full, 1);
    mbar_init(r_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&c_full[i], 1);
      mbar_init(&c_empty[i], 1);
    }
    mbar_init(t_full, 1);
    mbar_init(u_ready, 8);
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
    if (lane == 0) {
      uint32_t r_ph = 0;
      int ct = 0;  // running column-tile counter: buffer = ct & 1, phase = (ct >> 1) & 1
      for (int n = 0; n * static_cast<int>(gridDim.x) < units; ++n) {
        const int unit = bt_unit(n, units);
        if (unit >= units) continue;
        int b, h, rt, c_begin, c_end;
        bt_unit_coords<TRANSPOSED>(p, unit, b, h, rt, c_begin, c_end);
        if (c_begin >= c_end) continue;
        mbar_wait(r_empty, r_ph ^ 1u);
        r_ph ^= 1u;
        mbar_expect_tx(r_full, 2 * n_chunks * kBtChunk);
        const int r_row = b * rows_total + rt * kBtEdge;
        for (int c = 0; c < n_chunks; ++c) {
          tma_load_3d(sR1 + c * kBtChunk, &tmap_r1, r_full, c * 64, h, r_row);
          tma_load_3d(sR2 + c * kBtChunk, &tmap_r2, r_full, c * 64, h, r_row);
        }
        for (int t = c_begin; t < c_end; ++t, ++ct) {
          const int buf = ct & 1;
          mbar_wait(&c_empty[buf], ((ct >> 1) & 1) ^ 1u);
          mbar_expect_tx(&c_full[buf], 2 * n_chunks * kBtChunk);
          const int c_row = b * cols_total + t * kBtEdge;
          uint8_t* dst = sC + buf * 2 * kBtTile;
          for (int c = 0; c < n_chunks; ++c) {
            tma_load_3d(dst + c * kBtChunk, &tmap_c1, &c_full[buf], c * 64, h, c_row);
            tma_load_3d(dst + kBtTile + c * kBtChunk, &tmap_c2, &c_full[buf], c * 64, h, c_row);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_acc = umma_idesc_bf16(128, static_cast<uint32_t>(p.dpad)) | (1u << 16);  // B MN-major
      uint32_t r_ph = 0, u_ph = 0, free_ph = 0;
      int ct = 0;
      auto issue_scores = [&](int t) {
        const int buf = ct & 1;
        mbar_wait(&c_full[buf], (ct >> 1) & 1);
        tc_fence_after();
        // columns that hold data, rounded up to the instruction granularity (a short last tile: cross-attention
        // with 32 queries multiplies 32 columns, not 128)
        int n_valid = cols_total - t * kBtEdge;
        if (n_valid > kBtEdge) n_valid = kBtEdge;
        const uint32_t idesc_t = umma_idesc_bf16(128, static_cast<uint32_t>((n_valid + 15) / 16 * 16));
        const uint8_t* c1 = sC + buf * 2 * kBtTile;
        const uint8_t* c2 = c1 + kBtTile;
        for (int ks = 0; ks < k_steps; ++ks) {
          const int c = ks >> 2, kk = ks & 3;
          umma_bf16(tmem_base, umma_desc_k_sw128(smem_u32(sR1 + c * kBtChunk)) + 2 * kk,
                    umma_desc_k_sw128(smem_u32(c1 + c * kBtChunk)) + 2 * kk, idesc_t, ks != 0 ? 1u : 0u);
        }
        for (int ks = 0; ks < k_steps; ++ks) {
          const int c = ks >> 2, kk = ks & 3;
          umma_bf16(tmem_base + 128, umma_desc_k_sw128(smem_u32(sR2 + c * kBtChunk)) + 2 * kk,
                    umma_desc_k_sw128(smem_u32(c2 + c * kBtChunk)) + 2 * kk, idesc_t, ks != 0 ? 1u : 0u);
        }
        umma_commit(t_full);
      };
      for (int n = 0; n * static_cast<int>(gridDim.x) < units; ++n) {
        const int unit = bt_unit(n, units);
        if (unit >= units) continue;
        int b, h, rt, c_begin, c_end;
        bt_unit_coords<TRANSPOSED>(p, unit, b, h, rt, c_begin, c_end);
        if (c_begin >= c_end) continue;
        mbar_wait(r_full, r_ph);
        r_ph ^= 1u;
        issue_scores(c_begin);
        for (int t = c_begin; t < c_end; ++t) {
          const int buf = ct & 1;
          mbar_wait(u_ready, u_ph);
          u_ph ^= 1u;
          if (t == c_begin) {  // the previous unit's accumulators have been read out
            mbar_wait(acc_free, free_ph ^ 1u);
            free_ph ^= 1u;
          }
          tc_fence_after();
          int n_valid = cols_total - t * kBtEdge;
          if (n_valid > kBtEdge) n_valid = kBtEdge;
          const int u_steps = (n_valid + 15) / 16;
          const uint8_t* c1 = sC + buf * 2 * kBtTile;
          const uint8_t* c2 = c1 + kBtTile;
          for (int js = 0; js < u_steps; ++js) {
            const uint32_t a_col = static_cast<uint32_t>(64 * (js >> 2) + 8 * (js & 3));  // bf16 pairs of 16 columns
            const uint32_t accum = (t != c_begin || js != 0) ? 1u : 0u;
            if (TRANSPOSED)
              umma_bf16_ts(tmem_base + col_acc1, tmem_base + a_col,
                           umma_desc_mn_sw128(smem_u32(c2 + js * 16 * 128), kBtChunk), idesc_acc, accum);
            umma_bf16_ts(tmem_base + col_acc2, tmem_base + 128 + a_col,
                         umma_desc_mn_sw128(smem_u32(c1 + js * 16 * 128), kBtChunk), idesc_acc, accum);
          }
          umma_commit(&c_empty[buf]);
          ++ct;
          if (t + 1 < c_end) issue_scores(t + 1);
        }
        umma_commit(acc_full);
        umma_commit(r_empty);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ elementwise + epilogue
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r_in = quarter * 32 + lane;                        // row (TMEM lane) inside the tile
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int off = p.skv - p.sq;
    const bool slow_always = p.drop_thresh != 0u || p.rel_bias != nullptr;
    uint32_t t_ph = 0, acc_ph = 0;
    int tt = 0;  // running column-tile counter (per-column smem buffer = tt & 1)
    for (int n = 0; n * static_cast<int>(gridDim.x) < units; ++n) {
      const int unit = bt_unit(n, units);
      if (unit >= units) continue;
      int b, h, rt, c_begin, c_end;
      bt_unit_coords<TRANSPOSED>(p, unit, b, h, rt, c_begin, c_end);
      const int row_g = rt * kBtEdge + r_in;   // key (TRANSPOSED) or query index of this lane
      const float* lse_bh = p.lse + (static_cast<long long>(b) * p.heads + h) * p.sq;
      const float* delta_bh = p.delta + (static_cast<long long>(b) * p.heads + h) * p.sq;
      const uint8_t* km = p.key_mask != nullptr ? p.key_mask + static_cast<long long>(b) * p.skv : nullptr;
      const float* rb = p.rel_bias != nullptr ? p.rel_bias + h * p.rel_bias_stride + (p.sq - 1) : nullptr;
      // per-row terms: +inf in neg_row makes P = 0 for the whole row
      float neg_row, d_row = 0.0f;
      if (TRANSPOSED) {
        neg_row = (row_g < p.skv && (km == nullptr || km[row_g] != 0)) ? 0.0f : INFINITY;
      } else {
        neg_row = INFINITY;
        if (row_g < p.sq) {
          const float l = lse_bh[row_g];
          if (l != -INFINITY) neg_row = l * 1.4426950408889634f;
          d_row = delta_bh[row_g];
        }
      }
      for (int t = c_begin; t < c_end; ++t, ++tt) {
        // ---- per-column terms of this tile -> shared memory (threads of column half 0, one column each)
        float2* col = sCol + (tt & 1) * kBtEdge;
        if (half == 0) {
          const int cg = t * kBtEdge + r_in;
          float2 v = make_float2(INFINITY, 0.0f);
          if (TRANSPOSED) {
            if (cg < p.sq) {
              const float l = lse_bh[cg];
              if (l != -INFINITY) v.x = l * 1.4426950408889634f;
              v.y = delta_bh[cg];
            }
          } else {
            if (cg < p.skv && (km == nullptr || km[cg] != 0)) v.x = 0.0f;
          }
          col[r_in] = v;
        }
        named_bar_sync(1, 256);
        int n_valid = cols_total - t * kBtEdge;
        if (n_valid > kBtEdge) n_valid = kBtEdge;
        const int n_used = (n_valid + 15) / 16 * 16;  // columns the accumulate instructions read
        // causal: is every (row, column) pair of the tile visible?
        bool slow = slow_always;
        if (p.causal) {
          const int key_max = TRANSPOSED ? rt * kBtEdge + kBtEdge - 1 : t * kBtEdge + kBtEdge - 1;
          const int q_min = TRANSPOSED ? t * kBtEdge : rt * kBtEdge;
          slow = slow || key_max > q_min + off;
        }
        mbar_wait(t_full, t_ph);
        t_ph ^= 1u;
        tc_fence_after();
#pragma unroll 1
        for (int c2 = 0; c2 < 2; ++c2) {
          const int col0 = 64 * half + 32 * c2;
          if (col0 >= n_used) continue;
          uint32_t t1[32], t2[32];
          tmem_ld_32(t_row + col0, t1);
          tmem_ld_32(t_row + 128 + col0, t2);
          tmem_ld_wait();
          uint32_t u1[16], u2[16];
          if (!slow) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 ca = col[col0 + j], cb = col[col0 + j + 1];
              const float pa = exp2f(fmaf(__uint_as_float(t1[j]), p.scale_log2, -(neg_row + ca.x)));
              const float pb = exp2f(fmaf(__uint_as_float(t1[j + 1]), p.scale_log2, -(neg_row + cb.x)));
              const float da = pa * (__uint_as_float(t2[j]) - (d_row + ca.y));
              const float db = pb * (__uint_as_float(t2[j + 1]) - (d_row + cb.y));
              if (TRANSPOSED) u1[j >> 1] = pack_bf16x2(pa, pb);
              u2[j >> 1] = pack_bf16x2(da, db);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float pr[2], ds[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int cg = t * kBtEdge + col0 + j + e;
                const int key = TRANSPOSED ? row_g : cg;
                const int qi = TRANSPOSED ? cg : row_g;
                const float2 cc = col[col0 + j + e];
                float sv = __uint_as_float(t1[j + e]) * p.scale_log2;
                const bool vis = !p.causal || key <= qi + off;
                const bool inside = key < p.skv && qi < p.sq;
                if (rb != nullptr && inside) sv += rb[key - qi] * 1.4426950408889634f;
                const float pe = (vis && inside) ? exp2f(sv - (neg_row + cc.x)) : 0.0f;
                float p_used = pe, dpv = __uint_as_float(t2[j + e]);
                if (p.drop_thresh != 0u) {  // regenerate the forward mask
                  const uint64_t idx = ((static_cast<uint64_t>(b) * p.heads + h) * p.sq + qi) * p.skv + key;
                  const bool keep = dropout_keep(*p.drop_seed + p.drop_salt, idx, p.drop_thresh);
                  p_used = keep ? pe * p.drop_scale : 0.0f;
                  dpv = keep ? dpv * p.drop_scale : 0.0f;
                }
                pr[e] = p_used;
                ds[e] = pe * (dpv - (d_row + cc.y));
              }
              if (TRANSPOSED) u1[j >> 1] = pack_bf16x2(pr[0], pr[1]);
              u2[j >> 1] = pack_bf16x2(ds[0], ds[1]);
            }
          }
          // bf16 pairs of columns [col0, col0 + 32) -> the first half of the columns just consumed
          if (TRANSPOSED) tmem_st_16(t_row + 64 * half + 16 * c2, u1);
          tmem_st_16(t_row + 128 + 64 * half + 16 * c2, u2);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(u_ready);
      }
      // ---- epilogue: accumulators -> global
      const bool have_acc = c_begin < c_end;
      if (have_acc) {
        mbar_wait(acc_full, acc_ph);
        acc_ph ^= 1u;
        tc_fence_after();
      }
      const int n16 = p.dpad / 16;
      if (TRANSPOSED) {
        // column half 0 writes dV, half 1 writes dK
        __nv_bfloat16* out = half == 0 ? p.out1 : p.out2;
        const long long bs = half == 0 ? p.out1_bs : p.out2_bs, rs = half == 0 ? p.out1_rs : p.out2_rs;
        const float mul = half == 0 ? 1.0f : p.out2_mul;
        const uint32_t acc_col = half == 0 ? col_acc1 : col_acc2;
        __nv_bfloat16* orow = out + b * bs + static_cast<long long>(row_g) * rs + h * p.d;
        for (int gi = 0; gi < n16; ++gi) {
          uint32_t r[16];
          if (have_acc) {
            tmem_ld_16(t_row + acc_col + gi * 16, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = 0u;
          }
          if (row_g < p.skv) {
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
      } else {
        // dQ: the two column halves split the 16-column groups
        const int g_lo = half == 0 ? 0 : (n16 + 1) / 2, g_hi = half == 0 ? (n16 + 1) / 2 : n16;
        __nv_bfloat16* orow = p.out2 + b * p.out2_bs + static_cast<long long>(row_g) * p.out2_rs + h * p.d;
        for (int gi = g_lo; gi < g_hi; ++gi) {
          uint32_t r[16];
          if (have_acc) {
            tmem_ld_16(t_row + col_acc2 + gi * 16, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = 0u;
          }
          if (row_g < p.sq) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              const int c0 = gi * 16 + j;
              if (c0 < p.d) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(r[j]) * p.out2_mul, __uint_as_float(r[j + 1]) * p.out2_mul);
                u.y = pack_bf16x2(__uint_as_float(r[j + 2]) * p.out2_mul, __uint_as_float(r[j + 3]) * p.out2_mul);
                u.z = pack_bf16x2(__uint_as_float(r[j + 4]) * p.out2_mul, __uint_as_float(r[j + 5]) * p.out2_mul);
                u.w = pack_bf16x2(__uint_as_float(r[j + 6]) * p.out2_mul, __uint_as_float(r[j + 7]) * p.out2_mul);
                *reinterpret_cast<uint4*>(orow + c0) = u;
              }
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
bool attention_bwd_tcgen05_eligible(const vb_attn_bwd_args& a) {
  static const bool on = [] {
    const char* e = std::getenv("VB_ATTN_BWD_TC");  // 0: keep the mma.sync kernel (A/B measurements)
    return e == nullptr || e[0] != '0';
  }();
  if (!on) return false;
  const vb_attn_args& f = a.fwd;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if (f.d % 16 != 0 || f.d < 16 || f.d > 128) return false;
  if (f.lse == nullptr || a.delta == nullptr) return false;
  if (!al(f.q) || !al(f.k) || !al(f.v) || !al(a.d_o) || !al(a.dq) || !al(a.dk) || !al(a.dv)) return false;
  if (f.q_rs % 8 || f.k_rs % 8 || f.v_rs % 8 || f.o_rs % 8 || a.dq_rs % 8 || a.dk_rs % 8 || a.dv_rs % 8) return false;
  if (a.dq_bs % 8 || a.dk_bs % 8 || a.dv_bs % 8) return false;
  // batches back to back: row of (b, s) = b * S + s (the 3-D tensor maps index tokens by one coordinate)
  if (f.batch > 1 && (f.q_bs != f.sq * f.q_rs || f.k_bs != f.skv * f.k_rs || f.v_bs != f.skv * f.v_rs ||
                      f.o_bs != f.sq * f.o_rs))
    return false;
  if (f.batch * f.heads * ((f.skv + kBtEdge - 1) / kBtEdge) > (1ll << 30)) return false;
  if (f.batch * f.sq > (1ll << 30) || f.batch * f.skv > (1ll << 30)) return false;
  return true;
}

cudaError_t attention_bwd_tcgen05_launch(const vb_attn_bwd_args& a, cudaStream_t stream) {
  const vb_attn_args& f = a.fwd;
  CUtensorMap tq, tk, tv, tdo;
  if (!make_tmap_heads(&tq, f.q, f.batch * f.sq, f.heads, f.d, f.q_rs, kBtEdge)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&tk, f.k, f.batch * f.skv, f.heads, f.d, f.k_rs, kBtEdge)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&tv, f.v, f.batch * f.skv, f.heads, f.d, f.v_rs, kBtEdge)) return cudaErrorInvalidValue;
  if (!make_tmap_heads(&tdo, a.d_o, f.batch * f.sq, f.heads, f.d, f.o_rs, kBtEdge)) return cudaErrorInvalidValue;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBtSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBtSmem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  BtParams p;
  p.lse = f.lse;
  p.delta = a.delta;
  p.key_mask = f.key_mask;
  p.batch = static_cast<int>(f.batch); p.heads = static_cast<int>(f.heads);
  p.sq = static_cast<int>(f.sq); p.skv = static_cast<int>(f.skv);
  p.d = static_cast<int>(f.d); p.dpad = (p.d + 15) / 16 * 16;
  p.causal = f.causal;
  p.scale_log2 = f.scale * 1.4426950408889634f;
  const bool drop = f.dropout_p > 0.0f && f.dropout_seed != nullptr;
  p.drop_seed = reinterpret_cast<const unsigned long long*>(f.dropout_seed);
  p.drop_salt = f.dropout_salt;
  p.drop_thresh = drop ? dropout_threshold(f.dropout_p) : 0u;
  p.drop_scale = drop ? 1.0f / (1.0f - f.dropout_p) : 1.0f;
  p.rel_bias = f.rel_bias;
  p.rel_bias_stride = f.rel_bias_stride;
  const int q_tiles = (p.sq + kBtEdge - 1) / kBtEdge, k_tiles = (p.skv + kBtEdge - 1) / kBtEdge;
  const float dq_scale = a.dq_scale == 0.0f ? 1.0f : a.dq_scale;

  // dK, dV: rows = keys
  p.out1 = reinterpret_cast<__nv_bfloat16*>(a.dv); p.out1_bs = a.dv_bs; p.out1_rs = a.dv_rs;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a.dk); p.out2_bs = a.dk_bs; p.out2_rs = a.dk_rs;
  p.out2_mul = f.scale;
  p.row_tiles = k_tiles; p.col_tiles = q_tiles;
  long long units = static_cast<long long>(p.batch) * p.heads * p.row_tiles;
  cudaError_t e = launch_pdl(attn_bwd_tc_kernel<true>, dim3(static_cast<unsigned>(units < sms ? units : sms)),
                             dim3(kBtThreads), kBtSmem, stream, tk, tv, tq, tdo, p);
  if (e != cudaSuccess) return e;
  // dQ: rows = queries
  p.out1 = nullptr; p.out1_bs = 0; p.out1_rs = 0;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a.dq); p.out2_bs = a.dq_bs; p.out2_rs = a.dq_rs;
  p.out2_mul = f.scale * dq_scale;
  p.row_tiles = q_tiles; p.col_tiles = k_tiles;
  units = static_cast<long long>(p.batch) * p.heads * p.row_tiles;
  return launch_pdl(attn_bwd_tc_kernel<false>, dim3(static_cast<unsigned>(units < sms ? units : sms)),
                    dim3(kBtThreads), kBtSmem, stream, tq, tdo, tk, tv, p);
}

}  // namespace vb
