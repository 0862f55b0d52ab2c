// CTA-pair variant of the tcgen05 GEMM (tcgen05.mma.cta_group::2): two CTAs of one cluster
// (one TPC) compute a 256 x tn tile together.  tn (<= 256, a multiple of 16) is a RUN-TIME width: for launches
// of two or three waves the host picks the width that fills whole waves of the 74 pairs (OPT at M = 976:
// N = 7680 -> 208, 148 tiles in two waves; N = 10240 -> 192); the TMA box, the transaction bytes, the
// shared-memory plan and the UMMA instruction descriptor follow it.  Each CTA stages its own 128 rows of A and
// HALF of the B tile; the pair's tensor cores read both halves, so per-SM shared-memory fill
// traffic and L2->SM traffic drop by a third against the single-CTA 128 x BN tile and the
// freed shared memory buys a deeper TMA ring (6-7 stages).  Used for the large-M launches
// (the ViT and the cross-attention K/V projection); small-M launches keep the 1-CTA kernel.
//
//   warp 0   TMA producer (both CTAs; transaction bytes are credited to the leader's barrier)
//   warp 1   MMA issuer   (leader CTA only, one thread; commits multicast to both CTAs)
//   warp 2   TMEM allocator (cta_group::2)
//   warps 4-19 epilogue (each CTA drains its own 128 accumulator rows; four warps per TMEM lane quarter,
//              one per 64-column slab of the tile: the K = 1408 launches of the ViT are epilogue-bound, and an
//              epilogue of dependent MUFU / FMA chains needs four warps per scheduler to hide its latencies —
//              profiles/r02_ncu_gemm_shapes.txt)
#include "common.cuh"
#include "gemm.h"
#include "gemm_epilogue.cuh"

namespace vb {

// -DVB_GEMM_TRACE (scripts/micro/gemm_trace.sh): per-CTA clock64 stamps of the kernel's phases, read back through
// vb_debug_gemm_trace.  Off in the product build.
#ifdef VB_GEMM_TRACE
__device__ unsigned long long g_trace[296 * 16];
#define VB_TRACE(i) g_trace[blockIdx.x * 16 + (i)] = clock64()
#else
#define VB_TRACE(i)
#endif

constexpr int k2BM = 128;  // rows per CTA (256 per pair)
constexpr int k2BK = 64;
constexpr int k2EpiWarps = 16;
constexpr int k2Threads = (4 + k2EpiWarps) * 32;  // 640: <= 96 registers per thread

// Shared-memory plan of one launch (all sizes follow the run-time tile width tn):
//   [stages x A tile 16 KB][stages x this CTA's half of B, tn/2 rows x 128 B][tn/64 output slabs x 16 KB][barriers]
// The output staging of the TMA-store epilogue takes one [128 rows][64 cols] bf16 slab (SWIZZLE_128B) per whole
// 64-column group of the tile; what a narrow tile does not need buys TMA stages (tn = 144: 7 stages, 256: 5).
struct Gemm2Plan {
  static constexpr int kABytes = k2BM * k2BK * 2;
  static constexpr int kSlabBytes = k2BM * 128;
  static constexpr int kMaxStages = 8;
  static constexpr int kSmemBytes = 227 * 1024;             // the whole opt-in maximum
  static constexpr int kBudget = kSmemBytes - 1024 - 256;   // alignment slack + barriers
  static int b_bytes(int tn) { return (tn / 2) * k2BK * 2; }
  static int slabs(int tn) { return tn / 64; }
  static int stages(int tn) {
    const int s = (kBudget - slabs(tn) * kSlabBytes) / (kABytes + b_bytes(tn));
    return s > kMaxStages ? kMaxStages : s;
  }
};


template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_c, const EpiParams p,
                         const int num_k_blocks, const int m_tiles, const int n_tiles, const int tn,
                         const int kStages) {
  using Cfg = Gemm2Plan;
  static_assert(BN == 256, "one instantiation: the tile width is the run-time tn <= 256");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int b_bytes = (tn / 2) * (k2BK * 2);   // a multiple of 1024 (tn % 16 == 0)
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint8_t* smem_c = smem_b + kStages * b_bytes;  // 1024-aligned: stage sizes are multiples of 1024
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_c + (tn / 64) * Cfg::kSlabBytes);
  uint64_t* empty_bar = full_bar + Cfg::kMaxStages;
  uint64_t* tmem_full = empty_bar + Cfg::kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) VB_TRACE(0);
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int num_tiles = m_tiles * n_tiles;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  // width of the last column block: the real columns rounded up to 32 (UMMA N % 16 per CTA half), <= BN
  const int n_rem = static_cast<int>(p.n - static_cast<long long>(n_tiles - 1) * tn);
  const int n_last = (n_rem + 31) / 32 * 32 < tn ? (n_rem + 31) / 32 * 32 : tn;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (p.tma_store) prefetch_tmap(&tmap_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 2);   // one arrival per CTA of the pair (used on the leader)
      mbar_init(&empty_bar[s], 1);  // multicast commit from the leader's MMA thread
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * k2EpiWarps);  // epilogue warps of both CTAs (leader only)
    }
    fence_barrier_init();
  }
  cluster_sync_all();  // peer barriers are initialised before any remote arrive / multicast
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) VB_TRACE(1);
  // Programmatic dependent launch: the set-up above may overlap the tail of the preceding kernel
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) VB_TRACE(2);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    // (elect_one, not lane == 0: ptxas keeps the operands of a region guarded by elect.sync in uniform registers;
    // under a lane test every TMA / MMA instruction is wrapped in a waterfall loop that re-derives them)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = 2u * static_cast<uint32_t>(Cfg::kABytes + b_bytes);  // both CTAs
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const int row_a = m_blk * (2 * k2BM) + static_cast<int>(cta_rank) * k2BM;
        // the pair's B tile is split in halves along N; a narrower last column block (n_last < BN) splits
        // its own width, so each CTA's half starts n_last / 2 rows apart
        const int tile_n = (n_blk == n_tiles - 1) ? n_last : tn;
        const int row_b = n_blk * tn + static_cast<int>(cta_rank) * (tile_n / 2);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (leader) mbar_expect_tx(&full_bar[stage], stage_tx);
          else mbar_arrive_remote(&full_bar[stage], 0);
          tma_load_2d_2sm(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * k2BK, row_a);
          tma_load_2d_2sm(smem_b + stage * b_bytes, &tmap_b, &full_bar[stage], kb * k2BK, row_b);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader only)
    if (leader && elect_one()) {
      const uint32_t idesc_full = umma_idesc_bf16(2 * k2BM, static_cast<uint32_t>(tn));
      // last column block of an N that is not a multiple of BN (ViT: 1408 = 5 x 256 + 128): issue the MMA at
      // the width that holds real columns (rounded up to 32) instead of multiplying zero-filled rows
      const uint32_t idesc_last = umma_idesc_bf16(2 * k2BM, static_cast<uint32_t>(n_last));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        const uint32_t idesc = (tile % n_tiles == n_tiles - 1) ? idesc_last : idesc_full;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
#ifdef VB_GEMM_TRACE
          if (tile == pair && kb < 4) VB_TRACE(3 + kb);
#endif
          const uint64_t a_desc = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t b_desc = umma_desc_k_sw128(smem_u32(smem_b + stage * b_bytes));
#pragma unroll
          for (int k = 0; k < k2BK / 16; ++k)
            umma_bf16_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage]);  // frees the slot in both CTAs
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(&tmem_full[acc]);  // accumulators of both CTAs complete
        VB_TRACE(7);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (own 128 rows)
    const int ew = warp - 4;
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int cq = ew >> 2;         // column group: 64-column slab (staged path) / quarter of the chunks
    // 16-column chunks of this warp's column group.  With the TMA store the groups own whole 64-column slabs
    // (a tile width that is not a multiple of 64 leaves its last, partial slab to direct stores); without it
    // the chunks are split evenly.
    const int chunks = tn / 16;
    const int group_chunks = p.tma_store ? 4 : (chunks + 3) / 4;
    const int c_begin = cq * group_chunks < chunks ? cq * group_chunks : chunks;
    const int c_end = (c_begin + group_chunks < chunks) ? c_begin + group_chunks : chunks;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const long long row = static_cast<long long>(m_blk) * (2 * k2BM) + cta_rank * k2BM + quarter * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 256;
      if (p.tma_store && (cq + 1) * 64 <= tn) {
        // ---- staged path: the four warps of column group `cq` fill one [128 rows][64 cols] slab
        // (bias / LayerNorm fold / activation / residual), one thread stores it with TMA.
        const int row_in = quarter * 32 + lane;              // row inside this CTA's 128 rows
        const long long tile_row0 = static_cast<long long>(m_blk) * (2 * k2BM) + cta_rank * k2BM;
        const bool has_res = p.residual != nullptr;
        const long long col_slab = static_cast<long long>(n_blk) * tn + cq * 64;
        uint8_t* slab = smem_c + cq * Cfg::kSlabBytes;
        const bool slab_live = col_slab < p.n;
        // the TMA store that last read this slab (previous tile) must have drained
        if (quarter == 0 && lane == 0) bulk_wait_read<0>();
        named_bar_sync(1 + cq, 128);
        if (has_res && slab_live) {
          // coalesced residual fetch, in flight while the accumulator is still being computed:
          // 8 lanes cover one 128-byte row, 4 rows per instruction
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = quarter * 32 + i * 4 + (lane >> 3);
            const int c16b = lane & 7;
            const long long grow = tile_row0 + r, gcol = col_slab + c16b * 8;
            uint8_t* dst = slab + r * 128 + ((c16b ^ (r & 7)) << 4);
            if (grow < p.m && gcol + 8 <= p.n) cp_async_16(dst, p.residual + grow * p.ldr + gcol);
            else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
          }
        }
        // per-row LayerNorm coefficients of a folded LayerNorm: fetched while the accumulator is in flight
        float2 ln_c = make_float2(1.0f, 0.0f);
        if (p.ln_stats != nullptr && row < p.m) ln_c = ln_fold_coeffs(p, row);
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        if (ew == 0 && lane == 0) VB_TRACE(8);
        if (has_res && slab_live) {
          cp_async_commit_wait_all();
          __syncwarp();  // each warp reads back only rows it fetched itself
        }
        float st_s = 0.0f, st_q = 0.0f;  // row statistics of this thread's 64 output columns
        uint8_t* slab_row = slab + row_in * 128;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t r0[16], r1[16];
          tmem_ld_16(t_row + cq * 64 + hf * 32, r0);
          tmem_ld_16(t_row + cq * 64 + hf * 32 + 16, r1);
          tmem_ld_wait();
          if (hf == 1) {  // last TMEM read of this warp for this tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&tmem_empty[acc], 0);
          }
          if (slab_live) {
            epilogue_row16_staged(p, row, col_slab + hf * 32, r0, slab_row, row_in & 7, 2 * hf, has_res, ln_c,
                                  st_s, st_q);
            epilogue_row16_staged(p, row, col_slab + hf * 32 + 16, r1, slab_row, row_in & 7, 2 * hf + 1, has_res,
                                  ln_c, st_s, st_q);
          }
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA engine
        named_bar_sync(1 + cq, 128);
        if (quarter == 0 && elect_one()) {
          if (slab_live) tma_store_2d(&tmap_c, slab, static_cast<int>(col_slab), static_cast<int>(tile_row0));
          bulk_commit();  // (an empty group keeps the wait_group bookkeeping uniform)
          if (cq == 0) VB_TRACE(9);
        }
        if (p.stats_zero != nullptr && n_blk == 0 && cq == 0 && row < p.m)
          *reinterpret_cast<double2*>(p.stats_zero + 2 * row) = make_double2(0.0, 0.0);
        if (p.stats_out != nullptr && row < p.m && slab_live) {
          atomicAdd(p.stats_out + 2 * row, static_cast<double>(st_s));
          atomicAdd(p.stats_out + 2 * row + 1, static_cast<double>(st_q));
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        continue;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      int ch = c_begin;
      for (; ch + 1 < c_end; ch += 2) {  // two 16-column loads in flight per wait
        uint32_t r0[16], r1[16];
        tmem_ld_16(t_row + ch * 16, r0);
        tmem_ld_16(t_row + (ch + 1) * 16, r1);
        tmem_ld_wait();
        if (ch + 2 >= c_end) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&tmem_empty[acc], 0);
        }
        epilogue_row16(p, row, static_cast<long long>(n_blk) * tn + ch * 16, r0);
        epilogue_row16(p, row, static_cast<long long>(n_blk) * tn + (ch + 1) * 16, r1);
      }
      if (ch < c_end) {
        uint32_t r0[16];
        tmem_ld_16(t_row + ch * 16, r0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&tmem_empty[acc], 0);
        epilogue_row16(p, row, static_cast<long long>(n_blk) * tn + ch * 16, r0);
      } else if (c_begin >= c_end) {  // a column group without chunks (narrow tiles) still releases the stage
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&tmem_empty[acc], 0);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  if (warp >= 4 && ((warp - 4) & 3) == 0 && lane == 0) bulk_wait_read<0>();  // smem must outlive the stores
  if (threadIdx.x == 128) VB_TRACE(10);
  tc_fence_before();
  cluster_sync_all();  // both CTAs are done with TMEM / remote barriers
  if (threadIdx.x == 0) VB_TRACE(11);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
    if (lane == 0) VB_TRACE(12);
  }
}

#ifdef VB_GEMM_TRACE
extern "C" int vb_debug_gemm_trace(unsigned long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, g_trace, sizeof(g_trace)));
}
#endif

bool make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, long long rows, long long cols,
                       long long ld, int box_rows);
void fill_epi_params(EpiParams& ep, const vb_gemm_args& a);

static int sm_count() { return device_sm_count(); }

// tn: tile width, a multiple of 16 in [32, 256]
cudaError_t gemm_tcgen05_2cta_launch(const vb_gemm_args& a, int tn, cudaStream_t stream) {
  constexpr int BN = 256;
  using Cfg = Gemm2Plan;
  if (tn < 32 || tn > BN || tn % 16 != 0) return cudaErrorInvalidValue;
  CUtensorMap ta, tb;
  if (!make_tmap_bf16_2d(&ta, a.a, a.m, a.k, a.lda, k2BM)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16_2d(&tb, a.b, a.n, a.k, a.ldb, tn / 2)) return cudaErrorInvalidValue;
  static DeviceOnce attr_set_once;
  bool& attr_set = attr_set_once();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  EpiParams ep;
  fill_epi_params(ep, a);
  // TMA-store epilogue: bf16 output, plain row mapping, no accumulation into C
  CUtensorMap tc = ta;
  ep.tma_store = 0;
  if (tn >= 64 && a.out_dtype == VB_BF16 && a.beta == 0.0f && a.row_group == 0) {
    if (make_tmap_bf16_2d(&tc, a.c, a.m, a.n, a.ldc, k2BM)) ep.tma_store = 1;
  }
  const int m_tiles = static_cast<int>((a.m + 2 * k2BM - 1) / (2 * k2BM));
  const int n_tiles = static_cast<int>((a.n + tn - 1) / tn);
  const int k_blocks = static_cast<int>((a.k + k2BK - 1) / k2BK);
  const int sms = sm_count();
  const long long tiles = static_cast<long long>(m_tiles) * n_tiles;
  long long pairs = sms / 2;
  if (tiles < pairs) pairs = tiles;
  return launch_pdl(gemm_tcgen05_2cta_kernel<BN>, dim3(static_cast<unsigned>(2 * pairs)), dim3(k2Threads),
                    Cfg::kSmemBytes, stream, ta, tb, tc, ep, k_blocks, m_tiles, n_tiles, tn, Cfg::stages(tn));
}

}  // namespace vb
