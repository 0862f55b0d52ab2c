// bf16 GEMM on the 5th-generation tensor cores:  C = epi(A . B^T)
//   A (M,K) and B (N,K) K-major bf16, fp32 accumulation in TMEM.
//
// Persistent, warp-specialised CTA (one per SM):
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128-byte swizzle, kStages ring)
//   warp 1      MMA issuer     (one thread, tcgen05.mma cta_group::1, 128 x BN x 16)
//   warp 2      TMEM allocator (512 columns = two accumulator stages)
//   warps 4-11  epilogue       (tcgen05.ld -> bias / alpha / GELU|ReLU / residual -> global)
// The two TMEM accumulator stages let the epilogue of tile i overlap the main loop of
// tile i+1.  Used for every dense projection of the VideoBLIP path (see
// include/videoblip_b200.h, vb_gemm).
#include <mutex>

#include "common.cuh"
#include "gemm.h"
#include "gemm_epilogue.cuh"
#include "tc_attention.cuh"

namespace vb {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM column offset between the two accumulator stages

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "invalid UMMA N");
  static_assert(kBBytes % 1024 == 0, "B stage must keep 1024B alignment");
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b, const EpiParams p,
                    const int num_k_blocks, const int m_tiles, const int n_tiles, const int transposed) {
  // transposed = 1: both operands are given TRANSPOSED — a = A^T stored (K, M), b = B^T stored (K, N), row-major —
  // and enter the instruction as MN-major tiles (rows = K index, 64 contiguous M / N elements per swizzled row):
  // the wgrad product dW = dY^T X straight from the (tokens, features) activations, no transpose pass.
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled operand tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch)
  // may run under the tail of the preceding kernel; its results are visible from here on.
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {  // (not lane == 0: see gemm_tcgen05_2cta.cu)
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (!transposed) {
            tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * kBK,
                        m_blk * kBM);
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * kBK,
                        n_blk * BN);
          } else {  // boxes of 64 (M or N, contiguous) x 64 (K rows): one per 64-wide chunk of the tile
            for (int c = 0; c < kBM / 64; ++c)
              tma_load_2d(smem_a + stage * Cfg::kABytes + c * (kBK * 128), &tmap_a, &full_bar[stage],
                          m_blk * kBM + 64 * c, kb * kBK);
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(smem_b + stage * Cfg::kBBytes + c * (kBK * 128), &tmap_b, &full_bar[stage],
                          n_blk * BN + 64 * c, kb * kBK);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (!transposed) {
            const uint64_t a_desc = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
            const uint64_t b_desc = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in >>4 units
              umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          } else {
            // MN-major: 16 K elements = 16 rows of 128 B; the next 64 M / N elements sit one chunk (8 KB) further
            constexpr uint32_t idesc_mn = idesc | (1u << 15) | (1u << 16);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t a_desc = umma_desc_mn_sw128(smem_u32(smem_a + stage * Cfg::kABytes + k * 16 * 128), kBK * 128);
              const uint64_t b_desc = umma_desc_mn_sw128(smem_u32(smem_b + stage * Cfg::kBBytes + k * 16 * 128), kBK * 128);
              umma_bf16(d_tmem, a_desc, b_desc, idesc_mn, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = ew >> 2;      // which half of the column chunks
    constexpr int kChunks = BN / 16;
    constexpr int kHalfChunks = (kChunks + 1) / 2;
    const int c_begin = half * kHalfChunks;
    const int c_end = (c_begin + kHalfChunks < kChunks) ? c_begin + kHalfChunks : kChunks;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const long long row = static_cast<long long>(m_blk) * kBM + quarter * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             acc * kAccStride;
      for (int ch = c_begin; ch < c_end; ++ch) {
        uint32_t r[16];
        tmem_ld_16(t_row + ch * 16, r);
        tmem_ld_wait();
        if (ch == c_end - 1) {
          // all TMEM reads of this warp for this tile are done: release the stage
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        epilogue_row16(p, row, static_cast<long long>(n_blk) * BN + ch * 16, r);
      }
      if (c_begin >= c_end) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

namespace {
struct TmapSlot {
  TmapKey key;
  CUtensorMap map;
  bool valid = false;
};
constexpr int kTmapSlots = 512;
TmapSlot g_tmap_slots[kTmapSlots];
std::mutex g_tmap_mutex;
unsigned tmap_hash(const TmapKey& k) {
  unsigned long long h = 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < 8; ++i) h = (h ^ k.w[i]) * 0xBF58476D1CE4E5B9ull + (h >> 29);
  return static_cast<unsigned>(h >> 40) % kTmapSlots;
}
}  // namespace

bool tmap_cache_lookup(const TmapKey& key, CUtensorMap* out) {
  std::lock_guard<std::mutex> lock(g_tmap_mutex);
  const TmapSlot& s = g_tmap_slots[tmap_hash(key)];
  if (!s.valid || !(s.key == key)) return false;
  *out = s.map;
  return true;
}
void tmap_cache_store(const TmapKey& key, const CUtensorMap& map) {
  std::lock_guard<std::mutex> lock(g_tmap_mutex);
  TmapSlot& s = g_tmap_slots[tmap_hash(key)];
  s.key = key;
  s.map = map;
  s.valid = true;
}

// (rows, cols) bf16 row-major matrix with row stride ld; box = 64 columns x box_rows rows.
bool make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, long long rows, long long cols,
                       long long ld, int box_rows) {
  const TmapKey key = {{2ull, reinterpret_cast<unsigned long long>(ptr), static_cast<unsigned long long>(rows),
                        static_cast<unsigned long long>(cols), static_cast<unsigned long long>(ld),
                        static_cast<unsigned long long>(box_rows), 0ull, 0ull}};
  if (tmap_cache_lookup(key, map)) return true;
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) tmap_cache_store(key, *map);
  return r == CUDA_SUCCESS;
}

bool gemm_tcgen05_eligible(const vb_gemm_args& a) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (a.m <= 0 || a.n <= 0 || a.k <= 0) return false;
  if (a.reserved2 == 1) {  // transposed operands: the contiguous dims are M and N
    if (a.m % 8 != 0 || a.n % 8 != 0 || a.lda % 8 != 0 || a.ldb % 8 != 0) return false;
  } else if (a.k % 8 != 0 || a.lda % 8 != 0 || a.ldb % 8 != 0) {
    return false;
  }
  if (!al16(a.a) || !al16(a.b) || !al16(a.c)) return false;
  if (a.n % 8 != 0) return false;
  if (a.out_dtype == VB_BF16 ? (a.ldc % 8 != 0) : (a.ldc % 4 != 0)) return false;
  if (a.bias != nullptr && !al16(a.bias)) return false;
  if (a.residual != nullptr && (!al16(a.residual) || a.ldr % 8 != 0)) return false;
  if (a.m > (1ll << 31) - 256 || a.n > (1ll << 31) - 256 || a.k > (1ll << 31) - 256) return false;
  return true;
}

static int pick_block_n(long long m, long long n) {
  // Prefer the widest tile that divides N; 176 covers the ViT widths 1408 / 4224.
  const int cands[4] = {256, 176, 128, 64};
  int best = 64;
  double best_cost = 1e300;
  const long long m_tiles = (m + kBM - 1) / kBM;
  for (int c = 0; c < 4; ++c) {
    const int bn = cands[c];
    const long long n_tiles = (n + bn - 1) / bn;
    const long long tiles = m_tiles * n_tiles;
    const long long waves = (tiles + 147) / 148;
    // time ~ waves * tile cost; narrow tiles run at lower tensor efficiency (smem-bound)
    const double eff = bn >= 176 ? 1.0 : (bn == 128 ? 0.9 : 0.6);
    const double cost = static_cast<double>(waves) * bn / eff;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

template <int BN>
static cudaError_t launch_bn(const vb_gemm_args& a, const EpiParams& ep, cudaStream_t stream,
                             int force_grid) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb;
  const int transposed = a.reserved2 == 1 ? 1 : 0;  // vb_gemm_args.operand_layout
  if (!transposed) {
    if (!make_tmap_bf16_2d(&ta, a.a, a.m, a.k, a.lda, kBM)) return cudaErrorInvalidValue;
    if (!make_tmap_bf16_2d(&tb, a.b, a.n, a.k, a.ldb, BN)) return cudaErrorInvalidValue;
  } else {  // a = A^T (K, M), b = B^T (K, N): boxes of 64 contiguous M / N elements x 64 K rows
    if (BN % 64 != 0) return cudaErrorInvalidValue;
    if (!make_tmap_bf16_2d(&ta, a.a, a.k, a.m, a.lda, kBK)) return cudaErrorInvalidValue;
    if (!make_tmap_bf16_2d(&tb, a.b, a.k, a.n, a.ldb, kBK)) return cudaErrorInvalidValue;
  }
  static DeviceOnce attr_set_once;
  bool& attr_set = attr_set_once();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int m_tiles = static_cast<int>((a.m + kBM - 1) / kBM);
  const int n_tiles = static_cast<int>((a.n + BN - 1) / BN);
  const int k_blocks = static_cast<int>((a.k + kBK - 1) / kBK);
  const int sms = device_sm_count();
  long long tiles = static_cast<long long>(m_tiles) * n_tiles;
  int grid = static_cast<int>(tiles < sms ? tiles : sms);
  if (force_grid > 0 && force_grid < grid) grid = force_grid;
  return launch_pdl(gemm_tcgen05_kernel<BN>, dim3(static_cast<unsigned>(grid)), dim3(kGemmThreads), Cfg::kSmemBytes,
                    stream, ta, tb, ep, k_blocks, m_tiles, n_tiles, transposed);
}

void fill_epi_params(EpiParams& ep, const vb_gemm_args& a) {
  ep.c = a.c;
  ep.bias = a.bias;
  ep.residual = reinterpret_cast<const __nv_bfloat16*>(a.residual);
  ep.m = a.m; ep.n = a.n; ep.ldc = a.ldc; ep.ldr = a.ldr;
  ep.alpha = a.alpha; ep.beta = a.beta;
  ep.alpha_cols = a.alpha_cols;
  ep.row_group = a.row_group;
  ep.epilogue = a.epilogue;
  ep.out_f32 = (a.out_dtype == VB_F32) ? 1 : 0;
  ep.tma_store = 0;
  const bool drop = a.dropout_p > 0.0f && a.dropout_seed != nullptr;
  ep.drop_seed = reinterpret_cast<const unsigned long long*>(a.dropout_seed);
  ep.drop_salt = a.dropout_salt;
  ep.drop_thresh = drop ? dropout_threshold(a.dropout_p) : 0u;
  ep.drop_scale = drop ? 1.0f / (1.0f - a.dropout_p) : 1.0f;
  ep.ln_stats = a.ln_colsum != nullptr ? a.ln_stats : nullptr;
  ep.ln_colsum = a.ln_colsum;
  ep.ln_inv_k = 1.0f / static_cast<float>(a.k);
  ep.ln_eps = a.ln_eps;
  ep.stats_out = a.stats_out;
  ep.stats_zero = a.stats_zero;
}

cudaError_t gemm_tcgen05_2cta_launch(const vb_gemm_args& a, int bn, cudaStream_t stream);

// CTA-pair 256-row tiles win on every large launch measured on B200 (sweep in
// profiles/r01_gemm_tile_sweep.txt: +6..+12 % over the best 1-CTA tile, including N = 1408
// where 8 % of the tile is padding); they need a few waves of pair-tiles to pay off.
// Returns the tile WIDTH (0 = use the 1-CTA kernel).  Many waves (the ViT, M = 34952): 256 wide, the best
// arithmetic intensity.  Two or three waves (OPT qkv / fc1 at M = 976: four row tiles): the width in 192..256
// whose tile count fills whole waves of the 74 pairs (N = 7680 -> 208: 148 tiles; N = 10240 -> 192), +5 % in
// profiles/r02_gemm_opt_widths.txt.  Narrower pair tiles lose to the 1-CTA kernel: per k-block a CTA pulls
// (128 + width/2) rows from L2 whatever the width.
static int pick_2cta_block_n(long long m, long long n, long long k) {
  if (m < 512) return 0;
  const long long pairs = 74;
  const long long m_tiles = (m + 255) / 256;
  const long long tiles256 = m_tiles * ((n + 255) / 256);
  if (tiles256 >= 4 * pairs) return 256;
  if (tiles256 < 100) {
    // Less than a wave and a half of 256-wide tiles (OPT N = 2560 at M = 976).  With a long reduction (fc2: K =
    // 10 240) the narrowest width that still fits ONE wave of pairs beats the 1-CTA kernel (width 144: 72 tiles,
    // 983 vs 880 TFLOP/s in profiles/r02_gemm_opt_widths.txt); with a short one the 1-CTA kernel's 120 CTAs win.
    if (k < 8192) return 0;
    for (int tn = 128; tn <= 256; tn += 16)
      if (m_tiles * ((n + tn - 1) / tn) <= pairs) return tn;
    return 0;
  }
  int best = 256;
  double best_cost = 1e300;
  for (int tn = 256; tn >= 192; tn -= 16) {
    const long long tiles = m_tiles * ((n + tn - 1) / tn);
    const long long waves = (tiles + pairs - 1) / pairs;
    const double cost = static_cast<double>(waves) * (tn + 24);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = tn; }
  }
  return best;
}

// vb_gemm_args.reserved: 0 = automatic; 64/128/176/256 = force the 1-CTA kernel with that
// BLOCK_N; 1000 + width (a multiple of 16, 32..256) = force the CTA-pair kernel at that tile width.
cudaError_t gemm_tcgen05_launch(const vb_gemm_args& a, cudaStream_t stream) {
  if (a.reserved2 == 1) {  // transposed operands: 1-CTA kernel, tile widths that are whole 64-element chunks
    EpiParams ep;
    fill_epi_params(ep, a);
    if (a.reserved == 64 || (a.reserved == 0 && a.n <= 64)) return launch_bn<64>(a, ep, stream, 0);
    return launch_bn<128>(a, ep, stream, 0);
  }
  if (a.reserved >= 1000) return gemm_tcgen05_2cta_launch(a, a.reserved - 1000, stream);
  if (a.reserved == 0) {
    const int bn2 = pick_2cta_block_n(a.m, a.n, a.k);
    if (bn2 > 0) return gemm_tcgen05_2cta_launch(a, bn2, stream);
  }
  EpiParams ep;
  fill_epi_params(ep, a);
  int bn = a.reserved > 0 ? a.reserved : pick_block_n(a.m, a.n);
  switch (bn) {
    case 256: return launch_bn<256>(a, ep, stream, 0);
    case 176: return launch_bn<176>(a, ep, stream, 0);
    case 128: return launch_bn<128>(a, ep, stream, 0);
    case 64: return launch_bn<64>(a, ep, stream, 0);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace vb
