// Shared device helpers for the sm_100a kernels of the VideoBLIP hot path:
// mbarrier / TMA / tcgen05 / TMEM inline-PTX wrappers, bf16 packing, epilogue math.
// Everything here is header-only and compiled with
//   nvcc -gencode arch=compute_100a,code=sm_100a
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define VB_DEVICE __device__ __forceinline__

namespace vb {

// ---------------------------------------------------------------- misc
VB_DEVICE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
VB_DEVICE uint32_t lane_id() { return threadIdx.x & 31u; }

VB_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
VB_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
VB_DEVICE void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
VB_DEVICE void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
VB_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
VB_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
VB_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe (no hardware suspend): for threads that poll several barriers at once.
VB_DEVICE bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a mis-programmed pipeline traps (surfacing as a CUDA error on the
// host) instead of hanging the GPU box.
VB_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      __trap();
    }
  }
}

// ---------------------------------------------------------------- TMA
VB_DEVICE void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: c0 = innermost (contiguous) coordinate, c1 = row coordinate.
VB_DEVICE void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 2-D tiled store smem -> global (bulk async group); rows/cols outside the tensor are clipped.
VB_DEVICE void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
VB_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// wait until at most N committed bulk groups of this thread still READ their smem source
template <int N>
VB_DEVICE void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
VB_DEVICE void named_bar_sync(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory");
}
VB_DEVICE void cp_async_16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
VB_DEVICE void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;\n" ::: "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
VB_DEVICE void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
VB_DEVICE void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
VB_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols)
               : "memory");
}
VB_DEVICE void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
VB_DEVICE void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; bf16 inputs, fp32 accumulate, one issuing thread.
VB_DEVICE void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on `bar` once all previously issued tcgen05.mma of this thread completed
// (implies tcgen05.fence::before_thread_sync).
VB_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
          smem_u32(bar))
      : "memory");
}
VB_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base+i).
VB_DEVICE void tmem_ld_16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 operand tile laid out by TMA with
// the 128-byte swizzle: rows of 64 bf16 (128 B), 8-row groups 1024 B apart.
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (unused: 1)
//   bits [32,46) stride byte offset >> 4   bits [46,48) descriptor version (1 on sm_100)
//   bits [61,64) layout type (2 = SWIZZLE_128B)
VB_DEVICE uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, dense.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---------------------------------------------------------------- CTA-pair (cta_group::2)
VB_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
VB_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// Arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster.
VB_DEVICE void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// 2-SM TMA load: data lands in THIS CTA's shared memory, the transaction bytes are
// credited to the mbarrier of the pair's leader CTA (peer bit cleared).
VB_DEVICE void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  const uint32_t bar_leader = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1)
      : "memory");
}
VB_DEVICE void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
VB_DEVICE void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
VB_DEVICE void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem of both CTAs] (+)= A * B over the CTA pair: M = 256 (128 rows per CTA), B split in
// halves along N between the two CTAs' shared memories.  Issued by the leader CTA only.
VB_DEVICE void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Commit of the pair's MMAs, arriving on the barrier at this offset in BOTH CTAs.
VB_DEVICE void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;\n" ::"r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// ---------------------------------------------------------------- math / packing
VB_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
VB_DEVICE float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
VB_DEVICE float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// GEMM-epilogue GELU (round 2): Phi(x) = sigmoid(2 g(x)) with g the odd quintic minimax fit of
// atanh(erf(x / sqrt2)) on [-8, 8] (x^2 clamped at 64: beyond it the sigmoid has saturated to
// 0 / 1 in fp32).  |gelu - exact erf GELU| <= 2.6e-5 everywhere (fit in tests/test_host_cpu.py;
// bf16 output rounding is >= 4e-5 for |y| >= 0.01), and the negative tail keeps its relative
// accuracy because nothing is subtracted.  9 instructions, 2 of them MUFU (ex2, rcp) — half of
// the Abramowitz-Stegun form above, which kept the fc1 epilogue longer than its mainloop
// (profiles/r01_ncu_gemm2cta_fc1.txt: tensor pipe 60 % active).  Constants carry -2 log2(e).
VB_DEVICE float gelu_fast(float x) {
  const float x2 = fminf(x * x, 64.0f);
  float p = fmaf(x2, 0.0010142630553f, -0.1067757240029f);
  p = fmaf(x2, p, -2.3011213394573f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * p));  // exp(-2 g(x)); +inf for very negative x
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}
VB_DEVICE float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Counter-based dropout: keep(seed, idx) is a pure function, so the backward pass regenerates
// the forward mask instead of storing it.  `thresh` = p * 2^32; an element is kept when the
// 32-bit hash of (seed, idx) is >= thresh.
VB_DEVICE bool dropout_keep(uint64_t seed, uint64_t idx, uint32_t thresh) {
  uint32_t h = static_cast<uint32_t>(idx) * 0x9E3779B1u ^ static_cast<uint32_t>(idx >> 32) * 0x85EBCA77u ^
               static_cast<uint32_t>(seed);
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  h += static_cast<uint32_t>(seed >> 32);
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
  return h >= thresh;
}
__host__ __device__ inline uint32_t dropout_threshold(float p) {
  if (!(p > 0.0f)) return 0u;
  if (p >= 1.0f) return 0xFFFFFFFFu;
  return static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0);
}

VB_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
VB_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- tensor-map cache
// cuTensorMapEncodeTiled costs about a microsecond of host time per map and an eager launch needs two to four of
// them (CUDA graphs hide it, short generate() calls and beam search run eagerly).  A map is a pure function of
// (address, extents, strides, box), so a small direct-mapped table of the last encodings is always valid.
struct TmapKey {
  unsigned long long w[8];
  bool operator==(const TmapKey& o) const {
    for (int i = 0; i < 8; ++i)
      if (w[i] != o.w[i]) return false;
    return true;
  }
};
bool tmap_cache_lookup(const TmapKey& key, CUtensorMap* out);   // gemm_tcgen05.cu
void tmap_cache_store(const TmapKey& key, const CUtensorMap& map);

// ---------------------------------------------------------------- per-device launch-time state
// One process may drive several devices: function attributes (dynamic shared memory opt-in) are per device, and
// so is the SM count a persistent grid is sized by.  (Round 1 kept these in per-process statics.)
constexpr int kMaxDevices = 64;
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}
struct DeviceOnce {
  bool done[kMaxDevices] = {};
  bool& operator()() { return done[current_device()]; }
};
inline int device_sm_count() {
  static int sms[kMaxDevices] = {};
  const int d = current_device();
  if (sms[d] == 0) {
    cudaDeviceGetAttribute(&sms[d], cudaDevAttrMultiProcessorCount, d);
    if (sms[d] <= 0) sms[d] = 148;
  }
  return sms[d];
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-
// serialization attribute may become resident while its predecessor is still running.
// Everything before pdl_wait() must touch only data no earlier kernel of the stream writes
// (weights); pdl_wait() returns once the predecessor grid has completed and flushed.
VB_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
VB_DEVICE void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
}  // namespace vb

#include <cstdlib>
namespace vb {
// VB_PDL=0 turns programmatic dependent launch off (A/B measurements).
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("VB_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}
// Launch `kern` allowing it to overlap the tail of the previous kernel of the stream.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(static_cast<Args&&>(args))...);
}

}  // namespace vb

// Epilogue selector shared by the tcgen05 and the generic GEMM (values are part of the
// C ABI, see include/videoblip_b200.h).
enum VbEpilogue : int {
  VB_EPI_NONE = 0,       // C = alpha*(A.B^T) (+bias)
  VB_EPI_GELU = 1,       // exact (erf) GELU
  VB_EPI_RELU = 2,
  VB_EPI_GELU_BWD = 3,   // C = acc * gelu'(residual): activation backward fused into the dgrad GEMM
  VB_EPI_RELU_BWD = 4,   // C = residual > 0 ? acc : 0
};
