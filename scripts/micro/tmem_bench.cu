// Micro-benchmark: TMEM read bandwidth (tcgen05.ld) and MUFU.EX2 throughput per SM on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tmem_bench scripts/micro/tmem_bench.cu
// Prints cycles per 128-lane x 256-column fp32 tile read (128 KB) for 4 / 8 reading warps and the
// .x32 / .x64 / .x128 shapes, and cycles per 32768 ex2 (one 128 x 256 tile) for 4 / 8 warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define DEV __device__ __forceinline__
DEV uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int N> struct Ld;
template <> struct Ld<32> {
  DEV static void go(uint32_t a, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
        "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
      : "r"(a) : "memory");
  }
};

DEV void st16(uint32_t a, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(a),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
               "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
DEV float ex2f_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
DEV uint32_t packbf(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

template <int mode>
__global__ void __launch_bounds__(384, 1) tmem_read_kernel(int warps, int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t base = slot;
  float acc = 0.f, acc2 = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp >= 4 && warp < 4 + warps) {
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    const uint32_t row = base + (static_cast<uint32_t>(quarter * 32) << 16);
    // 8 warps: each reads half of the 256 columns; 4 warps: all 256
    const int c0 = warps == 8 ? half * 128 : 0, c1 = warps == 8 ? c0 + 128 : 256;
    asm volatile("bar.sync 1, %0;\n" ::"r"(warps * 32) : "memory");
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {          // one x32 load per wait
        for (int c = c0; c < c1; c += 32) {
          uint32_t r[32];
          Ld<32>::go(row + c, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
        }
      } else if (mode == 1) {   // two x32 loads in flight per wait
        for (int c = c0; c < c1; c += 64) {
          uint32_t r[32], q[32];
          Ld<32>::go(row + c, r);
          Ld<32>::go(row + c + 32, q);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          acc += __uint_as_float(r[0]) + __uint_as_float(r[31]) + __uint_as_float(q[3]) + __uint_as_float(q[30]);
        }
      } else if (mode == 3) {   // the softmax exp pass without the math: load x32, wait, store x16 (P in place)
        for (int c = c0; c < c1; c += 32) {
          uint32_t r[32], pk[16];
          Ld<32>::go(row + c, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          for (int j = 0; j < 16; ++j) pk[j] = r[2 * j] ^ r[2 * j + 1];
          st16(row + c / 2, pk);
        }
      } else if (mode == 4) {   // the softmax exp pass: load x32, wait, 32 x (fma, ex2, add) + 16 packs, store x16
        for (int c = c0; c < c1; c += 32) {
          uint32_t r[32], pk[16];
          Ld<32>::go(row + c, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          for (int j = 0; j < 32; j += 2) {
            const float p0 = ex2f_(fmaf(__uint_as_float(r[j]), 1.0001f, -0.5f)), p1 = ex2f_(fmaf(__uint_as_float(r[j + 1]), 1.0001f, -0.5f));
            acc += p0; acc2 += p1;
            pk[j / 2] = packbf(p0, p1);
          }
          st16(row + c / 2, pk);
        }
      } else {                  // four x32 loads in flight per wait
        for (int c = c0; c < c1; c += 128) {
          uint32_t r[32], q[32], s[32], u[32];
          Ld<32>::go(row + c, r);
          Ld<32>::go(row + c + 32, q);
          Ld<32>::go(row + c + 64, s);
          Ld<32>::go(row + c + 96, u);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          acc += __uint_as_float(r[0]) + __uint_as_float(r[31]) + __uint_as_float(q[3]) + __uint_as_float(s[5]) + __uint_as_float(u[7]) + __uint_as_float(u[31]);
        }
      }
    }
    asm volatile("bar.sync 1, %0;\n" ::"r"(warps * 32) : "memory");
    t1 = clock64();
    if (threadIdx.x == 128 && blockIdx.x == 0) out[0] = (t1 - t0);
    sink[blockIdx.x * 384 + threadIdx.x] = acc + acc2;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(base) : "memory");
  }
}

__global__ void __launch_bounds__(384, 1) mufu_kernel(int warps, int iters, long long* out, float* sink) {
  const int warp = threadIdx.x >> 5;
  if (warp >= warps) return;
  float x[16];
  for (int j = 0; j < 16; ++j) x[j] = -0.001f * (threadIdx.x + j);
  asm volatile("bar.sync 1, %0;\n" ::"r"(warps * 32) : "memory");
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
  }
  asm volatile("bar.sync 1, %0;\n" ::"r"(warps * 32) : "memory");
  long long t1 = clock64();
  float acc = 0.f;
  for (int j = 0; j < 16; ++j) acc += x[j];
  sink[blockIdx.x * 384 + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 148 * 384 * 4);
  const int iters = 200;
  for (int warps : {4, 8}) for (int mode : {3, 4}) {
    long long h = 0;
    if (mode == 3) tmem_read_kernel<3><<<148, 384>>>(warps, iters, out, sink);
    else tmem_read_kernel<4><<<148, 384>>>(warps, iters, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%s warps=%d : %.1f clk per 128x256 tile  [%s]\n", mode == 3 ? "ld x32 + wait + st x16         " : "ld x32 + wait + exp math + st x16", warps,
           double(h) / iters, cudaGetErrorString(e));
  }
  for (int warps : {4, 8}) for (int mode : {0, 1, 2}) {
    long long h = 0;
    if (mode == 0) tmem_read_kernel<0><<<148, 384>>>(warps, iters, out, sink);
    else if (mode == 1) tmem_read_kernel<1><<<148, 384>>>(warps, iters, out, sink);
    else tmem_read_kernel<2><<<148, 384>>>(warps, iters, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("tmem_read warps=%d loads_in_flight=%d : %.1f clk per 128x256 fp32 tile (128 KB) -> %.1f B/clk/SM  [%s]\n", warps,
           mode == 0 ? 1 : (mode == 1 ? 2 : 4), double(h) / iters, 131072.0 * iters / double(h), cudaGetErrorString(e));
  }
  for (int warps : {4, 8, 12}) {
    long long h = 0;
    mufu_kernel<<<148, 384>>>(warps, 2000, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    const double per_sm = double(warps) * 32 * 16 * 2000;
    printf("mufu ex2 warps=%d : %.2f ex2/clk/SM  [%s]\n", warps, per_sm / double(h), cudaGetErrorString(e));
  }
  return 0;
}
