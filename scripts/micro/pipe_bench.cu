// Micro-benchmark: issue rates of the instructions in the attention softmax on sm_100a, per SM:
// MUFU.EX2, F2FP.BF16.F32.PACK_AB (cvt.rn.bf16x2.f32), PRMT, FMNMX3, FFMA and the mixes the kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/pipe_bench scripts/micro/pipe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(384, 1) pipe_kernel(int warps, int iters, long long* out, float* sink) {
  const int warp = threadIdx.x >> 5;
  if (warp >= warps) return;
  float x[16];
  uint32_t u[8];
  for (int j = 0; j < 16; ++j) x[j] = -0.001f * (threadIdx.x + j + 1);
  for (int j = 0; j < 8; ++j) u[j] = threadIdx.x * 2654435761u + j;
  const float c = 1.0001f, m = -0.0003f;
  asm volatile("bar.sync 1, %0;\n" ::"r"(warps * 32) : "memory");
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {        // 16 ex2
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
    } else if (MODE == 1) { // 16 cvt.bf16x2 (each packs two floats)
#pragma unroll
      for (int j = 0; j < 16; ++j)
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[j & 7]) : "f"(x[j]), "f"(x[(j + 1) & 15]));
    } else if (MODE == 2) { // 16 prmt
#pragma unroll
      for (int j = 0; j < 16; ++j)
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[j & 7]) : "r"(__float_as_uint(x[j])), "r"(u[(j + 1) & 7]));
    } else if (MODE == 3) { // 16 3-input max
#pragma unroll
      for (int j = 0; j < 16; ++j)
        asm volatile("max.ftz.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(x[(j + 1) & 15]), "f"(x[(j + 5) & 15]));
    } else if (MODE == 4) { // 16 fma
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(c), "f"(m));
    } else if (MODE == 5) { // softmax element mix: 16 x (fma, ex2, add) + 8 cvt
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(c), "f"(m));
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("add.ftz.f32 %0, %0, %1;" : "+f"(s) : "f"(x[j]));
#pragma unroll
      for (int j = 0; j < 16; j += 2)
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[j >> 1]) : "f"(x[j]), "f"(x[j + 1]));
      x[0] += s * 1e-30f;
    } else if (MODE == 6) { // the same mix with prmt instead of cvt
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(c), "f"(m));
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("add.ftz.f32 %0, %0, %1;" : "+f"(s) : "f"(x[j]));
#pragma unroll
      for (int j = 0; j < 16; j += 2)
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[j >> 1]) : "r"(__float_as_uint(x[j])), "r"(__float_as_uint(x[j + 1])));
      x[0] += s * 1e-30f;
    } else if (MODE == 7) { // 16 ex2 + 8 cvt only (do they share a pipe?)
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
#pragma unroll
      for (int j = 0; j < 16; j += 2)
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[j >> 1]) : "f"(x[j]), "f"(x[j + 1]));
    }
  }
  asm volatile("bar.sync 1, %0;\n" ::"r"(warps * 32) : "memory");
  long long t1 = clock64();
  float acc = 0.f;
  for (int j = 0; j < 16; ++j) acc += x[j];
  for (int j = 0; j < 8; ++j) acc += __uint_as_float(u[j] & 0x3fffffffu);
  sink[blockIdx.x * 384 + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, long long* out, float* sink) {
  for (int warps : {4, 8}) {
    long long h = 0;
    pipe_kernel<MODE><<<148, 384>>>(warps, 2000, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-44s warps=%d : %7.1f clk per iteration per warp-set  (%.2f clk per warp per iteration)  [%s]\n", name, warps,
           double(h) / 2000, double(h) / 2000 / (warps / 4), cudaGetErrorString(e));
  }
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 148 * 384 * 4);
  run<0>("16 ex2", out, sink);
  run<1>("16 cvt.rn.bf16x2.f32", out, sink);
  run<2>("16 prmt", out, sink);
  run<3>("16 max3", out, sink);
  run<4>("16 fma", out, sink);
  run<5>("16 (fma, ex2, add) + 8 cvt", out, sink);
  run<6>("16 (fma, ex2, add) + 8 prmt", out, sink);
  run<7>("16 ex2 + 8 cvt", out, sink);
  return 0;
}
