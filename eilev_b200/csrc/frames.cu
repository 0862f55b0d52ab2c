// Training-time frame transform on the device (scripts/general/train_v2.py:143-167):
//   ConvertUint8ToFloat (/255) -> Normalize(mean, std) -> RandomResizedCrop(out_h, out_w, bicubic) ->
//   RandomHorizontalFlip
// fused into one pass over the decoded uint8 clip: the crop box and the flip decision are drawn on the
// host (integer bookkeeping), the kernel reads the uint8 crop once and writes the normalised float clip.
// pytorchvideo's RandomResizedCrop crops and then calls torch.nn.functional.interpolate(mode="bicubic")
// (align_corners=False, NOT antialiased): the cubic-convolution kernel with A = -0.75 over a 4 x 4
// neighbourhood whose indices are clamped to the CROP (the reference interpolates the cropped tensor).
// The interpolation weights sum to 1, so the affine /255 + normalise commutes with it and is applied
// after the 16-tap sum.  HBM-bound: 1 B read per source sample of the crop (L2-cached taps), 4 B (f32)
// or 2 B (bf16) written per output sample.
#include "common.cuh"
#include "internal.h"

namespace vb {

VB_DEVICE float cubic1(float x, float a) { return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f; }
VB_DEVICE float cubic2(float x, float a) { return ((a * x - 5.0f * a) * x + 8.0f * a) * x - 4.0f * a; }
// ATen get_cubic_upsample_coefficients
VB_DEVICE void cubic_coeffs(float t, float (&c)[4]) {
  const float a = -0.75f;
  c[0] = cubic2(t + 1.0f, a);
  c[1] = cubic1(t, a);
  c[2] = cubic1(1.0f - t, a);
  c[3] = cubic2(2.0f - t, a);
}

struct CropParams {
  const unsigned char* in;
  void* out;
  long long c, t, h, w;
  long long top, left, ch, cw;
  long long oh, ow;
  float sy, sx;                 // crop / out scale per axis
  float rescale;
  float mean[4], inv_std[4];
  int flip, out_bf16;
};

__global__ void __launch_bounds__(256) crop_resize_normalize_kernel(const CropParams p) {
  pdl_wait();     // programmatic dependent launch: results of the preceding kernels are visible from here on
  pdl_trigger();  // the next kernel of the stream may start its prologue
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per_plane = p.oh * p.ow;
  if (idx >= p.c * p.t * per_plane) return;
  const long long plane = idx / per_plane;        // c * T + t
  const long long rem = idx % per_plane;
  const long long oy = rem / p.ow, ox = rem % p.ow;
  const long long dx = p.flip ? (p.ow - 1 - ox) : ox;  // flip AFTER the resize: output x reads resized column W-1-x
  // ATen area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=true)
  const float fy = p.sy * (static_cast<float>(oy) + 0.5f) - 0.5f;
  const float fx = p.sx * (static_cast<float>(dx) + 0.5f) - 0.5f;
  const float fy0 = floorf(fy), fx0 = floorf(fx);
  float cy[4], cx[4];
  cubic_coeffs(fy - fy0, cy);
  cubic_coeffs(fx - fx0, cx);
  const long long iy = static_cast<long long>(fy0), ix = static_cast<long long>(fx0);
  const unsigned char* src = p.in + plane * p.h * p.w;
  float acc = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long y = iy - 1 + i;
    y = y < 0 ? 0 : (y > p.ch - 1 ? p.ch - 1 : y);
    const unsigned char* row = src + (p.top + y) * p.w + p.left;
    float r = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long x = ix - 1 + j;
      x = x < 0 ? 0 : (x > p.cw - 1 ? p.cw - 1 : x);
      r = fmaf(cx[j], static_cast<float>(__ldg(row + x)), r);
    }
    acc = fmaf(cy[i], r, acc);
  }
  const int ch = static_cast<int>(plane / p.t);
  const float v = (acc * p.rescale - p.mean[ch]) * p.inv_std[ch];
  if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.out)[idx] = __float2bfloat16(v);
  else reinterpret_cast<float*>(p.out)[idx] = v;
}

cudaError_t crop_resize_normalize_launch(const void* in, long long c, long long t, long long h, long long w,
                                         long long top, long long left, long long ch, long long cw, int flip,
                                         void* out, int out_bf16, long long oh, long long ow, float rescale,
                                         const float* mean, const float* stdv, cudaStream_t s) {
  const long long total = c * t * oh * ow;
  if (total <= 0) return cudaSuccess;
  CropParams p;
  p.in = reinterpret_cast<const unsigned char*>(in);
  p.out = out;
  p.c = c; p.t = t; p.h = h; p.w = w;
  p.top = top; p.left = left; p.ch = ch; p.cw = cw;
  p.oh = oh; p.ow = ow;
  p.sy = static_cast<float>(ch) / static_cast<float>(oh);
  p.sx = static_cast<float>(cw) / static_cast<float>(ow);
  p.rescale = rescale;
  for (int i = 0; i < 4; ++i) {
    p.mean[i] = i < c ? mean[i] : 0.0f;
    p.inv_std[i] = i < c ? 1.0f / stdv[i] : 1.0f;
  }
  p.flip = flip;
  p.out_bf16 = out_bf16;
  launch_pdl(crop_resize_normalize_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, p);
  return cudaGetLastError();
}

}  // namespace vb
