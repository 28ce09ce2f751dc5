// Input pipeline on the GPU: F.to_tensor + Normalize (datasets/transforms.py:233-250) and the pad-to-batch
// NestedTensor build of util/collate_fn.py:24-41 in ONE launch, from the raw uint8 HWC images -- the host ships a quarter of the
// bytes (uint8 instead of normalised fp32) and no padded batch is ever built on the CPU.
#include "common.cuh"
#include "host.h"

namespace rb {

struct Norm3 { float mean[3], std[3]; };

// table[b] = {byte offset of image b in `packed`, h_b, w_b};  out fp32 [B,3,H,W], mask u8/bool [B,H,W] (1 = padding)
__global__ void __launch_bounds__(256) collate_u8_kernel(const uint8_t* __restrict__ packed, const long long* __restrict__ table, int H, int W, Norm3 n,
                                                         float* __restrict__ out, uint8_t* __restrict__ mask) {
  const int b = blockIdx.y;
  const long long off = table[3 * b];
  const int h = static_cast<int>(table[3 * b + 1]), w = static_cast<int>(table[3 * b + 2]);
  const long long HW = static_cast<long long>(H) * W;
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < HW; p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int y = static_cast<int>(p / W), x = static_cast<int>(p - static_cast<long long>(y) * W);
    const bool inside = y < h && x < w;
    float v[3] = {0.f, 0.f, 0.f};
    if (inside) {
      const uint8_t* src = packed + off + (static_cast<long long>(y) * w + x) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = (static_cast<float>(src[c]) / 255.f - n.mean[c]) / n.std[c];  // to_tensor().sub(mean).div(std), same rounding
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(static_cast<long long>(b) * 3 + c) * HW + p] = v[c];
    mask[static_cast<long long>(b) * HW + p] = inside ? 0 : 1;
  }
}

// ------------------------------------------------------------------------------------------------ resize (Pillow bilinear, bit-exact)
// torchvision's F.resize on a PIL image (datasets/transforms.py:111) is Pillow's ImagingResample: two separable passes over uint8
// with 22-bit fixed-point coefficients (computed in double on the host, data.py:pil_bilinear_coeffs, exactly as Pillow's
// precompute_coeffs / normalize_coeffs_8bpc do) and a uint8 intermediate: out = clip8((2^21 + sum_x pix[x] * k[x]) >> 22).
constexpr int RS_PREC = 22;
__device__ __forceinline__ uint8_t rs_clip8(int v) {
  v >>= RS_PREC;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}
// src [h, w, 3] -> dst [h, ow, 3]; bounds [ow, 2] = (xmin, count), kk [ow, ksize]
__global__ void __launch_bounds__(128) resize_h_u8_kernel(const uint8_t* __restrict__ src, int h, int w, uint8_t* __restrict__ dst, int ow,
                                                          const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const int ox = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y;
  if (ox >= ow) return;
  const int x0 = bounds[2 * ox], n = bounds[2 * ox + 1];
  const int* k = kk + static_cast<long long>(ox) * ksize;
  const uint8_t* row = src + (static_cast<long long>(y) * w + x0) * 3;
  int s0 = 1 << (RS_PREC - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < n; ++x) {
    const int c = __ldg(k + x);
    s0 += row[3 * x] * c; s1 += row[3 * x + 1] * c; s2 += row[3 * x + 2] * c;
  }
  uint8_t* o = dst + (static_cast<long long>(y) * ow + ox) * 3;
  o[0] = rs_clip8(s0); o[1] = rs_clip8(s1); o[2] = rs_clip8(s2);
}
// src [h, w, 3] -> dst [oh, w, 3]; bounds [oh, 2] = (ymin, count), kk [oh, ksize]
__global__ void __launch_bounds__(128) resize_v_u8_kernel(const uint8_t* __restrict__ src, int h, int w, uint8_t* __restrict__ dst, int oh,
                                                          const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const int i = blockIdx.x * 128 + threadIdx.x, oy = blockIdx.y;   // i over the w * 3 bytes of a row
  if (i >= w * 3) return;
  const int y0 = bounds[2 * oy], n = bounds[2 * oy + 1];
  const int* k = kk + static_cast<long long>(oy) * ksize;
  int s = 1 << (RS_PREC - 1);
  for (int y = 0; y < n; ++y) s += src[(static_cast<long long>(y0 + y) * w) * 3 + i] * __ldg(k + y);
  dst[(static_cast<long long>(oy) * w) * 3 + i] = rs_clip8(s);
}

}  // namespace rb

using namespace rb;

extern "C" int rb_resize_u8(const void* src, int h, int w, void* dst, int oh, int ow, const int* bounds_h, const int* kk_h, int ksize_h,
                            const int* bounds_v, const int* kk_v, int ksize_v, void* tmp, void* stream) {
  if (!src || !dst || h <= 0 || w <= 0 || oh <= 0 || ow <= 0) return rb_fail("rb_resize_u8: bad arguments");
  if (h > 65535 || oh > 65535) return rb_fail("rb_resize_u8: image too tall for a grid dimension");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool need_h = ow != w, need_v = oh != h;   // Pillow skips a pass whose size does not change
  if (need_h && (!bounds_h || !kk_h || ksize_h <= 0)) return rb_fail("rb_resize_u8: horizontal coefficients missing");
  if (need_v && (!bounds_v || !kk_v || ksize_v <= 0)) return rb_fail("rb_resize_u8: vertical coefficients missing");
  if (need_h && need_v && !tmp) return rb_fail("rb_resize_u8: tmp [h, ow, 3] missing");
  if (!need_h && !need_v) {
    RB_CUDA(cudaMemcpyAsync(dst, src, static_cast<size_t>(h) * w * 3, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  const uint8_t* cur = static_cast<const uint8_t*>(src);
  if (need_h) {
    uint8_t* o = need_v ? static_cast<uint8_t*>(tmp) : static_cast<uint8_t*>(dst);
    resize_h_u8_kernel<<<dim3((ow + 127) / 128, h), 128, 0, st>>>(cur, h, w, o, ow, bounds_h, kk_h, ksize_h);
    RB_CUDA(cudaGetLastError());
    cur = o;
  }
  if (need_v) {
    resize_v_u8_kernel<<<dim3((ow * 3 + 127) / 128, oh), 128, 0, st>>>(cur, h, ow, static_cast<uint8_t*>(dst), oh, bounds_v, kk_v, ksize_v);
    RB_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int rb_collate_u8(const void* packed, const long long* table, int B, int H, int W, float mean0, float mean1, float mean2, float std0, float std1,
                             float std2, float* out, void* mask, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0) return rb_fail("rb_collate_u8: empty batch");
  if (std0 == 0.f || std1 == 0.f || std2 == 0.f) return rb_fail("rb_collate_u8: zero std");
  Norm3 n = {{mean0, mean1, mean2}, {std0, std1, std2}};
  long long blocks = (static_cast<long long>(H) * W + 255) / 256;
  if (blocks > 592) blocks = 592;
  collate_u8_kernel<<<dim3(static_cast<unsigned>(blocks), B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(packed), table, H, W, n, out, static_cast<uint8_t*>(mask));
  RB_CUDA(cudaGetLastError());
  return 0;
}
