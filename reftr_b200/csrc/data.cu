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

}  // namespace rb

using namespace rb;

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
