// Optimizer step on device-flat buffers: the reference's engine_vg.py:62-67 runs
// torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1) and torch.optim.AdamW(param_dicts).step() (main_vg.py:234-268) over
// ~700 tensors; here the parameters, gradients and both moments of all of them live in four flat fp32 buffers, so the global
// gradient norm is ONE reduction and clip + decoupled weight decay + Adam update ONE pass (7 floats of HBM traffic per element).
#include "common.cuh"
#include "host.h"

namespace rb {

__global__ void __launch_bounds__(256) sumsq_kernel(const float4* __restrict__ x, long long n4, float* __restrict__ out) {
  float s = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(x + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

// Hand-over of the flat gradient buffer (engine.run_backward): dst = src * scale (undoes the static loss scale of the 16-bit
// backward, and gives autograd fresh storage) and, in the same pass, a non-finite detector: *flag is set to 1 when any element of
// src is inf / NaN.  No extra memory traffic; the decision is taken by zero_if_kernel below.
__global__ void __launch_bounds__(256) scale_copy_check_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long long n4, float scale,
                                                               int* __restrict__ flag) {
  bool bad = false;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = __ldg(src + i);
    // x - x is 0 for finite x and NaN for inf / NaN
    bad |= !((v.x - v.x) == 0.f) | !((v.y - v.y) == 0.f) | !((v.z - v.z) == 0.f) | !((v.w - v.w) == 0.f);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    dst[i] = v;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

// If *flag != 0: x[0..n) = 0 (a skipped step) and, once, *counter += 1.  If *flag == 0 (every normal step) each block reads one word
// and exits.
__global__ void __launch_bounds__(256) zero_if_kernel(float4* __restrict__ x, long long n4, const int* __restrict__ flag, float* __restrict__ counter) {
  if (*flag == 0) return;
  if (counter && blockIdx.x == 0 && threadIdx.x == 0) *counter += 1.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x)
    x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

struct AdamSegs {  // segment s covers elements [s == 0 ? 0 : end[s-1], end[s]) and belongs to parameter group `group[s]`
  long long end[RB_ADAMW_MAX_SEGMENTS];
  int group[RB_ADAMW_MAX_SEGMENTS];
  float lr[RB_ADAMW_MAX_GROUPS], wd[RB_ADAMW_MAX_GROUPS];
  int nseg;
};

// torch.optim.AdamW (single-tensor formula, amsgrad off, maximize off):
//   p *= 1 - lr*wd;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// with g pre-multiplied by the clip coefficient min(1, max_norm / (||g|| + 1e-6)) of torch.nn.utils.clip_grad_norm_.
__global__ void __launch_bounds__(256) adamw_flat_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                                                         long long n4, const AdamSegs segs, float b1, float b2, float eps, float inv_bc1, float inv_sqrt_bc2,
                                                         const float* __restrict__ sumsq, float max_norm) {
  float clip = 1.f;
  if (sumsq && max_norm > 0.f) clip = fminf(1.f, max_norm / (sqrtf(*sumsq) + 1e-6f));
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = i * 4;
    int s = 0;
    while (s < segs.nseg - 1 && e >= segs.end[s]) ++s;  // a handful of segments; slots are 64-element aligned so a float4 never straddles
    const int grp = segs.group[s];
    const float lr = segs.lr[grp], decay = 1.f - lr * segs.wd[grp], step = lr * inv_bc1;
    float4 pv = p[i], mv = m[i], vv = v[i];
    const float4 gv = __ldg(g + i);
    float* pp = &pv.x; float* mm = &mv.x; float* vq = &vv.x; const float* gg = &gv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gg[k] * clip;
      mm[k] = b1 * mm[k] + (1.f - b1) * gk;
      vq[k] = b2 * vq[k] + (1.f - b2) * gk * gk;
      pp[k] = pp[k] * decay - step * mm[k] / (sqrtf(vq[k]) * inv_sqrt_bc2 + eps);
    }
    p[i] = pv; m[i] = mv; v[i] = vv;
  }
}

}  // namespace rb

using namespace rb;

extern "C" int rb_sumsq(const float* x, long long n, float* out, void* stream) {
  if (n % 4 || (reinterpret_cast<uintptr_t>(x) & 15)) return rb_fail("rb_sumsq: n %% 4 != 0 or unaligned buffer");
  if (n <= 0) return 0;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sumsq_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(x), n / 4, out);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_scale_copy_check(const float* src, float* dst, long long n, float scale, int* flag, void* stream) {
  if (n % 4 || ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15)) return rb_fail("rb_scale_copy_check: n %% 4 != 0 or unaligned buffers");
  if (!flag) return rb_fail("rb_scale_copy_check: flag is NULL");
  if (n <= 0) return 0;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  scale_copy_check_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(src),
                                                                                                    reinterpret_cast<float4*>(dst), n / 4, scale, flag);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_zero_if(float* x, long long n, const int* flag, float* counter, void* stream) {
  if (n % 4 || (reinterpret_cast<uintptr_t>(x) & 15)) return rb_fail("rb_zero_if: n %% 4 != 0 or unaligned buffer");
  if (!flag) return rb_fail("rb_zero_if: flag is NULL");
  if (n <= 0) return 0;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  zero_if_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<float4*>(x), n / 4, flag, counter);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_adamw_flat(float* p, const float* g, float* m, float* v, long long n, const rb_adamw_segments* segs, float beta1, float beta2, float eps,
                             int step, const float* sumsq, float max_norm, void* stream) {
  if (!segs || segs->nseg < 1 || segs->nseg > RB_ADAMW_MAX_SEGMENTS) return rb_fail("rb_adamw_flat: 1..%d segments", RB_ADAMW_MAX_SEGMENTS);
  if (n % 4 || ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15))
    return rb_fail("rb_adamw_flat: n %% 4 != 0 or unaligned buffers");
  if (step < 1) return rb_fail("rb_adamw_flat: step counts from 1");
  if (n <= 0) return 0;
  AdamSegs s;
  s.nseg = segs->nseg;
  for (int i = 0; i < RB_ADAMW_MAX_SEGMENTS; ++i) {
    s.end[i] = i < segs->nseg ? segs->end[i] : n;
    s.group[i] = i < segs->nseg ? segs->group[i] : 0;
    if (s.group[i] < 0 || s.group[i] >= RB_ADAMW_MAX_GROUPS) return rb_fail("rb_adamw_flat: group index out of range");
    if (i < segs->nseg && (s.end[i] % 4)) return rb_fail("rb_adamw_flat: segment ends must be multiples of 4");
  }
  for (int i = 0; i < RB_ADAMW_MAX_GROUPS; ++i) { s.lr[i] = segs->lr[i]; s.wd[i] = segs->weight_decay[i]; }
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step), bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adamw_flat_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n / 4, s, beta1, beta2, eps,
      static_cast<float>(1.0 / bc1), static_cast<float>(1.0 / sqrt(bc2)), sumsq, max_norm);
  RB_CUDA(cudaGetLastError());
  return 0;
}
