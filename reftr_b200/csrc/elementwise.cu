// Memory-bound support kernels of the backbone path: stem im2col, max-pool, parity split/merge for stride-2 blocks,
// weight packing (FrozenBN fold, OIHW -> [O,(r,s),I] bf16, transposes), gradient unpacking, casts and reductions.
// All are HBM-bound: 128-bit accesses, one 16-byte chunk (8 bf16) per thread, grids sized from the element count.
#include "common.cuh"
#include "host.h"

namespace rb {

static inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------ stem im2col
// img fp32 NCHW [B,3,H,W] -> out bf16 [B*H1*W1, 160]; column k = (r*7+s)*3+c for the 7x7/stride-2/pad-3 stem
// (torchvision ResNet conv1, reached from backbone.py:99-102); columns 147..159 are zero.
__global__ void stem_im2col_kernel(const float* __restrict__ img, uint4* __restrict__ out, int B, int H, int W, int H1, int W1) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * H1 * W1 * 20;
  if (idx >= total) return;
  const int chunk = static_cast<int>(idx % 20);
  const long long pix = idx / 20;
  const int wo = static_cast<int>(pix % W1);
  const int ho = static_cast<int>((pix / W1) % H1);
  const int b = static_cast<int>(pix / (static_cast<long long>(W1) * H1));
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = chunk * 8 + e;
    float x = 0.f;
    if (k < 147) {
      const int c = k % 3, rs = k / 3, s = rs % 7, r = rs / 7;
      const int y = 2 * ho - 3 + r, xx = 2 * wo - 3 + s;
      if (y >= 0 && y < H && xx >= 0 && xx < W) x = __ldg(img + ((static_cast<long long>(b) * 3 + c) * H + y) * W + xx);
    }
    v[e] = x;
  }
  uint4 o;
  o.x = pack_t2(v[0], v[1]); o.y = pack_t2(v[2], v[3]); o.z = pack_t2(v[4], v[5]); o.w = pack_t2(v[6], v[7]);
  out[idx] = o;
}

// ------------------------------------------------------------------------------------------------ max-pool 3x3/2 pad 1
// in: bf16 NHWC [B,H1,W1,C] (un-padded, post-ReLU so >= 0) -> out: padded NHWC [B,H2+2,W2+2,C] with zero border.
__global__ void maxpool_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H1, int W1, int C8, int H2, int W2) {
  const int Hp = H2 + 2, Wp = W2 + 2;
  const unsigned i = blockIdx.y * 256u + threadIdx.x;  // grid: x = padded output row (b, u), y = chunks of the row's Wp*C8 vectors
  if (i >= static_cast<unsigned>(Wp * C8)) return;
  const int u = static_cast<int>(blockIdx.x % static_cast<unsigned>(Hp)), b = static_cast<int>(blockIdx.x / static_cast<unsigned>(Hp));
  const int v = static_cast<int>(i / static_cast<unsigned>(C8)), c = static_cast<int>(i - static_cast<unsigned>(v) * C8);
  const long long idx = static_cast<long long>(blockIdx.x) * (Wp * C8) + i;
  uint4 o = make_uint4(0, 0, 0, 0);
  if (u >= 1 && u <= H2 && v >= 1 && v <= W2) {
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -3.0e38f;
    const int y0 = 2 * (u - 1) - 1, x0 = 2 * (v - 1) - 1;
    for (int dy = 0; dy < 3; ++dy) {
      const int y = y0 + dy;
      if (y < 0 || y >= H1) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int x = x0 + dx;
        if (x < 0 || x >= W1) continue;
        const uint4 t = __ldg(in + ((static_cast<long long>(b) * H1 + y) * W1 + x) * C8 + c);
        m[0] = fmaxf(m[0], t_lo(t.x)); m[1] = fmaxf(m[1], t_hi(t.x)); m[2] = fmaxf(m[2], t_lo(t.y)); m[3] = fmaxf(m[3], t_hi(t.y));
        m[4] = fmaxf(m[4], t_lo(t.z)); m[5] = fmaxf(m[5], t_hi(t.z)); m[6] = fmaxf(m[6], t_lo(t.w)); m[7] = fmaxf(m[7], t_hi(t.w));
      }
    }
    o.x = pack_t2(m[0], m[1]); o.y = pack_t2(m[2], m[3]); o.z = pack_t2(m[4], m[5]); o.w = pack_t2(m[6], m[7]);
  }
  out[idx] = o;
}

// ------------------------------------------------------------------------------------------------ parity split / merge
// x padded [B,H+2,W+2,C] -> xs [4,B,Hs,Ws,C] (Hs=Ho+2, Ws=Wo+2); plane (p,q) cell (u,v) = x[2u+p, 2v+q] or 0.
// only_plane < 0: all four planes (xs = [4, ...]); otherwise only that plane is produced (xs = that plane's [B, Hs, Ws, C] block)
// grid: x = row (plane, b, u), y = 256-element chunks of the row's Ws*C8 vectors -- no per-thread 64-bit divisions (those made
// these copy kernels ALU-bound: ~200 instructions per 16 bytes)
__global__ void __launch_bounds__(256) parity_split_kernel(const uint4* __restrict__ x, uint4* __restrict__ xs, int B, int H, int W, int C8, int Hs, int Ws, int only_plane) {
  const unsigned i = blockIdx.y * 256u + threadIdx.x;
  if (i >= static_cast<unsigned>(Ws * C8)) return;
  unsigned row = blockIdx.x;                      // (plane', b, u), plane' = 0 when a single plane is produced
  const int u = static_cast<int>(row % static_cast<unsigned>(Hs)); row /= static_cast<unsigned>(Hs);
  const int b = static_cast<int>(row % static_cast<unsigned>(B));
  const int plane = only_plane < 0 ? static_cast<int>(row / static_cast<unsigned>(B)) : only_plane;
  const int v = static_cast<int>(i / static_cast<unsigned>(C8)), c = static_cast<int>(i - static_cast<unsigned>(v) * C8);
  const int y = 2 * u + (plane >> 1), xx = 2 * v + (plane & 1);
  uint4 o = make_uint4(0, 0, 0, 0);
  if (y <= H + 1 && xx <= W + 1) o = __ldg(x + ((static_cast<long long>(b) * (H + 2) + y) * (W + 2) + xx) * C8 + c);
  xs[static_cast<long long>(blockIdx.x) * (Ws * C8) + i] = o;
}

// dx[b,y,x,:] = relu_mask(dxs[plane(y&1,x&1), b, y>>1, x>>1, :] (+ add[b,y,x,:])) on interior pixels, 0 on the border.
// only_plane >= 0: dxs is that single plane's block and the three other planes are zero
__global__ void __launch_bounds__(256) parity_merge_kernel(const uint4* __restrict__ dxs, const uint4* __restrict__ add, const uint4* __restrict__ mask_src,
                                    uint4* __restrict__ dx, int B, int H, int W, int C8, int Hs, int Ws, int only_plane) {
  const int Hp = H + 2, Wp = W + 2;
  const unsigned i = blockIdx.y * 256u + threadIdx.x;  // grid: x = padded row (b, y), y = chunks of the row's Wp*C8 vectors
  if (i >= static_cast<unsigned>(Wp * C8)) return;
  const int y = static_cast<int>(blockIdx.x % static_cast<unsigned>(Hp)), b = static_cast<int>(blockIdx.x / static_cast<unsigned>(Hp));
  const int xx = static_cast<int>(i / static_cast<unsigned>(C8)), c = static_cast<int>(i - static_cast<unsigned>(xx) * C8);
  const long long idx = static_cast<long long>(blockIdx.x) * (Wp * C8) + i;
  uint4 o = make_uint4(0, 0, 0, 0);
  if (y >= 1 && y <= H && xx >= 1 && xx <= W) {
    const int plane = ((y & 1) << 1) | (xx & 1);
    uint4 t = make_uint4(0, 0, 0, 0);
    if (only_plane < 0) t = __ldg(dxs + (((static_cast<long long>(plane) * B + b) * Hs + (y >> 1)) * Ws + (xx >> 1)) * C8 + c);
    else if (plane == only_plane) t = __ldg(dxs + ((static_cast<long long>(b) * Hs + (y >> 1)) * Ws + (xx >> 1)) * C8 + c);
    float f[8] = {t_lo(t.x), t_hi(t.x), t_lo(t.y), t_hi(t.y), t_lo(t.z), t_hi(t.z), t_lo(t.w), t_hi(t.w)};
    if (add) {
      const uint4 a = __ldg(add + idx);
      f[0] += t_lo(a.x); f[1] += t_hi(a.x); f[2] += t_lo(a.y); f[3] += t_hi(a.y);
      f[4] += t_lo(a.z); f[5] += t_hi(a.z); f[6] += t_lo(a.w); f[7] += t_hi(a.w);
    }
    if (mask_src) {
      const uint4 m = __ldg(mask_src + idx);
      const float g[8] = {t_lo(m.x), t_hi(m.x), t_lo(m.y), t_hi(m.y), t_lo(m.z), t_hi(m.z), t_lo(m.w), t_hi(m.w)};
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (!(g[e] > 0.f)) f[e] = 0.f;
    }
    o.x = pack_t2(f[0], f[1]); o.y = pack_t2(f[2], f[3]); o.z = pack_t2(f[4], f[5]); o.w = pack_t2(f[6], f[7]);
  }
  dx[idx] = o;
}

// ------------------------------------------------------------------------------------------------ weight packing
// Conv weight fp32 [Cout,Cin,kh,kw] (+ FrozenBN buffers, backbone.py:70-80) ->
//   fwd  bf16 [Cout, ldk]            column (r*kw+s)*Cin + ci            (zero padded to ldk)
//   dgr  bf16 [Cin, kh*kw*Cout]      column ((kh-1-r)*kw+(kw-1-s))*Cout + co   (taps flipped: dgrad == conv with same shifts)
//   scale[co] = w*rsqrt(rv+eps), bias[co] = b - rm*scale  (scale = 1, bias = conv bias when there is no BN)
__global__ void pack_conv_kernel(const float* __restrict__ w, int Cout, int Cin, int kh, int kw, const float* bn_w, const float* bn_b,
                                 const float* bn_rm, const float* bn_rv, float eps, const float* conv_bias, rb_t* fwd, int ldk,
                                 rb_t* dgr, float* scale_out, float* bias_out) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(Cout) * ldk;
  if (idx >= total) return;
  const int co = static_cast<int>(idx / ldk), k = static_cast<int>(idx % ldk);
  float sc = 1.f, bi = conv_bias ? conv_bias[co] : 0.f;
  if (bn_w) {
    sc = bn_w[co] * rsqrtf(bn_rv[co] + eps);
    bi = bn_b[co] - bn_rm[co] * sc;
  }
  if (k == 0) {
    if (scale_out) scale_out[co] = sc;
    if (bias_out) bias_out[co] = bi;
  }
  float v = 0.f;
  const int taps = kh * kw;
  if (k < taps * Cin) {
    const int ci = k % Cin, t = k / Cin, s = t % kw, r = t / kw;
    v = w[((static_cast<long long>(co) * Cin + ci) * kh + r) * kw + s] * sc;
    if (dgr) dgr[static_cast<long long>(ci) * taps * Cout + (static_cast<long long>((kh - 1 - r) * kw + (kw - 1 - s))) * Cout + co] = f2t(v);
  }
  fwd[idx] = f2t(v);
}

// Linear weight fp32 [N,K] -> wb bf16 [N,K] and (optionally) wt bf16 [K,N].
__global__ void pack_linear_kernel(const float* __restrict__ w, int N, int K, rb_t* wb, long long ldwb, rb_t* wt, long long ldwt) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    float v = 0.f;
    if (n < N && k < K) {
      v = w[static_cast<long long>(n) * K + k];
      if (wb) wb[static_cast<long long>(n) * ldwb + k] = f2t(v);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  if (wt) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int k = k0 + i, n = n0 + threadIdx.x;
      if (n < N && k < K) wt[static_cast<long long>(k) * ldwt + n] = f2t(tile[threadIdx.x][i]);
    }
  }
}

// Linear weight fp32 [N,K] -> w2 [N, 2K] 16-bit: columns [0,K) = the weight rounded to 16 bits, columns [K,2K) = the rounding residual
// (w - hi) rounded to 16 bits.  A GEMM over both halves (two taps reading the same A) multiplies by the weight to ~2^-22.
__global__ void pack_linear_hilo_kernel(const float* __restrict__ w, int N, int K, rb_t* __restrict__ w2, long long ld) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= static_cast<long long>(N) * K) return;
  const int n = static_cast<int>(i / K), k = static_cast<int>(i - static_cast<long long>(n) * K);
  const float v = w[i];
  const rb_t hi = f2t(v);
  w2[n * ld + k] = hi;
  w2[n * ld + K + k] = f2t(v - t2f(hi));
}

// dWfold fp32 [Cout, taps, Cin] -> grad fp32 [Cout, Cin, kh, kw] * scale[co]
__global__ void unpack_conv_grad_kernel(const float* __restrict__ dwf, const float* __restrict__ scale, float* __restrict__ grad, int Cout,
                                        int Cin, int taps) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(Cout) * Cin * taps;
  if (idx >= total) return;
  const int t = static_cast<int>(idx % taps);
  const int ci = static_cast<int>((idx / taps) % Cin);
  const int co = static_cast<int>(idx / (static_cast<long long>(taps) * Cin));
  grad[idx] = dwf[(static_cast<long long>(co) * taps + t) * Cin + ci] * (scale ? scale[co] : 1.f);
}

// ------------------------------------------------------------------------------------------------ casts / reductions
__global__ void cast_bf16_kernel(const float* __restrict__ in, rb_t* __restrict__ out, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    uint2 o;
    o.x = pack_t2(v.x, v.y); o.y = pack_t2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + i) = o;
  } else {
    for (long long j = i; j < n; ++j) out[j] = f2t(in[j]);
  }
}

// out[n] (+)= sum over rows of x[row, n]; x is bf16 or fp32; rows are split over blockIdx.y, partials added atomically.
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long long ld, long long rows, int N, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float acc = 0.f;
  for (long long r = r0; r < r1; ++r) acc += static_cast<float>(x[r * ld + n]);
  atomicAdd(out + n, acc);
}

// Vectorised variant: a block is 32 column-lanes (VEC columns each) x 8 row-lanes; each thread walks its rows with 4 independent
// loads in flight, the 8 row-lanes are reduced in shared memory, one atomic per column per block.
template <typename T, int VEC>
__global__ void colsum_vec_kernel(const T* __restrict__ x, long long ld, long long rows, int N, float* __restrict__ out, int rows_per_block) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n0 = (blockIdx.x * 32 + tx) * VEC;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  if (n0 < N) {
    for (long long r = r0 + ty; r < r1; r += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long rr = r + 8 * u;
        if (rr < r1) {
          if (sizeof(T) == 2) {
            const uint4 v = *reinterpret_cast<const uint4*>(x + rr * ld + n0);
            acc[0] += t_lo(v.x); acc[1] += t_hi(v.x); acc[2] += t_lo(v.y); acc[3] += t_hi(v.y);
            acc[4 % VEC] += t_lo(v.z); acc[5 % VEC] += t_hi(v.z); acc[6 % VEC] += t_lo(v.w); acc[7 % VEC] += t_hi(v.w);
          } else {
            const float4 v = *reinterpret_cast<const float4*>(x + rr * ld + n0);
            acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
          }
        }
      }
    }
  }
  __shared__ float red[8][32 * VEC + 1];
#pragma unroll
  for (int i = 0; i < VEC; ++i) red[ty][tx * VEC + i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VEC; c += blockDim.x) {
    const int n = blockIdx.x * 32 * VEC + c;
    if (n >= N) continue;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][c];
    atomicAdd(out + n, v);
  }
}

// y = a + b (fp32), optional bf16 copy
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, rb_t* __restrict__ yb, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = a[i] + (b ? b[i] : 0.f);
  if (y) y[i] = v;
  if (yb) yb[i] = f2t(v);
}

}  // namespace rb

using namespace rb;

#define RB_CHECK_LAUNCH() RB_CUDA(cudaGetLastError())

extern "C" int rb_stem_im2col(const float* img, void* out, int B, int H, int W, int H1, int W1, void* stream) {
  const long long total = static_cast<long long>(B) * H1 * W1 * 20;
  stem_im2col_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, static_cast<uint4*>(out), B, H, W, H1, W1);
  RB_CHECK_LAUNCH();
  return 0;
}

extern "C" int rb_maxpool_3x3s2(const void* in, void* out, int B, int H1, int W1, int C, int H2, int W2, void* stream) {
  if (C % 8) return rb_fail("rb_maxpool_3x3s2: C must be a multiple of 8");
  maxpool_kernel<<<dim3(static_cast<unsigned>(B * (H2 + 2)), blocks_for(static_cast<long long>(W2 + 2) * (C / 8), 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), B, H1, W1, C / 8, H2, W2);
  RB_CHECK_LAUNCH();
  return 0;
}

extern "C" int rb_parity_split(const void* x, void* xs, int B, int H, int W, int C, int Ho, int Wo, void* stream) {
  if (C % 8) return rb_fail("rb_parity_split: C must be a multiple of 8");
  parity_split_kernel<<<dim3(static_cast<unsigned>(4 * B * (Ho + 2)), blocks_for(static_cast<long long>(Wo + 2) * (C / 8), 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(x), static_cast<uint4*>(xs), B, H, W, C / 8, Ho + 2, Wo + 2, -1);
  RB_CHECK_LAUNCH();
  return 0;
}

extern "C" int rb_parity_split_plane(const void* x, void* xs_plane, int B, int H, int W, int C, int Ho, int Wo, int plane, void* stream) {
  if (C % 8) return rb_fail("rb_parity_split_plane: C must be a multiple of 8");
  if (plane < 0 || plane > 3) return rb_fail("rb_parity_split_plane: plane must be 0..3");
  parity_split_kernel<<<dim3(static_cast<unsigned>(B * (Ho + 2)), blocks_for(static_cast<long long>(Wo + 2) * (C / 8), 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(x), static_cast<uint4*>(xs_plane), B, H, W, C / 8, Ho + 2, Wo + 2, plane);
  RB_CHECK_LAUNCH();
  return 0;
}

extern "C" int rb_parity_merge_plane(const void* dxs_plane, int plane, const void* add, const void* mask_src, void* dx, int B, int H, int W, int C, int Ho, int Wo,
                                     void* stream) {
  if (C % 8) return rb_fail("rb_parity_merge_plane: C must be a multiple of 8");
  if (plane < 0 || plane > 3) return rb_fail("rb_parity_merge_plane: plane must be 0..3");
  parity_merge_kernel<<<dim3(static_cast<unsigned>(B * (H + 2)), blocks_for(static_cast<long long>(W + 2) * (C / 8), 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dxs_plane), static_cast<const uint4*>(add), static_cast<const uint4*>(mask_src), static_cast<uint4*>(dx), B, H, W, C / 8, Ho + 2, Wo + 2, plane);
  RB_CHECK_LAUNCH();
  return 0;
}

extern "C" int rb_parity_merge(const void* dxs, const void* add, const void* mask_src, void* dx, int B, int H, int W, int C, int Ho, int Wo, void* stream) {
  if (C % 8) return rb_fail("rb_parity_merge: C must be a multiple of 8");
  parity_merge_kernel<<<dim3(static_cast<unsigned>(B * (H + 2)), blocks_for(static_cast<long long>(W + 2) * (C / 8), 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dxs), static_cast<const uint4*>(add), static_cast<const uint4*>(mask_src), static_cast<uint4*>(dx), B, H, W, C / 8, Ho + 2, Wo + 2, -1);
  RB_CHECK_LAUNCH();
  return 0;
}

extern "C" int rb_pack_conv(const float* w, int Cout, int Cin, int kh, int kw, const float* bn_w, const float* bn_b, const float* bn_rm,
                            const float* bn_rv, float eps, const float* conv_bias, void* fwd, int ldk, void* dgr, float* scale_out, float* bias_out,
                            void* stream) {
  if (ldk < kh * kw * Cin) return rb_fail("rb_pack_conv: ldk too small");
  const long long total = static_cast<long long>(Cout) * ldk;
  pack_conv_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(w, Cout, Cin, kh, kw, bn_w, bn_b, bn_rm, bn_rv, eps, conv_bias,
                                                                                          static_cast<rb_t*>(fwd), ldk, static_cast<rb_t*>(dgr), scale_out, bias_out);
  RB_CHECK_LAUNCH();
  return 0;
}

extern "C" int rb_pack_linear(const float* w, int N, int K, void* wb, long long ldwb, void* wt, long long ldwt, void* stream) {
  dim3 grid((K + 31) / 32, (N + 31) / 32);
  pack_linear_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(w, N, K, static_cast<rb_t*>(wb), ldwb, static_cast<rb_t*>(wt), ldwt);
  RB_CHECK_LAUNCH();
  return 0;
}

extern "C" int rb_pack_linear_hilo(const float* w, int N, int K, void* w2, long long ld, void* stream) {
  if (!w || !w2 || N <= 0 || K <= 0 || ld < 2LL * K) return rb_fail("rb_pack_linear_hilo: bad arguments");
  pack_linear_hilo_kernel<<<blocks_for(static_cast<long long>(N) * K, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(w, N, K, static_cast<rb_t*>(w2), ld);
  RB_CHECK_LAUNCH();
  return 0;
}

extern "C" int rb_unpack_conv_grad(const float* dwf, const float* scale, float* grad, int Cout, int Cin, int taps, void* stream) {
  const long long total = static_cast<long long>(Cout) * Cin * taps;
  unpack_conv_grad_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dwf, scale, grad, Cout, Cin, taps);
  RB_CHECK_LAUNCH();
  return 0;
}

extern "C" int rb_cast_bf16(const float* in, void* out, long long n, void* stream) {
  if (n <= 0) return 0;
  cast_bf16_kernel<<<blocks_for((n + 3) / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, static_cast<rb_t*>(out), n);
  RB_CHECK_LAUNCH();
  return 0;
}

extern "C" int rb_colsum(const void* x, int is_bf16, long long ld, long long rows, int N, float* out, void* stream) {
  if (rows <= 0 || N <= 0) return 0;
  const int vec = is_bf16 ? 8 : 4;
  if (N % vec == 0 && ld % vec == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // rows per block: enough blocks to fill the machine, at least 64 rows each
    const int gx = (N + 32 * vec - 1) / (32 * vec);
    long long rpb = (rows * gx + 591) / 592;
    rpb = rpb < 64 ? 64 : ((rpb + 31) / 32) * 32;
    dim3 g(gx, static_cast<unsigned>((rows + rpb - 1) / rpb));
    if (is_bf16)
      colsum_vec_kernel<rb_t, 8><<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const rb_t*>(x), ld, rows, N, out, static_cast<int>(rpb));
    else
      colsum_vec_kernel<float, 4><<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const float*>(x), ld, rows, N, out, static_cast<int>(rpb));
    RB_CHECK_LAUNCH();
    return 0;
  }
  int ysplit = static_cast<int>(rows / 256);
  ysplit = ysplit < 1 ? 1 : (ysplit > 64 ? 64 : ysplit);
  dim3 grid((N + 127) / 128, ysplit);
  if (is_bf16)
    colsum_kernel<rb_t><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const rb_t*>(x), ld, rows, N, out);
  else
    colsum_kernel<float><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const float*>(x), ld, rows, N, out);
  RB_CHECK_LAUNCH();
  return 0;
}

extern "C" int rb_add(const float* a, const float* b, float* y, void* yb, long long n, void* stream) {
  if (n <= 0) return 0;
  add_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, y, static_cast<rb_t*>(yb), n);
  RB_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------ fused box losses
// L1 + GIoU of paired cxcywh boxes for all decoder layers at once (criterion.py:113-153, :189-201; box_ops.py:9-13, :52-77),
// forward values and the gradient w.r.t. the predicted boxes in one pass.
namespace rb {
__global__ void box_loss_kernel(const float* __restrict__ boxes, const float* __restrict__ tgt, const uint8_t* __restrict__ valid, int n_layers, int N,
                                float inv_norm, const float* __restrict__ inv_norm_dev, float* __restrict__ losses, float* __restrict__ dl1,
                                float* __restrict__ dgiou) {
  const int layer = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const float inv = inv_norm_dev ? *inv_norm_dev : inv_norm;
  float l1 = 0.f, lg = 0.f;
  if (n < N) {
    const long long o = (static_cast<long long>(layer) * N + n) * 4;
    float g1[4] = {0.f, 0.f, 0.f, 0.f}, gg[4] = {0.f, 0.f, 0.f, 0.f};
    if (!valid || valid[n]) {
      const float4 p = *reinterpret_cast<const float4*>(boxes + o);
      const float4 t = *reinterpret_cast<const float4*>(tgt + static_cast<long long>(n) * 4);
      const float pv[4] = {p.x, p.y, p.z, p.w}, tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float d = pv[i] - tv[i];
        l1 += fabsf(d);
        g1[i] = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      }
      const float a[4] = {p.x - 0.5f * p.z, p.y - 0.5f * p.w, p.x + 0.5f * p.z, p.y + 0.5f * p.w};
      const float b[4] = {t.x - 0.5f * t.z, t.y - 0.5f * t.w, t.x + 0.5f * t.z, t.y + 0.5f * t.w};
      const float Aa = (a[2] - a[0]) * (a[3] - a[1]), Ab = (b[2] - b[0]) * (b[3] - b[1]);
      const float w0 = fmaxf(fminf(a[2], b[2]) - fmaxf(a[0], b[0]), 0.f), w1 = fmaxf(fminf(a[3], b[3]) - fmaxf(a[1], b[1]), 0.f);
      const float I = w0 * w1, U = Aa + Ab - I;
      const float c0 = fmaxf(fmaxf(a[2], b[2]) - fminf(a[0], b[0]), 0.f), c1 = fmaxf(fmaxf(a[3], b[3]) - fminf(a[1], b[1]), 0.f);
      const float C = c0 * c1;
      const float giou = I / U - (C - U) / C;
      lg = 1.f - giou;
      // derivatives w.r.t. the corners a0..a3
      const float dAa[4] = {-(a[3] - a[1]), -(a[2] - a[0]), (a[3] - a[1]), (a[2] - a[0])};
      const float dI[4] = {(w0 > 0.f && a[0] > b[0]) ? -w1 : 0.f, (w1 > 0.f && a[1] > b[1]) ? -w0 : 0.f, (w0 > 0.f && a[2] < b[2]) ? w1 : 0.f,
                           (w1 > 0.f && a[3] < b[3]) ? w0 : 0.f};
      const float dC[4] = {(a[0] < b[0]) ? -c1 : 0.f, (a[1] < b[1]) ? -c0 : 0.f, (a[2] > b[2]) ? c1 : 0.f, (a[3] > b[3]) ? c0 : 0.f};
      float da[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dU = dAa[i] - dI[i];
        const float diou = (dI[i] * U - I * dU) / (U * U);
        const float dgi = diou + (dU * C - U * dC[i]) / (C * C);
        da[i] = -dgi;  // d(1 - giou)
      }
      gg[0] = da[0] + da[2];
      gg[1] = da[1] + da[3];
      gg[2] = 0.5f * (da[2] - da[0]);
      gg[3] = 0.5f * (da[3] - da[1]);
    }
    *reinterpret_cast<float4*>(dl1 + o) = make_float4(g1[0] * inv, g1[1] * inv, g1[2] * inv, g1[3] * inv);
    *reinterpret_cast<float4*>(dgiou + o) = make_float4(gg[0] * inv, gg[1] * inv, gg[2] * inv, gg[3] * inv);
  }
  l1 = warp_sum(l1);
  lg = warp_sum(lg);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(losses + layer * 2, l1 * inv);
    atomicAdd(losses + layer * 2 + 1, lg * inv);
  }
}
}  // namespace rb

extern "C" int rb_box_loss(const float* boxes, const float* tgt, const void* valid, int n_layers, int N, float inv_norm, const float* inv_norm_dev,
                           float* losses, float* dl1, float* dgiou, void* stream) {
  if (n_layers <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RB_CUDA(cudaMemsetAsync(losses, 0, sizeof(float) * 2 * n_layers, st));
  if (N <= 0) return 0;  // no boxes: every loss is exactly 0 (the caller allocates `losses` uninitialised)
  rb::box_loss_kernel<<<dim3((N + 127) / 128, n_layers), 128, 0, st>>>(boxes, tgt, static_cast<const uint8_t*>(valid), n_layers, N, inv_norm, inv_norm_dev, losses,
                                                                     dl1, dgiou);
  RB_CHECK_LAUNCH();
  return 0;
}
