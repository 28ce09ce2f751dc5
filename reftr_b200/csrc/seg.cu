// Segmentation-head support kernels (HBM / latency bound): token <-> grid re-layout, MHAttentionMap softmax
// (reftr_segmentation.py:196-207), GroupNorm on padded NHWC grids, nearest upsample + add (reftr_segmentation.py:243-280).
// The convolutions of the mask head themselves run on rb_gemm (9 shifted taps over padded NHWC).
#include "common.cuh"
#include "host.h"

namespace rb {

// ------------------------------------------------------------------------------------------------ tokens <-> grid
__global__ void tokens_to_grid_kernel(const float* __restrict__ tok, int S, int L, int h, int w, int C, rb_t* __restrict__ grid, long long ld,
                                      int col0, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = static_cast<int>(i % (C / 4));
  long long t = i / (C / 4);
  const int p = static_cast<int>(t % (h * w));
  const int b = static_cast<int>(t / (h * w));
  const int y = p / w, x = p - y * w;
  const float4 v = *reinterpret_cast<const float4*>(tok + (static_cast<long long>(b) * S + L + p) * C + c4 * 4);
  rb_t* dst = grid + ((static_cast<long long>(b) * (h + 2) + y + 1) * (w + 2) + x + 1) * ld + col0 + c4 * 4;
  uint2 o;
  o.x = pack_t2(v.x, v.y);
  o.y = pack_t2(v.z, v.w);
  *reinterpret_cast<uint2*>(dst) = o;
}

__global__ void grid_to_tokens_kernel(const rb_t* __restrict__ grid, long long ld, int col0, int S, int L, int h, int w, int C,
                                      float* __restrict__ dtok, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = static_cast<int>(i % (C / 4));
  long long t = i / (C / 4);
  const int p = static_cast<int>(t % (h * w));
  const int b = static_cast<int>(t / (h * w));
  const int y = p / w, x = p - y * w;
  const uint2 v = *reinterpret_cast<const uint2*>(grid + ((static_cast<long long>(b) * (h + 2) + y + 1) * (w + 2) + x + 1) * ld + col0 + c4 * 4);
  *reinterpret_cast<float4*>(dtok + (static_cast<long long>(b) * S + L + p) * C + c4 * 4) = make_float4(t_lo(v.x), t_hi(v.x), t_lo(v.y), t_hi(v.y));
}

// ------------------------------------------------------------------------------------------------ attention map
// one CTA per sample; 8 warps = 8 heads; logits live in shared memory [8][hw]
__global__ void __launch_bounds__(256)
attn_map_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const uint8_t* __restrict__ kpm, int S, int L, int hw, int w, float scale,
                    float* __restrict__ att, rb_t* __restrict__ grid, long long ld, int col0) {
  extern __shared__ float sm[];
  float* lg = sm;  // [8][hw]
  __shared__ float red[8];
  __shared__ float bcast;
  const int b = blockIdx.x, n = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float qv = q[static_cast<long long>(b) * 256 + n * 32 + lane] * scale;
  float mx = -INFINITY;
  for (int p = 0; p < hw; ++p) {
    float v = qv * k[(static_cast<long long>(b) * S + L + p) * 256 + n * 32 + lane];
    v = warp_sum(v);
    if (kpm[static_cast<long long>(b) * S + L + p]) v = -INFINITY;
    if (lane == 0) lg[n * hw + p] = v;
    mx = fmaxf(mx, v);
  }
  if (lane == 0) red[n] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    bcast = (m == -INFINITY) ? 0.f : m;
  }
  __syncthreads();
  const float m = bcast;
  float sum = 0.f;
  for (int i = threadIdx.x; i < 8 * hw; i += 256) {
    const float e = __expf(lg[i] - m);
    lg[i] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[n] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    bcast = s > 0.f ? 1.f / s : 0.f;
  }
  __syncthreads();
  const float inv = bcast;
  const int h = hw / w;
  for (int i = threadIdx.x; i < 8 * hw; i += 256) {
    const int nn = i / hw, p = i - nn * hw;
    const float a = lg[i] * inv;
    att[(static_cast<long long>(b) * 8 + nn) * hw + p] = a;
    const int y = p / w, x = p - y * w;
    grid[((static_cast<long long>(b) * (h + 2) + y + 1) * (w + 2) + x + 1) * ld + col0 + nn] = f2t(a);
  }
}

__global__ void __launch_bounds__(256)
attn_map_bwd_kernel(const float* __restrict__ datt_ext, const rb_t* __restrict__ dgrid, long long ld, int col0, const float* __restrict__ att,
                    const float* __restrict__ q, const float* __restrict__ k, int S, int L, int hw, int w, float scale, float* __restrict__ dq,
                    float* __restrict__ dk) {
  extern __shared__ float sm[];
  float* dl = sm;  // [8][hw]: d(att), then d(logit)
  __shared__ float red[8];
  __shared__ float bcast;
  const int b = blockIdx.x, n = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = hw / w;
  float dot = 0.f;
  for (int i = threadIdx.x; i < 8 * hw; i += 256) {
    const int nn = i / hw, p = i - nn * hw;
    const int y = p / w, x = p - y * w;
    float d = t2f(dgrid[((static_cast<long long>(b) * (h + 2) + y + 1) * (w + 2) + x + 1) * ld + col0 + nn]);
    if (datt_ext) d += datt_ext[(static_cast<long long>(b) * 8 + nn) * hw + p];
    dl[i] = d;
    dot += d * att[static_cast<long long>(b) * 8 * hw + i];
  }
  dot = warp_sum(dot);
  if (lane == 0) red[n] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    bcast = s;
  }
  __syncthreads();
  const float tot = bcast;
  for (int i = threadIdx.x; i < 8 * hw; i += 256) dl[i] = att[static_cast<long long>(b) * 8 * hw + i] * (dl[i] - tot) * scale;
  __syncthreads();
  // language rows of dk are zero
  for (int i = threadIdx.x; i < L * 256; i += 256) dk[static_cast<long long>(b) * S * 256 + i] = 0.f;
  const float qv = q[static_cast<long long>(b) * 256 + n * 32 + lane];
  float acc = 0.f;
  for (int p = 0; p < hw; ++p) {
    const float g = dl[n * hw + p];
    const long long row = (static_cast<long long>(b) * S + L + p) * 256 + n * 32 + lane;
    acc += g * k[row];
    dk[row] = g * qv;
  }
  dq[static_cast<long long>(b) * 256 + n * 32 + lane] = acc;
}

// ------------------------------------------------------------------------------------------------ GroupNorm NHWC
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}

// one CTA per (sample, group); thread t walks pixels t, t+blockDim, ... and the Cg channels of the group
__global__ void __launch_bounds__(256)
groupnorm_nhwc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int H, int W, int C, int G,
                          float eps, int relu, rb_t* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  __shared__ float red[8];
  const int b = blockIdx.x, g = blockIdx.y;
  const int Cg = C / G, Wp = W + 2, Hp = H + 2;
  const long long base = static_cast<long long>(b) * Hp * Wp;
  float s = 0.f, qq = 0.f;
  for (int p = threadIdx.x; p < H * W; p += blockDim.x) {
    const int yy = p / W, xx = p - yy * W;
    const float* r = x + (base + static_cast<long long>(yy + 1) * Wp + xx + 1) * C + g * Cg;
    for (int c = 0; c < Cg; ++c) {
      const float v = r[c];
      s += v;
      qq += v * v;
    }
  }
  const float n = static_cast<float>(H) * W * Cg;
  const float mu = block_sum(s, red) / n;
  const float var = block_sum(qq, red) / n - mu * mu;
  const float rs = rsqrtf(fmaxf(var, 0.f) + eps);
  if (threadIdx.x == 0) {
    mean_out[b * G + g] = mu;
    rstd_out[b * G + g] = rs;
  }
  for (int p = threadIdx.x; p < Hp * Wp; p += blockDim.x) {
    const int yy = p / Wp, xx = p - yy * Wp;
    const bool interior = yy >= 1 && yy <= H && xx >= 1 && xx <= W;
    const long long o = (base + p) * C + g * Cg;
    for (int c = 0; c < Cg; ++c) {
      float v = 0.f;
      if (interior) {
        v = (x[o + c] - mu) * rs * gamma[g * Cg + c] + beta[g * Cg + c];
        if (relu) v = fmaxf(v, 0.f);
      }
      y[o + c] = f2t(v);
    }
  }
}

__global__ void __launch_bounds__(256)
groupnorm_nhwc_bwd_kernel(const rb_t* __restrict__ dy, const rb_t* __restrict__ y, const float* __restrict__ x,
                          const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd, int H, int W, int C, int G,
                          int relu, rb_t* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[8];
  extern __shared__ float csum[];  // [2][Cg] per-channel partial sums (dgamma, dbeta) of this CTA
  const int b = blockIdx.x, g = blockIdx.y;
  const int Cg = C / G, Wp = W + 2, Hp = H + 2;
  const long long base = static_cast<long long>(b) * Hp * Wp;
  const float mu = mean[b * G + g], rs = rstd[b * G + g];
  for (int c = threadIdx.x; c < 2 * Cg; c += blockDim.x) csum[c] = 0.f;
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  // pass 1: sums over the group; per-channel dgamma / dbeta with the channel loop OUTSIDE so one smem atomic per (thread, channel)
  for (int c = 0; c < Cg; ++c) {
    float dg = 0.f, db = 0.f;
    const float gm = gamma[g * Cg + c];
    for (int p = threadIdx.x; p < H * W; p += blockDim.x) {
      const int yy = p / W, xx = p - yy * W;
      const long long o = (base + static_cast<long long>(yy + 1) * Wp + xx + 1) * C + g * Cg + c;
      float d = t2f(dy[o]);
      if (relu && !(t2f(y[o]) > 0.f)) d = 0.f;
      const float xh = (x[o] - mu) * rs;
      dg += d * xh;
      db += d;
      s1 += d * gm;
      s2 += d * gm * xh;
    }
    dg = warp_sum(dg);
    db = warp_sum(db);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&csum[c], dg);
      atomicAdd(&csum[Cg + c], db);
    }
  }
  const float n = static_cast<float>(H) * W * Cg;
  const float m1 = block_sum(s1, red) / n;
  const float m2 = block_sum(s2, red) / n;
  for (int c = threadIdx.x; c < Cg; c += blockDim.x) {
    atomicAdd(dgamma + g * Cg + c, csum[c]);
    atomicAdd(dbeta + g * Cg + c, csum[Cg + c]);
  }
  for (int p = threadIdx.x; p < Hp * Wp; p += blockDim.x) {
    const int yy = p / Wp, xx = p - yy * Wp;
    const bool interior = yy >= 1 && yy <= H && xx >= 1 && xx <= W;
    const long long o = (base + p) * C + g * Cg;
    for (int c = 0; c < Cg; ++c) {
      float v = 0.f;
      if (interior) {
        float d = t2f(dy[o + c]);
        if (relu && !(t2f(y[o + c]) > 0.f)) d = 0.f;
        const float xh = (x[o + c] - mu) * rs;
        v = rs * (d * gamma[g * Cg + c] - m1 - xh * m2);
      }
      dx[o + c] = f2t(v);
    }
  }
}

// ------------------------------------------------------------------------------------------------ nearest upsample
__device__ __forceinline__ int nearest_src(int dst, int in, int out) {
  // F.interpolate(mode="nearest"): src = floor(dst * in / out), computed in float like ATen (scale = in / out)
  const float scale = static_cast<float>(in) / static_cast<float>(out);
  const int sidx = static_cast<int>(floorf(dst * scale));
  return sidx < in - 1 ? sidx : in - 1;
}

__global__ void upsample_add_kernel(const uint4* __restrict__ lo, const uint4* __restrict__ cur, uint4* __restrict__ y, int h, int w, int H, int W, int C8,
                                    long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C8);
  long long t = i / C8;
  const int X = static_cast<int>(t % (W + 2));
  t /= (W + 2);
  const int Y = static_cast<int>(t % (H + 2));
  const int b = static_cast<int>(t / (H + 2));
  uint4 o = make_uint4(0, 0, 0, 0);
  if (Y >= 1 && Y <= H && X >= 1 && X <= W) {
    const int sy = nearest_src(Y - 1, h, H), sx = nearest_src(X - 1, w, W);
    const uint4 a = lo[((static_cast<long long>(b) * (h + 2) + sy + 1) * (w + 2) + sx + 1) * C8 + c];
    const uint4 d = cur[i];
    o.x = pack_t2(t_lo(a.x) + t_lo(d.x), t_hi(a.x) + t_hi(d.x));
    o.y = pack_t2(t_lo(a.y) + t_lo(d.y), t_hi(a.y) + t_hi(d.y));
    o.z = pack_t2(t_lo(a.z) + t_lo(d.z), t_hi(a.z) + t_hi(d.z));
    o.w = pack_t2(t_lo(a.w) + t_lo(d.w), t_hi(a.w) + t_hi(d.w));
  }
  y[i] = o;
}

__global__ void upsample_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dlo, int h, int w, int H, int W, int C8, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C8);
  long long t = i / C8;
  const int x = static_cast<int>(t % (w + 2));
  t /= (w + 2);
  const int y = static_cast<int>(t % (h + 2));
  const int b = static_cast<int>(t / (h + 2));
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (y >= 1 && y <= h && x >= 1 && x <= w) {
    const int sy = y - 1, sx = x - 1;
    // candidate children: a window around sy * H / h (scale factors are small); membership is re-checked with nearest_src
    const int Y0 = max(0, static_cast<int>((static_cast<long long>(sy) * H) / h) - 1), Y1 = min(H - 1, static_cast<int>((static_cast<long long>(sy + 1) * H) / h) + 1);
    const int X0 = max(0, static_cast<int>((static_cast<long long>(sx) * W) / w) - 1), X1 = min(W - 1, static_cast<int>((static_cast<long long>(sx + 1) * W) / w) + 1);
    for (int Y = Y0; Y <= Y1; ++Y) {
      if (nearest_src(Y, h, H) != sy) continue;
      for (int X = X0; X <= X1; ++X) {
        if (nearest_src(X, w, W) != sx) continue;
        const uint4 v = dy[((static_cast<long long>(b) * (H + 2) + Y + 1) * (W + 2) + X + 1) * C8 + c];
        acc[0] += t_lo(v.x); acc[1] += t_hi(v.x); acc[2] += t_lo(v.y); acc[3] += t_hi(v.y);
        acc[4] += t_lo(v.z); acc[5] += t_hi(v.z); acc[6] += t_lo(v.w); acc[7] += t_hi(v.w);
      }
    }
  }
  uint4 o;
  o.x = pack_t2(acc[0], acc[1]); o.y = pack_t2(acc[2], acc[3]); o.z = pack_t2(acc[4], acc[5]); o.w = pack_t2(acc[6], acc[7]);
  dlo[i] = o;
}

}  // namespace rb

using namespace rb;

static unsigned nblocks(long long total, int threads) { return static_cast<unsigned>((total + threads - 1) / threads); }

extern "C" int rb_tokens_to_grid(const float* tok, int B, int S, int L, int h, int w, int C, void* grid, long long ld, int col0, void* stream) {
  if (C % 4 || col0 % 4 || ld % 4) return rb_fail("rb_tokens_to_grid: C, col0 and ld must be multiples of 4");
  const long long total = static_cast<long long>(B) * h * w * (C / 4);
  tokens_to_grid_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(tok, S, L, h, w, C, static_cast<rb_t*>(grid), ld, col0, total);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_grid_to_tokens(const void* grid, long long ld, int col0, int B, int S, int L, int h, int w, int C, float* dtok, void* stream) {
  if (C % 4 || col0 % 4 || ld % 4) return rb_fail("rb_grid_to_tokens: C, col0 and ld must be multiples of 4");
  const long long total = static_cast<long long>(B) * h * w * (C / 4);
  grid_to_tokens_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const rb_t*>(grid), ld, col0, S, L, h, w, C, dtok, total);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_attn_map_fwd(const float* q, const float* k, const void* kpm, int B, int S, int L, int hw, int w, float scale, float* att, void* grid,
                               long long ld, int col0, void* stream) {
  const size_t sm = static_cast<size_t>(8) * hw * sizeof(float);
  if (sm > 200 * 1024) return rb_fail("rb_attn_map_fwd: %d visual tokens exceed the shared-memory plan", hw);
  static bool cfg = false;
  if (!cfg) { RB_CUDA(cudaFuncSetAttribute(attn_map_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg = true; }
  attn_map_fwd_kernel<<<B, 256, sm, static_cast<cudaStream_t>(stream)>>>(q, k, static_cast<const uint8_t*>(kpm), S, L, hw, w, scale, att,
                                                                        static_cast<rb_t*>(grid), ld, col0);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_attn_map_bwd(const float* datt_ext, const void* dgrid, long long ld, int col0, const float* att, const float* q, const float* k, int B,
                               int S, int L, int hw, int w, float scale, float* dq, float* dk, void* stream) {
  const size_t sm = static_cast<size_t>(8) * hw * sizeof(float);
  if (sm > 200 * 1024) return rb_fail("rb_attn_map_bwd: %d visual tokens exceed the shared-memory plan", hw);
  static bool cfg = false;
  if (!cfg) { RB_CUDA(cudaFuncSetAttribute(attn_map_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg = true; }
  attn_map_bwd_kernel<<<B, 256, sm, static_cast<cudaStream_t>(stream)>>>(datt_ext, static_cast<const rb_t*>(dgrid), ld, col0, att, q, k, S, L, hw, w,
                                                                        scale, dq, dk);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_groupnorm_nhwc_fwd(const float* x, const float* gamma, const float* beta, int B, int H, int W, int C, int G, float eps, int relu, void* y,
                                     float* mean, float* rstd, void* stream) {
  if (C % G) return rb_fail("rb_groupnorm_nhwc_fwd: C %% G != 0");
  groupnorm_nhwc_fwd_kernel<<<dim3(B, G), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, gamma, beta, H, W, C, G, eps, relu, static_cast<rb_t*>(y),
                                                                                      mean, rstd);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_groupnorm_nhwc_bwd(const void* dy, const void* y, const float* x, const float* gamma, const float* mean, const float* rstd, int B, int H,
                                     int W, int C, int G, int relu, void* dx, float* dgamma, float* dbeta, void* stream) {
  if (C % G) return rb_fail("rb_groupnorm_nhwc_bwd: C %% G != 0");
  const size_t sm = static_cast<size_t>(2) * (C / G) * sizeof(float);
  groupnorm_nhwc_bwd_kernel<<<dim3(B, G), 256, sm, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const rb_t*>(dy), static_cast<const rb_t*>(y), x, gamma, mean, rstd, H, W, C, G, relu, static_cast<rb_t*>(dx),
      dgamma, dbeta);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_upsample_add(const void* lo, const void* cur, void* y, int B, int h, int w, int H, int W, int C, void* stream) {
  if (C % 8) return rb_fail("rb_upsample_add: C must be a multiple of 8");
  const long long total = static_cast<long long>(B) * (H + 2) * (W + 2) * (C / 8);
  upsample_add_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(lo), static_cast<const uint4*>(cur),
                                                                                         static_cast<uint4*>(y), h, w, H, W, C / 8, total);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_upsample_bwd(const void* dy, void* dlo, int B, int h, int w, int H, int W, int C, void* stream) {
  if (C % 8) return rb_fail("rb_upsample_bwd: C must be a multiple of 8");
  const long long total = static_cast<long long>(B) * (h + 2) * (w + 2) * (C / 8);
  upsample_bwd_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(dy), static_cast<uint4*>(dlo), h, w, H, W,
                                                                                         C / 8, total);
  RB_CUDA(cudaGetLastError());
  return 0;
}
