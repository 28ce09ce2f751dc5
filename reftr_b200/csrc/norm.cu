// LayerNorm / GroupNorm (forward + backward), position-embedding / key-padding-mask construction and the embedding
// gradient reduction of the VL transformer.  fp32 statistics, warp-shuffle reductions, one warp per 256-wide row.
#include "common.cuh"
#include "host.h"

namespace rb {

struct RowMap {  // out_row = (r / group) * stride + (r % group) + offset   (group == 0: identity)
  int group, stride, offset;
  __device__ __forceinline__ long long operator()(long long r) const {
    return group ? (r / group) * stride + (r % group) + offset : r;
  }
};

constexpr int LN_D = 256;

// ------------------------------------------------------------------------------------------------ LayerNorm fwd
// y = LN(x)*gamma+beta (optionally ReLU'd); writes fp32 y, bf16 y and bf16 (y + pos) at mapped rows; saves mean/rstd.
// Replaces nn.LayerNorm at transformer.py:176/:181/:242/:248/:252, reftr_transformer.py:17/:21/:38.
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, long long rows,
                                     float eps, float* __restrict__ y32, rb_t* __restrict__ yb, const float* __restrict__ pos32,
                                     rb_t* __restrict__ ypb, int relu, float* __restrict__ mean_out, float* __restrict__ rstd_out, RowMap map,
                                     DropK drop) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * LN_D + lane * 8);
  float4 a = xr[0], b = xr[1];
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.f / LN_D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] -= mean; q += v[i] * v[i]; }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / LN_D) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const float4 g0 = *reinterpret_cast<const float4*>(gamma + lane * 8), g1 = *reinterpret_cast<const float4*>(gamma + lane * 8 + 4);
  const float4 b0 = *reinterpret_cast<const float4*>(beta + lane * 8), b1 = *reinterpret_cast<const float4*>(beta + lane * 8 + 4);
  const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = v[i] * rstd * g[i] + bb[i];
    if (relu) v[i] = fmaxf(v[i], 0.f);
  }
  if (drop.seed) {  // nn.Dropout after the ReLU of mlp_mapping (reftr_transformer.py:19); site tensor = the compact [rows, 256]
    const uint32_t key = drop_key(drop);
    const uint32_t c0 = static_cast<uint32_t>(row) * (LN_D / 2) + lane * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t w = drop_word(key, c0 + i);
      v[2 * i] = drop_keep(w, 0, drop.thr) ? v[2 * i] * drop.scale : 0.f;
      v[2 * i + 1] = drop_keep(w, 1, drop.thr) ? v[2 * i + 1] * drop.scale : 0.f;
    }
  }
  const long long orow = map(row);
  if (y32) {
    float4* o = reinterpret_cast<float4*>(y32 + orow * LN_D + lane * 8);
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (yb) {
    uint4 o;
    o.x = pack_t2(v[0], v[1]); o.y = pack_t2(v[2], v[3]); o.z = pack_t2(v[4], v[5]); o.w = pack_t2(v[6], v[7]);
    *reinterpret_cast<uint4*>(yb + orow * LN_D + lane * 8) = o;
  }
  if (ypb) {
    const float4 p0 = *reinterpret_cast<const float4*>(pos32 + orow * LN_D + lane * 8), p1 = *reinterpret_cast<const float4*>(pos32 + orow * LN_D + lane * 8 + 4);
    uint4 o;
    o.x = pack_t2(v[0] + p0.x, v[1] + p0.y); o.y = pack_t2(v[2] + p0.z, v[3] + p0.w);
    o.z = pack_t2(v[4] + p1.x, v[5] + p1.y); o.w = pack_t2(v[6] + p1.z, v[7] + p1.w);
    *reinterpret_cast<uint4*>(ypb + orow * LN_D + lane * 8) = o;
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm bwd
// dy is read at mapped rows (+ optional second addend dy2 at the same rows); y (mapped rows) gives the ReLU mask.
// dx = rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*gamma; dgamma += dy*xhat, dbeta += dy (atomics, once per warp).
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dy2, const float* __restrict__ y_relu, float relu_scale,
                                     const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, long long rows, float* __restrict__ dx32, rb_t* __restrict__ dxb,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, RowMap map, DropK odrop) {
  const int lane = threadIdx.x & 31;
  const uint32_t okey = odrop.seed ? drop_key(odrop) : 0u;
  const int warps_per_block = blockDim.x >> 5;
  const long long warp_global = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  const long long n_warps = static_cast<long long>(gridDim.x) * warps_per_block;
  const float4 g0 = *reinterpret_cast<const float4*>(gamma + lane * 8), g1 = *reinterpret_cast<const float4*>(gamma + lane * 8 + 4);
  const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  float dg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, db[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long row = warp_global; row < rows; row += n_warps) {
    const long long irow = map(row);
    const float4* dr = reinterpret_cast<const float4*>(dy + irow * LN_D + lane * 8);
    float4 a = dr[0], b = dr[1];
    float d[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (dy2) {
      const float4* d2 = reinterpret_cast<const float4*>(dy2 + irow * LN_D + lane * 8);
      a = d2[0]; b = d2[1];
      d[0] += a.x; d[1] += a.y; d[2] += a.z; d[3] += a.w; d[4] += b.x; d[5] += b.y; d[6] += b.z; d[7] += b.w;
    }
    if (y_relu) {
      const float4* yr = reinterpret_cast<const float4*>(y_relu + irow * LN_D + lane * 8);
      a = yr[0]; b = yr[1];
      const float yy[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = (yy[i] > 0.f) ? d[i] * relu_scale : 0.f;
    }
    const float4* xr = reinterpret_cast<const float4*>(x + row * LN_D + lane * 8);
    a = xr[0]; b = xr[1];
    const float m = mean[row], rs = rstd[row];
    float xh[8] = {(a.x - m) * rs, (a.y - m) * rs, (a.z - m) * rs, (a.w - m) * rs, (b.x - m) * rs, (b.y - m) * rs, (b.z - m) * rs, (b.w - m) * rs};
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dg[i] += d[i] * xh[i];
      db[i] += d[i];
      d[i] *= g[i];
      s1 += d[i];
      s2 += d[i] * xh[i];
    }
    s1 = warp_sum(s1) * (1.f / LN_D);
    s2 = warp_sum(s2) * (1.f / LN_D);
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = rs * (d[i] - s1 - xh[i] * s2);
    if (dx32) {
      float4* p = reinterpret_cast<float4*>(dx32 + row * LN_D + lane * 8);
      p[0] = make_float4(o[0], o[1], o[2], o[3]);
      p[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
    if (dxb) {
      if (odrop.seed) {  // the bf16 copy feeds the backward of the layer whose output was dropped before this LN's residual sum
        const uint32_t c0 = static_cast<uint32_t>(row) * (LN_D / 2) + lane * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t w = drop_word(okey, c0 + i);
          o[2 * i] = drop_keep(w, 0, odrop.thr) ? o[2 * i] * odrop.scale : 0.f;
          o[2 * i + 1] = drop_keep(w, 1, odrop.thr) ? o[2 * i + 1] * odrop.scale : 0.f;
        }
      }
      uint4 p;
      p.x = pack_t2(o[0], o[1]); p.y = pack_t2(o[2], o[3]); p.z = pack_t2(o[4], o[5]); p.w = pack_t2(o[6], o[7]);
      *reinterpret_cast<uint4*>(dxb + row * LN_D + lane * 8) = p;
    }
  }
  // block-level reduction of the per-warp partial dgamma / dbeta, then ONE atomic per column per block
  __shared__ float red[2][8][LN_D + 8];
  const int wib = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[0][wib][lane * 8 + i] = dg[i];
    red[1][wib][lane * 8 + i] = db[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * LN_D; c += blockDim.x) {
    const int which = c / LN_D, col = c - which * LN_D;
    float* dst = which ? dbeta : dgamma;
    if (!dst) continue;
    float v = 0.f;
    for (int w = 0; w < warps_per_block; ++w) v += red[which][w][col];
    atomicAdd(dst + col, v);
  }
}

// ------------------------------------------------------------------------------------------------ GroupNorm -> tokens
// input_proj's GroupNorm(32, 256) (reftr_transformer.py:121-125): x is the 1x1-conv output in padded NHWC fp32
// [B, h+2, w+2, 256]; one CTA per (sample, group of 8 channels) normalises over the h*w interior pixels and writes
// token rows b*S + L + (y*w + x): fp32, bf16 and bf16(+pos).
__global__ void groupnorm_tokens_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int h, int w,
                                            int S, int L, float eps, float* __restrict__ y32, rb_t* __restrict__ yb,
                                            const float* __restrict__ pos32, rb_t* __restrict__ ypb, float* __restrict__ mean_out,
                                            float* __restrict__ rstd_out) {
  const int b = blockIdx.x, grp = blockIdx.y;
  const int hw = h * w, wp = w + 2;
  const float* xb = x + static_cast<long long>(b) * (h + 2) * wp * LN_D + grp * 8;
  __shared__ float red[2][32];
  float s = 0.f, q = 0.f;
  for (int p = threadIdx.x; p < hw; p += blockDim.x) {
    const int yy = p / w, xx = p - yy * w;
    const float4* r = reinterpret_cast<const float4*>(xb + (static_cast<long long>(yy + 1) * wp + xx + 1) * LN_D);
    const float4 a = r[0], c = r[1];
    s += a.x + a.y + a.z + a.w + c.x + c.y + c.z + c.w;
    q += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w;
  }
  s = warp_sum(s); q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    s = threadIdx.x < nw ? red[0][threadIdx.x] : 0.f;
    q = threadIdx.x < nw ? red[1][threadIdx.x] : 0.f;
    s = warp_sum(s); q = warp_sum(q);
    if (threadIdx.x == 0) { red[0][0] = s; red[1][0] = q; }
  }
  __syncthreads();
  const float n = static_cast<float>(hw) * 8.f;
  const float mean = red[0][0] / n;
  const float var = fmaxf(red[1][0] / n - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) { mean_out[b * 32 + grp] = mean; rstd_out[b * 32 + grp] = rstd; }
  float g[8], be[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { g[i] = gamma[grp * 8 + i]; be[i] = beta[grp * 8 + i]; }
  for (int p = threadIdx.x; p < hw; p += blockDim.x) {
    const int yy = p / w, xx = p - yy * w;
    const float4* r = reinterpret_cast<const float4*>(xb + (static_cast<long long>(yy + 1) * wp + xx + 1) * LN_D);
    const float4 a = r[0], c = r[1];
    float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * g[i] + be[i];
    const long long o = (static_cast<long long>(b) * S + L + p) * LN_D + grp * 8;
    float4* o4 = reinterpret_cast<float4*>(y32 + o);
    o4[0] = make_float4(v[0], v[1], v[2], v[3]);
    o4[1] = make_float4(v[4], v[5], v[6], v[7]);
    uint4 t;
    t.x = pack_t2(v[0], v[1]); t.y = pack_t2(v[2], v[3]); t.z = pack_t2(v[4], v[5]); t.w = pack_t2(v[6], v[7]);
    *reinterpret_cast<uint4*>(yb + o) = t;
    if (ypb) {
      const float4 p0 = *reinterpret_cast<const float4*>(pos32 + o), p1 = *reinterpret_cast<const float4*>(pos32 + o + 4);
      t.x = pack_t2(v[0] + p0.x, v[1] + p0.y); t.y = pack_t2(v[2] + p0.z, v[3] + p0.w);
      t.z = pack_t2(v[4] + p1.x, v[5] + p1.y); t.w = pack_t2(v[6] + p1.z, v[7] + p1.w);
      *reinterpret_cast<uint4*>(ypb + o) = t;
    }
  }
}

// Backward: dy at token rows (fp32, + optional dy2) -> dx bf16 in padded NHWC [B,h+2,w+2,256] (interior only; the
// caller zero-fills the buffer once), dgamma/dbeta accumulated atomically.
__global__ void groupnorm_tokens_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dy2, const float* __restrict__ x,
                                            const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd, int h, int w,
                                            int S, int L, rb_t* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int b = blockIdx.x, grp = blockIdx.y;
  const int hw = h * w, wp = w + 2;
  const long long img_off = static_cast<long long>(b) * (h + 2) * wp * LN_D + grp * 8;
  const float m = mean[b * 32 + grp], rs = rstd[b * 32 + grp];
  float g[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) g[i] = gamma[grp * 8 + i];
  __shared__ float red[2][32];
  __shared__ float redc[16][8];
  float dg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, db[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float s1 = 0.f, s2 = 0.f;
  for (int p = threadIdx.x; p < hw; p += blockDim.x) {
    const int yy = p / w, xx = p - yy * w;
    const long long o = (static_cast<long long>(b) * S + L + p) * LN_D + grp * 8;
    float4 a = *reinterpret_cast<const float4*>(dy + o), c = *reinterpret_cast<const float4*>(dy + o + 4);
    float d[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    if (dy2) {
      a = *reinterpret_cast<const float4*>(dy2 + o); c = *reinterpret_cast<const float4*>(dy2 + o + 4);
      d[0] += a.x; d[1] += a.y; d[2] += a.z; d[3] += a.w; d[4] += c.x; d[5] += c.y; d[6] += c.z; d[7] += c.w;
    }
    const float4* r = reinterpret_cast<const float4*>(x + img_off + (static_cast<long long>(yy + 1) * wp + xx + 1) * LN_D);
    a = r[0]; c = r[1];
    const float xh[8] = {(a.x - m) * rs, (a.y - m) * rs, (a.z - m) * rs, (a.w - m) * rs, (c.x - m) * rs, (c.y - m) * rs, (c.z - m) * rs, (c.w - m) * rs};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dg[i] += d[i] * xh[i];
      db[i] += d[i];
      const float t = d[i] * g[i];
      s1 += t;
      s2 += t * xh[i];
    }
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    s1 = threadIdx.x < nw ? red[0][threadIdx.x] : 0.f;
    s2 = threadIdx.x < nw ? red[1][threadIdx.x] : 0.f;
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (threadIdx.x == 0) { red[0][0] = s1; red[1][0] = s2; }
  }
  __syncthreads();
  const float n = static_cast<float>(hw) * 8.f;
  const float ms1 = red[0][0] / n, ms2 = red[1][0] / n;
  // channel-wise dgamma / dbeta
#pragma unroll
  for (int i = 0; i < 8; ++i) { dg[i] = warp_sum(dg[i]); db[i] = warp_sum(db[i]); }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { redc[threadIdx.x >> 5][i] = dg[i]; redc[8 + (threadIdx.x >> 5)][i] = db[i]; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    const int i = threadIdx.x & 7, which = threadIdx.x >> 3;
    float acc = 0.f;
    for (int wv = 0; wv < (blockDim.x >> 5); ++wv) acc += redc[which * 8 + wv][i];
    atomicAdd((which ? dbeta : dgamma) + grp * 8 + i, acc);
  }
  for (int p = threadIdx.x; p < hw; p += blockDim.x) {
    const int yy = p / w, xx = p - yy * w;
    const long long o = (static_cast<long long>(b) * S + L + p) * LN_D + grp * 8;
    float4 a = *reinterpret_cast<const float4*>(dy + o), c = *reinterpret_cast<const float4*>(dy + o + 4);
    float d[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    if (dy2) {
      a = *reinterpret_cast<const float4*>(dy2 + o); c = *reinterpret_cast<const float4*>(dy2 + o + 4);
      d[0] += a.x; d[1] += a.y; d[2] += a.z; d[3] += a.w; d[4] += c.x; d[5] += c.y; d[6] += c.z; d[7] += c.w;
    }
    const long long xi = img_off + (static_cast<long long>(yy + 1) * wp + xx + 1) * LN_D;
    const float4* r = reinterpret_cast<const float4*>(x + xi);
    a = r[0]; c = r[1];
    const float xh[8] = {(a.x - m) * rs, (a.y - m) * rs, (a.z - m) * rs, (a.w - m) * rs, (c.x - m) * rs, (c.y - m) * rs, (c.z - m) * rs, (c.w - m) * rs};
    float ov[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ov[i] = rs * (d[i] * g[i] - ms1 - xh[i] * ms2);
    uint4 t;
    t.x = pack_t2(ov[0], ov[1]); t.y = pack_t2(ov[2], ov[3]); t.z = pack_t2(ov[4], ov[5]); t.w = pack_t2(ov[6], ov[7]);
    *reinterpret_cast<uint4*>(dx + xi) = t;
  }
}

// ------------------------------------------------------------------------------------------------ positions + masks
// Builds, per sample: key-padding mask [B,S] (1 = ignore) and pos32 [B*S,256]:
//   language rows l<L : lang_pos[l] + token_type[0]              (reftr.py:82-89), mask = !sentence_mask
//   visual rows       : sine(y,x) + level_embed[0] + token_type[1] (position_encoding.py:36-56, reftr.py:57-73),
//                       mask = nearest-downsampled image padding mask (backbone.py:107)
__global__ void build_pos_mask_kernel(const uint8_t* __restrict__ img_mask, int H, int W, int h, int w, const long long* __restrict__ sent_mask, int L,
                                      const float* __restrict__ lang_pos, const float* __restrict__ token_type, const float* __restrict__ level_embed,
                                      float* __restrict__ pos32, uint8_t* __restrict__ kpm) {
  const int b = blockIdx.x;
  const int S = L + h * w;
  const int hw = h * w;
  extern __shared__ uint8_t sm[];
  uint8_t* nm = sm;                                                        // not_mask [hw]
  float* ey = reinterpret_cast<float*>(sm + ((hw + 15) & ~15));            // normalised y_embed [hw]
  float* ex = ey + hw;                                                     // normalised x_embed [hw]
  const float sh = static_cast<float>(H) / h, sw = static_cast<float>(W) / w;
  const bool first = blockIdx.y == 0;
  for (int p = threadIdx.x; p < hw; p += blockDim.x) {
    const int yy = p / w, xx = p - yy * w;
    const int ys = min(static_cast<int>(floorf(yy * sh)), H - 1), xs = min(static_cast<int>(floorf(xx * sw)), W - 1);
    const uint8_t m = img_mask[(static_cast<long long>(b) * H + ys) * W + xs];
    nm[p] = m ? 0 : 1;
    if (first) kpm[static_cast<long long>(b) * S + L + p] = m ? 1 : 0;
  }
  if (first) {
    for (int l = threadIdx.x; l < L; l += blockDim.x) kpm[static_cast<long long>(b) * S + l] = sent_mask[static_cast<long long>(b) * L + l] ? 0 : 1;
    for (int i = threadIdx.x; i < L * LN_D; i += blockDim.x) {
      const int l = i / LN_D, c = i - l * LN_D;
      pos32[(static_cast<long long>(b) * S + l) * LN_D + c] = lang_pos[l * LN_D + c] + token_type[c];
    }
  }
  __syncthreads();
  const float two_pi = 6.283185307179586f;
  // cumulative sums of not_mask along y (one thread per column) and along x (one thread per row), normalised to 2*pi
  for (int t = threadIdx.x; t < w + h; t += blockDim.x) {
    if (t < w) {
      int all = 0;
      for (int yy = 0; yy < h; ++yy) all += nm[yy * w + t];
      int cs = 0;
      for (int yy = 0; yy < h; ++yy) { cs += nm[yy * w + t]; ey[yy * w + t] = (static_cast<float>(cs) - 0.5f) / (static_cast<float>(all) + 1e-6f) * two_pi; }
    } else {
      const int yy = t - w;
      int all = 0;
      for (int xx = 0; xx < w; ++xx) all += nm[yy * w + xx];
      int cs = 0;
      for (int xx = 0; xx < w; ++xx) { cs += nm[yy * w + xx]; ex[yy * w + xx] = (static_cast<float>(cs) - 0.5f) / (static_cast<float>(all) + 1e-6f) * two_pi; }
    }
  }
  __syncthreads();
  // this CTA's slice of the visual tokens; thread = channel
  const int per = (hw + gridDim.y - 1) / gridDim.y;
  const int p0 = blockIdx.y * per, p1 = min(hw, p0 + per);
  const int c = threadIdx.x;  // blockDim.x == LN_D
  const int ci = c & 127;
  const float dim_t = powf(10000.f, static_cast<float>(2 * (ci / 2)) / 128.f);
  const float add = level_embed[c] + token_type[LN_D + c];
  for (int p = p0; p < p1; ++p) {
    const float e = (c < 128) ? ey[p] : ex[p];
    const float a = e / dim_t;
    const float v = (ci & 1) ? cosf(a) : sinf(a);
    pos32[(static_cast<long long>(b) * S + L + p) * LN_D + c] = v + add;
  }
}

// d(lang_pos)[l] = sum_b dpos[b,l]; d(token_type)[0] = sum_{b,l<L}; d(token_type)[1] = d(level_embed)[0] = sum_{b, visual}
__global__ void embed_grad_kernel(const float* __restrict__ dpos, int B, int S, int L, float* __restrict__ d_lang_pos, float* __restrict__ d_token_type,
                                  float* __restrict__ d_level) {
  const int c = threadIdx.x;  // 256 threads
  const int blk = blockIdx.x;
  if (blk < L) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += dpos[(static_cast<long long>(b) * S + blk) * LN_D + c];
    d_lang_pos[blk * LN_D + c] = acc;
    atomicAdd(d_token_type + c, acc);
  } else {
    const int chunk = blk - L, nchunk = gridDim.x - L;
    const long long total = static_cast<long long>(B) * (S - L);
    float acc = 0.f;
    for (long long i = chunk; i < total; i += nchunk) {
      const long long b = i / (S - L), p = i - b * (S - L);
      acc += dpos[(b * S + L + p) * LN_D + c];
    }
    atomicAdd(d_token_type + LN_D + c, acc);
    atomicAdd(d_level + c, acc);
  }
}

}  // namespace rb

using namespace rb;

extern "C" int rb_layernorm_fwd(const float* x, const float* gamma, const float* beta, long long rows, int D, float eps, float* y32, void* yb,
                                const float* pos32, void* ypb, int relu, float* mean, float* rstd, int map_group, int map_stride, int map_offset,
                                const rb_dropout* drop, void* stream) {
  if (D != LN_D) return rb_fail("rb_layernorm_fwd: only D == 256 is built (got %d)", D);
  if (rows <= 0) return 0;
  if (ypb && !pos32) return rb_fail("rb_layernorm_fwd: ypb needs pos32");
  RowMap m{map_group, map_stride, map_offset};
  layernorm_fwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, gamma, beta, rows, eps, y32, static_cast<rb_t*>(yb), pos32, static_cast<rb_t*>(ypb), relu, mean, rstd, m, make_dropk(drop));
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_layernorm_bwd(const float* dy, const float* dy2, const float* y_relu, float relu_scale, const float* x, const float* gamma,
                                const float* mean, const float* rstd, long long rows, int D, float* dx32, void* dxb, float* dgamma, float* dbeta,
                                int map_group, int map_stride, int map_offset, const rb_dropout* dxb_drop, void* stream) {
  if (D != LN_D) return rb_fail("rb_layernorm_bwd: only D == 256 is built (got %d)", D);
  if (rows <= 0) return 0;
  RowMap m{map_group, map_stride, map_offset};
  long long blocks = (rows + 7) / 8;
  if (blocks > 148) blocks = 148;  // one per SM; warps stride over rows, dgamma/dbeta cost one atomic per column per block
  layernorm_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dy, dy2, y_relu, relu_scale == 0.f ? 1.f : relu_scale, x, gamma, mean, rstd, rows, dx32, static_cast<rb_t*>(dxb), dgamma, dbeta, m,
      make_dropk(dxb_drop));
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_groupnorm_tokens_fwd(const float* x, const float* gamma, const float* beta, int B, int h, int w, int S, int L, float eps, float* y32,
                                       void* yb, const float* pos32, void* ypb, float* mean, float* rstd, void* stream) {
  groupnorm_tokens_fwd_kernel<<<dim3(B, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, gamma, beta, h, w, S, L, eps, y32, static_cast<rb_t*>(yb),
                                                                                       pos32, static_cast<rb_t*>(ypb), mean, rstd);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_groupnorm_tokens_bwd(const float* dy, const float* dy2, const float* x, const float* gamma, const float* mean, const float* rstd,
                                       int B, int h, int w, int S, int L, void* dx, float* dgamma, float* dbeta, void* stream) {
  groupnorm_tokens_bwd_kernel<<<dim3(B, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, dy2, x, gamma, mean, rstd, h, w, S, L,
                                                                                       static_cast<rb_t*>(dx), dgamma, dbeta);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_build_pos_mask(const void* img_mask, int B, int H, int W, int h, int w, const long long* sent_mask, int L, const float* lang_pos,
                                 const float* token_type, const float* level_embed, float* pos32, void* kpm, void* stream) {
  const size_t smem = ((static_cast<size_t>(h) * w + 15) & ~static_cast<size_t>(15)) + 2 * static_cast<size_t>(h) * w * sizeof(float);
  if (smem > 48 * 1024) return rb_fail("rb_build_pos_mask: %d x %d tokens exceed the shared-memory plan", h, w);
  build_pos_mask_kernel<<<dim3(B, 10), LN_D, smem, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(img_mask), H, W, h, w, sent_mask, L, lang_pos,
                                                                            token_type, level_embed, pos32, static_cast<uint8_t*>(kpm));
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_embed_grad(const float* dpos, int B, int S, int L, float* d_lang_pos, float* d_token_type, float* d_level, void* stream) {
  embed_grad_kernel<<<L + 32, LN_D, 0, static_cast<cudaStream_t>(stream)>>>(dpos, B, S, L, d_lang_pos, d_token_type, d_level);
  RB_CUDA(cudaGetLastError());
  return 0;
}
