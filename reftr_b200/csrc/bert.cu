// BERT-base support kernels (the language backbone on the same kernel family; call sites
// reftr_transformer.py:200, :217).  The GEMMs of BERT run on rb_gemm; this file holds what is specific to it:
// embedding gather + gradient scatter, LayerNorm over 768-wide rows, exact (erf) GELU, tanh, and the 64-wide-head
// attention over short token sequences (<= 128 tokens) with a key-padding mask.
#include "common.cuh"
#include "host.h"

namespace rb {

// ------------------------------------------------------------------------------------------------ embeddings
// out[r, :] = word[ids[r]] + pos[r % L] + type[0]       (token_type_ids = 0, position_ids = arange(L): HF defaults)
__global__ void bert_embed_fwd_kernel(const long long* __restrict__ ids, int L, int D, const float* __restrict__ word, const float* __restrict__ pos,
                                      const float* __restrict__ type0, float* __restrict__ out, long long rows) {
  const long long r = blockIdx.x;
  if (r >= rows) return;
  const long long id = ids[r];
  const int t = static_cast<int>(r % L);
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    const float4 a = *reinterpret_cast<const float4*>(word + id * D + c);
    const float4 b = *reinterpret_cast<const float4*>(pos + static_cast<long long>(t) * D + c);
    const float4 e = *reinterpret_cast<const float4*>(type0 + c);
    *reinterpret_cast<float4*>(out + r * D + c) = make_float4(a.x + b.x + e.x, a.y + b.y + e.y, a.z + b.z + e.z, a.w + b.w + e.w);
  }
}

__global__ void bert_embed_bwd_kernel(const float* __restrict__ d, const long long* __restrict__ ids, int L, int D, float* __restrict__ dword,
                                      float* __restrict__ dpos, float* __restrict__ dtype0, long long rows) {
  const long long r = blockIdx.x;
  if (r >= rows) return;
  const long long id = ids[r];
  const int t = static_cast<int>(r % L);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float v = d[r * D + c];
    if (dword) atomicAdd(dword + id * D + c, v);
    if (dpos) atomicAdd(dpos + static_cast<long long>(t) * D + c, v);
    if (dtype0) atomicAdd(dtype0 + c, v);
  }
}

// dropout on 4 consecutive elements starting at an even column: word counters c0, c0 + 1
__device__ __forceinline__ void drop4(float4& o, uint32_t key, const DropK& d, uint32_t c0) {
  const uint32_t w0 = drop_word(key, c0), w1 = drop_word(key, c0 + 1);
  o.x = drop_keep(w0, 0, d.thr) ? o.x * d.scale : 0.f;
  o.y = drop_keep(w0, 1, d.thr) ? o.y * d.scale : 0.f;
  o.z = drop_keep(w1, 0, d.thr) ? o.z * d.scale : 0.f;
  o.w = drop_keep(w1, 1, d.thr) ? o.w * d.scale : 0.f;
}

// ------------------------------------------------------------------------------------------------ LayerNorm, D = 32*4*NV
template <int NV>
__global__ void __launch_bounds__(256)
ln_wide_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, long long rows, float eps,
                   float* __restrict__ y32, rb_t* __restrict__ yb, float* __restrict__ mean_out, float* __restrict__ rstd_out, DropK drop) {
  constexpr int D = NV * 128;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = *reinterpret_cast<const float4*>(x + row * D + i * 128 + lane * 4);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const uint32_t dkey = drop.seed ? drop_key(drop) : 0u;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = *reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4);
    const float4 b = *reinterpret_cast<const float4*>(beta + i * 128 + lane * 4);
    float4 o = make_float4(v[i].x * rstd * g.x + b.x, v[i].y * rstd * g.y + b.y, v[i].z * rstd * g.z + b.z, v[i].w * rstd * g.w + b.w);
    if (drop.seed) drop4(o, dkey, drop, static_cast<uint32_t>(row) * (D / 2) + i * 64 + lane * 2);
    if (y32) *reinterpret_cast<float4*>(y32 + row * D + i * 128 + lane * 4) = o;
    if (yb) {
      uint2 p;
      p.x = pack_t2(o.x, o.y);
      p.y = pack_t2(o.z, o.w);
      *reinterpret_cast<uint2*>(yb + row * D + i * 128 + lane * 4) = p;
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
ln_wide_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dy2, const float* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ mean, const float* __restrict__ rstd, long long rows, float* __restrict__ dx32, rb_t* __restrict__ dxb,
                   float* __restrict__ dgamma, float* __restrict__ dbeta, DropK idrop, DropK odrop) {
  constexpr int D = NV * 128;
  __shared__ float red[2][D];
  const uint32_t ikey = idrop.seed ? drop_key(idrop) : 0u, okey = odrop.seed ? drop_key(odrop) : 0u;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long warp_global = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  const long long n_warps = static_cast<long long>(gridDim.x) * warps_per_block;
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) (&red[0][0])[c] = 0.f;
  __syncthreads();
  float4 dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = warp_global; row < rows; row += n_warps) {
    const float m = mean[row], rs = rstd[row];
    float4 d[NV], xh[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      d[i] = *reinterpret_cast<const float4*>(dy + row * D + i * 128 + lane * 4);
      if (dy2) {
        const float4 e = *reinterpret_cast<const float4*>(dy2 + row * D + i * 128 + lane * 4);
        d[i].x += e.x; d[i].y += e.y; d[i].z += e.z; d[i].w += e.w;
      }
      if (idrop.seed) drop4(d[i], ikey, idrop, static_cast<uint32_t>(row) * (D / 2) + i * 64 + lane * 2);
      const float4 xv = *reinterpret_cast<const float4*>(x + row * D + i * 128 + lane * 4);
      xh[i] = make_float4((xv.x - m) * rs, (xv.y - m) * rs, (xv.z - m) * rs, (xv.w - m) * rs);
      dg[i].x += d[i].x * xh[i].x; dg[i].y += d[i].y * xh[i].y; dg[i].z += d[i].z * xh[i].z; dg[i].w += d[i].w * xh[i].w;
      db[i].x += d[i].x; db[i].y += d[i].y; db[i].z += d[i].z; db[i].w += d[i].w;
      const float4 g = *reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4);
      d[i].x *= g.x; d[i].y *= g.y; d[i].z *= g.z; d[i].w *= g.w;
      s1 += d[i].x + d[i].y + d[i].z + d[i].w;
      s2 += d[i].x * xh[i].x + d[i].y * xh[i].y + d[i].z * xh[i].z + d[i].w * xh[i].w;
    }
    s1 = warp_sum(s1) * (1.f / D);
    s2 = warp_sum(s2) * (1.f / D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o = make_float4(rs * (d[i].x - s1 - xh[i].x * s2), rs * (d[i].y - s1 - xh[i].y * s2), rs * (d[i].z - s1 - xh[i].z * s2),
                             rs * (d[i].w - s1 - xh[i].w * s2));
      if (dx32) *reinterpret_cast<float4*>(dx32 + row * D + i * 128 + lane * 4) = o;
      if (dxb) {
        if (odrop.seed) drop4(o, okey, odrop, static_cast<uint32_t>(row) * (D / 2) + i * 64 + lane * 2);
        uint2 p;
        p.x = pack_t2(o.x, o.y);
        p.y = pack_t2(o.z, o.w);
        *reinterpret_cast<uint2*>(dxb + row * D + i * 128 + lane * 4) = p;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    atomicAdd(&red[0][c], dg[i].x); atomicAdd(&red[0][c + 1], dg[i].y); atomicAdd(&red[0][c + 2], dg[i].z); atomicAdd(&red[0][c + 3], dg[i].w);
    atomicAdd(&red[1][c], db[i].x); atomicAdd(&red[1][c + 1], db[i].y); atomicAdd(&red[1][c + 2], db[i].z); atomicAdd(&red[1][c + 3], db[i].w);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + c, red[0][c]);
    if (dbeta) atomicAdd(dbeta + c, red[1][c]);
  }
}

// ------------------------------------------------------------------------------------------------ GELU (erf) / tanh
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.7071067811865476f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.f + erff(x * 0.7071067811865476f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

__global__ void gelu_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 v = x[i];
  uint4 o;
  o.x = pack_t2(gelu_f(t_lo(v.x)), gelu_f(t_hi(v.x)));
  o.y = pack_t2(gelu_f(t_lo(v.y)), gelu_f(t_hi(v.y)));
  o.z = pack_t2(gelu_f(t_lo(v.z)), gelu_f(t_hi(v.z)));
  o.w = pack_t2(gelu_f(t_lo(v.w)), gelu_f(t_hi(v.w)));
  y[i] = o;
}

__global__ void gelu_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, uint4* __restrict__ dx, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 v = x[i], d = dy[i];
  uint4 o;
  o.x = pack_t2(t_lo(d.x) * gelu_grad_f(t_lo(v.x)), t_hi(d.x) * gelu_grad_f(t_hi(v.x)));
  o.y = pack_t2(t_lo(d.y) * gelu_grad_f(t_lo(v.y)), t_hi(d.y) * gelu_grad_f(t_hi(v.y)));
  o.z = pack_t2(t_lo(d.z) * gelu_grad_f(t_lo(v.z)), t_hi(d.z) * gelu_grad_f(t_hi(v.z)));
  o.w = pack_t2(t_lo(d.w) * gelu_grad_f(t_lo(v.w)), t_hi(d.w) * gelu_grad_f(t_hi(v.w)));
  dx[i] = o;
}

__global__ void tanh_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = tanhf(x[i]);
}

__global__ void tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, rb_t* __restrict__ dxb, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = dy[i] * (1.f - y[i] * y[i]);
  if (dx) dx[i] = v;
  if (dxb) dxb[i] = f2t(v);
}

// ------------------------------------------------------------------------------------------------ attention, head_dim 64, S <= 128
// one CTA per (batch, head); Q, K, V (and dO) staged in shared memory as fp32 [S][65]; P fp32 [B,H,S,S] is saved for the backward.
constexpr int AS_DH = 64;
constexpr int AS_PAD = 65;

__global__ void __launch_bounds__(128)
attn_small_fwd_kernel(const rb_t* __restrict__ Q, const rb_t* __restrict__ K, const rb_t* __restrict__ V,
                      const uint8_t* __restrict__ mask, rb_t* __restrict__ O, float* __restrict__ P, int H, int S, long long ldq, long long ldk,
                      long long ldv, long long ldo, float scale, DropK drop) {
  extern __shared__ float sm[];
  float* q = sm;
  float* k = q + S * AS_PAD;
  float* v = k + S * AS_PAD;
  float* p = v + S * AS_PAD;  // [S][S]
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  for (int i = threadIdx.x; i < S * AS_DH; i += blockDim.x) {
    const int r = i / AS_DH, c = i - r * AS_DH;
    const long long row = static_cast<long long>(b) * S + r;
    q[r * AS_PAD + c] = t2f(Q[row * ldq + h * AS_DH + c]) * scale;
    k[r * AS_PAD + c] = t2f(K[row * ldk + h * AS_DH + c]);
    v[r * AS_PAD + c] = t2f(V[row * ldv + h * AS_DH + c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
    const int r = i / S, c = i - r * S;
    float acc = 0.f;
#pragma unroll 16
    for (int d = 0; d < AS_DH; ++d) acc = fmaf(q[r * AS_PAD + d], k[c * AS_PAD + d], acc);
    p[i] = (mask && mask[static_cast<long long>(b) * S + c]) ? -INFINITY : acc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t dkey = drop.seed ? drop_key(drop) : 0u;
  for (int r = warp; r < S; r += blockDim.x >> 5) {
    float mx = -INFINITY;
    for (int c = lane; c < S; c += 32) mx = fmaxf(mx, p[r * S + c]);
    mx = warp_max(mx);
    if (mx == -INFINITY) mx = 0.f;
    float sum = 0.f;
    for (int c = lane; c < S; c += 32) {
      const float e = __expf(p[r * S + c] - mx);
      p[r * S + c] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    const uint32_t wrow = (static_cast<uint32_t>(blockIdx.x) * S + r) * ((S + 1) >> 1);
    for (int c = lane; c < S; c += 32) {
      const float a = p[r * S + c] * inv;
      P[(static_cast<long long>(blockIdx.x) * S + r) * S + c] = a;  // the UNDROPPED softmax is saved; the backward recomputes the mask
      p[r * S + c] = (!drop.seed) ? a : (drop_keep(drop_word(dkey, wrow + (c >> 1)), c & 1, drop.thr) ? a * drop.scale : 0.f);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * AS_DH; i += blockDim.x) {
    const int r = i / AS_DH, c = i - r * AS_DH;
    float acc = 0.f;
    for (int j = 0; j < S; ++j) acc = fmaf(p[r * S + j], v[j * AS_PAD + c], acc);
    O[(static_cast<long long>(b) * S + r) * ldo + h * AS_DH + c] = f2t(acc);
  }
}

__global__ void __launch_bounds__(128)
attn_small_bwd_kernel(const rb_t* __restrict__ Q, const rb_t* __restrict__ K, const rb_t* __restrict__ V,
                      const rb_t* __restrict__ dO, const float* __restrict__ P, rb_t* __restrict__ dQ, rb_t* __restrict__ dK,
                      rb_t* __restrict__ dV, int H, int S, long long ldq, long long ldk, long long ldv, long long lddo, long long lddq,
                      long long lddk, long long lddv, float scale, DropK drop) {
  extern __shared__ float sm[];
  float* q = sm;
  float* k = q + S * AS_PAD;
  float* v = k + S * AS_PAD;
  float* go = v + S * AS_PAD;
  float* p = go + S * AS_PAD;  // [S][S] probabilities
  float* ds = p + S * S;       // [S][S] d(scores)
  float* pm = drop.seed ? ds + S * S : p;  // [S][S] dropped probabilities (aliases p without dropout: same thread, read before write)
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  for (int i = threadIdx.x; i < S * AS_DH; i += blockDim.x) {
    const int r = i / AS_DH, c = i - r * AS_DH;
    const long long row = static_cast<long long>(b) * S + r;
    q[r * AS_PAD + c] = t2f(Q[row * ldq + h * AS_DH + c]);
    k[r * AS_PAD + c] = t2f(K[row * ldk + h * AS_DH + c]);
    v[r * AS_PAD + c] = t2f(V[row * ldv + h * AS_DH + c]);
    go[r * AS_PAD + c] = t2f(dO[row * lddo + h * AS_DH + c]);
  }
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) p[i] = P[static_cast<long long>(blockIdx.x) * S * S + i];
  __syncthreads();
  // dP = dO V^T
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
    const int r = i / S, c = i - r * S;
    float acc = 0.f;
#pragma unroll 16
    for (int d = 0; d < AS_DH; ++d) acc = fmaf(go[r * AS_PAD + d], v[c * AS_PAD + d], acc);
    ds[i] = acc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t dkey = drop.seed ? drop_key(drop) : 0u;
  for (int r = warp; r < S; r += blockDim.x >> 5) {
    // with dropout: O = (P .* M) V, M = keep * scale.  dP = (dO V^T) .* M, dS = P .* (dP - rowsum(P .* dP)); dV uses P .* M
    const uint32_t wrow = (static_cast<uint32_t>(blockIdx.x) * S + r) * ((S + 1) >> 1);
    float dot = 0.f;
    for (int c = lane; c < S; c += 32) {
      float mk = 1.f;
      if (drop.seed) mk = drop_keep(drop_word(dkey, wrow + (c >> 1)), c & 1, drop.thr) ? drop.scale : 0.f;
      const float dp = ds[r * S + c] * mk;
      ds[r * S + c] = dp;
      dot += dp * p[r * S + c];
      pm[r * S + c] = p[r * S + c] * mk;
    }
    dot = warp_sum(dot);
    for (int c = lane; c < S; c += 32) ds[r * S + c] = p[r * S + c] * (ds[r * S + c] - dot) * scale;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * AS_DH; i += blockDim.x) {
    const int r = i / AS_DH, c = i - r * AS_DH;
    float aq = 0.f, ak = 0.f, av = 0.f;
    for (int j = 0; j < S; ++j) {
      aq = fmaf(ds[r * S + j], k[j * AS_PAD + c], aq);   // dQ[r] = sum_j dS[r,j] K[j]
      ak = fmaf(ds[j * S + r], q[j * AS_PAD + c], ak);   // dK[r] = sum_j dS[j,r] Q[j]
      av = fmaf(pm[j * S + r], go[j * AS_PAD + c], av);  // dV[r] = sum_j (P .* M)[j,r] dO[j]
    }
    const long long row = static_cast<long long>(b) * S + r;
    dQ[row * lddq + h * AS_DH + c] = f2t(aq);
    dK[row * lddk + h * AS_DH + c] = f2t(ak);
    dV[row * lddv + h * AS_DH + c] = f2t(av);
  }
}

}  // namespace rb

using namespace rb;

static unsigned nblk(long long total, int threads) { return static_cast<unsigned>((total + threads - 1) / threads); }

extern "C" int rb_bert_embed_fwd(const long long* ids, long long rows, int L, int D, const float* word, const float* pos, const float* type0, float* out,
                                 void* stream) {
  if (D % 4) return rb_fail("rb_bert_embed_fwd: D %% 4 != 0");
  if (rows <= 0) return 0;
  bert_embed_fwd_kernel<<<static_cast<unsigned>(rows), 192, 0, static_cast<cudaStream_t>(stream)>>>(ids, L, D, word, pos, type0, out, rows);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_bert_embed_bwd(const float* d, const long long* ids, long long rows, int L, int D, float* dword, float* dpos, float* dtype0, void* stream) {
  if (rows <= 0) return 0;
  bert_embed_bwd_kernel<<<static_cast<unsigned>(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(d, ids, L, D, dword, dpos, dtype0, rows);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_ln_wide_fwd(const float* x, const float* gamma, const float* beta, long long rows, int D, float eps, float* y32, void* yb, float* mean,
                              float* rstd, const rb_dropout* drop, void* stream) {
  if (rows <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  if (D == 768) ln_wide_fwd_kernel<6><<<grid, 256, 0, st>>>(x, gamma, beta, rows, eps, y32, static_cast<rb_t*>(yb), mean, rstd, make_dropk(drop));
  else if (D == 1024) ln_wide_fwd_kernel<8><<<grid, 256, 0, st>>>(x, gamma, beta, rows, eps, y32, static_cast<rb_t*>(yb), mean, rstd, make_dropk(drop));
  else return rb_fail("rb_ln_wide_fwd: D must be 768 or 1024 (got %d)", D);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_ln_wide_bwd(const float* dy, const float* dy2, const float* x, const float* gamma, const float* mean, const float* rstd, long long rows,
                              int D, float* dx32, void* dxb, float* dgamma, float* dbeta, const rb_dropout* dy_drop, const rb_dropout* dxb_drop, void* stream) {
  if (rows <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long blocks = (rows + 7) / 8;
  if (blocks > 148) blocks = 148;
  const unsigned grid = static_cast<unsigned>(blocks);
  if (D == 768) ln_wide_bwd_kernel<6><<<grid, 256, 0, st>>>(dy, dy2, x, gamma, mean, rstd, rows, dx32, static_cast<rb_t*>(dxb), dgamma, dbeta,
                                                                  make_dropk(dy_drop), make_dropk(dxb_drop));
  else if (D == 1024) ln_wide_bwd_kernel<8><<<grid, 256, 0, st>>>(dy, dy2, x, gamma, mean, rstd, rows, dx32, static_cast<rb_t*>(dxb), dgamma, dbeta,
                                                                   make_dropk(dy_drop), make_dropk(dxb_drop));
  else return rb_fail("rb_ln_wide_bwd: D must be 768 or 1024 (got %d)", D);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_gelu_fwd(const void* x, void* y, long long n, void* stream) {
  if (n % 8) return rb_fail("rb_gelu_fwd: n %% 8 != 0");
  if (n <= 0) return 0;
  gelu_fwd_kernel<<<nblk(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(x), static_cast<uint4*>(y), n / 8);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_gelu_bwd(const void* dy, const void* x, void* dx, long long n, void* stream) {
  if (n % 8) return rb_fail("rb_gelu_bwd: n %% 8 != 0");
  if (n <= 0) return 0;
  gelu_bwd_kernel<<<nblk(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(dy), static_cast<const uint4*>(x),
                                                                                  static_cast<uint4*>(dx), n / 8);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_tanh_fwd(const float* x, float* y, long long n, void* stream) {
  if (n <= 0) return 0;
  tanh_fwd_kernel<<<nblk(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_tanh_bwd(const float* dy, const float* y, float* dx, void* dxb, long long n, void* stream) {
  if (n <= 0) return 0;
  tanh_bwd_kernel<<<nblk(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, y, dx, static_cast<rb_t*>(dxb), n);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_attn_small_fwd(const void* Q, const void* K, const void* V, const void* mask, void* O, float* P, int B, int H, int dh, int S, long long ldq,
                                 long long ldk, long long ldv, long long ldo, float scale, const rb_dropout* drop, void* stream) {
  if (dh != AS_DH) return rb_fail("rb_attn_small_fwd: head_dim must be 64 (got %d)", dh);
  if (S < 1 || S > 128) return rb_fail("rb_attn_small_fwd: 1 <= S <= 128 tokens (got %d)", S);
  const size_t smem = (static_cast<size_t>(3) * S * AS_PAD + static_cast<size_t>(S) * S) * sizeof(float);
  static bool cfg = false;
  if (!cfg) { RB_CUDA(cudaFuncSetAttribute(attn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); cfg = true; }
  attn_small_fwd_kernel<<<B * H, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const rb_t*>(Q), static_cast<const rb_t*>(K), static_cast<const rb_t*>(V), static_cast<const uint8_t*>(mask),
      static_cast<rb_t*>(O), P, H, S, ldq, ldk, ldv, ldo, scale, make_dropk(drop));
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_attn_small_bwd(const void* Q, const void* K, const void* V, const void* dO, const float* P, void* dQ, void* dK, void* dV, int B, int H,
                                 int dh, int S, long long ldq, long long ldk, long long ldv, long long lddo, long long lddq, long long lddk, long long lddv,
                                 float scale, const rb_dropout* drop, void* stream) {
  if (dh != AS_DH) return rb_fail("rb_attn_small_bwd: head_dim must be 64 (got %d)", dh);
  if (S < 1 || S > 128) return rb_fail("rb_attn_small_bwd: 1 <= S <= 128 tokens (got %d)", S);
  const DropK dk = make_dropk(drop);
  const size_t smem = (static_cast<size_t>(4) * S * AS_PAD + static_cast<size_t>(dk.seed ? 3 : 2) * S * S) * sizeof(float);
  if (smem > 220 * 1024) return rb_fail("rb_attn_small_bwd: S = %d exceeds the shared-memory plan", S);
  static bool cfg = false;
  if (!cfg) { RB_CUDA(cudaFuncSetAttribute(attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); cfg = true; }
  attn_small_bwd_kernel<<<B * H, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const rb_t*>(Q), static_cast<const rb_t*>(K), static_cast<const rb_t*>(V), static_cast<const rb_t*>(dO), P,
      static_cast<rb_t*>(dQ), static_cast<rb_t*>(dK), static_cast<rb_t*>(dV), H, S, ldq, ldk, ldv, lddo, lddq, lddk, lddv, scale, dk);
  RB_CUDA(cudaGetLastError());
  return 0;
}
