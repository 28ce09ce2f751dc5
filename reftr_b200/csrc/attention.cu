// Multi-head attention core (head_dim 32), forward and backward, with key-padding mask -- round-1 SIMT version.
// Replaces the bmm/softmax/bmm of torch's F.multi_head_attention_forward as invoked at transformer.py:174, :239,
// :243 (scores are never materialised in HBM; the reference writes [B*h,S,S] fp32 per layer).
//
// Layout: Q [B*Tq, ldq] bf16 with head h at columns [32h, 32h+32); K, V [B*Sk, ld] likewise; O like Q.
// One CTA per (b, h, chunk of 64 queries): K^T (bf16x2-packed along d) and V live in shared memory, each warp owns
// 2 queries at a time, lanes own keys for QK^T (scores in registers) and own the output dim for PV.
// (The tcgen05 version of this kernel is the next step; see DESIGN.md.)
#include "common.cuh"
#include "host.h"

namespace rb {

constexpr int DH = 32;
constexpr int NQ = 2;

// s[q][c] = sum_d vec[q][d] * T2[d/2][lane + 32c]  (T2 packs dims (2d', 2d'+1) of row j into one 32-bit word)
template <int SCH>
__device__ __forceinline__ void dot_phase(const float* __restrict__ vec, const uint32_t* __restrict__ T2, int SP, int lane, float (&s)[NQ][SCH]) {
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int c = 0; c < SCH; ++c) s[q][c] = 0.f;
#pragma unroll 4
  for (int dp = 0; dp < DH / 2; ++dp) {
    float2 qv[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) qv[q] = *reinterpret_cast<const float2*>(vec + q * DH + 2 * dp);
    const uint32_t* row = T2 + dp * SP + lane;
#pragma unroll
    for (int c = 0; c < SCH; ++c) {
      const uint32_t kk = row[32 * c];
      const float k0 = t_lo(kk), k1 = t_hi(kk);
#pragma unroll
      for (int q = 0; q < NQ; ++q) s[q][c] = fmaf(qv[q].x, k0, fmaf(qv[q].y, k1, s[q][c]));
    }
  }
}

// acc[q] = sum_j w[q][j] * M[j][lane]   (M row-major bf16 [n][32], w fp32 in shared memory, n multiple of 4)
__device__ __forceinline__ void mix_phase(const float* __restrict__ w, int SP, const rb_t* __restrict__ M, int n, int lane, float (&acc)[NQ]) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = 0.f;
  for (int j = 0; j < n; j += 4) {
    float4 p[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) p[q] = *reinterpret_cast<const float4*>(w + q * SP + j);
    const float v0 = t2f(M[(j + 0) * DH + lane]), v1 = t2f(M[(j + 1) * DH + lane]);
    const float v2 = t2f(M[(j + 2) * DH + lane]), v3 = t2f(M[(j + 3) * DH + lane]);
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = fmaf(p[q].x, v0, fmaf(p[q].y, v1, fmaf(p[q].z, v2, fmaf(p[q].w, v3, acc[q]))));
  }
}

// Stage rows [0,n) of a [*, ld] bf16 matrix (columns col0..col0+31) into shared memory: row-major copy `rm` ([SP][32] bf16)
// and/or transposed-packed copy `t2` ([16][SP] uint32).  Rows >= n are zero-filled up to SP.
__device__ __forceinline__ void stage_rows(const rb_t* __restrict__ g, long long ld, int n, int SP, rb_t* rm, uint32_t* t2) {
  for (int idx = threadIdx.x; idx < SP * 4; idx += blockDim.x) {
    const int j = idx >> 2, part = idx & 3;  // 4 x 16B per row
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < n) v = *reinterpret_cast<const uint4*>(g + static_cast<long long>(j) * ld + part * 8);
    if (rm) *reinterpret_cast<uint4*>(rm + j * DH + part * 8) = v;
    if (t2) {
      t2[(part * 4 + 0) * SP + j] = v.x; t2[(part * 4 + 1) * SP + j] = v.y;
      t2[(part * 4 + 2) * SP + j] = v.z; t2[(part * 4 + 3) * SP + j] = v.w;
    }
  }
}

template <int SCH, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
attn_fwd_kernel(const rb_t* __restrict__ Q, const rb_t* __restrict__ K, const rb_t* __restrict__ V,
                const uint8_t* __restrict__ kpm, rb_t* __restrict__ O, float* __restrict__ LSE, int H, int Tq, int Sk, long long ldq,
                long long ldk, long long ldv, long long ldo, float scale, int q_per_block, DropK drop) {
  constexpr int SP = SCH * 32;
  const uint32_t dkey = drop.seed ? drop_key(drop) : 0u;
  const uint32_t wpr = static_cast<uint32_t>((Sk + 1) >> 1);
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* Kt2 = reinterpret_cast<uint32_t*>(smem);                          // [16][SP]
  rb_t* Vs = reinterpret_cast<rb_t*>(Kt2 + 16 * SP);        // [SP][32]
  float* ps = reinterpret_cast<float*>(Vs + SP * DH);                         // [NWARPS][NQ][SP]
  float* qs = ps + NWARPS * NQ * SP;                                          // [NWARPS][NQ][32]
  uint8_t* msk = reinterpret_cast<uint8_t*>(qs + NWARPS * NQ * DH);           // [SP]
  const int b = blockIdx.y / H, h = blockIdx.y - b * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows(K + static_cast<long long>(b) * Sk * ldk + h * DH, ldk, Sk, SP, nullptr, Kt2);
  stage_rows(V + static_cast<long long>(b) * Sk * ldv + h * DH, ldv, Sk, SP, Vs, nullptr);
  for (int j = threadIdx.x; j < SP; j += blockDim.x) msk[j] = (j >= Sk) || (kpm && kpm[static_cast<long long>(b) * Sk + j]);
  __syncthreads();
  float* myps = ps + warp * NQ * SP;
  float* myq = qs + warp * NQ * DH;
  const int q_begin = blockIdx.x * q_per_block, q_end = min(Tq, q_begin + q_per_block);
  const int n4 = (Sk + 3) & ~3;
  for (int t0 = q_begin + warp * NQ; t0 < q_end; t0 += NWARPS * NQ) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int t = t0 + q;
      myq[q * DH + lane] = t < q_end ? t2f(Q[(static_cast<long long>(b) * Tq + t) * ldq + h * DH + lane]) * scale : 0.f;
    }
    __syncwarp();
    float s[NQ][SCH];
    dot_phase<SCH>(myq, Kt2, SP, lane, s);
    float mx[NQ], sum[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float m = -INFINITY;
#pragma unroll
      for (int c = 0; c < SCH; ++c) {
        if (msk[lane + 32 * c]) s[q][c] = -INFINITY;
        m = fmaxf(m, s[q][c]);
      }
      m = warp_max(m);
      if (m == -INFINITY) m = 0.f;  // every key masked: the reference yields NaN; we yield zeros
      float acc = 0.f;
      const uint32_t wrow = (static_cast<uint32_t>(blockIdx.y) * Tq + (t0 + q)) * wpr;
#pragma unroll
      for (int c = 0; c < SCH; ++c) {
        const float e = __expf(s[q][c] - m);
        acc += e;
        const int j = lane + 32 * c;
        // dropout on the probabilities: the row sum stays that of the undropped softmax; 1/(1-p) is folded into `inv` below
        myps[q * SP + j] = (drop.seed && !drop_keep(drop_word(dkey, wrow + (j >> 1)), j & 1, drop.thr)) ? 0.f : e;
      }
      mx[q] = m;
      sum[q] = warp_sum(acc);
    }
    __syncwarp();
    float acc[NQ];
    mix_phase(myps, SP, Vs, n4, lane, acc);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int t = t0 + q;
      if (t < q_end) {
        const float inv = sum[q] > 0.f ? drop.scale / sum[q] : 0.f;
        O[(static_cast<long long>(b) * Tq + t) * ldo + h * DH + lane] = f2t(acc[q] * inv);
        if (lane == 0 && LSE) LSE[(static_cast<long long>(b) * H + h) * Tq + t] = mx[q] + __logf(fmaxf(sum[q], 1e-30f));
      }
    }
    __syncwarp();
  }
}

// dQ (and D = rowsum(dO*O)) -- same loop structure as the forward pass.
template <int SCH, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
attn_bwd_dq_kernel(const rb_t* __restrict__ Q, const rb_t* __restrict__ K, const rb_t* __restrict__ V,
                   const uint8_t* __restrict__ kpm, const rb_t* __restrict__ O, const rb_t* __restrict__ dO,
                   const float* __restrict__ LSE, rb_t* __restrict__ dQ, float* __restrict__ Dbuf, int H, int Tq, int Sk, long long ldq,
                   long long ldk, long long ldv, long long ldo, long long lddo, long long lddq, float scale, int q_per_block, DropK drop) {
  constexpr int SP = SCH * 32;
  const uint32_t dkey = drop.seed ? drop_key(drop) : 0u;
  const uint32_t wpr = static_cast<uint32_t>((Sk + 1) >> 1);
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* Kt2 = reinterpret_cast<uint32_t*>(smem);
  uint32_t* Vt2 = Kt2 + 16 * SP;
  rb_t* Ks = reinterpret_cast<rb_t*>(Vt2 + 16 * SP);
  float* ps = reinterpret_cast<float*>(Ks + SP * DH);
  float* qs = ps + NWARPS * NQ * SP;        // [NWARPS][2][NQ][32]: q then dO
  uint8_t* msk = reinterpret_cast<uint8_t*>(qs + NWARPS * 2 * NQ * DH);
  const int b = blockIdx.y / H, h = blockIdx.y - b * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows(K + static_cast<long long>(b) * Sk * ldk + h * DH, ldk, Sk, SP, Ks, Kt2);
  stage_rows(V + static_cast<long long>(b) * Sk * ldv + h * DH, ldv, Sk, SP, nullptr, Vt2);
  for (int j = threadIdx.x; j < SP; j += blockDim.x) msk[j] = (j >= Sk) || (kpm && kpm[static_cast<long long>(b) * Sk + j]);
  __syncthreads();
  float* myps = ps + warp * NQ * SP;
  float* myq = qs + warp * 2 * NQ * DH;
  float* mydo = myq + NQ * DH;
  const int q_begin = blockIdx.x * q_per_block, q_end = min(Tq, q_begin + q_per_block);
  const int n4 = (Sk + 3) & ~3;
  for (int t0 = q_begin + warp * NQ; t0 < q_end; t0 += NWARPS * NQ) {
    float Dq[NQ], lse[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int t = t0 + q;
      float qv = 0.f, dov = 0.f, ov = 0.f;
      if (t < q_end) {
        const long long r = static_cast<long long>(b) * Tq + t;
        qv = t2f(Q[r * ldq + h * DH + lane]) * scale;
        dov = t2f(dO[r * lddo + h * DH + lane]);
        ov = t2f(O[r * ldo + h * DH + lane]);
        lse[q] = LSE[(static_cast<long long>(b) * H + h) * Tq + t];
      } else {
        lse[q] = 0.f;
      }
      myq[q * DH + lane] = qv;
      mydo[q * DH + lane] = dov;
      Dq[q] = warp_sum(dov * ov);
      if (t < q_end && lane == 0) Dbuf[(static_cast<long long>(b) * H + h) * Tq + t] = Dq[q];
    }
    __syncwarp();
    float s[NQ][SCH], dp[NQ][SCH];
    dot_phase<SCH>(myq, Kt2, SP, lane, s);
    dot_phase<SCH>(mydo, Vt2, SP, lane, dp);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const uint32_t wrow = (static_cast<uint32_t>(blockIdx.y) * Tq + (t0 + q)) * wpr;
#pragma unroll
      for (int c = 0; c < SCH; ++c) {
        const int j = lane + 32 * c;
        const float p = msk[j] ? 0.f : __expf(s[q][c] - lse[q]);
        float dpv = dp[q][c];  // gradient w.r.t. the DROPPED probabilities -> w.r.t. the softmax output
        if (drop.seed) dpv = drop_keep(drop_word(dkey, wrow + (j >> 1)), j & 1, drop.thr) ? dpv * drop.scale : 0.f;
        myps[q * SP + j] = p * (dpv - Dq[q]) * scale;
      }
    }
    __syncwarp();
    float acc[NQ];
    mix_phase(myps, SP, Ks, n4, lane, acc);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int t = t0 + q;
      if (t < q_end) dQ[(static_cast<long long>(b) * Tq + t) * lddq + h * DH + lane] = f2t(acc[q]);
    }
    __syncwarp();
  }
}

// dK, dV: roles swapped -- each warp owns 2 keys, loops over all queries of (b, h) staged in shared memory.
template <int TCH, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
attn_bwd_dkv_kernel(const rb_t* __restrict__ Q, const rb_t* __restrict__ K, const rb_t* __restrict__ V,
                    const uint8_t* __restrict__ kpm, const rb_t* __restrict__ dO, const float* __restrict__ LSE, const float* __restrict__ Dbuf,
                    rb_t* __restrict__ dK, rb_t* __restrict__ dV, int H, int Tq, int Sk, long long ldq, long long ldk, long long ldv,
                    long long lddo, long long lddk, long long lddv, float scale, int k_per_block, DropK drop) {
  constexpr int TP = TCH * 32;
  const uint32_t dkey = drop.seed ? drop_key(drop) : 0u;
  const uint32_t wpr = static_cast<uint32_t>((Sk + 1) >> 1);
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* Qt2 = reinterpret_cast<uint32_t*>(smem);
  uint32_t* dOt2 = Qt2 + 16 * TP;
  rb_t* Qs = reinterpret_cast<rb_t*>(dOt2 + 16 * TP);
  rb_t* dOs = Qs + TP * DH;
  float* ps = reinterpret_cast<float*>(dOs + TP * DH);  // [NWARPS][2][NQ][TP]
  float* ks = ps + NWARPS * 2 * NQ * TP;                // [NWARPS][2][NQ][32]
  float* lse_s = ks + NWARPS * 2 * NQ * DH;             // [TP]
  float* D_s = lse_s + TP;                              // [TP]
  const int b = blockIdx.y / H, h = blockIdx.y - b * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows(Q + static_cast<long long>(b) * Tq * ldq + h * DH, ldq, Tq, TP, Qs, Qt2);
  stage_rows(dO + static_cast<long long>(b) * Tq * lddo + h * DH, lddo, Tq, TP, dOs, dOt2);
  for (int t = threadIdx.x; t < TP; t += blockDim.x) {
    lse_s[t] = t < Tq ? LSE[(static_cast<long long>(b) * H + h) * Tq + t] : INFINITY;  // exp(s - inf) = 0 for padding queries
    D_s[t] = t < Tq ? Dbuf[(static_cast<long long>(b) * H + h) * Tq + t] : 0.f;
  }
  __syncthreads();
  float* myp = ps + warp * 2 * NQ * TP;
  float* myds = myp + NQ * TP;
  float* myk = ks + warp * 2 * NQ * DH;
  float* myv = myk + NQ * DH;
  const int k_begin = blockIdx.x * k_per_block, k_end = min(Sk, k_begin + k_per_block);
  const int n4 = (Tq + 3) & ~3;
  for (int j0 = k_begin + warp * NQ; j0 < k_end; j0 += NWARPS * NQ) {
    bool dead[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int j = j0 + q;
      dead[q] = (j >= k_end) || (kpm && kpm[static_cast<long long>(b) * Sk + j]);
      float kv = 0.f, vv = 0.f;
      if (j < k_end) {
        kv = t2f(K[(static_cast<long long>(b) * Sk + j) * ldk + h * DH + lane]) * scale;
        vv = t2f(V[(static_cast<long long>(b) * Sk + j) * ldv + h * DH + lane]);
      }
      myk[q * DH + lane] = kv;
      myv[q * DH + lane] = vv;
    }
    __syncwarp();
    float s[NQ][TCH], dp[NQ][TCH];
    dot_phase<TCH>(myk, Qt2, TP, lane, s);
    dot_phase<TCH>(myv, dOt2, TP, lane, dp);
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int c = 0; c < TCH; ++c) {
        const int t = lane + 32 * c;
        const float p = dead[q] ? 0.f : __expf(s[q][c] - lse_s[t]);
        float mk = 1.f;
        if (drop.seed) {
          const int j = j0 + q;
          mk = drop_keep(drop_word(dkey, (static_cast<uint32_t>(blockIdx.y) * Tq + t) * wpr + (j >> 1)), j & 1, drop.thr) ? drop.scale : 0.f;
        }
        myp[q * TP + t] = p * mk;                                   // dV uses the dropped probabilities
        myds[q * TP + t] = p * (dp[q][c] * mk - D_s[t]) * scale;
      }
    __syncwarp();
    float accv[NQ], acck[NQ];
    mix_phase(myp, TP, dOs, n4, lane, accv);
    mix_phase(myds, TP, Qs, n4, lane, acck);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int j = j0 + q;
      if (j < k_end) {
        dV[(static_cast<long long>(b) * Sk + j) * lddv + h * DH + lane] = f2t(accv[q]);
        dK[(static_cast<long long>(b) * Sk + j) * lddk + h * DH + lane] = f2t(acck[q]);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ QueryEncoder pooling
// reftr_transformer.py:47-55: att[b,ph,l] = softmax_l(mask(k[b] . q[b,l])) (no 1/sqrt(d)); c[b,ph] = sum_l att * v[b,l].
// One CTA (256 threads = channels) per sample.  k: [B,256] fp32, q, v: [B*L,256] fp32, mask: [B,n_ph,L] u8 (1 = ignore).
__global__ void qenc_pool_fwd_kernel(const float* __restrict__ k, const float* __restrict__ q, const float* __restrict__ v, const uint8_t* __restrict__ mask,
                                     int L, int n_ph, float* __restrict__ att, float* __restrict__ c) {
  const int b = blockIdx.x, ch = threadIdx.x;
  extern __shared__ float sm[];
  float* sc = sm;           // [L] raw scores
  float* pa = sm + L;       // [L] probabilities for the current phrase
  __shared__ float red[8];
  const float kv = k[b * 256 + ch];
  for (int l = 0; l < L; ++l) {
    float p = warp_sum(kv * q[(static_cast<long long>(b) * L + l) * 256 + ch]);
    if ((ch & 31) == 0) red[ch >> 5] = p;
    __syncthreads();
    if (ch == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; sc[l] = t; }
    __syncthreads();
  }
  for (int ph = 0; ph < n_ph; ++ph) {
    const uint8_t* mk = mask + (static_cast<long long>(b) * n_ph + ph) * L;
    if (ch == 0) {
      float m = -INFINITY;
      for (int l = 0; l < L; ++l) if (!mk[l]) m = fmaxf(m, sc[l]);
      float s = 0.f;
      for (int l = 0; l < L; ++l) { const float e = mk[l] ? 0.f : __expf(sc[l] - m); pa[l] = e; s += e; }
      const float inv = 1.f / s;  // an all-masked row gives NaN exactly like the reference (softmax of all -inf)
      for (int l = 0; l < L; ++l) { pa[l] *= inv; att[(static_cast<long long>(b) * n_ph + ph) * L + l] = pa[l]; }
    }
    __syncthreads();
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc += pa[l] * v[(static_cast<long long>(b) * L + l) * 256 + ch];
    c[(static_cast<long long>(b) * n_ph + ph) * 256 + ch] = acc;
    __syncthreads();
  }
}

// Backward: dc [B*n_ph,256] -> dk [B,256], dq [B*L,256], dv [B*L,256] (all overwritten)
__global__ void qenc_pool_bwd_kernel(const float* __restrict__ dc, const float* __restrict__ k, const float* __restrict__ q, const float* __restrict__ v,
                                     const float* __restrict__ att, int L, int n_ph, float* __restrict__ dk, float* __restrict__ dq, float* __restrict__ dv) {
  const int b = blockIdx.x, ch = threadIdx.x;
  extern __shared__ float sm[];
  float* ds = sm;  // [n_ph][L]
  __shared__ float red[8];
  // datt[ph,l] = dc[ph] . v[l]
  for (int ph = 0; ph < n_ph; ++ph) {
    const float dcv = dc[(static_cast<long long>(b) * n_ph + ph) * 256 + ch];
    for (int l = 0; l < L; ++l) {
      float p = warp_sum(dcv * v[(static_cast<long long>(b) * L + l) * 256 + ch]);
      if ((ch & 31) == 0) red[ch >> 5] = p;
      __syncthreads();
      if (ch == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; ds[ph * L + l] = t; }
      __syncthreads();
    }
  }
  if (ch < n_ph) {  // softmax backward per phrase
    const float* a = att + (static_cast<long long>(b) * n_ph + ch) * L;
    float dot = 0.f;
    for (int l = 0; l < L; ++l) dot += a[l] * ds[ch * L + l];
    for (int l = 0; l < L; ++l) ds[ch * L + l] = a[l] * (ds[ch * L + l] - dot);
  }
  __syncthreads();
  const float kv = k[b * 256 + ch];
  float dkv = 0.f;
  for (int l = 0; l < L; ++l) {
    float dsl = 0.f, dvl = 0.f;
    for (int ph = 0; ph < n_ph; ++ph) {
      dsl += ds[ph * L + l];
      dvl += att[(static_cast<long long>(b) * n_ph + ph) * L + l] * dc[(static_cast<long long>(b) * n_ph + ph) * 256 + ch];
    }
    const long long o = (static_cast<long long>(b) * L + l) * 256 + ch;
    dq[o] = dsl * kv;
    dv[o] = dvl;
    dkv += dsl * q[o];
  }
  dk[b * 256 + ch] = dkv;
}

// ------------------------------------------------------------------------------------------------ host dispatch
template <int SCH>
static int launch_fwd(const void* Q, const void* K, const void* V, const void* kpm, void* O, float* LSE, int B, int H, int Tq, int Sk, long long ldq,
                      long long ldk, long long ldv, long long ldo, float scale, DropK drop, cudaStream_t st) {
  constexpr int NW = 8, SP = SCH * 32;
  constexpr int SMEM = 16 * SP * 4 + SP * DH * 2 + NW * NQ * SP * 4 + NW * NQ * DH * 4 + SP;
  auto kern = attn_fwd_kernel<SCH, NW>;
  static bool cfg = false;
  if (!cfg) { RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); cfg = true; }
  const int qpb = 64;
  kern<<<dim3((Tq + qpb - 1) / qpb, B * H), NW * 32, SMEM, st>>>(static_cast<const rb_t*>(Q), static_cast<const rb_t*>(K),
                                                              static_cast<const rb_t*>(V), static_cast<const uint8_t*>(kpm),
                                                              static_cast<rb_t*>(O), LSE, H, Tq, Sk, ldq, ldk, ldv, ldo, scale, qpb, drop);
  RB_CUDA(cudaGetLastError());
  return 0;
}

template <int SCH>
static int launch_dq(const void* Q, const void* K, const void* V, const void* kpm, const void* O, const void* dO, const float* LSE, void* dQ, float* Dbuf,
                     int B, int H, int Tq, int Sk, long long ldq, long long ldk, long long ldv, long long ldo, long long lddo, long long lddq, float scale,
                     DropK drop, cudaStream_t st) {
  constexpr int NW = 8, SP = SCH * 32;
  constexpr int SMEM = 2 * 16 * SP * 4 + SP * DH * 2 + NW * NQ * SP * 4 + NW * 2 * NQ * DH * 4 + SP;
  auto kern = attn_bwd_dq_kernel<SCH, NW>;
  static bool cfg = false;
  if (!cfg) { RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); cfg = true; }
  const int qpb = 64;
  kern<<<dim3((Tq + qpb - 1) / qpb, B * H), NW * 32, SMEM, st>>>(
      static_cast<const rb_t*>(Q), static_cast<const rb_t*>(K), static_cast<const rb_t*>(V), static_cast<const uint8_t*>(kpm),
      static_cast<const rb_t*>(O), static_cast<const rb_t*>(dO), LSE, static_cast<rb_t*>(dQ), Dbuf, H, Tq, Sk, ldq, ldk, ldv,
      ldo, lddo, lddq, scale, qpb, drop);
  RB_CUDA(cudaGetLastError());
  return 0;
}

template <int TCH, int NW>
static int launch_dkv(const void* Q, const void* K, const void* V, const void* kpm, const void* dO, const float* LSE, const float* Dbuf, void* dK, void* dV,
                      int B, int H, int Tq, int Sk, long long ldq, long long ldk, long long ldv, long long lddo, long long lddk, long long lddv, float scale,
                      DropK drop, cudaStream_t st) {
  constexpr int TP = TCH * 32;
  constexpr int SMEM = 2 * 16 * TP * 4 + 2 * TP * DH * 2 + NW * 2 * NQ * TP * 4 + NW * 2 * NQ * DH * 4 + 2 * TP * 4;
  auto kern = attn_bwd_dkv_kernel<TCH, NW>;
  static bool cfg = false;
  if (!cfg) { RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); cfg = true; }
  const int kpb = 64;
  kern<<<dim3((Sk + kpb - 1) / kpb, B * H), NW * 32, SMEM, st>>>(
      static_cast<const rb_t*>(Q), static_cast<const rb_t*>(K), static_cast<const rb_t*>(V), static_cast<const uint8_t*>(kpm),
      static_cast<const rb_t*>(dO), LSE, Dbuf, static_cast<rb_t*>(dK), static_cast<rb_t*>(dV), H, Tq, Sk, ldq, ldk, ldv, lddo,
      lddk, lddv, scale, kpb, drop);
  RB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rb

using namespace rb;

#define RB_ATTN_CHECK(name)                                                                                 \
  if (dh != DH) return rb_fail(name ": only head_dim 32 is built (got %d)", dh);                            \
  if (Tq <= 0 || Sk <= 0 || B <= 0 || H <= 0) return rb_fail(name ": empty problem");

int rb::attn_fwd_simt(const void* Q, const void* K, const void* V, const void* kpm, void* O, float* LSE, int B, int H, int dh, int Tq, int Sk,
                           long long ldq, long long ldk, long long ldv, long long ldo, float scale, const rb_dropout* drop, void* stream) {
  RB_ATTN_CHECK("rb_attn_fwd");
  const DropK dk = make_dropk(drop);
  if ((ldk % 8) || (ldv % 8)) return rb_fail("rb_attn_fwd: K/V pitch must be a multiple of 8 elements");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (Sk <= 32) return launch_fwd<1>(Q, K, V, kpm, O, LSE, B, H, Tq, Sk, ldq, ldk, ldv, ldo, scale, dk, st);
  if (Sk <= 128) return launch_fwd<4>(Q, K, V, kpm, O, LSE, B, H, Tq, Sk, ldq, ldk, ldv, ldo, scale, dk, st);
  if (Sk <= 448) return launch_fwd<14>(Q, K, V, kpm, O, LSE, B, H, Tq, Sk, ldq, ldk, ldv, ldo, scale, dk, st);
  if (Sk <= 672) return launch_fwd<21>(Q, K, V, kpm, O, LSE, B, H, Tq, Sk, ldq, ldk, ldv, ldo, scale, dk, st);
  return rb_fail("rb_attn_fwd: Sk = %d > 672 keys is not built yet", Sk);
}

int rb::attn_bwd_simt(const void* Q, const void* K, const void* V, const void* kpm, const void* O, const void* dO, const float* LSE, void* dQ, void* dK,
                           void* dV, float* Dbuf, int B, int H, int dh, int Tq, int Sk, long long ldq, long long ldk, long long ldv, long long ldo,
                           long long lddo, long long lddq, long long lddk, long long lddv, float scale, const rb_dropout* drop, void* stream) {
  RB_ATTN_CHECK("rb_attn_bwd");
  const DropK dk = make_dropk(drop);
  if ((ldk % 8) || (ldv % 8) || (ldq % 8) || (lddo % 8)) return rb_fail("rb_attn_bwd: pitches must be multiples of 8 elements");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if (Sk <= 32) rc = launch_dq<1>(Q, K, V, kpm, O, dO, LSE, dQ, Dbuf, B, H, Tq, Sk, ldq, ldk, ldv, ldo, lddo, lddq, scale, dk, st);
  else if (Sk <= 128) rc = launch_dq<4>(Q, K, V, kpm, O, dO, LSE, dQ, Dbuf, B, H, Tq, Sk, ldq, ldk, ldv, ldo, lddo, lddq, scale, dk, st);
  else if (Sk <= 448) rc = launch_dq<14>(Q, K, V, kpm, O, dO, LSE, dQ, Dbuf, B, H, Tq, Sk, ldq, ldk, ldv, ldo, lddo, lddq, scale, dk, st);
  else if (Sk <= 672) rc = launch_dq<21>(Q, K, V, kpm, O, dO, LSE, dQ, Dbuf, B, H, Tq, Sk, ldq, ldk, ldv, ldo, lddo, lddq, scale, dk, st);
  else return rb_fail("rb_attn_bwd: Sk = %d > 672 keys is not built yet", Sk);
  if (rc) return rc;
  if (Tq <= 32) return launch_dkv<1, 8>(Q, K, V, kpm, dO, LSE, Dbuf, dK, dV, B, H, Tq, Sk, ldq, ldk, ldv, lddo, lddk, lddv, scale, dk, st);
  if (Tq <= 128) return launch_dkv<4, 8>(Q, K, V, kpm, dO, LSE, Dbuf, dK, dV, B, H, Tq, Sk, ldq, ldk, ldv, lddo, lddk, lddv, scale, dk, st);
  if (Tq <= 448) return launch_dkv<14, 8>(Q, K, V, kpm, dO, LSE, Dbuf, dK, dV, B, H, Tq, Sk, ldq, ldk, ldv, lddo, lddk, lddv, scale, dk, st);
  if (Tq <= 672) return launch_dkv<21, 4>(Q, K, V, kpm, dO, LSE, Dbuf, dK, dV, B, H, Tq, Sk, ldq, ldk, ldv, lddo, lddk, lddv, scale, dk, st);
  return rb_fail("rb_attn_bwd: Tq = %d > 672 queries is not built yet", Tq);
}

extern "C" int rb_qenc_pool_fwd(const float* k, const float* q, const float* v, const void* mask, int B, int L, int n_ph, float* att, float* c, void* stream) {
  qenc_pool_fwd_kernel<<<B, 256, 2 * L * sizeof(float), static_cast<cudaStream_t>(stream)>>>(k, q, v, static_cast<const uint8_t*>(mask), L, n_ph, att, c);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_qenc_pool_bwd(const float* dc, const float* k, const float* q, const float* v, const float* att, int B, int L, int n_ph, float* dk, float* dq,
                                float* dv, void* stream) {
  if (n_ph > 256) return rb_fail("rb_qenc_pool_bwd: n_ph > 256");
  qenc_pool_bwd_kernel<<<B, 256, n_ph * L * sizeof(float), static_cast<cudaStream_t>(stream)>>>(dc, k, q, v, att, L, n_ph, dk, dq, dv);
  RB_CUDA(cudaGetLastError());
  return 0;
}
