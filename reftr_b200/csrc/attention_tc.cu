// Multi-head attention core on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM), head_dim 32.
// Replaces the bmm / softmax / bmm of torch's F.multi_head_attention_forward as invoked at transformer.py:174, :239, :243
// for query counts >= 64 (the encoder self-attention; the decoder's 1..16 queries stay on the SIMT kernels of attention.cu).
//
// Forward (one CTA = 128 queries of one (batch, head); loop over blocks of 64 keys; online softmax):
//     S   = Q K_j^T          UMMA 128 x 64 x 32      A = Q  [128 q x 32]  K-major SW64,  B = K_j [64 keys x 32] K-major SW64
//     P   = exp2(S*c - m)    4 warps, one query row per thread, S read with tcgen05.ld, P written to smem as bf16 (SW128)
//     O  += P V_j            UMMA 128 x 32 x 64      A = P  [128 q x 64]  K-major SW128, B = V_j [64 keys x 32] MN-major SW64
// Backward is two kernels that recompute P from the saved log-sum-exp, so that no transposes and no atomics are needed:
//     dQ  kernel (128 queries per CTA, loop over key blocks):   S, dP = dO V_j^T, dS = c P (dP - D),  dQ += dS K_j
//     dKV kernel (128 keys per CTA, loop over query blocks):    S^T = K Q_i^T, dP^T = V dO_i^T, dV += P^T dO_i, dK += dS^T Q_i
// Warp roles (192 threads): warps 0-3 softmax / element-wise (TMEM lane quarter = warp), warp 4 TMA producer, warp 5 MMA
// issuer + TMEM allocator.  Several CTAs are resident per SM (<= 45 KB smem, 128 / 256 TMEM columns), which hides the
// MMA -> softmax -> MMA dependency chain of one CTA behind the others.
#include "common.cuh"
#include "host.h"

#include <initializer_list>

namespace rb {

constexpr int ATC_THREADS = 192;
constexpr int ATC_ROWS = 128;  // rows (queries, or keys in the dKV kernel) per CTA
constexpr int ATC_BLK = 64;    // columns (keys, or queries in the dKV kernel) per block
constexpr int ATC_DH = 32;
constexpr int ATC_MAXBLK = 64;  // up to 4096 columns
// resident CTAs per SM the register allocation is held to (these kernels are latency chains MMA -> softmax -> MMA: other CTAs on the SM
// are what hides them; un-bounded, ptxas took 119 registers for the forward = 2 CTAs and 185 for dK/dV = ONE CTA, i.e. 6 warps per SM)
#ifndef ATC_FWD_CTAS
#define ATC_FWD_CTAS 3
#endif
#ifndef ATC_DKV_CTAS
#define ATC_DKV_CTAS 2
#endif

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// K-major operand with 64-byte rows (32 bf16), 64B swizzle: 8-row groups are 512 B apart; k-step of 16 elements = +32 B
__device__ __forceinline__ uint64_t desc_k64(uint32_t saddr, int kstep) { return umma_smem_desc(saddr + kstep * 32, 16, 512, SWZ_64B); }
// K-major operand with 128-byte rows (64 bf16), 128B swizzle (the P / dS tiles we write ourselves)
__device__ __forceinline__ uint64_t desc_k128(uint32_t saddr, int kstep) { return umma_smem_desc(saddr + kstep * 32, 16, 1024, SWZ_128B); }
// MN-major operand [64 k-rows x 32 n] with 64-byte rows, 64B swizzle: 8-row (k) groups are 512 B apart; k-step of 16 rows = +1024 B
__device__ __forceinline__ uint64_t desc_mn64(uint32_t saddr, int kstep) { return umma_smem_desc(saddr + kstep * 1024, 512, 512, SWZ_64B); }

// thread `r` writes 32 consecutive bf16 (columns [32*half, 32*half+32)) of row r of a [128 x 64] bf16 SW128 K-major tile
__device__ __forceinline__ void store_row_half_sw128(uint8_t* tile, int r, int half, const float (&v)[32]) {
  uint8_t* row = tile + r * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 t;
    t.x = pack_t2(v[8 * c], v[8 * c + 1]); t.y = pack_t2(v[8 * c + 2], v[8 * c + 3]);
    t.z = pack_t2(v[8 * c + 4], v[8 * c + 5]); t.w = pack_t2(v[8 * c + 6], v[8 * c + 7]);
    const int chunk = half * 4 + c;
    *reinterpret_cast<uint4*>(row + ((chunk ^ (r & 7)) << 4)) = t;
  }
}

__device__ __forceinline__ void build_maskbits(uint32_t* bits, const uint8_t* kpm_row, int n_valid, int n_words, int warp, int lane, int nwarps) {
  for (int w = warp; w < n_words; w += nwarps) {
    const int j = w * 32 + lane;
    const bool m = (j >= n_valid) || (kpm_row && kpm_row[j]);
    const uint32_t b = __ballot_sync(0xffffffffu, m);
    if (lane == 0) bits[w] = b;
  }
}

struct AtcSmem {
  uint8_t* tileA;   // 16 KB SW128 tile written by the element-wise warps (P / dS / P^T)
  uint8_t* tileB;   // second such tile (dS^T in the dKV kernel)
  uint8_t* rowA;    // 8 KB: the CTA's 128-row operand (Q, or K in dKV)
  uint8_t* rowB;    // 8 KB: second 128-row operand (dO in dQ, V in dKV)
  uint8_t* blk0;    // 2 x 4 KB: per-block operand 0 (K_j, or Q_i)
  uint8_t* blk1;    // 2 x 4 KB: per-block operand 1 (V_j, or dO_i)
  uint64_t* bars;   // [0] rows, [1..2] blk_full, [3..4] blk_free, [5] acc_full, [6] ew_ready
  uint32_t* tmem_slot;
  uint32_t* maskbits;  // [ATC_MAXBLK * 2]
  float* colL;      // [2][64] per-block column LSE (dKV)
  float* colD;      // [2][64] per-block column D (dKV)
};
// layout (offsets from a 1024-aligned base): tileA 0 | rowA 16K | blk0 24K | blk1 32K | misc 40K (2 KB) | rowB 42K | tileB 50K
constexpr int ATC_SMEM_FWD = 43008 + 1024;   // forward: no rowB / tileB
constexpr int ATC_SMEM_DQ = 51200 + 1024;    // dQ: + rowB (dO)
constexpr int ATC_SMEM_DKV = 67584 + 1024;   // dKV: + tileB (dS^T)

__device__ __forceinline__ AtcSmem carve(uint8_t* raw) {
  uint8_t* p = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  AtcSmem s;
  s.tileA = p;
  s.rowA = p + 16384;
  s.blk0 = p + 24576;
  s.blk1 = p + 32768;
  uint8_t* misc = p + 40960;
  s.bars = reinterpret_cast<uint64_t*>(misc);
  s.tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  s.maskbits = reinterpret_cast<uint32_t*>(misc + 128);          // ATC_MAXBLK * 2 words = 512 B
  s.colL = reinterpret_cast<float*>(misc + 640);                  // 2 x 64 floats
  s.colD = reinterpret_cast<float*>(misc + 640 + 512);            // 2 x 64 floats  (ends at 1664 < 2048)
  s.rowB = p + 43008;
  s.tileB = p + 51200;
  return s;
}

enum { BAR_ROWS = 0, BAR_FULL = 1, BAR_FREE = 3, BAR_ACC = 5, BAR_EW = 6 };

__device__ __forceinline__ void atc_init(const AtcSmem& s, int warp, int tmem_cols_is_256) {
  if (threadIdx.x == 0) {
    mbar_init(&s.bars[BAR_ROWS], 1);
    mbar_init(&s.bars[BAR_FULL], 1); mbar_init(&s.bars[BAR_FULL + 1], 1);
    mbar_init(&s.bars[BAR_FREE], 1); mbar_init(&s.bars[BAR_FREE + 1], 1);
    mbar_init(&s.bars[BAR_ACC], 1);
    mbar_init(&s.bars[BAR_EW], 128);
    fence_barrier_init();
  }
  if (warp == 5) {
    if (tmem_cols_is_256) tmem_alloc<256>(s.tmem_slot); else tmem_alloc<128>(s.tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// =====================================================================================================================
// forward
// =====================================================================================================================
__global__ void __launch_bounds__(ATC_THREADS, ATC_FWD_CTAS)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const uint8_t* __restrict__ kpm, rb_t* __restrict__ O, float* __restrict__ LSE, int H, int Tq, int Sk, long long ldo,
                   float scale, DropK drop) {
  extern __shared__ uint8_t smem_raw[];
  const AtcSmem s = carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int q0 = blockIdx.x * ATC_ROWS;
  const int nb = (Sk + ATC_BLK - 1) / ATC_BLK;
  build_maskbits(s.maskbits, kpm ? kpm + static_cast<long long>(b) * Sk : nullptr, Sk, nb * 2, warp, lane, ATC_THREADS / 32);
  atc_init(s, warp, 0);
  const uint32_t tmem = *s.tmem_slot;
  const uint32_t T_S = tmem, T_O = tmem + 64;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(&s.bars[BAR_ROWS], 8192);
      tma_load_2d(s.rowA, &tmQ, &s.bars[BAR_ROWS], h * ATC_DH, b * Tq + q0);
      for (int j = 0; j < nb; ++j) {
        const int st = j & 1;
        if (j >= 2) mbar_wait(&s.bars[BAR_FREE + st], ((j >> 1) - 1) & 1);
        mbar_expect_tx(&s.bars[BAR_FULL + st], 8192);
        tma_load_2d(s.blk0 + st * 4096, &tmK, &s.bars[BAR_FULL + st], h * ATC_DH, b * Sk + j * ATC_BLK);
        tma_load_2d(s.blk1 + st * 4096, &tmV, &s.bars[BAR_FULL + st], h * ATC_DH, b * Sk + j * ATC_BLK);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc_t(128, 64, 0, 0);
      constexpr uint32_t idO = umma_idesc_t(128, 32, 0, 1);
      const uint32_t aQ = smem_u32(s.rowA), aP = smem_u32(s.tileA);
      mbar_wait(&s.bars[BAR_ROWS], 0);
      for (int j = 0; j < nb; ++j) {
        const int st = j & 1;
        mbar_wait(&s.bars[BAR_FULL + st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(s.blk0 + st * 4096), aV = smem_u32(s.blk1 + st * 4096);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16_ss(T_S, desc_k64(aQ, k), desc_k64(aK, k), idS, k);
        umma_commit(&s.bars[BAR_ACC]);
        mbar_wait(&s.bars[BAR_EW], j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(T_O, desc_k128(aP, k), desc_mn64(aV, k), idO, (j | k) != 0);
        umma_commit(&s.bars[BAR_FREE + st]);
      }
    }
    __syncwarp();
  } else {
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const float c = scale * 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;
    const bool use_drop = drop.seed != nullptr;
    const uint32_t dkey = use_drop ? drop_key(drop) : 0u;
    const uint32_t wrow = (static_cast<uint32_t>(bh) * Tq + (q0 + r)) * static_cast<uint32_t>((Sk + 1) >> 1);
    for (int j = 0; j < nb; ++j) {
      mbar_wait(&s.bars[BAR_ACC], j & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(T_S + lane_addr, v0);
      tmem_ld_32x32(T_S + lane_addr + 32, v1);
      tmem_ld_wait();
      const uint32_t mb0 = s.maskbits[2 * j], mb1 = s.maskbits[2 * j + 1];
      float p0[32], p1[32];
      float mx = -INFINITY;
#ifdef RB_ATTN_FAST
      if ((mb0 | mb1) == 0u) {  // EXPERIMENTAL variant (tools/build_variant.sh): no key of this block is masked -> no selects
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          p0[i] = __uint_as_float(v0[i]) * c;
          p1[i] = __uint_as_float(v1[i]) * c;
          mx = fmaxf(mx, fmaxf(p0[i], p1[i]));
        }
      } else
#endif
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        p0[i] = ((mb0 >> i) & 1u) ? -INFINITY : __uint_as_float(v0[i]) * c;
        p1[i] = ((mb1 >> i) & 1u) ? -INFINITY : __uint_as_float(v1[i]) * c;
        mx = fmaxf(mx, fmaxf(p0[i], p1[i]));
      }
      const float m_new = fmaxf(m, mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m == -INFINITY) ? 1.f : ex2(m - m_safe);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        p0[i] = ex2(p0[i] - m_safe);
        p1[i] = ex2(p1[i] - m_safe);
        sum += p0[i] + p1[i];
      }
      l = l * alpha + sum;
      m = m_new;
      if (use_drop) {  // dropout on the probabilities (l stays the undropped row sum; 1/(1-p) is applied with 1/l at the end)
        const uint32_t cb = wrow + j * (ATC_BLK / 2);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t w0 = drop_word(dkey, cb + i), w1 = drop_word(dkey, cb + 16 + i);
          if (!drop_keep(w0, 0, drop.thr)) p0[2 * i] = 0.f;
          if (!drop_keep(w0, 1, drop.thr)) p0[2 * i + 1] = 0.f;
          if (!drop_keep(w1, 0, drop.thr)) p1[2 * i] = 0.f;
          if (!drop_keep(w1, 1, drop.thr)) p1[2 * i + 1] = 0.f;
        }
      }
      if (j > 0) {
        // the previous block's P V must have completed before P is overwritten and O is rescaled
        mbar_wait(&s.bars[BAR_FREE + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
          uint32_t o[32];
          tmem_ld_32x32(T_O + lane_addr, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32(T_O + lane_addr, o);
          tmem_st_wait();
        }
      }
      store_row_half_sw128(s.tileA, r, 0, p0);
      store_row_half_sw128(s.tileA, r, 1, p1);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&s.bars[BAR_EW]);
    }
    mbar_wait(&s.bars[BAR_FREE + ((nb - 1) & 1)], ((nb - 1) >> 1) & 1);
    tc_fence_after();
    uint32_t o[32];
    tmem_ld_32x32(T_O + lane_addr, o);
    tmem_ld_wait();
    const int q = q0 + r;
    if (q < Tq) {
      const float inv = l > 0.f ? drop.scale / l : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(O + (static_cast<long long>(b) * Tq + q) * ldo + h * ATC_DH);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 t;
        t.x = pack_t2(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
        t.y = pack_t2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
        t.z = pack_t2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
        t.w = pack_t2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
        dst[i] = t;
      }
      if (LSE) LSE[static_cast<long long>(bh) * Tq + q] = (m == -INFINITY) ? -69.07755279f /* log(1e-30), as the SIMT kernel */
                                                                          : (m + __log2f(fmaxf(l, 1e-30f))) * 0.6931471805599453f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<128>(tmem);
}

// =====================================================================================================================
// backward: dQ (and D = rowsum(dO * O))
// =====================================================================================================================
__global__ void __launch_bounds__(ATC_THREADS)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                      const __grid_constant__ CUtensorMap tmdO, const uint8_t* __restrict__ kpm, const rb_t* __restrict__ O,
                      const rb_t* __restrict__ dO, const float* __restrict__ LSE, rb_t* __restrict__ dQ, float* __restrict__ Dbuf, int H,
                      int Tq, int Sk, long long ldo, long long lddo, long long lddq, float scale, DropK drop) {
  extern __shared__ uint8_t smem_raw[];
  const AtcSmem s = carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int q0 = blockIdx.x * ATC_ROWS;
  const int nb = (Sk + ATC_BLK - 1) / ATC_BLK;
  build_maskbits(s.maskbits, kpm ? kpm + static_cast<long long>(b) * Sk : nullptr, Sk, nb * 2, warp, lane, ATC_THREADS / 32);
  atc_init(s, warp, 1);
  const uint32_t tmem = *s.tmem_slot;
  const uint32_t T_S = tmem, T_dP = tmem + 64, T_dQ = tmem + 128;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(&s.bars[BAR_ROWS], 16384);
      tma_load_2d(s.rowA, &tmQ, &s.bars[BAR_ROWS], h * ATC_DH, b * Tq + q0);
      tma_load_2d(s.rowB, &tmdO, &s.bars[BAR_ROWS], h * ATC_DH, b * Tq + q0);
      for (int j = 0; j < nb; ++j) {
        const int st = j & 1;
        if (j >= 2) mbar_wait(&s.bars[BAR_FREE + st], ((j >> 1) - 1) & 1);
        mbar_expect_tx(&s.bars[BAR_FULL + st], 8192);
        tma_load_2d(s.blk0 + st * 4096, &tmK, &s.bars[BAR_FULL + st], h * ATC_DH, b * Sk + j * ATC_BLK);
        tma_load_2d(s.blk1 + st * 4096, &tmV, &s.bars[BAR_FULL + st], h * ATC_DH, b * Sk + j * ATC_BLK);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc_t(128, 64, 0, 0);
      constexpr uint32_t idQ = umma_idesc_t(128, 32, 0, 1);
      const uint32_t aQ = smem_u32(s.rowA), adO = smem_u32(s.rowB), aDS = smem_u32(s.tileA);
      mbar_wait(&s.bars[BAR_ROWS], 0);
      for (int j = 0; j < nb; ++j) {
        const int st = j & 1;
        mbar_wait(&s.bars[BAR_FULL + st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(s.blk0 + st * 4096), aV = smem_u32(s.blk1 + st * 4096);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16_ss(T_S, desc_k64(aQ, k), desc_k64(aK, k), idS, k);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16_ss(T_dP, desc_k64(adO, k), desc_k64(aV, k), idS, k);
        umma_commit(&s.bars[BAR_ACC]);
        mbar_wait(&s.bars[BAR_EW], j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(T_dQ, desc_k128(aDS, k), desc_mn64(aK, k), idQ, (j | k) != 0);
        umma_commit(&s.bars[BAR_FREE + st]);
      }
    }
    __syncwarp();
  } else {
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const float c = scale * 1.4426950408889634f;
    const int q = q0 + r;
    const bool q_ok = q < Tq;
    float lse2 = INFINITY, D = 0.f;
    const bool use_drop = drop.seed != nullptr;
    const uint32_t dkey = use_drop ? drop_key(drop) : 0u;
    const uint32_t wrow = (static_cast<uint32_t>(bh) * Tq + q) * static_cast<uint32_t>((Sk + 1) >> 1);
    if (q_ok) {
      lse2 = LSE[static_cast<long long>(bh) * Tq + q] * 1.4426950408889634f;
      const uint4* po = reinterpret_cast<const uint4*>(O + (static_cast<long long>(b) * Tq + q) * ldo + h * ATC_DH);
      const uint4* pd = reinterpret_cast<const uint4*>(dO + (static_cast<long long>(b) * Tq + q) * lddo + h * ATC_DH);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 a = po[i], d = pd[i];
        D += t_lo(a.x) * t_lo(d.x) + t_hi(a.x) * t_hi(d.x) + t_lo(a.y) * t_lo(d.y) + t_hi(a.y) * t_hi(d.y) +
             t_lo(a.z) * t_lo(d.z) + t_hi(a.z) * t_hi(d.z) + t_lo(a.w) * t_lo(d.w) + t_hi(a.w) * t_hi(d.w);
      }
      Dbuf[static_cast<long long>(bh) * Tq + q] = D;
    }
    for (int j = 0; j < nb; ++j) {
      mbar_wait(&s.bars[BAR_ACC], j & 1);
      tc_fence_after();
      if (j > 0) {  // dS tile is free once the previous block's dQ MMA has completed
        mbar_wait(&s.bars[BAR_FREE + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(T_S + lane_addr + half * 32, sv);
        tmem_ld_32x32(T_dP + lane_addr + half * 32, dv);
        tmem_ld_wait();
        const uint32_t mb = s.maskbits[2 * j + half];
        float ds[32];
        if (use_drop) {  // dv is the gradient w.r.t. the DROPPED probabilities
          const uint32_t cb = wrow + j * (ATC_BLK / 2) + half * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t w = drop_word(dkey, cb + i);
            dv[2 * i] = drop_keep(w, 0, drop.thr) ? __float_as_uint(__uint_as_float(dv[2 * i]) * drop.scale) : 0u;
            dv[2 * i + 1] = drop_keep(w, 1, drop.thr) ? __float_as_uint(__uint_as_float(dv[2 * i + 1]) * drop.scale) : 0u;
          }
        }
#ifdef RB_ATTN_FAST
        if (mb == 0u) {
#pragma unroll
          for (int i = 0; i < 32; ++i) ds[i] = ex2(__uint_as_float(sv[i]) * c - lse2) * (__uint_as_float(dv[i]) - D) * scale;
        } else
#endif
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float p = ((mb >> i) & 1u) ? 0.f : ex2(__uint_as_float(sv[i]) * c - lse2);
          ds[i] = p * (__uint_as_float(dv[i]) - D) * scale;
        }
        store_row_half_sw128(s.tileA, r, half, ds);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&s.bars[BAR_EW]);
    }
    mbar_wait(&s.bars[BAR_FREE + ((nb - 1) & 1)], ((nb - 1) >> 1) & 1);
    tc_fence_after();
    uint32_t o[32];
    tmem_ld_32x32(T_dQ + lane_addr, o);
    tmem_ld_wait();
    if (q_ok) {
      uint4* dst = reinterpret_cast<uint4*>(dQ + (static_cast<long long>(b) * Tq + q) * lddq + h * ATC_DH);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 t;
        t.x = pack_t2(__uint_as_float(o[8 * i]), __uint_as_float(o[8 * i + 1]));
        t.y = pack_t2(__uint_as_float(o[8 * i + 2]), __uint_as_float(o[8 * i + 3]));
        t.z = pack_t2(__uint_as_float(o[8 * i + 4]), __uint_as_float(o[8 * i + 5]));
        t.w = pack_t2(__uint_as_float(o[8 * i + 6]), __uint_as_float(o[8 * i + 7]));
        dst[i] = t;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<256>(tmem);
}

// =====================================================================================================================
// backward: dK, dV  (rows of the CTA = 128 keys; blocks = 64 queries)
// =====================================================================================================================
__global__ void __launch_bounds__(ATC_THREADS, ATC_DKV_CTAS)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                       const __grid_constant__ CUtensorMap tmdO, const uint8_t* __restrict__ kpm, const float* __restrict__ LSE,
                       const float* __restrict__ Dbuf, rb_t* __restrict__ dK, rb_t* __restrict__ dV, int H, int Tq, int Sk, long long lddk,
                       long long lddv, float scale, DropK drop) {
  extern __shared__ uint8_t smem_raw[];
  const AtcSmem s = carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int k0 = blockIdx.x * ATC_ROWS;
  const int nb = (Tq + ATC_BLK - 1) / ATC_BLK;
  atc_init(s, warp, 1);
  const uint32_t tmem = *s.tmem_slot;
  const uint32_t T_S = tmem, T_dP = tmem + 64, T_dV = tmem + 128, T_dK = tmem + 160;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(&s.bars[BAR_ROWS], 16384);
      tma_load_2d(s.rowA, &tmK, &s.bars[BAR_ROWS], h * ATC_DH, b * Sk + k0);
      tma_load_2d(s.rowB, &tmV, &s.bars[BAR_ROWS], h * ATC_DH, b * Sk + k0);
    }
    for (int i = 0; i < nb; ++i) {
      const int st = i & 1;
      if (i >= 2) mbar_wait(&s.bars[BAR_FREE + st], ((i >> 1) - 1) & 1);
      // per-query LSE (log2 units; +inf beyond Tq so that P = 0) and D of this block
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int cq = i * ATC_BLK + lane + 32 * t;
        const bool ok = cq < Tq;
        s.colL[st * 64 + lane + 32 * t] = ok ? LSE[static_cast<long long>(bh) * Tq + cq] * 1.4426950408889634f : INFINITY;
        s.colD[st * 64 + lane + 32 * t] = ok ? Dbuf[static_cast<long long>(bh) * Tq + cq] : 0.f;
      }
      __syncwarp();
      if (lane == 0) {
        mbar_expect_tx(&s.bars[BAR_FULL + st], 8192);
        tma_load_2d(s.blk0 + st * 4096, &tmQ, &s.bars[BAR_FULL + st], h * ATC_DH, b * Tq + i * ATC_BLK);
        tma_load_2d(s.blk1 + st * 4096, &tmdO, &s.bars[BAR_FULL + st], h * ATC_DH, b * Tq + i * ATC_BLK);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc_t(128, 64, 0, 0);
      constexpr uint32_t idG = umma_idesc_t(128, 32, 0, 1);
      const uint32_t aK = smem_u32(s.rowA), aV = smem_u32(s.rowB), aPT = smem_u32(s.tileA), aDST = smem_u32(s.tileB);
      mbar_wait(&s.bars[BAR_ROWS], 0);
      for (int i = 0; i < nb; ++i) {
        const int st = i & 1;
        mbar_wait(&s.bars[BAR_FULL + st], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t aQ = smem_u32(s.blk0 + st * 4096), adO = smem_u32(s.blk1 + st * 4096);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16_ss(T_S, desc_k64(aK, k), desc_k64(aQ, k), idS, k);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16_ss(T_dP, desc_k64(aV, k), desc_k64(adO, k), idS, k);
        umma_commit(&s.bars[BAR_ACC]);
        mbar_wait(&s.bars[BAR_EW], i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(T_dV, desc_k128(aPT, k), desc_mn64(adO, k), idG, (i | k) != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(T_dK, desc_k128(aDST, k), desc_mn64(aQ, k), idG, (i | k) != 0);
        umma_commit(&s.bars[BAR_FREE + st]);
      }
    }
    __syncwarp();
  } else {
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const float c = scale * 1.4426950408889634f;
    const int key = k0 + r;
    const bool key_ok = key < Sk && !(kpm && kpm[static_cast<long long>(b) * Sk + key]);
    const bool use_drop = drop.seed != nullptr;
    const uint32_t dkey = use_drop ? drop_key(drop) : 0u;
    const uint32_t wpr = static_cast<uint32_t>((Sk + 1) >> 1);
    for (int i = 0; i < nb; ++i) {
      const int st = i & 1;
      mbar_wait(&s.bars[BAR_FULL + st], (i >> 1) & 1);  // colL / colD of this block are visible
      mbar_wait(&s.bars[BAR_ACC], i & 1);
      tc_fence_after();
      if (i > 0) {
        mbar_wait(&s.bars[BAR_FREE + ((i - 1) & 1)], ((i - 1) >> 1) & 1);
        tc_fence_after();
      }
      const float* cl = s.colL + st * 64;
      const float* cd = s.colD + st * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(T_S + lane_addr + half * 32, sv);
        tmem_ld_32x32(T_dP + lane_addr + half * 32, dv);
        tmem_ld_wait();
        float p[32], ds[32];
        // element (query i*64 + half*32 + t, key): this thread owns one key, so one random word per element
        uint32_t ctr = (static_cast<uint32_t>(bh) * Tq + i * ATC_BLK + half * 32) * wpr + static_cast<uint32_t>(key >> 1);
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const float pv = key_ok ? ex2(__uint_as_float(sv[t]) * c - cl[half * 32 + t]) : 0.f;
          float mk = 1.f;
#ifdef RB_ATTN_FAST
          if (use_drop) {  // the even thread hashes the even query rows, the odd thread the odd ones; one shuffle hands the word over
            uint32_t w = 0;
            if ((t & 1) == (lane & 1)) w = drop_word(dkey, ctr);
            const uint32_t wo = __shfl_xor_sync(0xffffffffu, w, 1);
            if ((t & 1) != (lane & 1)) w = wo;
            mk = drop_keep(w, key & 1, drop.thr) ? drop.scale : 0.f;
            ctr += wpr;
          }
#else
          if (use_drop) { mk = drop_keep(drop_word(dkey, ctr), key & 1, drop.thr) ? drop.scale : 0.f; ctr += wpr; }
#endif
          p[t] = pv * mk;  // dV uses the dropped probabilities
          ds[t] = pv * (__uint_as_float(dv[t]) * mk - cd[half * 32 + t]) * scale;
        }
        store_row_half_sw128(s.tileA, r, half, p);
        store_row_half_sw128(s.tileB, r, half, ds);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&s.bars[BAR_EW]);
    }
    mbar_wait(&s.bars[BAR_FREE + ((nb - 1) & 1)], ((nb - 1) >> 1) & 1);
    tc_fence_after();
    uint32_t gv[32], gk[32];
    tmem_ld_32x32(T_dV + lane_addr, gv);
    tmem_ld_32x32(T_dK + lane_addr, gk);
    tmem_ld_wait();
    if (key < Sk) {
      uint4* pv = reinterpret_cast<uint4*>(dV + (static_cast<long long>(b) * Sk + key) * lddv + h * ATC_DH);
      uint4* pk = reinterpret_cast<uint4*>(dK + (static_cast<long long>(b) * Sk + key) * lddk + h * ATC_DH);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 t;
        t.x = pack_t2(__uint_as_float(gv[8 * i]), __uint_as_float(gv[8 * i + 1]));
        t.y = pack_t2(__uint_as_float(gv[8 * i + 2]), __uint_as_float(gv[8 * i + 3]));
        t.z = pack_t2(__uint_as_float(gv[8 * i + 4]), __uint_as_float(gv[8 * i + 5]));
        t.w = pack_t2(__uint_as_float(gv[8 * i + 6]), __uint_as_float(gv[8 * i + 7]));
        pv[i] = t;
        t.x = pack_t2(__uint_as_float(gk[8 * i]), __uint_as_float(gk[8 * i + 1]));
        t.y = pack_t2(__uint_as_float(gk[8 * i + 2]), __uint_as_float(gk[8 * i + 3]));
        t.z = pack_t2(__uint_as_float(gk[8 * i + 4]), __uint_as_float(gk[8 * i + 5]));
        t.w = pack_t2(__uint_as_float(gk[8 * i + 6]), __uint_as_float(gk[8 * i + 7]));
        pk[i] = t;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------------------------------------ host side
static bool tc_eligible(int dh, int Tq, int Sk, std::initializer_list<const void*> ptrs, std::initializer_list<long long> lds) {
  if (dh != ATC_DH || Tq < 64 || Sk < 1 || Sk > ATC_MAXBLK * ATC_BLK || Tq > ATC_MAXBLK * ATC_BLK) return false;
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 15) return false;
  for (long long ld : lds)
    if (ld % 8) return false;
  return true;
}

template <typename K>
static int set_smem(K kern, int bytes) {
  RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

}  // namespace rb

using namespace rb;

extern "C" int rb_attn_fwd(const void* Q, const void* K, const void* V, const void* kpm, void* O, float* LSE, int B, int H, int dh, int Tq, int Sk,
                           long long ldq, long long ldk, long long ldv, long long ldo, float scale, const rb_dropout* drop, void* stream) {
  if (static_cast<long long>(B) * H * Tq * ((Sk + 1) / 2) >= (1LL << 32)) return rb_fail("rb_attn_fwd: dropout counter overflow");
  if (!tc_eligible(dh, Tq, Sk, {Q, K, V, O}, {ldq, ldk, ldv, ldo}))
    return attn_fwd_simt(Q, K, V, kpm, O, LSE, B, H, dh, Tq, Sk, ldq, ldk, ldv, ldo, scale, drop, stream);
  CUtensorMap tmQ, tmK, tmV;
  if (make_tmap_2d(&tmQ, Q, static_cast<uint64_t>(H) * dh, static_cast<uint64_t>(B) * Tq, ldq * 2, ATC_DH, ATC_ROWS)) return 1;
  if (make_tmap_2d(&tmK, K, static_cast<uint64_t>(H) * dh, static_cast<uint64_t>(B) * Sk, ldk * 2, ATC_DH, ATC_BLK)) return 1;
  if (make_tmap_2d(&tmV, V, static_cast<uint64_t>(H) * dh, static_cast<uint64_t>(B) * Sk, ldv * 2, ATC_DH, ATC_BLK)) return 1;
  static bool cfg = false;
  if (!cfg) { if (set_smem(attn_fwd_tc_kernel, ATC_SMEM_FWD)) return 1; cfg = true; }
  dim3 grid((Tq + ATC_ROWS - 1) / ATC_ROWS, B * H);
  attn_fwd_tc_kernel<<<grid, ATC_THREADS, ATC_SMEM_FWD, static_cast<cudaStream_t>(stream)>>>(
      tmQ, tmK, tmV, static_cast<const uint8_t*>(kpm), static_cast<rb_t*>(O), LSE, H, Tq, Sk, ldo, scale, make_dropk(drop));
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_attn_bwd(const void* Q, const void* K, const void* V, const void* kpm, const void* O, const void* dO, const float* LSE, void* dQ, void* dK,
                           void* dV, float* Dbuf, int B, int H, int dh, int Tq, int Sk, long long ldq, long long ldk, long long ldv, long long ldo,
                           long long lddo, long long lddq, long long lddk, long long lddv, float scale, const rb_dropout* drop, void* stream) {
  if (!tc_eligible(dh, Tq, Sk, {Q, K, V, O, dO, dQ, dK, dV}, {ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv}))
    return attn_bwd_simt(Q, K, V, kpm, O, dO, LSE, dQ, dK, dV, Dbuf, B, H, dh, Tq, Sk, ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, scale, drop, stream);
  const DropK dk = make_dropk(drop);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool cfg = false;
  if (!cfg) { if (set_smem(attn_bwd_dq_tc_kernel, ATC_SMEM_DQ) || set_smem(attn_bwd_dkv_tc_kernel, ATC_SMEM_DKV)) return 1; cfg = true; }
  {
    CUtensorMap tmQ, tmK, tmV, tmdO;
    if (make_tmap_2d(&tmQ, Q, static_cast<uint64_t>(H) * dh, static_cast<uint64_t>(B) * Tq, ldq * 2, ATC_DH, ATC_ROWS)) return 1;
    if (make_tmap_2d(&tmdO, dO, static_cast<uint64_t>(H) * dh, static_cast<uint64_t>(B) * Tq, lddo * 2, ATC_DH, ATC_ROWS)) return 1;
    if (make_tmap_2d(&tmK, K, static_cast<uint64_t>(H) * dh, static_cast<uint64_t>(B) * Sk, ldk * 2, ATC_DH, ATC_BLK)) return 1;
    if (make_tmap_2d(&tmV, V, static_cast<uint64_t>(H) * dh, static_cast<uint64_t>(B) * Sk, ldv * 2, ATC_DH, ATC_BLK)) return 1;
    dim3 grid((Tq + ATC_ROWS - 1) / ATC_ROWS, B * H);
    attn_bwd_dq_tc_kernel<<<grid, ATC_THREADS, ATC_SMEM_DQ, st>>>(tmQ, tmK, tmV, tmdO, static_cast<const uint8_t*>(kpm),
                                                                    static_cast<const rb_t*>(O), static_cast<const rb_t*>(dO), LSE,
                                                                    static_cast<rb_t*>(dQ), Dbuf, H, Tq, Sk, ldo, lddo, lddq, scale, dk);
    RB_CUDA(cudaGetLastError());
  }
  {
    CUtensorMap tmQ, tmK, tmV, tmdO;
    if (make_tmap_2d(&tmQ, Q, static_cast<uint64_t>(H) * dh, static_cast<uint64_t>(B) * Tq, ldq * 2, ATC_DH, ATC_BLK)) return 1;
    if (make_tmap_2d(&tmdO, dO, static_cast<uint64_t>(H) * dh, static_cast<uint64_t>(B) * Tq, lddo * 2, ATC_DH, ATC_BLK)) return 1;
    if (make_tmap_2d(&tmK, K, static_cast<uint64_t>(H) * dh, static_cast<uint64_t>(B) * Sk, ldk * 2, ATC_DH, ATC_ROWS)) return 1;
    if (make_tmap_2d(&tmV, V, static_cast<uint64_t>(H) * dh, static_cast<uint64_t>(B) * Sk, ldv * 2, ATC_DH, ATC_ROWS)) return 1;
    dim3 grid((Sk + ATC_ROWS - 1) / ATC_ROWS, B * H);
    attn_bwd_dkv_tc_kernel<<<grid, ATC_THREADS, ATC_SMEM_DKV, st>>>(tmQ, tmK, tmV, tmdO, static_cast<const uint8_t*>(kpm), LSE, Dbuf,
                                                                     static_cast<rb_t*>(dK), static_cast<rb_t*>(dV), H, Tq, Sk, lddk, lddv,
                                                                     scale, dk);
    RB_CUDA(cudaGetLastError());
  }
  return 0;
}
