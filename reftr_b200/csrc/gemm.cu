// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
// One CTA computes one 128 x BN output tile.  Warp roles (192 threads):
//   warp 0      : TMA producer (one elected lane) -- fills a STAGES-deep ring of {A tile, B tile} in shared memory
//   warp 1      : TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma, accumulators live in TMEM)
//   warps 2..5  : epilogue -- tcgen05.ld the accumulator (one thread per output row), fused bias / residual /
//                 ReLU / ReLU-mask / border-zero, vectorised stores (bf16 and/or fp32) or fp32 atomics (split-K)
// Two to three CTAs are resident per SM (<= 100 KB shared memory, <= 256 TMEM columns each) so one CTA's epilogue
// overlaps another CTA's main loop.
//
// mode NT: A [rows, K] and B [N, K] are K-major; tiles are [128|BN rows] x 64 k (128 B per row), 128B-swizzled by TMA.
//          Convolution taps are row shifts of A (padded-NHWC layout, see include/reftr_b200.h).
// mode TN: the contraction runs over rows (pixels/tokens): A [rows, M] and B [rows, N] are "MN-major"; tiles are
//          64 rows x 64 channels boxes, consumed through MN-major UMMA descriptors (no transposes anywhere).
#include "common.cuh"
#include "host.h"

namespace rb {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;

struct GemmKParams {
  int M, N;
  int kblocks;  // NT: k-blocks per tap; TN: total row-blocks
  int taps;
  int a_rowoff[16];
  int b_koff[16];
  int splits;
  long long out_row_off;
  const float* bias;
  const __nv_bfloat16* res; long long ldres;
  const float* res32; long long ldres32;
  const __nv_bfloat16* mask_src; long long ldmask;
  __nv_bfloat16* out; long long ldo;
  float* out32; long long ldo32;
  long long out32_z_stride;
  int relu, atomic;
  rb_geom geom;
};

__device__ __forceinline__ bool row_is_interior(const rb_geom& g, long long row) {
  if (g.mode == 0) return true;
  if (g.mode == 1) {
    const int t = static_cast<int>(row % g.HpWp);
    const int u = t / g.Wp, v = t - u * g.Wp;
    return (u >= 1) && (u <= g.H) && (v >= 1) && (v <= g.W);
  }
  // parity planes: cell (u,v) of plane (p,q) holds padded-input pixel (2u+p, 2v+q); interior iff 1 <= . <= H (resp. W)
  const int plane = static_cast<int>(row / g.Rs);
  row -= static_cast<long long>(plane) * g.Rs;
  const int t = static_cast<int>(row % g.HpWp);
  const int u = t / g.Wp, v = t - u * g.Wp;
  const int y = 2 * u + (plane >> 1), x = 2 * v + (plane & 1);
  return (y >= 1) && (y <= g.H) && (x >= 1) && (x <= g.W);
}

template <int BN, int MODE, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ GemmKParams p) {
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* accum = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // k range of this CTA
  int it_begin = 0, it_end = p.taps * p.kblocks;
  int z_tap = 0;
  if (MODE == 1) {
    z_tap = blockIdx.z / p.splits;
    int split = blockIdx.z - z_tap * p.splits;
    int per = (p.kblocks + p.splits - 1) / p.splits;
    it_begin = split * per;
    it_end = min(p.kblocks, it_begin + per);
    if (it_end <= it_begin) return;  // uniform per CTA: nothing to add
  }
  const int n_it = it_end - it_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < n_it; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
        uint8_t* a_dst = sA + s * A_BYTES;
        uint8_t* b_dst = sB + s * B_BYTES;
        const int it = it_begin + i;
        if (MODE == 0) {
          const int tap = it / p.kblocks;
          const int kc = it - tap * p.kblocks;
          tma_load_2d(a_dst, &tmA, &full[s], kc * BK, m0 + p.a_rowoff[tap]);
          tma_load_2d(b_dst, &tmB, &full[s], p.b_koff[tap] + kc * BK, n0);
        } else {
          const int r0 = it * BK;
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d(a_dst + j * 8192, &tmA, &full[s], m0 + 64 * j, r0 + p.a_rowoff[z_tap]);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * 8192, &tmB, &full[s], n0 + 64 * j, r0 + p.b_koff[z_tap]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, MODE, MODE);
      for (int i = 0; i < n_it; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + s * A_BYTES);
        const uint32_t b_base = smem_u32(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          uint64_t ad, bd;
          if (MODE == 0) {
            ad = umma_smem_desc(a_base + k * 32, 16, 1024, SWZ_128B);
            bd = umma_smem_desc(b_base + k * 32, 16, 1024, SWZ_128B);
          } else {
            ad = umma_smem_desc(a_base + k * 2048, 8192, 1024, SWZ_128B);
            bd = umma_smem_desc(b_base + k * 2048, 8192, 1024, SWZ_128B);
          }
          umma_bf16_ss(tmem_base, ad, bd, idesc, (i | k) != 0);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(accum);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const long long gm = static_cast<long long>(m0) + r;
    const bool row_ok = gm < p.M;
    const long long orow = gm + p.out_row_off;
    const bool interior = row_ok && row_is_interior(p.geom, orow);
    float* o32 = p.out32 ? p.out32 + static_cast<long long>(z_tap) * p.out32_z_stride + orow * p.ldo32 : nullptr;
    mbar_wait(accum, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent epilogue body
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      const int col0 = n0 + c * 32;
      if (!row_ok || col0 >= p.N) continue;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      const bool full_chunk = (col0 + 32 <= p.N);
      if (p.atomic) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) atomicAdd(o32 + col0 + j, f[j]);
        continue;
      }
      if (full_chunk) {
        if (p.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 b = __ldg(b4 + j);
            f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
          }
        }
        if (p.res) {
          const uint4* r4 = reinterpret_cast<const uint4*>(p.res + orow * p.ldres + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 t = __ldg(r4 + j);
            f[8 * j] += bf16_lo(t.x); f[8 * j + 1] += bf16_hi(t.x); f[8 * j + 2] += bf16_lo(t.y); f[8 * j + 3] += bf16_hi(t.y);
            f[8 * j + 4] += bf16_lo(t.z); f[8 * j + 5] += bf16_hi(t.z); f[8 * j + 6] += bf16_lo(t.w); f[8 * j + 7] += bf16_hi(t.w);
          }
        }
        if (p.res32) {
          const float4* r4 = reinterpret_cast<const float4*>(p.res32 + orow * p.ldres32 + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 b = __ldg(r4 + j);
            f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (p.mask_src) {
          const uint4* m4 = reinterpret_cast<const uint4*>(p.mask_src + orow * p.ldmask + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 t = __ldg(m4 + j);
            uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (!(bf16_lo(w[e]) > 0.f)) f[8 * j + 2 * e] = 0.f;
              if (!(bf16_hi(w[e]) > 0.f)) f[8 * j + 2 * e + 1] = 0.f;
            }
          }
        }
        if (!interior) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = 0.f;
        }
        if (p.out) {
          uint4* o4 = reinterpret_cast<uint4*>(p.out + orow * p.ldo + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 t;
            t.x = pack_bf16x2(f[8 * j], f[8 * j + 1]); t.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
            t.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]); t.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
            o4[j] = t;
          }
        }
        if (o32) {
          float4* o4 = reinterpret_cast<float4*>(o32 + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) o4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
      } else {
        // ragged N tail (N not a multiple of 32): scalar, guarded
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          if (col >= p.N) continue;
          float x = f[j];
          if (p.bias) x += p.bias[col];
          if (p.res) x += __bfloat162float(p.res[orow * p.ldres + col]);
          if (p.res32) x += p.res32[orow * p.ldres32 + col];
          if (p.relu) x = fmaxf(x, 0.f);
          if (p.mask_src && !(__bfloat162float(p.mask_src[orow * p.ldmask + col]) > 0.f)) x = 0.f;
          if (!interior) x = 0.f;
          if (p.out) p.out[orow * p.ldo + col] = __float2bfloat16(x);
          if (o32) o32[col] = x;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------ host side
template <int BN, int MODE, int STAGES>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& kp, dim3 grid, cudaStream_t st) {
  constexpr int SMEM = STAGES * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 1) * 8 + 16 + 1024;
  static bool configured = false;
  auto kern = umma_gemm_kernel<BN, MODE, STAGES>;
  if (!configured) {
    RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  kern<<<grid, GEMM_THREADS, SMEM, st>>>(tmA, tmB, kp);
  RB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rb

using namespace rb;

extern "C" int rb_gemm(const rb_gemm_args* a, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!a || !a->A || !a->B) return rb_fail("rb_gemm: null operand");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return rb_fail("rb_gemm: empty problem (M,N,K must be > 0)");
  if (a->taps < 1 || a->taps > 16) return rb_fail("rb_gemm: taps must be in [1,16]");
  if ((a->lda % 8) || (a->ldb % 8)) return rb_fail("rb_gemm: operand row pitch must be a multiple of 8 elements (16 B) for TMA");
  if ((reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->B) & 15)) return rb_fail("rb_gemm: operands must be 16-byte aligned");
  if (!a->out && !a->out32) return rb_fail("rb_gemm: no output");
  if (a->atomic && !a->out32) return rb_fail("rb_gemm: atomic accumulation needs out32");
  GemmKParams kp;
  kp.M = a->M; kp.N = a->N; kp.taps = a->taps; kp.splits = a->splits < 1 ? 1 : a->splits;
  for (int i = 0; i < 16; ++i) { kp.a_rowoff[i] = a->a_rowoff[i]; kp.b_koff[i] = a->b_koff[i]; }
  kp.out_row_off = a->out_row_off;
  kp.bias = a->bias;
  kp.res = static_cast<const __nv_bfloat16*>(a->res); kp.ldres = a->ldres;
  kp.res32 = a->res32; kp.ldres32 = a->ldres32;
  kp.mask_src = static_cast<const __nv_bfloat16*>(a->mask_src); kp.ldmask = a->ldmask;
  kp.out = static_cast<__nv_bfloat16*>(a->out); kp.ldo = a->ldo;
  kp.out32 = a->out32; kp.ldo32 = a->ldo32; kp.out32_z_stride = a->out32_z_stride;
  kp.relu = a->relu; kp.atomic = a->atomic; kp.geom = a->geom;
  kp.kblocks = (a->K + BK - 1) / BK;
  // vector epilogue alignment
  if (a->out && ((a->ldo % 8) || (reinterpret_cast<uintptr_t>(a->out) & 15))) return rb_fail("rb_gemm: out must be 16-byte aligned with pitch % 8 == 0");
  if (a->out32 && !a->atomic && ((a->ldo32 % 4) || (reinterpret_cast<uintptr_t>(a->out32) & 15))) return rb_fail("rb_gemm: out32 must be 16-byte aligned with pitch % 4 == 0");
  if (a->res && ((a->ldres % 8) || (reinterpret_cast<uintptr_t>(a->res) & 15))) return rb_fail("rb_gemm: res alignment");
  if (a->res32 && ((a->ldres32 % 4) || (reinterpret_cast<uintptr_t>(a->res32) & 15))) return rb_fail("rb_gemm: res32 alignment");
  if (a->mask_src && ((a->ldmask % 8) || (reinterpret_cast<uintptr_t>(a->mask_src) & 15))) return rb_fail("rb_gemm: mask_src alignment");
  if (a->bias && (reinterpret_cast<uintptr_t>(a->bias) & 15)) return rb_fail("rb_gemm: bias must be 16-byte aligned");

  int bn = a->block_n;
  const long long tiles_m = (a->M + BM - 1) / BM;
  if (bn == 0) {
    if (a->mode == 1) {
      bn = a->N >= 128 ? 128 : 64;
    } else {
      // widest tile that still gives every SM a few CTAs
      bn = 32;
      if (a->N > 32) bn = 64;
      if (a->N > 64) bn = 128;
      if (a->N >= 256 && tiles_m * ((a->N + 255) / 256) >= 296) bn = 256;
    }
  }
  if (a->mode == 1 && bn < 64) return rb_fail("rb_gemm: TN mode needs block_n >= 64");
  CUtensorMap tmA, tmB;
  dim3 grid;
  if (a->mode == 0) {
    if (make_tmap_2d(&tmA, a->A, static_cast<uint64_t>(a->a_cols), static_cast<uint64_t>(a->a_rows), a->lda * 2, 64, BM)) return 1;
    if (make_tmap_2d(&tmB, a->B, static_cast<uint64_t>(a->b_cols), static_cast<uint64_t>(a->b_rows), a->ldb * 2, 64, bn)) return 1;
    grid = dim3(static_cast<unsigned>(tiles_m), (a->N + bn - 1) / bn, 1);
  } else if (a->mode == 1) {
    if (make_tmap_2d(&tmA, a->A, static_cast<uint64_t>(a->a_cols), static_cast<uint64_t>(a->a_rows), a->lda * 2, 64, 64)) return 1;
    if (make_tmap_2d(&tmB, a->B, static_cast<uint64_t>(a->b_cols), static_cast<uint64_t>(a->b_rows), a->ldb * 2, 64, 64)) return 1;
    grid = dim3(static_cast<unsigned>(tiles_m), (a->N + bn - 1) / bn, a->taps * kp.splits);
  } else {
    return rb_fail("rb_gemm: mode must be 0 (NT) or 1 (TN)");
  }
  if (a->mode == 0) {
    switch (bn) {
      case 32: return launch_gemm<32, 0, 4>(tmA, tmB, kp, grid, st);
      case 64: return launch_gemm<64, 0, 4>(tmA, tmB, kp, grid, st);
      case 128: return launch_gemm<128, 0, 3>(tmA, tmB, kp, grid, st);
      case 256: return launch_gemm<256, 0, 4>(tmA, tmB, kp, grid, st);
    }
  } else {
    switch (bn) {
      case 64: return launch_gemm<64, 1, 4>(tmA, tmB, kp, grid, st);
      case 128: return launch_gemm<128, 1, 3>(tmA, tmB, kp, grid, st);
      case 256: return launch_gemm<256, 1, 4>(tmA, tmB, kp, grid, st);
    }
  }
  return rb_fail("rb_gemm: unsupported block_n");
}
