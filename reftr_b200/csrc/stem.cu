// ResNet stem: 7x7 / stride 2 / pad 3 convolution 3 -> 64 (+ folded FrozenBN + ReLU) as ONE kernel: the im2col tile is built in
// shared memory (never in HBM) and contracted on the tensor cores (tcgen05.mma, accumulator in TMEM).
// Replaces torchvision ResNet.conv1 / bn1 / relu reached from models/modeling/backbone.py:99-102.
//
// One CTA = 8 x 16 output pixels of one image (128 GEMM rows).  128 threads, no warp specialisation; 2-3 CTAs are resident per SM
// and overlap each other's phases:
//   1. thread 0: TMA-load the packed weights [64, 160] (three 64-column K blocks, 128B swizzle; columns >= 160 read as zero)
//   2. all: stage the fp32 input patch [3][21][37] in shared memory (zero outside the image)
//   3. thread r builds row r of the im2col tile A [128, 192] bf16 directly in the 128B-swizzled K-major layout UMMA expects
//      (column (r*7+s)*3+c, the order rb_pack_conv uses)
//   4. thread 0: 12 x tcgen05.mma (128 x 64 x 16) -> TMEM; commit
//   5. all: tcgen05.ld, + bias, ReLU, bf16, 128-byte row stores (NHWC, one pixel per thread)
#include "common.cuh"
#include "host.h"

namespace rb {

constexpr int ST_TH = 8, ST_TW = 16;            // output tile
constexpr int ST_PH = 2 * ST_TH + 5, ST_PW = 2 * ST_TW + 5, ST_PWP = ST_PW + 2;  // input patch 21 x 37 (+pad)
constexpr int ST_K = 147, ST_KB = 3;            // 147 real columns in three 64-wide K blocks
constexpr int ST_SMEM = ST_KB * 16384 + ST_KB * 8192 + 3 * ST_PH * ST_PWP * 4 + 64 + 1024;

__global__ void __launch_bounds__(128)
stem_conv_kernel(const __grid_constant__ CUtensorMap tmW, const float* __restrict__ img, const float* __restrict__ bias, rb_t* __restrict__ out,
                 int H, int W, int H1, int W1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                  // 3 x [128 x 64] bf16, SW128
  uint8_t* sW = smem + ST_KB * 16384;                  // 3 x [64 x 64] bf16, SW128
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + ST_KB * 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  float* patch = reinterpret_cast<float*>(sW + ST_KB * 8192 + 64);  // [3][21][39]
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.z, oy0 = blockIdx.y * ST_TH, ox0 = blockIdx.x * ST_TW;

  if (tid == 0) {
    tma_prefetch_desc(&tmW);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<64>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(&bars[0], ST_KB * 8192);
#pragma unroll
    for (int kb = 0; kb < ST_KB; ++kb) tma_load_2d(sW + kb * 8192, &tmW, &bars[0], kb * 64, 0);
  }
  // ---- input patch -------------------------------------------------------------------------------------------------
  const int iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;
  constexpr int PATCH_N = 3 * ST_PH * ST_PW, PATCH_IT = (PATCH_N + 127) / 128;
  float pv[PATCH_IT];
#pragma unroll
  for (int it = 0; it < PATCH_IT; ++it) {  // all loads in flight before the first store
    const int i = tid + it * 128;
    const int c = i / (ST_PH * ST_PW), rem = i - c * (ST_PH * ST_PW);
    const int py = rem / ST_PW, px = rem - py * ST_PW;
    const int iy = iy0 + py, ix = ix0 + px;
    pv[it] = 0.f;
    if (i < PATCH_N && iy >= 0 && iy < H && ix >= 0 && ix < W) pv[it] = __ldg(img + ((static_cast<long long>(b) * 3 + c) * H + iy) * W + ix);
  }
#pragma unroll
  for (int it = 0; it < PATCH_IT; ++it) {
    const int i = tid + it * 128;
    if (i < PATCH_N) {
      const int c = i / (ST_PH * ST_PW), rem = i - c * (ST_PH * ST_PW);
      const int py = rem / ST_PW, px = rem - py * ST_PW;
      patch[(c * ST_PH + py) * ST_PWP + px] = pv[it];
    }
  }
  __syncthreads();
  // ---- im2col row of this thread, written in the swizzled layout -------------------------------------------------------
  const int ly = tid / ST_TW, lx = tid - ly * ST_TW;
  const float* pbase = patch + (2 * ly) * ST_PWP + 2 * lx;
#pragma unroll
  for (int kb = 0; kb < ST_KB; ++kb) {
    uint8_t* row = sA + kb * 16384 + tid * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = kb * 64 + j * 8 + e;  // compile-time after unrolling
        if (k < ST_K) {
          const int r = k / 21, s = (k % 21) / 3, c = k % 3;
          v[e] = pbase[(c * ST_PH + r) * ST_PWP + s];
        } else {
          v[e] = 0.f;
        }
      }
      uint4 t;
      t.x = pack_t2(v[0], v[1]); t.y = pack_t2(v[2], v[3]); t.z = pack_t2(v[4], v[5]); t.w = pack_t2(v[6], v[7]);
      *reinterpret_cast<uint4*>(row + ((j ^ (tid & 7)) << 4)) = t;
    }
  }
  fence_proxy_async_smem();
  __syncthreads();
  // ---- MMA ------------------------------------------------------------------------------------------------------------
  if (tid == 0) {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_t(128, 64, 0, 0);
#pragma unroll
    for (int kb = 0; kb < ST_KB; ++kb) {
      const uint32_t a_base = smem_u32(sA + kb * 16384), b_base = smem_u32(sW + kb * 8192);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_f16_ss(tmem, umma_smem_desc(a_base + kk * 32, 16, 1024, SWZ_128B), umma_smem_desc(b_base + kk * 32, 16, 1024, SWZ_128B), idesc, (kb | kk) != 0);
    }
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  // ---- epilogue -------------------------------------------------------------------------------------------------------
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  const int oy = oy0 + ly, ox = ox0 + lx;
  const bool ok = oy < H1 && ox < W1;
  rb_t* orow = out + ((static_cast<long long>(b) * H1 + oy) * W1 + ox) * 64;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + lane_addr + half * 32, v);
    tmem_ld_wait();
    if (ok) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(v[8 * j + e]) + __ldg(bias + half * 32 + 8 * j + e), 0.f);
        uint4 t;
        t.x = pack_t2(f[0], f[1]); t.y = pack_t2(f[2], f[3]); t.z = pack_t2(f[4], f[5]); t.w = pack_t2(f[6], f[7]);
        *reinterpret_cast<uint4*>(orow + half * 32 + 8 * j) = t;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

}  // namespace rb

using namespace rb;

extern "C" int rb_stem_conv(const float* img, const void* wf, int ldk, const float* bias, void* out, int B, int H, int W, int H1, int W1, void* stream) {
  if (ldk < ST_K || (ldk % 8)) return rb_fail("rb_stem_conv: packed weight pitch must be >= 147 and a multiple of 8 (got %d)", ldk);
  CUtensorMap tmW;
  if (make_tmap_2d(&tmW, wf, static_cast<uint64_t>(ldk), 64, static_cast<uint64_t>(ldk) * 2, 64, 64)) return 1;
  static bool cfg = false;
  if (!cfg) { RB_CUDA(cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM)); cfg = true; }
  dim3 grid((W1 + ST_TW - 1) / ST_TW, (H1 + ST_TH - 1) / ST_TH, B);
  stem_conv_kernel<<<grid, 128, ST_SMEM, static_cast<cudaStream_t>(stream)>>>(tmW, img, bias, static_cast<rb_t*>(out), H, W, H1, W1);
  RB_CUDA(cudaGetLastError());
  return 0;
}
