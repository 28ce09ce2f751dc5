// ResNet stem in ONE pass: 7x7 / stride 2 / pad 3 convolution 3 -> 64 (+ folded FrozenBN + ReLU) AND the 3x3 / stride 2 / pad 1
// max-pool that follows it; the 64-channel stride-2 map (210 MB at 16 x 640 x 640) never reaches HBM.
// Replaces torchvision ResNet.conv1 / bn1 / relu / maxpool reached from models/modeling/backbone.py:99-102 (conv1 and layer1 are
// frozen there, :30-33, so no gradient passes through the pool).
//
// No im2col is ever built.  A first kernel rewrites the fp32 NCHW batch as 16-bit HWC4 pixels (8 bytes, 4th channel zero).  In that
// layout the 7 x 3 patch row that output pixel ox reads on tap row r is 28 CONSECUTIVE 16-bit values starting 16 bytes after the one
// of pixel ox-1 (stride 2 pixels x 8 bytes).  A K-major, un-swizzled UMMA operand is exactly that: 8-row x 16-byte core matrices whose
// rows are 16 bytes apart -- so with LBO = 16 B (next K chunk = next 16 bytes of the SAME bytes) and SBO = 128 B the descriptor
// reads the overlapping windows straight out of an input row that TMA dropped into shared memory, and one output row segment of
// 128 pixels is 7 taps x 2 tcgen05.mma (128 x 64 x 16; K = 8 pixels x 4 channels per tap row, the 8th pixel / 4th channel meet zero
// weights).  Out-of-image rows and columns are TMA's zero fill (signed coordinates).
//
// One CTA = (image, segment of 126 conv columns = 63 pooled columns, chunk of pooled rows); it walks down its conv rows:
//   warp 0      producer: two new input rows per conv row into a 16-slot ring (TMA)
//   warp 1      MMA: 14 tcgen05.mma per conv row into one of two 64-column TMEM accumulators
//   warps 2-5   TMEM -> + bias, ReLU, 16 bit -> conv-row ring in shared memory (4 rows; pixels outside the image become 0, which
//               is what the pool's padding is worth after a ReLU)
//   warps 6-9   pool rows 2k, 2k+1, 2k+2 of the ring -> pooled row k -> global (padded NHWC [B, H2+2, W2+2, 64], zero border)
#include "common.cuh"
#include "host.h"

namespace rb {

constexpr int SP_SEG = 126;                  // conv columns a segment advances by (63 pooled columns)
constexpr int SP_PX_A = 128, SP_PX_B = 136;  // input pixels per row, loaded as two boxes (2 * 127 + 7 = 261 <= 264)
constexpr int SP_SLOT = 2176;                // input-row slot pitch (264 px x 8 B = 2112, rounded to 128 B)
constexpr int SP_IN_SLOTS = 16;
constexpr int SP_CROWS = 4;                  // conv-row ring
constexpr int SP_CROW_BYTES = 128 * 128;     // 128 pixels x 64 channels x 2 B
constexpr int SP_W_BYTES = 7 * 4096;         // packed weights: [tap row 7][k chunk 4][n 64][8 x 16 bit]
constexpr int SP_THREADS = 320;
constexpr int SP_OFF_W = SP_IN_SLOTS * SP_SLOT;
constexpr int SP_OFF_CROW = SP_OFF_W + SP_W_BYTES;
constexpr int SP_OFF_BIAS = SP_OFF_CROW + SP_CROWS * SP_CROW_BYTES;
constexpr int SP_OFF_BARS = SP_OFF_BIAS + 256;
constexpr int SP_SMEM = SP_OFF_BARS + 512 + 1024;

// fp32 NCHW -> 16-bit HWC4 (channel 3 = 0), rows of W + 2 pixels: pixel x sits in column x + 1, columns 0 and W + 1 are zero.  (The
// shift makes the first pixel a segment stages, 2 * c0 - 3, land on an even column: TMA box starts stay 16-byte aligned.)
__global__ void __launch_bounds__(256) img_to_hwc4_kernel(const float* __restrict__ img, uint2* __restrict__ out, int H, int W) {
  const int col = blockIdx.x * 256 + threadIdx.x;   // grid: x = column chunks, y = row, z = image
  if (col >= W + 2) return;
  const int y = blockIdx.y, b = blockIdx.z;
  uint2 o = make_uint2(0, 0);
  if (col >= 1 && col <= W) {
    const long long HW = static_cast<long long>(H) * W;
    const float* src = img + static_cast<long long>(b) * 3 * HW + static_cast<long long>(y) * W + (col - 1);
    o.x = pack_t2(__ldg(src), __ldg(src + HW));
    o.y = pack_t2(__ldg(src + 2 * HW), 0.f);
  }
  out[(static_cast<long long>(b) * H + y) * (W + 2) + col] = o;
}

__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
#if defined(RB_ACT_BF16)
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
#else
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
#endif
  return *reinterpret_cast<uint32_t*>(&r);
}

__global__ void __launch_bounds__(SP_THREADS, 1)
stem_pool_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const uint4* __restrict__ wpk,
                 const float* __restrict__ bias, uint4* __restrict__ out, int H1, int W1, int H2, int W2, int rows_per_chunk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_in = smem;
  uint8_t* s_w = smem + SP_OFF_W;
  uint8_t* s_crow = smem + SP_OFF_CROW;
  float* s_bias = reinterpret_cast<float*>(smem + SP_OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SP_OFF_BARS);
  uint64_t* in_full = bars;            // [16]
  uint64_t* mma_done = bars + 16;      // [8]
  uint64_t* acc_full = bars + 24;      // [2]
  uint64_t* acc_empty = bars + 26;     // [2]
  uint64_t* crow_full = bars + 28;     // [4]
  uint64_t* crow_free = bars + 32;     // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int seg = blockIdx.x, chunk = blockIdx.y, b = blockIdx.z;
  const int p0 = chunk * rows_per_chunk;
  const int p1 = (p0 + rows_per_chunk < H2) ? p0 + rows_per_chunk : H2;
  const int n_pool = p1 - p0;
  if (n_pool <= 0) return;                       // (whole CTA: nothing allocated yet)
  const int n_rows = 2 * n_pool + 1;             // conv rows 2*p0-1 .. 2*p1-1
  const int oy0 = 2 * p0 - 1;
  const int c0 = SP_SEG * seg - 1;               // first conv column of the segment (local column t <-> conv column c0 + t)
  const int x0 = 2 * c0 - 3 + 1;                 // first input pixel staged (2 * c0 - 3) as a column of the shifted HWC4 rows: even
  const int y0 = 2 * oy0 - 3;                    // input row of ring index 0

  if (tid == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < SP_IN_SLOTS; ++i) mbar_init(&in_full[i], 1);
    for (int i = 0; i < 8; ++i) mbar_init(&mma_done[i], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    for (int i = 0; i < SP_CROWS; ++i) { mbar_init(&crow_full[i], 128); mbar_init(&crow_free[i], 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  // packed weights + bias -> shared memory (generic proxy), visible to the tensor core after the proxy fence
  for (int i = tid; i < SP_W_BYTES / 16; i += SP_THREADS) reinterpret_cast<uint4*>(s_w)[i] = __ldg(wpk + i);
  if (tid < 64) s_bias[tid] = __ldg(bias + tid);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------------------------------------- producer
    if (lane == 0) {
      const int n_in = 2 * (n_rows - 1) + 7;     // ring indices 0 .. n_in-1 (input rows y0 + yy)
      for (int yy = 0; yy < n_in; ++yy) {
        const int slot = yy & (SP_IN_SLOTS - 1);
        if (yy >= SP_IN_SLOTS) {                 // the slot's previous row (yy-16) was last read by conv row (yy-16)/2
          const int il = (yy - SP_IN_SLOTS) >> 1;
          mbar_wait(&mma_done[il & 7], (il >> 3) & 1);
        }
        uint8_t* dst = s_in + slot * SP_SLOT;
        mbar_expect_tx(&in_full[slot], (SP_PX_A + SP_PX_B) * 8);
        tma_load_3d(dst, &tmA, &in_full[slot], x0, y0 + yy, b);
        tma_load_3d(dst + SP_PX_A * 8, &tmB, &in_full[slot], x0 + SP_PX_A, y0 + yy, b);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_t(128, 64, 0, 0);
      const uint32_t in_base = smem_u32(s_in), w_base = smem_u32(s_w);
      for (int i = 0; i < n_rows; ++i) {
        const int first = i == 0 ? 0 : 5;
        for (int r = first; r < 7; ++r) {
          const int yy = 2 * i + r;
          mbar_wait(&in_full[yy & (SP_IN_SLOTS - 1)], (yy >> 4) & 1);
        }
        mbar_wait(&acc_empty[i & 1], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem + (i & 1) * 64;
#pragma unroll
        for (int r = 0; r < 7; ++r) {
          const uint32_t a_row = in_base + ((2 * i + r) & (SP_IN_SLOTS - 1)) * SP_SLOT;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            umma_f16_ss(d, umma_smem_desc(a_row + kk * 32, 16, 128, SWZ_NONE), umma_smem_desc(w_base + r * 4096 + kk * 2048, 1024, 128, SWZ_NONE), idesc,
                        (r | kk) != 0);
        }
        umma_commit(&acc_full[i & 1]);
        umma_commit(&mma_done[i & 7]);
      }
    }
  } else if (warp < 6) {
    // ---------------------------------------------------------------------------------------------- TMEM -> conv-row ring
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int t = q * 32 + lane;                  // local conv column
    const bool col_ok = (c0 + t) >= 0 && (c0 + t) < W1;
    for (int i = 0; i < n_rows; ++i) {
      const int oy = oy0 + i;
      const bool ok = col_ok && oy >= 0 && oy < H1;
      if (i >= SP_CROWS) mbar_wait(&crow_free[i & (SP_CROWS - 1)], ((i >> 2) - 1) & 1);
      mbar_wait(&acc_full[i & 1], (i >> 1) & 1);
      tc_fence_after();
      uint8_t* dst = s_crow + (i & (SP_CROWS - 1)) * SP_CROW_BYTES + t * 128;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld_32x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + (i & 1) * 64 + half * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b0 = *reinterpret_cast<const float4*>(s_bias + half * 32 + 8 * j);
          const float4 b1 = *reinterpret_cast<const float4*>(s_bias + half * 32 + 8 * j + 4);
          uint4 o = make_uint4(0, 0, 0, 0);
          if (ok) {
            o.x = pack_t2(fmaxf(__uint_as_float(v[8 * j + 0]) + b0.x, 0.f), fmaxf(__uint_as_float(v[8 * j + 1]) + b0.y, 0.f));
            o.y = pack_t2(fmaxf(__uint_as_float(v[8 * j + 2]) + b0.z, 0.f), fmaxf(__uint_as_float(v[8 * j + 3]) + b0.w, 0.f));
            o.z = pack_t2(fmaxf(__uint_as_float(v[8 * j + 4]) + b1.x, 0.f), fmaxf(__uint_as_float(v[8 * j + 5]) + b1.y, 0.f));
            o.w = pack_t2(fmaxf(__uint_as_float(v[8 * j + 6]) + b1.z, 0.f), fmaxf(__uint_as_float(v[8 * j + 7]) + b1.w, 0.f));
          }
          const int ch = half * 4 + j;            // 16-byte chunk of the pixel's 128 bytes, XOR-swizzled against bank conflicts
          *reinterpret_cast<uint4*>(dst + ((ch ^ (t & 7)) << 4)) = o;
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[i & 1]);
      mbar_arrive(&crow_full[i & (SP_CROWS - 1)]);
    }
  } else {
    // ---------------------------------------------------------------------------------------------- pool + store
    const int pt = tid - 192;                     // 0..127
    const int Wp = W2 + 2;
    const int n_q = (W2 - 63 * seg) < 63 ? (W2 - 63 * seg) : 63;   // pooled columns of this segment
    for (int k = 0; k < n_pool; ++k) {
      for (int d = (k == 0 ? 0 : 1); d < 3; ++d) {
        const int i = 2 * k + d;
        mbar_wait(&crow_full[i & (SP_CROWS - 1)], (i >> 2) & 1);
      }
      const uint8_t* r0 = s_crow + ((2 * k) & (SP_CROWS - 1)) * SP_CROW_BYTES;
      const uint8_t* r1 = s_crow + ((2 * k + 1) & (SP_CROWS - 1)) * SP_CROW_BYTES;
      const uint8_t* r2 = s_crow + ((2 * k + 2) & (SP_CROWS - 1)) * SP_CROW_BYTES;
      uint4* orow = out + ((static_cast<long long>(b) * (H2 + 2) + (p0 + k + 1)) * Wp + (63 * seg + 1)) * 8;
      for (int it = pt; it < n_q * 8; it += 128) {
        const int qq = it >> 3, ch = it & 7;
        uint4 m = make_uint4(0, 0, 0, 0);         // post-ReLU values are >= 0
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int t = 2 * qq + dx;
          const int off = t * 128 + ((ch ^ (t & 7)) << 4);
          const uint4 a = *reinterpret_cast<const uint4*>(r0 + off);
          const uint4 c = *reinterpret_cast<const uint4*>(r1 + off);
          const uint4 e = *reinterpret_cast<const uint4*>(r2 + off);
          m.x = hmax2_u32(m.x, hmax2_u32(a.x, hmax2_u32(c.x, e.x)));
          m.y = hmax2_u32(m.y, hmax2_u32(a.y, hmax2_u32(c.y, e.y)));
          m.z = hmax2_u32(m.z, hmax2_u32(a.z, hmax2_u32(c.z, e.z)));
          m.w = hmax2_u32(m.w, hmax2_u32(a.w, hmax2_u32(c.w, e.w)));
        }
        orow[it] = m;
      }
      // zero border cells of this padded output row
      if (seg == 0 && pt < 8) out[((static_cast<long long>(b) * (H2 + 2) + (p0 + k + 1)) * Wp) * 8 + pt] = make_uint4(0, 0, 0, 0);
      if (63 * seg + n_q == W2 && pt >= 8 && pt < 16)
        out[((static_cast<long long>(b) * (H2 + 2) + (p0 + k + 1)) * Wp + (W2 + 1)) * 8 + (pt - 8)] = make_uint4(0, 0, 0, 0);
      mbar_arrive(&crow_free[(2 * k) & (SP_CROWS - 1)]);
      mbar_arrive(&crow_free[(2 * k + 1) & (SP_CROWS - 1)]);
    }
    // top / bottom border rows (the segment's share of their columns, border corners included by the first / last segment)
    const int cbeg = seg == 0 ? 0 : 63 * seg + 1;
    const int cend = (63 * seg + n_q == W2) ? W2 + 2 : 63 * seg + n_q + 1;
    if (p0 == 0)
      for (int it = pt; it < (cend - cbeg) * 8; it += 128) out[((static_cast<long long>(b) * (H2 + 2)) * Wp + cbeg) * 8 + it] = make_uint4(0, 0, 0, 0);
    if (p1 == H2)
      for (int it = pt; it < (cend - cbeg) * 8; it += 128)
        out[((static_cast<long long>(b) * (H2 + 2) + (H2 + 1)) * Wp + cbeg) * 8 + it] = make_uint4(0, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem);
}

}  // namespace rb

using namespace rb;

extern "C" int rb_stem_pool(const float* img, const void* wpk, const float* bias, void* hwc4, void* out, int B, int H, int W, int H1, int W1, int H2,
                            int W2, void* stream) {
  if (!img || !wpk || !bias || !hwc4 || !out) return rb_fail("rb_stem_pool: null pointer");
  if (B <= 0 || H <= 0 || W <= 0) return rb_fail("rb_stem_pool: empty batch");
  if (H1 != (H + 6 - 7) / 2 + 1 || W1 != (W + 6 - 7) / 2 + 1 || H2 != (H1 + 2 - 3) / 2 + 1 || W2 != (W1 + 2 - 3) / 2 + 1)
    return rb_fail("rb_stem_pool: inconsistent output geometry");
  if (W % 2) return rb_fail("rb_stem_pool: the image width must be even (16-byte row pitch of the HWC4 copy)");
  if ((reinterpret_cast<uintptr_t>(hwc4) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(wpk) & 15))
    return rb_fail("rb_stem_pool: buffers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (H > 65535 || B > 65535) return rb_fail("rb_stem_pool: H and B must fit a grid dimension");
  img_to_hwc4_kernel<<<dim3((W + 2 + 255) / 256, H, B), 256, 0, st>>>(img, static_cast<uint2*>(hwc4), H, W);
  RB_CUDA(cudaGetLastError());
  CUtensorMap tmA, tmB;
  const uint64_t Wc = static_cast<uint64_t>(W) + 2;
  if (make_tmap_3d_px8(&tmA, hwc4, Wc, H, B, Wc * 8, static_cast<uint64_t>(H) * Wc * 8, SP_PX_A)) return 1;
  if (make_tmap_3d_px8(&tmB, hwc4, Wc, H, B, Wc * 8, static_cast<uint64_t>(H) * Wc * 8, SP_PX_B)) return 1;
  static bool cfg = false;
  if (!cfg) { RB_CUDA(cudaFuncSetAttribute(stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM)); cfg = true; }
  const int n_seg = (W2 + 62) / 63;
  int chunks = sm_count() / (B * n_seg);
  if (chunks < 1) chunks = 1;
  if (chunks > H2) chunks = H2;
  const int rows_per_chunk = (H2 + chunks - 1) / chunks;
  chunks = (H2 + rows_per_chunk - 1) / rows_per_chunk;
  stem_pool_kernel<<<dim3(n_seg, chunks, B), SP_THREADS, SP_SMEM, st>>>(tmA, tmB, static_cast<const uint4*>(wpk), bias, static_cast<uint4*>(out), H1, W1, H2, W2,
                                                                        rows_per_chunk);
  RB_CUDA(cudaGetLastError());
  return 0;
}
