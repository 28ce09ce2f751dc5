// Skinny linear layers: M <= 128 rows (the decoder's 1..16 queries per sample, the query encoder, the box head, BERT's pooler;
// transformer.py:231-252, reftr_transformer.py:41-66, :287).  These are latency chains, not throughput problems: 16 x 256 x 256
// is 1 MFLOP, and the persistent tcgen05 kernel's fixed cost (TMEM allocation, barrier setup, tensor-map fetch, pipeline fill and
// drain: ~6.5 us in a graph) is the whole run time.  This kernel has no setup at all: one CTA per 8 output columns x 16 rows, its
// 8 warps split K, every lane streams its operand slices straight from L2 with 16-byte loads into mma.sync.m16n8k16 (fp32
// accumulate), partial sums meet in 4 KB of shared memory and 128 threads apply the same epilogue as the big kernel
// (bias / dropout / residuals / ReLU / ReLU-mask; rb_gemm dispatches here, the C ABI is unchanged).
#include "common.cuh"
#include "host.h"

namespace rb {

struct SkinnyParams {
  const rb_t* A; long long lda;
  const rb_t* B; long long ldb;
  int M, N, K;
  const float* bias;
  const rb_t* res; long long ldres;
  const float* res32; long long ldres32;
  const rb_t* mask; long long ldmask;
  rb_t* out; long long ldo;
  float* out32; long long ldo32;
  int relu;
  DropK drop; int drop_gshift; uint32_t drop_wpr;
  float mask_scale;
};

// D[16x8] += A[16x16] * B[16x8]; A row-major, B "column-major" (= rows of the K-major weight matrix), 16-bit inputs, fp32 accumulate
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
#ifdef RB_ACT_BF16
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#else
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#endif
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

constexpr int SK_WARPS = 8;

__global__ void __launch_bounds__(SK_WARPS * 32) gemm_skinny_kernel(const SkinnyParams p) {
  __shared__ float part[SK_WARPS][16 * 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;  // fragment row / column group
  const int m0 = blockIdx.y * 16, n0 = blockIdx.x * 8;
  // The k index inside an MMA is a dummy: any permutation works if A and B use the same one.  Per 32-wide k block, lane (g, q)
  // takes the 8 CONSECUTIVE elements k = 32*blk + 8q .. +7 of its rows with one 16-byte load and feeds them as
  // (k 2q,2q+1 | k 8+2q,8+2q+1) of two MMAs.
  const int ra = min(m0 + g, p.M - 1), rb2 = min(m0 + g + 8, p.M - 1), rn = min(n0 + g, p.N - 1);  // clamped: out-of-range rows are dropped at the end
  const rb_t* a_lo = p.A + static_cast<long long>(ra) * p.lda + q * 8;
  const rb_t* a_hi = p.A + static_cast<long long>(rb2) * p.lda + q * 8;
  const rb_t* b_r = p.B + static_cast<long long>(rn) * p.ldb + q * 8;
  float c[4] = {0.f, 0.f, 0.f, 0.f};
  const int nblk = p.K >> 5;
#pragma unroll 4
  for (int blk = warp; blk < nblk; blk += SK_WARPS) {
    const uint4 al = ldg_nc16(a_lo + blk * 32), ah = ldg_nc16(a_hi + blk * 32), bw = ldg_nc16(b_r + blk * 32);
    mma_16816(c, al.x, ah.x, al.y, ah.y, bw.x, bw.y);
    mma_16816(c, al.z, ah.z, al.w, ah.w, bw.z, bw.w);
  }
  // c0,c1: (row g, cols 2q, 2q+1); c2,c3: (row g+8, same cols)
  part[warp][g * 8 + 2 * q] = c[0];
  part[warp][g * 8 + 2 * q + 1] = c[1];
  part[warp][(g + 8) * 8 + 2 * q] = c[2];
  part[warp][(g + 8) * 8 + 2 * q + 1] = c[3];
  __syncthreads();
  if (threadIdx.x >= 128) return;
  const int r = threadIdx.x >> 3, cc = threadIdx.x & 7;
  const int row = m0 + r, col = n0 + cc;
  if (row >= p.M || col >= p.N) return;
  float v = 0.f;
#pragma unroll
  for (int w = 0; w < SK_WARPS; ++w) v += part[w][threadIdx.x];
  if (p.bias) v += __ldg(p.bias + col);
  if (p.drop.seed) {
    if (p.relu) v = fmaxf(v, 0.f);
    const int e = col >> p.drop_gshift;
    const uint32_t w = drop_word(drop_key(p.drop), static_cast<uint32_t>(row) * p.drop_wpr + static_cast<uint32_t>(e >> 1));
    v = drop_keep(w, e & 1, p.drop.thr) ? v * p.drop.scale : 0.f;
  }
  if (p.res) v += t2f(p.res[static_cast<long long>(row) * p.ldres + col]);
  if (p.res32) v += p.res32[static_cast<long long>(row) * p.ldres32 + col];
  if (p.relu && !p.drop.seed) v = fmaxf(v, 0.f);
  if (p.mask) v = t2f(p.mask[static_cast<long long>(row) * p.ldmask + col]) > 0.f ? v * p.mask_scale : 0.f;
  if (p.out) p.out[static_cast<long long>(row) * p.ldo + col] = f2t(v);
  if (p.out32) p.out32[static_cast<long long>(row) * p.ldo32 + col] = v;
}

bool gemm_skinny_eligible(const rb_gemm_args* a) {
  return a->mode == 0 && a->taps == 1 && a->a_rowoff[0] == 0 && a->b_koff[0] == 0 && !a->atomic && a->geom.mode == 0 && a->out_row_off == 0 &&
         a->M <= 128 && (a->K % 32) == 0 && a->block_n == 0;
}

int gemm_skinny_launch(const rb_gemm_args* a, const DropK& drop, int drop_wpr, cudaStream_t st) {
  SkinnyParams p;
  p.A = static_cast<const rb_t*>(a->A); p.lda = a->lda;
  p.B = static_cast<const rb_t*>(a->B); p.ldb = a->ldb;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bias = a->bias;
  p.res = static_cast<const rb_t*>(a->res); p.ldres = a->ldres;
  p.res32 = a->res32; p.ldres32 = a->ldres32;
  p.mask = static_cast<const rb_t*>(a->mask_src); p.ldmask = a->ldmask;
  p.out = static_cast<rb_t*>(a->out); p.ldo = a->ldo;
  p.out32 = a->out32; p.ldo32 = a->ldo32;
  p.relu = a->relu;
  p.drop = drop; p.drop_gshift = a->drop_gshift; p.drop_wpr = static_cast<uint32_t>(drop_wpr);
  p.mask_scale = a->mask_scale == 0.f ? 1.f : a->mask_scale;
  const dim3 grid((a->N + 7) / 8, (a->M + 15) / 16);
  gemm_skinny_kernel<<<grid, SK_WARPS * 32, 0, st>>>(p);
  RB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rb
