// Small row-mapped glue kernels of the VL transformer (fp32, 256-wide or narrower rows): gather / add / concat and
// scatter-add.  They replace the torch indexing, cat, repeat and broadcast adds of QueryEncoder.forward
// (reftr_transformer.py:41-66) and of VLTransformer.encode's memory slicing (reftr_transformer.py:261-267).
#include "common.cuh"
#include "host.h"

namespace rb {

struct RowMap4 {  // index(r) = (r / group) * stride + (r % group) * inner + offset   (group == 0: r + offset)
  int group, stride, inner, offset;
  __device__ __forceinline__ long long operator()(long long r) const {
    return group ? (r / group) * static_cast<long long>(stride) + (r % group) * static_cast<long long>(inner) + offset : r + offset;
  }
};

static RowMap4 host_map(const int* m) {
  RowMap4 r{0, 0, 0, 0};
  if (m) { r.group = m[0]; r.stride = m[1]; r.inner = m[2]; r.offset = m[3]; }
  return r;
}

// y[my(r), c] = a[ma(r), c] + b[mb(r), c]; one thread per 4 columns
__global__ void rows_add_kernel(const float* __restrict__ a, long long lda, RowMap4 ma, const float* __restrict__ b, long long ldb, RowMap4 mb,
                                float* __restrict__ y32, long long ldy, rb_t* __restrict__ yb, long long ldyb, RowMap4 my, long long rows,
                                int D4) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * D4) return;
  const long long r = idx / D4;
  const int c = static_cast<int>(idx - r * D4) * 4;
  float4 v = *reinterpret_cast<const float4*>(a + ma(r) * lda + c);
  if (b) {
    const float4 w = *reinterpret_cast<const float4*>(b + mb(r) * ldb + c);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  const long long o = my(r);
  if (y32) *reinterpret_cast<float4*>(y32 + o * ldy + c) = v;
  if (yb) {
    uint2 t;
    t.x = pack_t2(v.x, v.y); t.y = pack_t2(v.z, v.w);
    *reinterpret_cast<uint2*>(yb + o * ldyb + c) = t;
  }
}

// dst[md(r), c] += src[ms(r), c]  (fp32 atomics: several r may map to one destination row)
__global__ void rows_scatter_add_kernel(const float* __restrict__ src, long long lds, RowMap4 ms, float* __restrict__ dst, long long ldd, RowMap4 md,
                                        long long rows, int D) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * D) return;
  const long long r = idx / D;
  const int c = static_cast<int>(idx - r * D);
  atomicAdd(dst + md(r) * ldd + c, src[ms(r) * lds + c]);
}

}  // namespace rb

using namespace rb;

extern "C" int rb_rows_add(const float* a, long long lda, const int* map_a, const float* b, long long ldb, const int* map_b, float* y32, long long ldy,
                           void* yb, long long ldyb, const int* map_y, long long rows, int D, void* stream) {
  if (rows <= 0) return 0;
  if (D % 4) return rb_fail("rb_rows_add: D must be a multiple of 4");
  if ((lda % 4) || (b && (ldb % 4)) || (y32 && (ldy % 4)) || (yb && (ldyb % 4))) return rb_fail("rb_rows_add: pitches must be multiples of 4 elements");
  const long long total = rows * (D / 4);
  rows_add_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, lda, host_map(map_a), b, ldb, host_map(map_b), y32, ldy, static_cast<rb_t*>(yb), ldyb, host_map(map_y), rows, D / 4);
  RB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb_rows_scatter_add(const float* src, long long lds, const int* map_src, float* dst, long long ldd, const int* map_dst, long long rows,
                                   int D, void* stream) {
  if (rows <= 0) return 0;
  const long long total = rows * D;
  rows_scatter_add_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, lds, host_map(map_src), dst, ldd,
                                                                                                                 host_map(map_dst), rows, D);
  RB_CUDA(cudaGetLastError());
  return 0;
}
