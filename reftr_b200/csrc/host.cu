#include "host.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace rb {

static thread_local char g_err[512] = "";

int rb_fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    // resolved through the runtime so that the library does not link against libcuda.so (absent on CPU-only build boxes)
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_tmap_2d_impl(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t pitch_bytes, uint32_t box_inner,
                             uint32_t box_rows, CUtensorMapDataType dt, uint32_t esize) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return rb_fail("cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const uint32_t row_bytes = box_inner * esize;
  CUtensorMapSwizzle sw = row_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row_bytes >= 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return rb_fail("cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu pitch=%llu box=%ux%u esize=%u ptr=%p", static_cast<int>(r),
                   static_cast<unsigned long long>(inner), static_cast<unsigned long long>(rows),
                   static_cast<unsigned long long>(pitch_bytes), box_inner, box_rows, esize, ptr);
  return 0;
}

int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t pitch_bytes, uint32_t box_inner,
                 uint32_t box_rows) {
  return make_tmap_2d_impl(out, ptr, inner, rows, pitch_bytes, box_inner, box_rows, rb::RB_ACT_DTYPE ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2);
}

int make_tmap_2d_f32(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t pitch_bytes, uint32_t box_inner,
                     uint32_t box_rows) {
  return make_tmap_2d_impl(out, ptr, inner, rows, pitch_bytes, box_inner, box_rows, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4);
}

int make_tmap_3d_px8(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t imgs, uint64_t pitch_row_bytes,
                     uint64_t pitch_img_bytes, uint32_t box_inner) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return rb_fail("cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t dims[3] = {inner, rows, imgs};
  cuuint64_t strides[2] = {pitch_row_bytes, pitch_img_bytes};
  cuuint32_t box[3] = {box_inner, 1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return rb_fail("cuTensorMapEncodeTiled (3D px8) failed (%d): inner=%llu rows=%llu imgs=%llu box=%u ptr=%p", static_cast<int>(r),
                   static_cast<unsigned long long>(inner), static_cast<unsigned long long>(rows), static_cast<unsigned long long>(imgs), box_inner, ptr);
  return 0;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace rb

extern "C" const char* rb_last_error(void) { return rb::g_err; }
extern "C" int rb_version(void) { return 2; }
extern "C" int rb_act_dtype(void) { return rb::RB_ACT_DTYPE; }
