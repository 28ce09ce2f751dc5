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

int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t pitch_bytes, uint32_t box_inner,
                 uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return rb_fail("cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = box_inner * 2 >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : box_inner * 2 >= 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return rb_fail("cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu pitch=%llu box=%ux%u ptr=%p", static_cast<int>(r),
                   static_cast<unsigned long long>(inner), static_cast<unsigned long long>(rows),
                   static_cast<unsigned long long>(pitch_bytes), box_inner, box_rows, ptr);
  return 0;
}

}  // namespace rb

extern "C" const char* rb_last_error(void) { return rb::g_err; }
extern "C" int rb_version(void) { return 1; }
