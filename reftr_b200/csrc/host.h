// Host-side helpers shared by the C-ABI translation units: error reporting and TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/reftr_b200.h"
#include "common.cuh"

namespace rb {
int rb_fail(const char* fmt, ...);
// 2D bf16 row-major tensor [rows, inner] with `pitch_bytes` per row; box = [box_rows, box_inner]; 128B swizzle
// (box_inner*2 must be <= 128); out-of-bounds elements (including negative coordinates) read as zero.
int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t pitch_bytes, uint32_t box_inner,
                 uint32_t box_rows);
// same for an fp32 tensor (box_inner * 4 bytes <= 128; swizzle chosen by the box row bytes: 128 -> 128B, 64 -> 64B, else 32B)
int make_tmap_2d_f32(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t pitch_bytes, uint32_t box_inner,
                     uint32_t box_rows);
// 3D tensor [imgs, rows, inner] of 8-byte elements (one 4-channel 16-bit pixel each), no swizzle; box = [1, 1, box_inner];
// out-of-bounds elements (including negative coordinates) read as zero
int make_tmap_3d_px8(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t imgs, uint64_t pitch_row_bytes,
                     uint64_t pitch_img_bytes, uint32_t box_inner);
int sm_count();
// host rb_dropout (nullable) -> kernel parameter; off when NULL / seed == NULL / p <= 0
inline DropK make_dropk(const rb_dropout* d) {
  DropK k;
  k.seed = nullptr; k.site = 0; k.thr = 0; k.scale = 1.f;
  if (d && d->seed && d->p > 0.f) {
    uint32_t thr = static_cast<uint32_t>(d->p * 65536.f + 0.5f);
    if (thr > 65535u) thr = 65535u;
    k.seed = static_cast<const unsigned long long*>(d->seed);
    k.site = static_cast<uint32_t>(d->site);
    k.thr = thr;
    k.scale = 65536.f / static_cast<float>(65536u - thr);
  }
  return k;
}
// skinny (M <= 128) linear layers on mma.sync (gemm_skinny.cu); rb_gemm dispatches to it
bool gemm_skinny_eligible(const rb_gemm_args* a);
int gemm_skinny_launch(const rb_gemm_args* a, const DropK& drop, int drop_wpr, cudaStream_t st);
// SIMT attention (attention.cu); the C-ABI entry points in attention_tc.cu fall back to these for short query counts
int attn_fwd_simt(const void* Q, const void* K, const void* V, const void* kpm, void* O, float* LSE, int B, int H, int dh, int Tq, int Sk, long long ldq,
                  long long ldk, long long ldv, long long ldo, float scale, const rb_dropout* drop, void* stream);
int attn_bwd_simt(const void* Q, const void* K, const void* V, const void* kpm, const void* O, const void* dO, const float* LSE, void* dQ, void* dK,
                  void* dV, float* Dbuf, int B, int H, int dh, int Tq, int Sk, long long ldq, long long ldk, long long ldv, long long ldo,
                  long long lddo, long long lddq, long long lddk, long long lddv, float scale, const rb_dropout* drop, void* stream);
}  // namespace rb

#define RB_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) return ::rb::rb_fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
