/* reftr_b200 -- C ABI of the B200 (sm_100a) kernels behind the RefTR forward/backward hot path.
 *
 * The reference (ubc-vision/RefTR) is pure Python/PyTorch and has no FFI; every function below replaces
 * a group of ATen calls on the reference's hot path (file:line given per entry, paths relative to the
 * reference root).  All pointers are DEVICE pointers unless stated otherwise; `stream` is a cudaStream_t
 * passed as void*.  Every function returns 0 on success, non-zero on failure; rb_last_error() gives text.
 * No function allocates device memory, synchronises the device or touches the host heap after return,
 * so all of them are CUDA-graph capturable.  bf16 = __nv_bfloat16 (2 bytes), row-major everywhere.
 *
 * Activation layouts
 *   "padded NHWC": [N, H+2, W+2, C] bf16 with a one-pixel ZERO border; viewed as a matrix [R, C] with
 *                  R = N*(H+2)*(W+2).  A 3x3/stride-1 convolution tap (r,s) is then the same matrix with the
 *                  row index shifted by (r-1)*(W+2)+(s-1), so every convolution is a GEMM over shifted rows.
 *   "parity planes": for stride-2 blocks, 4 planes [4, N, Ho+2, Wo+2, C]; plane (p,q) cell (u,v) holds
 *                  padded-input pixel (2u+p, 2v+q).  Stride-2 taps become constant row shifts again.
 */
#ifndef REFTR_B200_H
#define REFTR_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* rb_last_error(void);
int rb_version(void);

/* Row geometry for border masking in GEMM epilogues (rows that are padding must be written as zero). */
typedef struct {
  int mode; /* 0 none; 1 padded NHWC grid; 2 parity planes */
  int Wp;   /* pixels per padded row (W+2, or Wo+2 for planes) */
  int HpWp; /* pixels per padded image */
  int H, W; /* interior size (Ho, Wo for planes) */
  int Rs;   /* rows per plane (mode 2) */
} rb_geom;

/* Generic tensor-core GEMM / implicit-GEMM convolution (tcgen05.mma, TMA operands, fp32 accumulate in TMEM).
 *
 * mode 0 ("NT"): D[m, n] = sum_tap sum_k A[m + a_rowoff[tap], k] * B[n, b_koff[tap] + k]
 *      A: [a_rows, K] bf16 (lda elements per row), B: [N, ...] bf16 (ldb), k in [0, K).  taps=1 is a plain
 *      linear layer x @ W^T (replaces F.linear at transformer.py:176-178, reftr_transformer.py:14-23, ...);
 *      taps=9 is a 3x3 convolution over padded NHWC (replaces torchvision Bottleneck conv2d, backbone.py:99-102).
 * mode 1 ("TN"): D[m, n] = sum_r A[r + a_rowoff[z], m] * B[r + b_rowoff[z], n], r in [0, K) ; used for
 *      weight gradients (contraction over pixels / tokens), one z per tap, split-K over `splits` CTAs with
 *      fp32 atomic accumulation into out32 + z * out32_z_stride (replaces autograd's conv/linear wgrad).
 * Epilogue (mode 0, and mode 1 when atomic=0): v = acc + bias[n] + res[row, n] + res32[row, n];
 *      relu; v = mask_src[row, n] > 0 ? v : 0; rows that are padding per `geom` -> 0; written as bf16 (out)
 *      and/or fp32 (out32).  row = m + out_row_off for out/res/mask_src addressing and geometry.
 */
typedef struct {
  int mode;
  const void* A; long long a_rows; int a_cols; long long lda;
  const void* B; long long b_rows; int b_cols; long long ldb;
  int M, N, K;
  int taps;
  int a_rowoff[16];
  int b_koff[16]; /* mode 0: k offset into B per tap; mode 1: row offset into B per z */
  int splits;     /* mode 1 */
  int block_n;    /* 0 = auto, else 32/64/128/256 */
  long long out_row_off;
  const float* bias;
  const void* res; long long ldres;
  const float* res32; long long ldres32;
  const void* mask_src; long long ldmask;
  void* out; long long ldo;
  float* out32; long long ldo32;
  long long out32_z_stride;
  int relu;
  int atomic;
  rb_geom geom;
} rb_gemm_args;

int rb_gemm(const rb_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif
