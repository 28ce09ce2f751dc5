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
/* the 16-bit type "t16" of tensor-core operands, saved activations and activation gradients this library was built for:
 * 0 = IEEE half (default), 1 = bfloat16 (-DRB_ACT_BF16).  Everything called "bf16" below is this type. */
int rb_act_dtype(void);

/* Row geometry for border masking in GEMM epilogues (rows that are padding must be written as zero). */
typedef struct {
  int mode; /* 0 none; 1 padded NHWC grid; 2 parity planes */
  int Wp;   /* pixels per padded row (W+2, or Wo+2 for planes) */
  int HpWp; /* pixels per padded image */
  int H, W; /* interior size of the grid (mode 2: of the FULL-resolution grid the planes were split from) */
  int Rs;   /* rows per plane (mode 2) */
} rb_geom;

/* Counter-based dropout (nn.Dropout / the attention-probability dropout of nn.MultiheadAttention in train mode;
 * transformer.py:151, :154-160, :211-223, reftr_transformer.py:19, HF BertModel).  Nothing is stored: the keep decision of
 * element (row, col) of a site's logical [rows, cols] tensor is recomputed wherever it is needed (forward and backward) as
 *     key  = mix32(lo32(seed) ^ mix32(hi32(seed) + site * 0x9E3779B9))
 *     word = mix32((row * ((cols + 1) / 2) + col / 2) * 0x9E3779B1 + key)
 *     keep = ((word >> 16 * (col & 1)) & 0xFFFF) >= thr,   thr = round(p * 65536);  kept values are scaled by 65536 / (65536 - thr)
 * with mix32(x): x ^= x >> 16; x *= 0x21f0aaad; x ^= x >> 15; x *= 0x735a2d97; x ^= x >> 15.
 * `seed` is a DEVICE pointer to one uint64 that the host rewrites between steps (so a captured CUDA graph draws a new mask
 * on every replay); the struct itself is passed by HOST pointer, NULL (or seed == NULL, or p <= 0) = no dropout. */
typedef struct {
  const void* seed;
  int site;
  float p;
} rb_dropout;

/* Generic tensor-core GEMM / implicit-GEMM convolution (tcgen05.mma, TMA operands, fp32 accumulate in TMEM).
 *
 * mode 0 ("NT"): D[m, n] = sum_tap sum_k A[m + a_rowoff[tap], k] * B[n, b_koff[tap] + k]
 *      A: [a_rows, K] bf16 (lda elements per row), B: [N, ...] bf16 (ldb), k in [0, K).  taps=1 is a plain
 *      linear layer x @ W^T (replaces F.linear at transformer.py:176-178, reftr_transformer.py:14-23, ...);
 *      taps=9 is a 3x3 convolution over padded NHWC (replaces torchvision Bottleneck conv2d, backbone.py:99-102).
 * mode 1 ("TN"): D[m, n] = sum_r A[r + a_rowoff[z], m] * B[r + b_rowoff[z], n], r in [0, K) ; used for
 *      weight gradients (contraction over pixels / tokens), one z per tap, split-K over `splits` CTAs with
 *      fp32 atomic accumulation into out32 + z * out32_z_stride (replaces autograd's conv/linear wgrad).
 *      Optional bias gradient / per-row scale: see bias_grad, row_scale, out_scale below.
 * Epilogue (mode 0, and mode 1 when atomic=0): v = acc + bias[n] + res[row, n] + res32[row, n];
 *      relu; v = mask_src[row, n] > 0 ? v : 0; rows that are padding per `geom` -> 0; written as bf16 (out)
 *      and/or fp32 (out32).  row = m + out_row_off for out/res/mask_src addressing and geometry.
 *      With `drop`: v = res + res32 + dropout(relu?(acc + bias)) (the residual dropouts and the FFN dropout of
 *      transformer.py:176-179); the site's logical tensor is [rows, N >> drop_gshift] with element (row, n >> drop_gshift), so
 *      drop_gshift = 5 drops whole 32-wide heads (attention-probability dropout of a single-key softmax).  `mask_scale`
 *      (0 = 1) multiplies the values that mask_src keeps (backward of ReLU followed by dropout).
 */
typedef struct {
  int mode;
  const void* A; long long a_rows; int a_cols; long long lda;
  const void* B; long long b_rows; int b_cols; long long ldb;
  int M, N, K;
  int taps;
  int a_rowoff[16];
  int b_koff[16]; /* mode 0: k offset into B per tap; mode 1: row offset into B per z */
  int splits;     /* mode 1: K splits; <= 0 with atomic accumulation = chosen by the library together with the tile width */
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
  const rb_dropout* drop; /* HOST pointer, nullable */
  int drop_gshift;
  float mask_scale;
  /* atomic accumulation only (weight gradients): everything accumulated is multiplied by out_scale (0 = 1) and row m by
   * row_scale[m] (nullable; the FrozenBatchNorm fold of a convolution's weight gradient, backbone.py:70-80); mode 1 only:
   * bias_grad[m] += out_scale * sum_r A[r + a_rowoff[0], m] (nullable) -- the bias gradient of the same layer, contracted on the
   * tensor cores against a block of ones next to the weight gradient (replaces a separate column-sum pass over dY) */
  float* bias_grad;
  const float* row_scale;
  float out_scale;
  /* 0 = all SMs.  > 0: the persistent grid (and the tile-shape / split choice) uses at most this many SMs -- for launches that are
   * OFF the critical path and run next to other kernels (the language backbone's chain beside the conv backbone): what they cost
   * the step is the SM time they hold, not their latency. */
  int sm_limit;
} rb_gemm_args;

int rb_gemm(const rb_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Backbone support kernels (HBM-bound).  Replace torchvision ResNet stem / pooling / stride handling reached from
 * models/modeling/backbone.py:99-102, and FrozenBatchNorm2d (backbone.py:70-80) which is folded into the weights.
 * ------------------------------------------------------------------------------------------------------------- */
/* img fp32 NCHW [B,3,H,W] -> out bf16 [B*H1*W1, 160]: im2col of the 7x7/2 pad-3 stem, column (r*7+s)*3+c, 147.. zero */
int rb_stem_im2col(const float* img, void* out, int B, int H, int W, int H1, int W1, void* stream);
/* the whole stem in one kernel: 7x7/2 pad-3 conv 3->64 (+ folded FrozenBN bias + ReLU), im2col built in shared memory;
 * img fp32 NCHW, wf = rb_pack_conv output bf16 [64, ldk] (column (r*7+s)*3+c), out bf16 NHWC [B*H1*W1, 64] */
int rb_stem_conv(const float* img, const void* wf, int ldk, const float* bias, void* out, int B, int H, int W, int H1, int W1, void* stream);
/* conv1 + bn1 + relu + maxpool (torchvision ResNet stem as reached from backbone.py:99-102) in one pass: img fp32 NCHW [B,3,H,W] ->
 * out 16-bit padded NHWC [B, H2+2, W2+2, 64] with a zero border (the layout rb_maxpool_3x3s2 writes); the 64-channel stride-2 map
 * is never written.  wpk: the folded weights as [7 tap rows][4 K chunks][64][8] 16-bit, K index = 4 * s + c (s: tap column, c:
 * channel; s = 7 and c = 3 are zero); hwc4: scratch of B*H*(W+2)*8 bytes (the batch as 16-bit HWC4 pixels, one zero pixel left and right of every row).  W must be even. */
int rb_stem_pool(const float* img, const void* wpk, const float* bias, void* hwc4, void* out, int B, int H, int W, int H1, int W1, int H2, int W2,
                 void* stream);
/* in bf16 NHWC [B,H1,W1,C] -> out padded NHWC [B,H2+2,W2+2,C]; 3x3 stride 2 pad 1 max-pool */
int rb_maxpool_3x3s2(const void* in, void* out, int B, int H1, int W1, int C, int H2, int W2, void* stream);
/* padded NHWC [B,H+2,W+2,C] <-> 4 parity planes [4,B,Ho+2,Wo+2,C] (see header comment); merge = backward of split,
 * with optional addend `add` (padded layout) and ReLU mask `mask_src` (padded layout, keep where > 0) */
int rb_parity_split(const void* x, void* xs, int B, int H, int W, int C, int Ho, int Wo, void* stream);
int rb_parity_merge(const void* dxs, const void* add, const void* mask_src, void* dx, int B, int H, int W, int C, int Ho, int Wo, void* stream);
/* single-plane variants for the 1x1 / stride-2 downsample convolution (torchvision Bottleneck.downsample), which reads plane 3
 * only: split produces just that plane's [B,Ho+2,Wo+2,C] block; merge treats the three other planes as zero */
int rb_parity_split_plane(const void* x, void* xs_plane, int B, int H, int W, int C, int Ho, int Wo, int plane, void* stream);
int rb_parity_merge_plane(const void* dxs_plane, int plane, const void* add, const void* mask_src, void* dx, int B, int H, int W, int C, int Ho, int Wo,
                          void* stream);
/* conv weight fp32 OIHW (+ FrozenBN buffers or conv bias) -> bf16 [Cout, ldk] ((r,s,ci) columns, BN scale folded),
 * flipped/transposed dgrad copy bf16 [Cin, kh*kw*Cout] (nullable), per-channel scale and bias (fp32, nullable) */
int rb_pack_conv(const float* w, int Cout, int Cin, int kh, int kw, const float* bn_w, const float* bn_b, const float* bn_rm, const float* bn_rv,
                 float eps, const float* conv_bias, void* fwd, int ldk, void* dgr, float* scale_out, float* bias_out, void* stream);
/* linear weight fp32 [N,K] (dense) -> bf16 [N,K] (wb, pitch ldwb, nullable) and bf16 [K,N] (wt, pitch ldwt, nullable) */
int rb_pack_linear(const float* w, int N, int K, void* wb, long long ldwb, void* wt, long long ldwt, void* stream);
/* w fp32 [N,K] -> w2 16-bit [N, 2K] (pitch ld): columns [0,K) the weight rounded to 16 bits, [K,2K) the rounding residual.  An rb_gemm
 * with taps {(0,0),(0,K)} over it multiplies by the fp32 weight to ~2^-22 (forward of the latency-bound transformer / BERT layers,
 * where 16-bit weight rounding dominated the box error: DESIGN.md section 2). */
int rb_pack_linear_hilo(const float* w, int N, int K, void* w2, long long ld, void* stream);
/* folded-layout weight gradient fp32 [Cout,taps,Cin] -> parameter layout fp32 [Cout,Cin,kh,kw], times scale[co] (nullable) */
int rb_unpack_conv_grad(const float* dwf, const float* scale, float* grad, int Cout, int Cin, int taps, void* stream);
int rb_cast_bf16(const float* in, void* out, long long n, void* stream);
/* out[n] += sum_rows x[row, n] (x bf16 or fp32 with pitch ld) -- bias gradients */
int rb_colsum(const void* x, int is_bf16, long long ld, long long rows, int N, float* out, void* stream);
/* y = a + b (b nullable); optional bf16 copy yb */
int rb_add(const float* a, const float* b, float* y, void* yb, long long n, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Normalisation, positions, masks.  Row maps: row r of the compact tensor lives at (r / group) * stride + r % group
 * + offset of the mapped tensor (group = 0: identity); used to write language rows into the [B*S,256] token matrix.
 * ------------------------------------------------------------------------------------------------------------- */
/* nn.LayerNorm(256) (+ optional ReLU, reftr_transformer.py:14-23): fp32 in, fp32 / bf16 / bf16(+pos) out, saves mean, rstd */
int rb_layernorm_fwd(const float* x, const float* gamma, const float* beta, long long rows, int D, float eps, float* y32, void* yb, const float* pos32,
                     void* ypb, int relu, float* mean, float* rstd, int map_group, int map_stride, int map_offset, const rb_dropout* drop, void* stream);
/* y_relu > 0 is the ReLU (and, after a dropout, the keep) mask of the forward output, kept gradients are multiplied by relu_scale;
 * `dxb_drop` applies the dropout of the layer that FED this LayerNorm's residual sum to the bf16 copy dxb only (dx32 stays the
 * undropped residual gradient) */
int rb_layernorm_bwd(const float* dy, const float* dy2, const float* y_relu, float relu_scale, const float* x, const float* gamma, const float* mean,
                     const float* rstd, long long rows, int D, float* dx32, void* dxb, float* dgamma, float* dbeta, int map_group, int map_stride,
                     int map_offset, const rb_dropout* dxb_drop, void* stream);
/* input_proj GroupNorm(32,256) (reftr_transformer.py:121-125) fused with flatten/transpose/concat (reftr.py:57,115-117):
 * x fp32 padded NHWC [B,h+2,w+2,256] -> token rows b*S+L+p of y32 / yb / ypb */
int rb_groupnorm_tokens_fwd(const float* x, const float* gamma, const float* beta, int B, int h, int w, int S, int L, float eps, float* y32, void* yb,
                            const float* pos32, void* ypb, float* mean, float* rstd, void* stream);
int rb_groupnorm_tokens_bwd(const float* dy, const float* dy2, const float* x, const float* gamma, const float* mean, const float* rstd, int B, int h, int w,
                            int S, int L, void* dx, float* dgamma, float* dbeta, void* stream);
/* PositionEmbeddingSine (position_encoding.py:36-56) + level/type/language-position embeddings (reftr.py:57-97) and the
 * key-padding mask [B,S] (u8, 1 = ignore) from the image padding mask (backbone.py:107) and sentence_mask (int64) */
int rb_build_pos_mask(const void* img_mask, int B, int H, int W, int h, int w, const long long* sent_mask, int L, const float* lang_pos,
                      const float* token_type, const float* level_embed, float* pos32, void* kpm, void* stream);
/* gradients of lang_pos_embeddings [L rows], token_type_embeddings [2,256], level_embed [1,256] from d(pos) [B*S,256] */
int rb_embed_grad(const float* dpos, int B, int S, int L, float* d_lang_pos, float* d_token_type, float* d_level, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Attention core, head_dim 32 (bmm / softmax / bmm of F.multi_head_attention_forward; transformer.py:174, :239, :243).
 * Q [B*Tq, ldq], K/V [B*Sk, ld], head h at columns [32h, 32h+32); kpm [B,Sk] u8 (1 = ignore, nullable);
 * LSE, Dbuf fp32 [B,H,Tq].  scale multiplies q before QK^T as torch does.
 * ------------------------------------------------------------------------------------------------------------- */
/* `drop`: dropout on the softmax probabilities (logical tensor [B*H*Tq, Sk]); LSE stays that of the undropped softmax */
int rb_attn_fwd(const void* Q, const void* K, const void* V, const void* kpm, void* O, float* LSE, int B, int H, int dh, int Tq, int Sk, long long ldq,
                long long ldk, long long ldv, long long ldo, float scale, const rb_dropout* drop, void* stream);
int rb_attn_bwd(const void* Q, const void* K, const void* V, const void* kpm, const void* O, const void* dO, const float* LSE, void* dQ, void* dK, void* dV,
                float* Dbuf, int B, int H, int dh, int Tq, int Sk, long long ldq, long long ldk, long long ldv, long long ldo, long long lddo, long long lddq,
                long long lddk, long long lddv, float scale, const rb_dropout* drop, void* stream);
/* QueryEncoder attended pooling (reftr_transformer.py:47-55): k [B,256], q,v [B*L,256] fp32, mask [B,n_ph,L] u8 */
int rb_qenc_pool_fwd(const float* k, const float* q, const float* v, const void* mask, int B, int L, int n_ph, float* att, float* c, void* stream);
int rb_qenc_pool_bwd(const float* dc, const float* k, const float* q, const float* v, const float* att, int B, int L, int n_ph, float* dk, float* dq,
                     float* dv, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Row-mapped glue (QueryEncoder.forward reftr_transformer.py:41-66; memory slicing :261-267).  A map is a HOST pointer
 * to 4 ints {group, stride, inner, offset}: index(r) = (r / group) * stride + (r % group) * inner + offset
 * (group == 0 or a NULL map: r + offset).
 * ------------------------------------------------------------------------------------------------------------- */
/* y[my(r), :D] = a[ma(r), :D] + b[mb(r), :D] (b nullable); fp32 in, fp32 (y32) and/or bf16 (yb) out */
int rb_rows_add(const float* a, long long lda, const int* map_a, const float* b, long long ldb, const int* map_b, float* y32, long long ldy,
                void* yb, long long ldyb, const int* map_y, long long rows, int D, void* stream);
/* dst[md(r), :D] += src[ms(r), :D] (fp32 atomics) */
int rb_rows_scatter_add(const float* src, long long lds, const int* map_src, float* dst, long long ldd, const int* map_dst, long long rows,
                        int D, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Segmentation head (reftr_segmentation.py:152-175, :196-207 MHAttentionMap, :243-280 MaskHeadSmallConv).
 * Grids are padded NHWC (see top); `ld` is the pixel pitch in elements, `col0` the first channel touched.
 * ------------------------------------------------------------------------------------------------------------- */
/* visual token rows b*S+L+p of tok fp32 [B*S, C] -> interior pixels of grid bf16, channels [col0, col0+C)  (:166, :243 cat) */
int rb_tokens_to_grid(const float* tok, int B, int S, int L, int h, int w, int C, void* grid, long long ld, int col0, void* stream);
/* reverse: dtok[b*S+L+p, :C] = grid[pixel, col0:col0+C] (fp32; language rows are not touched) */
int rb_grid_to_tokens(const void* grid, long long ld, int col0, int B, int S, int L, int h, int w, int C, float* dtok, void* stream);
/* MHAttentionMap: q fp32 [B,256] (q_linear output), k fp32 [B*S,256] (k_linear output, visual rows used), kpm u8 [B,S];
 * logits[b,n,p] = scale * <q[b,32n:32n+32], k[b*S+L+p, 32n:32n+32]>, masked where kpm, softmax over (n,p) jointly;
 * att fp32 [B,8,hw]; also written as bf16 into grid channels [col0, col0+8) */
int rb_attn_map_fwd(const float* q, const float* k, const void* kpm, int B, int S, int L, int hw, int w, float scale, float* att, void* grid,
                    long long ld, int col0, void* stream);
/* d(att) = datt_ext (nullable, fp32 [B,8,hw]) + grid gradient channels [col0, col0+8) (bf16); dq fp32 [B,256]; dk fp32 [B*S,256]
 * (language rows written as zero) */
int rb_attn_map_bwd(const float* datt_ext, const void* dgrid, long long ld, int col0, const float* att, const float* q, const float* k, int B, int S,
                    int L, int hw, int w, float scale, float* dq, float* dk, void* stream);
/* nn.GroupNorm(G, C) (+ReLU) over the interior of a padded NHWC fp32 grid [B,H+2,W+2,C] -> bf16 grid (border written as zero) */
int rb_groupnorm_nhwc_fwd(const float* x, const float* gamma, const float* beta, int B, int H, int W, int C, int G, float eps, int relu, void* y,
                          float* mean, float* rstd, void* stream);
/* dy, y bf16 grids (y > 0 is the ReLU mask when relu), x fp32 grid -> dx bf16 grid (border zero); dgamma/dbeta accumulated */
int rb_groupnorm_nhwc_bwd(const void* dy, const void* y, const float* x, const float* gamma, const float* mean, const float* rstd, int B, int H, int W,
                          int C, int G, int relu, void* dx, float* dgamma, float* dbeta, void* stream);
/* y = cur + nearest_upsample(lo) (F.interpolate(mode="nearest") to cur's size, :256-274); bf16 grids, C % 8 == 0 */
int rb_upsample_add(const void* lo, const void* cur, void* y, int B, int h, int w, int H, int W, int C, void* stream);
/* dlo[b,sy,sx,:] = sum of dy over the pixels that read (sy,sx) */
int rb_upsample_bwd(const void* dy, void* dlo, int B, int h, int w, int H, int W, int C, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * BERT-base language backbone (HuggingFace BertModel as called at reftr_transformer.py:200, :217; third party in the
 * reference).  Its linear layers run on rb_gemm; these are the remaining pieces.
 * ------------------------------------------------------------------------------------------------------------- */
/* BertEmbeddings before LayerNorm: out[r,:] = word[ids[r]] + pos[r % L] + type0   (token_type_ids = 0, position_ids = arange(L)) */
int rb_bert_embed_fwd(const long long* ids, long long rows, int L, int D, const float* word, const float* pos, const float* type0, float* out, void* stream);
/* scatter-add of d [rows, D] into the three embedding tables' gradients (any of them nullable) */
int rb_bert_embed_bwd(const float* d, const long long* ids, long long rows, int L, int D, float* dword, float* dpos, float* dtype0, void* stream);
/* nn.LayerNorm over D = 768 / 1024 wide rows: fp32 in, fp32 and/or bf16 out, saves mean / rstd */
/* `drop` (BertEmbeddings.dropout): applied to both outputs */
int rb_ln_wide_fwd(const float* x, const float* gamma, const float* beta, long long rows, int D, float eps, float* y32, void* yb, float* mean, float* rstd,
                   const rb_dropout* drop, void* stream);
/* `dy_drop`: the forward output had this dropout (dy + dy2 is masked and scaled first); `dxb_drop`: as in rb_layernorm_bwd */
int rb_ln_wide_bwd(const float* dy, const float* dy2, const float* x, const float* gamma, const float* mean, const float* rstd, long long rows, int D,
                   float* dx32, void* dxb, float* dgamma, float* dbeta, const rb_dropout* dy_drop, const rb_dropout* dxb_drop, void* stream);
/* exact (erf) GELU on bf16, n % 8 == 0; backward takes the PRE-activation x */
int rb_gelu_fwd(const void* x, void* y, long long n, void* stream);
int rb_gelu_bwd(const void* dy, const void* x, void* dx, long long n, void* stream);
/* BertPooler activation */
int rb_tanh_fwd(const float* x, float* y, long long n, void* stream);
int rb_tanh_bwd(const float* dy, const float* y, float* dx, void* dxb, long long n, void* stream);
/* BertSelfAttention core, head_dim 64, S <= 128 tokens: Q,K,V bf16 [B*S, ld] (head h at columns [64h, 64h+64)), mask u8 [B,S]
 * (1 = ignore key); P fp32 [B,H,S,S] is saved for the backward */
int rb_attn_small_fwd(const void* Q, const void* K, const void* V, const void* mask, void* O, float* P, int B, int H, int dh, int S, long long ldq,
                      long long ldk, long long ldv, long long ldo, float scale, const rb_dropout* drop, void* stream);
int rb_attn_small_bwd(const void* Q, const void* K, const void* V, const void* dO, const float* P, void* dQ, void* dK, void* dV, int B, int H, int dh, int S,
                      long long ldq, long long ldk, long long ldv, long long lddo, long long lddq, long long lddk, long long lddv, float scale,
                      const rb_dropout* drop, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Criterion (models/criterion.py:113-153, :189-201; util/box_ops.py): L1 + GIoU of paired cxcywh boxes for all decoder layers
 * at once.  boxes fp32 [n_layers, N, 4], tgt fp32 [N, 4], valid u8 [N] (nullable: all valid).  losses fp32 [n_layers, 2] =
 * (sum |p - t|, sum (1 - giou)) * inv_norm (inv_norm_dev, a device scalar, overrides inv_norm when non-NULL);
 * dl1 / dgiou fp32 [n_layers, N, 4] = gradients of the two terms w.r.t. boxes (already times inv_norm).
 * ------------------------------------------------------------------------------------------------------------- */
int rb_box_loss(const float* boxes, const float* tgt, const void* valid, int n_layers, int N, float inv_norm, const float* inv_norm_dev, float* losses,
                float* dl1, float* dgiou, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Optimizer step on device-flat buffers (engine_vg.py:62-67, main_vg.py:234-268): global-norm clip
 * (torch.nn.utils.clip_grad_norm_) + torch.optim.AdamW over ALL parameters in two launches.  p / g / m / v are flat fp32 buffers
 * of n elements (n % 4 == 0, 16-byte aligned) holding every parameter, its gradient and its two moments at the same offsets;
 * `segs` (HOST pointer) maps element ranges to parameter groups (per-group lr / weight decay, read on the host every step, so LR
 * schedulers work unchanged).
 * ------------------------------------------------------------------------------------------------------------- */
#define RB_ADAMW_MAX_SEGMENTS 32
#define RB_ADAMW_MAX_GROUPS 8
typedef struct {
  int nseg;
  long long end[RB_ADAMW_MAX_SEGMENTS]; /* segment s = elements [end[s-1], end[s]), multiples of 4 */
  int group[RB_ADAMW_MAX_SEGMENTS];
  float lr[RB_ADAMW_MAX_GROUPS];
  float weight_decay[RB_ADAMW_MAX_GROUPS];
} rb_adamw_segments;
/* *out += sum of squares of x[0..n)  (out: device scalar, zeroed by the caller) */
int rb_sumsq(const float* x, long long n, float* out, void* stream);
/* Hand-over of the flat gradient buffer at the end of the backward pass (replaces the reference's implicit per-tensor .grad
 * tensors, engine_vg.py:61): dst[i] = src[i] * scale (undoes the static loss scale of the 16-bit backward) and *flag |= 1 if any
 * src[i] is inf / NaN -- the finite check the reference only does on the loss (engine_vg.py:55-58), at no extra memory traffic.
 * n % 4 == 0, 16-byte aligned buffers; flag is a device int the caller zeroes. */
int rb_scale_copy_check(const float* src, float* dst, long long n, float scale, int* flag, void* stream);
/* If *flag != 0: x[0..n) = 0 and *counter += 1 (counter may be NULL).  If *flag == 0 the launch does nothing. */
int rb_zero_if(float* x, long long n, const int* flag, float* counter, void* stream);
/* One AdamW step (amsgrad off).  `step` counts from 1 (bias corrections are computed on the host in double).  If sumsq != NULL and
 * max_norm > 0 the gradient is multiplied by min(1, max_norm / (sqrt(*sumsq) + 1e-6)) first (clip_grad_norm_, no host sync). */
int rb_adamw_flat(float* p, const float* g, float* m, float* v, long long n, const rb_adamw_segments* segs, float beta1, float beta2, float eps, int step,
                  const float* sumsq, float max_norm, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Input pipeline on the GPU: torchvision to_tensor + Normalize (datasets/transforms.py:233-250) and the
 * pad-to-batch NestedTensor build of util/collate_fn.py:24-41 in one launch.  packed: the raw uint8 HWC images back to back;
 * table int64 [B,3] = {byte offset, h, w} (device); out fp32 [B,3,H,W] = ((x / 255) - mean) / std inside the image, 0 in the
 * padding; mask bool/u8 [B,H,W], 1 = padding.
 * ------------------------------------------------------------------------------------------------------------- */
int rb_collate_u8(const void* packed, const long long* table, int B, int H, int W, float mean0, float mean1, float mean2, float std0, float std1,
                  float std2, float* out, void* mask, void* stream);

/* Resize of one raw uint8 HWC image on the device, bit-exact with torchvision's F.resize on a PIL image (= Pillow's bilinear
 * ImagingResample; datasets/transforms.py:81-111, RandomResize :196-204): horizontal then vertical pass over uint8 with 22-bit
 * fixed-point coefficients.  bounds_* int32 [out, 2] = (first input index, count), kk_* int32 [out, ksize_*] -- computed on the host
 * exactly as Pillow does (reftr_b200/data.py:pil_bilinear_coeffs); a pass whose size does not change is skipped (its tables may
 * be NULL); tmp: [h, ow, 3] bytes, needed when both passes run. */
int rb_resize_u8(const void* src, int h, int w, void* dst, int oh, int ow, const int* bounds_h, const int* kk_h, int ksize_h, const int* bounds_v,
                 const int* kk_v, int ksize_v, void* tmp, void* stream);

#ifdef __cplusplus
}
#endif
#endif
