// tcgen05 GEMM / implicit-GEMM convolution for sm_100a -- persistent, warp-specialised, TMA in and TMA out.
//
// One CTA per SM walks a static list of 128 x BN output tiles (BN = 64..256, chosen per problem at run time).
// Warp roles (352 threads):
//   warp 0      : TMA producer for the main loop (one elected lane) -- STAGES-deep ring of {A tile, B tile} in shared memory
//   warp 1      : TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma); the accumulator is DOUBLE-BUFFERED in
//                 TMEM (2 x BN columns) so the main loop of tile i+1 overlaps the epilogue of tile i
//   warps 2..9  : epilogue, two teams of 4 warps that take alternate 64-column chunks -- tcgen05.ld the accumulator (one
//                 thread per output row), fused bias / residual / ReLU / ReLU-mask / border-zero; results are staged in shared
//                 memory (128B-swizzled, bank-conflict free) and written with TMA bulk stores (bf16 and/or fp32), or added
//                 with vector fp32 reductions (red.global.add.v4.f32) in split-K mode
//   warp 10     : TMA producer for the epilogue INPUTS (bf16 residual, fp32 residual, ReLU-mask source), 64-column chunks
//                 prefetched through a small ring, so that the epilogue never waits on a dependent global load
// All global traffic is therefore bulk, asynchronous and 128-byte coalesced; what bounds the HBM-bound layers is the
// number of bytes in flight (ring depths), not the number of resident warps.
//
// mode NT: A [rows, K] and B [N, K] are K-major; tiles are [128|BN rows] x 64 k (128 B per row), 128B-swizzled by TMA.
//          Convolution taps are row shifts of A (padded-NHWC layout, see include/reftr_b200.h).
// mode TN: the contraction runs over rows (pixels/tokens): A [rows, M] and B [rows, N] are "MN-major"; tiles are
//          64 rows x 64 channels boxes, consumed through MN-major UMMA descriptors (no transposes anywhere).
#include "common.cuh"
#include "host.h"

#include <stdlib.h>
#include <string.h>

namespace rb {

constexpr int BM = 128;
constexpr int BK = 64;
#ifdef RB_EPI16
// EXPERIMENTAL build variant (tools/build_variant.sh, not the default): 16 epilogue warps -- two teams of 8, the two 4-warp halves of
// a team take the two 32-column halves of the team's 64-column chunk -- because the epilogue is bound by the instruction stream of
// its warps.  Needs <= 96 registers per thread (640 threads).
constexpr int GEMM_THREADS = 640, EPI_WARP0 = 4, EIN_WARP = 2, TEAM_WARPS = 8;
#define RB_EPI_ROLLED
#else
constexpr int GEMM_THREADS = 352, EPI_WARP0 = 2, EIN_WARP = 10, TEAM_WARPS = 4;
#endif
constexpr int MAX_STAGES = 8;
constexpr int MAX_EIN = 8;
constexpr int A_BYTES = BM * BK * 2;
constexpr int SMEM_LIMIT = 232448;  // 227 KB opt-in maximum per CTA

struct GemmKParams {
  int mode;
  int M, N;
  int kblocks;  // NT: k-blocks per tap; TN: total row-blocks
  int taps;
  int a_rowoff[16];
  int b_koff[16];
  int splits;
  int bn, stages;
  int tiles_m, tiles_n, tiles_total;
  long long out_row_off;
  const float* bias;
  float* out32;  // atomic mode only (otherwise TMA store)
  long long ldo32;
  long long out32_z_stride;
  int relu, atomic;
  int has_out, has_out32, has_res, has_res32, has_mask;
  int ein_slots, out_slots;
  uint32_t stage_bytes;                                  // A_BYTES + bn * 128
  uint32_t off_ein, ein_slot_bytes, ein_off_mask, ein_off_res32;
  uint32_t off_out, out_slot_bytes, out_off_f32;         // out_slots per team, 2 teams
  uint32_t off_bars;
  uint32_t off_bias;   // bias staged in shared memory (bias_n floats, zero padded to whole tiles); bias_n == 0: read it from global
  int bias_n;
  rb_geom geom;
  DropK drop;          // dropout on relu?(acc + bias), before the residuals
  int drop_gshift;     // the site's element of output column n is n >> drop_gshift
  uint32_t drop_wpr;   // 32-bit random words per row of the site
  float mask_scale;    // multiplies what mask_src keeps
  float* bias_grad;    // TN + atomic only: bias_grad[m] += out_scale * sum_r A[r, m] (an extra N=16 MMA against a block of ones)
  const float* row_scale;  // atomic only: row m of the result is multiplied by row_scale[m] (FrozenBN fold of a conv weight gradient)
  float out_scale;     // atomic only: multiplies everything that is accumulated
  uint32_t off_ones;   // 8 KB of 1.0 (the B operand of the bias-gradient MMA)
  int cl;              // 1, or 2 = CTA pairs (cluster 2x1x1, tcgen05.mma.cta_group::2): a pair multiplies a 256 x bn tile; each CTA loads its
                       // 128 rows of A and HALF of the B tile, the leader issues the MMAs, each CTA runs the epilogue of its own 128 rows.
                       // The shared-memory fill per CTA and k-block shrinks from A + B to A + B/2 (these GEMMs are bound by
                       // the bytes they can keep in flight, not by the tensor pipe)
  int tiles_mp;        // cl == 2: pairs of row tiles
  int bm;              // rows per tile: 128, or 256 ("tall" tiles, cl == 1 only): two 128-row MMAs per k-step share the B tile and
                       // accumulate side by side in TMEM (no accumulator double buffering), so that the operand bytes pulled from L2
                       // per FLOP drop by a third against 128 x 256 -- the chip-wide L2 -> SM throughput (~6300 B/clk) is what bounds the
                       // long-K convolutions and weight gradients
  int debug;           // RB_GEMM_DEBUG (timing experiments only, results are wrong): 1 = no output stores, 4 = no epilogue math, 8 = no TMEM loads, 16 = no staging writes, 32 = no fence / barrier
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

struct TileCoord {
  int m0, n0, z_tap, it_begin, n_it;
};

template <int CL, int NH>
__device__ __forceinline__ TileCoord tile_coord(const GemmKParams& p, int t, int rank) {
  TileCoord c;
  const int tn = t % p.tiles_n;
  int rest = t / p.tiles_n;
  int tm, z;
  if constexpr (CL == 2) {  // t indexes PAIRS of row tiles; a row tile beyond tiles_m (odd tail) is all padding: loads read zeros, stores are clipped
    tm = 2 * (rest % p.tiles_mp) + rank;
    z = rest / p.tiles_mp;
  } else {
    tm = rest % p.tiles_m;
    z = rest / p.tiles_m;
  }
  c.m0 = tm * (NH * BM);
  c.n0 = tn * p.bn;
  if (p.mode == 0) {
    c.z_tap = 0;
    c.it_begin = 0;
    c.n_it = p.taps * p.kblocks;
  } else {
    c.z_tap = z / p.splits;
    const int split = z - c.z_tap * p.splits;
    const int per = (p.kblocks + p.splits - 1) / p.splits;
    c.it_begin = split * per;
    const int it_end = min(p.kblocks, c.it_begin + per);
    c.n_it = it_end - c.it_begin;  // may be <= 0: every role skips such a tile
  }
  return c;
}

// EPI selects how much of the epilogue is compiled in (the full body is ~40 KB of SASS, which thrashes the instruction cache of
// the 8 epilogue warps): 0 = every feature (linear layers: fp32 output / residual, dropout, ...); 1 = the convolution path
// (16-bit output; bias, 16-bit residual, ReLU, ReLU-mask, border zeroing only); 2 = split-K atomic accumulation only.
// CL = 1: one CTA per tile; CL = 2: CTA pairs (cluster 2x1x1, tcgen05 cta_group::2).  A compile-time parameter because a kernel that
// contains cta_group::2 instructions can only be launched with an even cluster size.
// NH = 128-row halves per tile: 1, or 2 = "tall" 256-row tiles (instantiated for the convolution path <0, 1, 1, 2> only).
template <int MODE, int EPI, int CL, int NH>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmOut32, const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmRes32,
                 const __grid_constant__ CUtensorMap tmMask, const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bars);
  uint64_t* full = bars;                       // [MAX_STAGES]
  uint64_t* empty = bars + MAX_STAGES;         // [MAX_STAGES]
  uint64_t* acc_full = bars + 2 * MAX_STAGES;  // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint64_t* ein_full = acc_empty + 2;          // [MAX_EIN]
  uint64_t* ein_empty = ein_full + MAX_EIN;    // [MAX_EIN]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ein_empty + MAX_EIN);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bn = p.bn;
  constexpr int cl = CL;
  int rank = 0;
  if constexpr (cl == 2) rank = static_cast<int>(cluster_ctarank());
  constexpr int nh = NH;  // 128-row halves per tile
  // (read blockIdx / gridDim at every tile loop instead of keeping them in variables: the compiler then proves the loop counters of the
  // single-thread roles warp-uniform and keeps the MMA descriptors in uniform registers; held in a vector register they cost an
  // ELECT + R2UR.BROADCAST waterfall per tcgen05.mma -- 10 % of the whole kernel, profiles/r02_gemm_uniform_regression.log)
#define RB_T_FIRST (cl == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x))
#define RB_T_STEP (cl == 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x))
  const bool f_atomic = EPI == 2 ? true : (EPI == 1 ? false : p.atomic != 0);
  const bool f_res = EPI != 2 && p.has_res != 0, f_mask = EPI != 2 && p.has_mask != 0, f_res32 = EPI == 0 && p.has_res32 != 0;
  const bool f_out = EPI == 1 ? true : (EPI == 2 ? false : p.has_out != 0), f_out32 = EPI == 0 && p.has_out32 != 0;
  const bool has_ein = f_res | f_res32 | f_mask;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);   // pairs: only the leader's is used; it collects the bytes of BOTH CTAs' loads
      mbar_init(&empty[s], 1);  // pairs: the leader's commit arrives on both CTAs' barriers
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], cl * 2 * TEAM_WARPS);  // pairs: the leader's barrier counts the epilogue warps of both CTAs
    }
    for (int s = 0; s < MAX_EIN; ++s) {
      mbar_init(&ein_full[s], 1);
      mbar_init(&ein_empty[s], TEAM_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (cl == 2) tmem_alloc_cg2<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  // the bias is read by every epilogue thread for every chunk: a global (L1-thrashed) load there costs an L2 round trip per chunk
  float* sbias = reinterpret_cast<float*>(smem + p.off_bias);
  for (int i = threadIdx.x; i < p.bias_n; i += GEMM_THREADS) sbias[i] = i < p.N ? __ldg(p.bias + i) : 0.f;
  if (MODE == 1 && p.bias_grad) {  // every element of the block is 1.0, so its (swizzled, MN-major) layout does not matter
    const rb_t one = f2t(1.f);
    const uint32_t one2 = static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&one)) * 0x10001u;
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + p.off_ones);
    for (int i = threadIdx.x; i < 8192 / 4; i += GEMM_THREADS) ones[i] = one2;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (cl == 2) cluster_sync_all();  // the peer's barriers exist before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ main-loop producer
    if (lane == 0) {
      const uint32_t tx_bytes = p.stage_bytes * static_cast<uint32_t>(cl);  // pairs: the leader's barrier expects both CTAs' stages
      uint32_t lead_full0 = 0u;  // pairs: the LEADER's full barriers (shared::cluster address)
      if constexpr (cl == 2) lead_full0 = mapa_u32(smem_u32(&full[0]), 0);
      int s = 0;         // ring position
      uint32_t ph = 0;   // ring phase
      for (int t = RB_T_FIRST; t < p.tiles_total; t += RB_T_STEP) {
        const TileCoord c = tile_coord<CL, NH>(p, t, rank);
        // incremental (tap, k-block) counters: this single thread is instruction-bound, so no divisions in the loop
        int tap = 0, kc = 0;
        if (MODE == 0 && c.it_begin) { tap = c.it_begin / p.kblocks; kc = c.it_begin - tap * p.kblocks; }
        int row_a = c.m0 + (MODE == 0 ? p.a_rowoff[tap] : 0), col_b = (MODE == 0 ? p.b_koff[tap] : 0);
        int r0 = c.it_begin * BK;
        const int ra = (MODE == 1) ? p.a_rowoff[c.z_tap] : 0, rb = (MODE == 1) ? p.b_koff[c.z_tap] : 0;
        for (int k = 0; k < c.n_it; ++k) {
          mbar_wait(&empty[s], ph ^ 1);
          if (rank == 0) mbar_expect_tx(&full[s], tx_bytes);
          uint8_t* a_dst = smem + s * p.stage_bytes;
          uint8_t* b_dst = a_dst + nh * A_BYTES;
          if (MODE == 0) {
            if constexpr (cl == 2) {
              tma_load_2d_cg2(a_dst, &tmA, (lead_full0 + 8u * s), kc * BK, row_a);
              tma_load_2d_cg2(b_dst, &tmB, (lead_full0 + 8u * s), col_b + kc * BK, c.n0 + rank * (bn >> 1));
            } else {
              tma_load_2d(a_dst, &tmA, &full[s], kc * BK, row_a);
              if (nh == 2) tma_load_2d(a_dst + A_BYTES, &tmA, &full[s], kc * BK, row_a + BM);
              tma_load_2d(b_dst, &tmB, &full[s], col_b + kc * BK, c.n0);
            }
            if (++kc == p.kblocks) {
              kc = 0;
              ++tap;
              row_a = c.m0 + p.a_rowoff[tap & 15];
              col_b = p.b_koff[tap & 15];
            }
          } else {
            if constexpr (cl == 2) {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d_cg2(a_dst + j * 8192, &tmA, (lead_full0 + 8u * s), c.m0 + 64 * j, r0 + ra);
              const int nb = bn >> 7, nc0 = c.n0 + rank * (bn >> 1);  // this CTA's half of the B tile: nb boxes of 64 columns
              for (int j = 0; j < nb; ++j) tma_load_2d_cg2(b_dst + j * 8192, &tmB, (lead_full0 + 8u * s), nc0 + 64 * j, r0 + rb);
            } else {
              for (int j = 0; j < nh * (BM / 64); ++j) tma_load_2d(a_dst + j * 8192, &tmA, &full[s], c.m0 + 64 * j, r0 + ra);
              for (int j = 0; j < bn / 64; ++j) tma_load_2d(b_dst + j * 8192, &tmB, &full[s], c.n0 + 64 * j, r0 + rb);
            }
            r0 += BK;
          }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0 && rank == 0) {  // pairs: the leader issues for both CTAs
      // The issuing thread is instruction-bound (one thread, dependent issue): descriptors are formed by ADDING 16-byte-unit
      // offsets to a base descriptor (the start-address field is the low 14 bits; shared memory is < 256 KB so no carry-out).
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t a_desc0 = (MODE == 0) ? umma_smem_desc(smem_base, 16, 1024, SWZ_128B) : umma_smem_desc(smem_base, 8192, 1024, SWZ_128B);
      const uint32_t stage_units = p.stage_bytes >> 4, b_units = static_cast<uint32_t>(nh * A_BYTES) >> 4;
      constexpr uint32_t kstep = (MODE == 0 ? 32 : 2048) >> 4;
      const uint64_t ones_desc = a_desc0 + static_cast<uint64_t>(p.off_ones >> 4);
      int s = 0, tcount = 0;
      uint32_t ph = 0;
      // the same loop twice (cta_group::1 / ::2) so that the per-MMA path carries no mode branch
#define RB_MMA_LOOP(MMA, COMMIT, ROWS)                                                                                             \
      {                                                                                                                            \
        const uint32_t idesc = umma_idesc_t(ROWS, bn, MODE, MODE), idesc_cs = umma_idesc_t(ROWS, 16, 1, 1);                        \
        for (int t = RB_T_FIRST; t < p.tiles_total; t += RB_T_STEP) {                                                                    \
          const TileCoord c = tile_coord<CL, NH>(p, t, rank);                                                                              \
          if (c.n_it <= 0) continue;                                                                                               \
          /* tall tiles: both halves' accumulators fill TMEM side by side, ONE accumulator stage */                                \
          const int as = nh == 2 ? 0 : (tcount & 1);                                                                               \
          mbar_wait(&acc_empty[as], ((nh == 2 ? tcount : (tcount >> 1)) & 1) ^ 1);                                                 \
          tc_fence_after();                                                                                                        \
          const uint32_t d_tmem = tmem_base + as * bn;                                                                             \
          /* bias gradient (TN): the tiles of the first column block also contract A against ones into 16 spare TMEM columns */    \
          const bool colsum = MODE == 1 && p.bias_grad != nullptr && c.n0 == 0 && c.z_tap == 0;                                    \
          const uint32_t cs_tmem = tmem_base + 2 * bn + as * 16;                                                                   \
          for (int k = 0; k < c.n_it; ++k) {                                                                                       \
            mbar_wait(&full[s], ph);                                                                                               \
            tc_fence_after();                                                                                                      \
            const uint64_t ad = a_desc0 + static_cast<uint64_t>(s * stage_units);                                                  \
            const uint64_t bd = ad + b_units;                                                                                      \
            MMA(d_tmem, ad, bd, idesc, k != 0);                                                                                    \
            _Pragma("unroll") for (int kk = 1; kk < BK / 16; ++kk) MMA(d_tmem, ad + kk * kstep, bd + kk * kstep, idesc, 1);        \
            if (nh == 2) {                                                                                                         \
              const uint64_t ad2 = ad + (A_BYTES >> 4);                                                                            \
              _Pragma("unroll") for (int kk = 0; kk < BK / 16; ++kk)                                                               \
                  MMA(d_tmem + bn, ad2 + kk * kstep, bd + kk * kstep, idesc, (k | kk) != 0);                                       \
            }                                                                                                                      \
            if (colsum) {                                                                                                          \
              _Pragma("unroll") for (int kk = 0; kk < BK / 16; ++kk)                                                               \
                  MMA(cs_tmem, ad + kk * kstep, ones_desc + kk * kstep, idesc_cs, (k | kk) != 0);                                  \
            }                                                                                                                      \
            COMMIT(&empty[s]);                                                                                                     \
            if (++s == p.stages) { s = 0; ph ^= 1; }                                                                               \
          }                                                                                                                        \
          COMMIT(&acc_full[as]);                                                                                                   \
          ++tcount;                                                                                                                \
        }                                                                                                                          \
      }
      if constexpr (cl == 2) RB_MMA_LOOP(umma_f16_ss_cg2, umma_commit_cg2, 2 * BM)
      else RB_MMA_LOOP(umma_f16_ss, umma_commit, BM)
#undef RB_MMA_LOOP
    }
    __syncwarp();
  } else if (warp == EIN_WARP) {
    // ------------------------------------------------------------------------------------------ epilogue-input producer
    if (lane == 0 && has_ein && !f_atomic) {
      const uint32_t tx = (f_res ? 16384u : 0u) + (f_mask ? 16384u : 0u) + (f_res32 ? 32768u : 0u);
      int g = 0;
      for (int t = RB_T_FIRST; t < p.tiles_total; t += RB_T_STEP) {
        const TileCoord c = tile_coord<CL, NH>(p, t, rank);
        if (c.n_it <= 0) continue;
        for (int sub = 0; sub < nh; ++sub) {
        const int m0s = c.m0 + sub * BM;
        for (int ch = 0; ch < bn / 64; ++ch, ++g) {
          const int col0 = c.n0 + ch * 64;
          if (col0 >= p.N) { g += bn / 64 - ch; break; }
          const int s = g % p.ein_slots;
          mbar_wait(&ein_empty[s], ((g / p.ein_slots) & 1) ^ 1);
          mbar_expect_tx(&ein_full[s], tx);
          uint8_t* dst = smem + p.off_ein + s * p.ein_slot_bytes;
          if (f_res) tma_load_2d(dst, &tmRes, &ein_full[s], col0, m0s);
          if (f_mask) tma_load_2d(dst + p.ein_off_mask, &tmMask, &ein_full[s], col0, m0s);
          if (f_res32) {
            tma_load_2d(dst + p.ein_off_res32, &tmRes32, &ein_full[s], col0, m0s);
            tma_load_2d(dst + p.ein_off_res32 + 16384, &tmRes32, &ein_full[s], col0 + 32, m0s);
          }
        }
        }
      }
    }
    __syncwarp();
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------------------------------------ epilogue (2 teams)
    const int team = (warp - EPI_WARP0) / TEAM_WARPS;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
#ifdef RB_EPI16
    const int half_lo = ((warp - EPI_WARP0) >> 2) & 1, half_hi = half_lo;  // this warp's 32-column half of the team's chunk
#elif defined(RB_EPI_ROLLED)
    const int half_lo = 0, half_hi = 1;
#endif
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const bool store_thread = (threadIdx.x == (EPI_WARP0 + team * TEAM_WARPS) * 32);
    const int sw128 = r & 7;
    const bool use_ein = has_ein && !f_atomic;
    const bool use_drop = EPI == 0 && p.drop.seed != nullptr;
    const uint32_t dkey = use_drop ? drop_key(p.drop) : 0u;
    int tcount = 0, g = 0, o = 0;
    for (int t = RB_T_FIRST; t < p.tiles_total; t += RB_T_STEP) {
      const TileCoord c = tile_coord<CL, NH>(p, t, rank);
      if (c.n_it <= 0) continue;
      const int as = nh == 2 ? 0 : (tcount & 1);
      mbar_wait(&acc_full[as], (nh == 2 ? tcount : (tcount >> 1)) & 1);
      tc_fence_after();
      for (int sub = 0; sub < nh; ++sub) {  // tall tiles: the second 128-row half follows the first, its accumulator bn columns further
      const int m0s = c.m0 + sub * BM;
      const uint32_t acc_col = static_cast<uint32_t>((nh == 2 ? sub : as) * bn);
      const long long gm = static_cast<long long>(m0s) + r;
      const bool row_ok = gm < p.M;
      const long long orow = gm + p.out_row_off;
      const bool interior = row_ok && row_is_interior(p.geom, orow);
#pragma unroll 1
      for (int ch = 0; ch < bn / 64; ++ch, ++g) {
        const int col0 = c.n0 + ch * 64;
        if (col0 >= p.N) { g += bn / 64 - ch; break; }
        if ((g & 1) != team) continue;
        if (MODE == 1 && EPI == 2 && ch == 0 && p.bias_grad != nullptr && c.n0 == 0 && c.z_tap == 0) {
          uint32_t cs;
          tmem_ld_32x1(tmem_base + lane_addr + 2 * bn + as * 16, cs);
          tmem_ld_wait();
          if (row_ok) atomicAdd(p.bias_grad + gm, __uint_as_float(cs) * p.out_scale);
        }
        const uint8_t* ein = nullptr;
        int es = 0;
        if (use_ein) {
          es = g % p.ein_slots;
          mbar_wait(&ein_full[es], (g / p.ein_slots) & 1);
          ein = smem + p.off_ein + es * p.ein_slot_bytes;
        }
        const bool two_slots = p.out_slots == 2;
        uint8_t* oslot = smem + p.off_out + (team * p.out_slots + (two_slots ? (o & 1) : 0)) * p.out_slot_bytes;
        ++o;
        if (!f_atomic && !two_slots) {
          if (store_thread) tma_store_wait_read<0>();  // the previous store of this team has finished reading the slot
          named_bar_sync(1 + team, TEAM_WARPS * 32);
        }
#ifdef RB_EPI_ROLLED
        // one 32-column half at a time, NOT unrolled: half the epilogue code (instruction-cache footprint) and fewer registers
#pragma unroll 1
        for (int half = half_lo; half <= half_hi; ++half) {
          const int hc0 = col0 + half * 32;
          if (hc0 >= p.N || (p.debug & 4)) continue;  // warp-uniform
          uint32_t v[32];
          if (p.debug & 8) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = static_cast<uint32_t>(r + j);
          } else {
            tmem_ld_32x32(tmem_base + lane_addr + acc_col + ch * 64 + half * 32, v);
            tmem_ld_wait();
          }
#else
        uint32_t vv[2][32];
        if (p.debug & (4 | 8)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) vv[0][j] = vv[1][j] = static_cast<uint32_t>(r + j);  // (not a constant: keeps the math alive)
        } else {
        tmem_ld_32x32(tmem_base + lane_addr + acc_col + ch * 64, vv[0]);
        tmem_ld_32x32(tmem_base + lane_addr + acc_col + ch * 64 + 32, vv[1]);
        tmem_ld_wait();
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int hc0 = col0 + half * 32;
          const uint32_t (&v)[32] = vv[half];
          if (hc0 >= p.N || (p.debug & 4)) continue;  // warp-uniform
#endif
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (f_atomic) {
            if (row_ok) {
              if (p.row_scale != nullptr || p.out_scale != 1.f) {  // warp-uniform
                const float rs = p.out_scale * (p.row_scale != nullptr ? __ldg(p.row_scale + gm) : 1.f);
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] *= rs;
              }
              float* dst = p.out32 + static_cast<long long>(c.z_tap) * p.out32_z_stride + orow * p.ldo32 + hc0;
              if (hc0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) red_add_v4(dst + 4 * j, f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (hc0 + j < p.N) atomicAdd(dst + j, f[j]);
              }
            }
            continue;
          }
          if (EPI != 2 && p.bias_n) {
            const float4* b4 = reinterpret_cast<const float4*>(sbias + hc0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = b4[j];
              f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
            }
          } else if (EPI != 2 && p.bias) {
            if (hc0 + 32 <= p.N) {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + hc0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(b4 + j);
                f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (hc0 + j < p.N) f[j] += __ldg(p.bias + hc0 + j);
            }
          }
          if (use_drop) {
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            const uint32_t wbase = static_cast<uint32_t>(orow) * p.drop_wpr;
            if (p.drop_gshift == 0) {
              const uint32_t c0 = wbase + static_cast<uint32_t>(hc0 >> 1);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const uint32_t w = drop_word(dkey, c0 + j);
                f[2 * j] = drop_keep(w, 0, p.drop.thr) ? f[2 * j] * p.drop.scale : 0.f;
                f[2 * j + 1] = drop_keep(w, 1, p.drop.thr) ? f[2 * j + 1] * p.drop.scale : 0.f;
              }
            } else {  // groups of >= 32 columns (whole heads): one decision for this 32-column half
              const int e = hc0 >> p.drop_gshift;
              const uint32_t w = drop_word(dkey, wbase + static_cast<uint32_t>(e >> 1));
              const float ks = drop_keep(w, e & 1, p.drop.thr) ? p.drop.scale : 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] *= ks;
            }
          }
          if (f_res) {
            const uint8_t* row = ein + r * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 tt = *reinterpret_cast<const uint4*>(row + (((half * 4 + j) ^ sw128) << 4));
              f[8 * j] += t_lo(tt.x); f[8 * j + 1] += t_hi(tt.x); f[8 * j + 2] += t_lo(tt.y); f[8 * j + 3] += t_hi(tt.y);
              f[8 * j + 4] += t_lo(tt.z); f[8 * j + 5] += t_hi(tt.z); f[8 * j + 6] += t_lo(tt.w); f[8 * j + 7] += t_hi(tt.w);
            }
          }
          if (f_res32) {
            const uint8_t* row = ein + p.ein_off_res32 + half * 16384 + r * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = *reinterpret_cast<const float4*>(row + ((j ^ sw128) << 4));
              f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
            }
          }
          // convolution path (16-bit output only): the ReLU rides on the f32 -> 16-bit conversion below
          const bool relu_in_pack = EPI == 1 && p.relu;
          if (p.relu && !use_drop && !relu_in_pack) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (f_mask) {
            const uint8_t* row = ein + p.ein_off_mask + r * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 tt = *reinterpret_cast<const uint4*>(row + (((half * 4 + j) ^ sw128) << 4));
              const uint32_t w[4] = {tt.x, tt.y, tt.z, tt.w};
#pragma unroll
              for (int ee = 0; ee < 4; ++ee) {
                f[8 * j + 2 * ee] = t_pos_lo(w[ee]) ? f[8 * j + 2 * ee] * p.mask_scale : 0.f;
                f[8 * j + 2 * ee + 1] = t_pos_hi(w[ee]) ? f[8 * j + 2 * ee + 1] * p.mask_scale : 0.f;
              }
            }
          }
          if (__any_sync(0xffffffffu, !interior)) {  // a real (warp-uniform) branch: 4 warps in 5 hold no border pixel and skip 32 selects
            if (!interior) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = 0.f;
            }
          }
          if (f_out && !(p.debug & 16)) {
            uint8_t* row = oslot + r * 128;
            if (relu_in_pack) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 tt;
                tt.x = pack_t2_relu(f[8 * j], f[8 * j + 1]); tt.y = pack_t2_relu(f[8 * j + 2], f[8 * j + 3]);
                tt.z = pack_t2_relu(f[8 * j + 4], f[8 * j + 5]); tt.w = pack_t2_relu(f[8 * j + 6], f[8 * j + 7]);
                *reinterpret_cast<uint4*>(row + (((half * 4 + j) ^ sw128) << 4)) = tt;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 tt;
                tt.x = pack_t2(f[8 * j], f[8 * j + 1]); tt.y = pack_t2(f[8 * j + 2], f[8 * j + 3]);
                tt.z = pack_t2(f[8 * j + 4], f[8 * j + 5]); tt.w = pack_t2(f[8 * j + 6], f[8 * j + 7]);
                *reinterpret_cast<uint4*>(row + (((half * 4 + j) ^ sw128) << 4)) = tt;
              }
            }
          }
          if (f_out32) {
            uint8_t* row = oslot + p.out_off_f32 + half * 16384 + r * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(row + ((j ^ sw128) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
        }
        if (f_atomic) continue;
        if (ein) {  // this warp is done with the input slot
          __syncwarp();
          if (lane == 0) mbar_arrive(&ein_empty[es]);
        }
        if (!(p.debug & 32)) fence_proxy_async_smem();
        // two slots: the store issued one chunk ago must have read ITS slot before the next chunk overwrites it; checking that
        // here (instead of before writing) needs a single barrier per chunk
        if (two_slots && store_thread && !(p.debug & 32)) tma_store_wait_read<0>();
        if (!(p.debug & 32)) named_bar_sync(1 + team, TEAM_WARPS * 32);
        if (store_thread && !(p.debug & 1)) {
          if (f_out) tma_store_2d(&tmOut, oslot, col0, m0s);
          if (f_out32) {
            tma_store_2d(&tmOut32, oslot + p.out_off_f32, col0, m0s);
            if (col0 + 32 < p.N) tma_store_2d(&tmOut32, oslot + p.out_off_f32 + 16384, col0 + 32, m0s);
          }
          tma_store_commit();
        }
      }
      }
      // every TMEM read of this accumulator stage by this warp has completed (tcgen05.wait::ld): release it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (cl == 2) {
          if (rank == 0) mbar_arrive(&acc_empty[as]);
          else mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[as]), 0));  // pairs: the leader's MMA issuer waits for both CTAs' epilogues
        } else {
          mbar_arrive(&acc_empty[as]);
        }
      }
      ++tcount;
    }
    if (store_thread) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (cl == 2) cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers / read its shared memory
  if (warp == 1) {
    if constexpr (cl == 2) tmem_dealloc_cg2<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace rb

using namespace rb;

static int pick_bn(int N, long long tiles_mz, long long k_iters, int nsm) {
  // cost model (arbitrary time units): waves x (fixed per-tile latency + part that grows with the tile width and the main-loop
  // length).  Few waves matter most for small problems; wide tiles re-read A less and amortise the epilogue barriers.
  const int cands[3] = {256, 128, 64};
  int best = 64;
  double best_cost = 1e300;
  for (int bn : cands) {
    if (bn >= 2 * N && bn > 64) continue;  // more than half of the tile would be padding
    const long long tiles = tiles_mz * ((N + bn - 1) / bn);
    const long long waves = (tiles + nsm - 1) / nsm;
    const double tile_cost = 1.0 + (bn / 64) * (0.1 + 0.02 * static_cast<double>(k_iters));
    const double cost = static_cast<double>(waves) * tile_cost;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

// Run-time model of one launch (clocks): the kernel is persistent with a static tile order, so it takes waves x (time of one tile); a
// k-block of a tile costs max(MMA time, its operand bytes / the tile's share of the chip-wide L2 -> SM throughput) -- ~6300 B/clk
// for the whole chip (measured: 128 x 128 tiles of the layer-3 3x3 convolution pull 32 KB per k-block in 795 clk on 148 SMs = 41 B/clk
// per SM; the 128 x 256 weight-gradient tiles 48 KB in 1250 clk), i.e. the long-K GEMMs are bound by bytes per FLOP, which only a
// larger tile lowers.  `overlap_epi`: the accumulator is double buffered (128-row tiles), so only part of the epilogue is exposed.
static double tile_clocks(int bm, int bn, int cl, long long k_it, long long active, bool atomic_epi) {
  // (the cap is per SM as much as chip-wide: a launch that leaves SMs idle does not speed up the busy ones -- 87 us against 53 us for
  // the layer-2 3x3 weight gradient on 81 against 144 tiles)
  double rate = 6300.0 / static_cast<double>(active < 1 ? 1 : active);
  if (rate > 43.0) rate = 43.0;
  const double mma = (bm / 128) * 2.0 * bn;
  const double l2 = (bm + bn / cl) * 128.0 / rate;
  const double epi = (bm / 128) * (bn / 64) * (atomic_epi ? 700.0 : 450.0);
  return static_cast<double>(k_it) * (mma > l2 ? mma : l2) + (bm == 256 ? epi : 0.35 * epi) + 2500.0;
}
static double launch_clocks(int bm, int bn, int cl, long long tiles, long long k_it, int nsm, bool atomic_epi) {
  const long long slots = cl == 2 ? nsm / 2 : nsm;
  const long long full = tiles / slots, rem = tiles % slots;
  double t = static_cast<double>(full) * tile_clocks(bm, bn, cl, k_it, nsm, atomic_epi);
  if (rem) t += tile_clocks(bm, bn, cl, k_it, rem * cl, atomic_epi);
  return t + 6000.0;  // launch, TMEM allocation, barrier setup, pipeline fill / drain
}

// Weight gradients (mode 1, split-K): tile width AND number of K splits together.  The kernel is persistent with a static tile
// order, so its run time is waves x (time of one tile); a tile's time is its share of the L2 -> shared-memory operand stream
// (k-blocks x (128 + bn) x 128 B: these GEMMs are bound by that stream) plus the atomic epilogue and a fixed
// fill / drain cost.  Minimising that over (bn, splits) lands on single-wave configurations with 128..148 tiles instead of
// e.g. 222 tiles in two half-empty waves.
static void pick_tn(int M, int N, int taps, int kblocks, int nsm, bool has_bias_grad, int fixed_bn, int cl_mode, int* bn_out, int* splits_out, int* cl_out) {
  const int cands[3] = {256, 128, 64};
  double best = 1e300;
  *bn_out = 64; *splits_out = 1; *cl_out = 1;
  const long long tiles_m = (M + BM - 1) / BM;
  for (int bn : cands) {
    if (fixed_bn && bn != fixed_bn) continue;
    if (!fixed_bn && ((bn >= 2 * N && bn > 64) || (has_bias_grad && bn == 256))) continue;
    for (int cl = 1; cl <= 2; ++cl) {  // cl_mode: 0 = never pairs, 1 = pairs wherever legal
      const bool legal = bn >= 128 && tiles_m >= 2;
      if (cl == 2 && (!legal || cl_mode == 0)) continue;
      if (cl == 1 && legal && cl_mode == 1) continue;
      const long long base = (cl == 2 ? (tiles_m + 1) / 2 : tiles_m) * ((N + bn - 1) / bn) * taps;
      const long long slots = cl == 2 ? nsm / 2 : nsm;
      const int smax = kblocks / 2 > 1 ? (kblocks / 2 < 512 ? kblocks / 2 : 512) : 1;
      for (int sp = 1; sp <= smax; ++sp) {
        const long long tiles = base * sp;
        const long long waves = (tiles + slots - 1) / slots;
        const int k_it = (kblocks + sp - 1) / sp;
        // per CTA and k-block: A (128 columns) + its share of B, 128 B per column; + the atomic epilogue and a fixed fill / drain cost
        const double cost = static_cast<double>(waves) * (k_it * (128.0 + bn / cl) * 0.125 + 1.5 * 0.5 * bn + 60.0);  // KB-equivalents
        if (cost < best - 1e-9) { best = cost; *bn_out = bn; *splits_out = sp; *cl_out = cl; }
        if (waves > 4 && sp > 1) break;  // more splits only add waves from here on
      }
    }
  }
}

extern "C" int rb_gemm(const rb_gemm_args* a, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!a || !a->A || !a->B) return rb_fail("rb_gemm: null operand");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return rb_fail("rb_gemm: empty problem (M,N,K must be > 0)");
  if (a->taps < 1 || a->taps > 16) return rb_fail("rb_gemm: taps must be in [1,16]");
  if ((a->lda % 8) || (a->ldb % 8)) return rb_fail("rb_gemm: operand row pitch must be a multiple of 8 elements (16 B) for TMA");
  if ((reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->B) & 15)) return rb_fail("rb_gemm: operands must be 16-byte aligned");
  if (!a->out && !a->out32) return rb_fail("rb_gemm: no output");
  if (a->atomic && !a->out32) return rb_fail("rb_gemm: atomic accumulation needs out32");
  if (a->mode != 0 && a->mode != 1) return rb_fail("rb_gemm: mode must be 0 (NT) or 1 (TN)");
  if (a->mode == 1 && !a->atomic && (a->taps != 1 || a->splits > 1)) return rb_fail("rb_gemm: TN mode with several taps / splits needs atomic accumulation");
  if (a->out && ((a->ldo % 8) || (reinterpret_cast<uintptr_t>(a->out) & 15))) return rb_fail("rb_gemm: out must be 16-byte aligned with pitch % 8 == 0");
  if (a->out32 && !a->atomic && ((a->ldo32 % 4) || (reinterpret_cast<uintptr_t>(a->out32) & 15))) return rb_fail("rb_gemm: out32 must be 16-byte aligned with pitch % 4 == 0");
  if (a->res && ((a->ldres % 8) || (reinterpret_cast<uintptr_t>(a->res) & 15))) return rb_fail("rb_gemm: res alignment");
  if (a->res32 && ((a->ldres32 % 4) || (reinterpret_cast<uintptr_t>(a->res32) & 15))) return rb_fail("rb_gemm: res32 alignment");
  if (a->mask_src && ((a->ldmask % 8) || (reinterpret_cast<uintptr_t>(a->mask_src) & 15))) return rb_fail("rb_gemm: mask_src alignment");
  if (a->bias && (reinterpret_cast<uintptr_t>(a->bias) & 15)) return rb_fail("rb_gemm: bias must be 16-byte aligned");

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.mode = a->mode;
  kp.M = a->M; kp.N = a->N; kp.taps = a->taps; kp.splits = a->splits < 1 ? 1 : a->splits;
  int auto_bn = 0, auto_cl = 0, auto_bm = 0;
  // RB_GEMM_CLUSTER: 0 = never use CTA pairs (cta_group::2), 1 = wherever legal, unset = by cost model
  // RB_GEMM_TALL:    0 = never use 256-row tiles, 1 = wherever legal, unset = by cost model
  // RB_GEMM_CLUSTER: 0 = never use CTA pairs (cta_group::2), 1 = wherever legal, unset = where they measured faster (below)
  // RB_GEMM_TALL:    1 = 256-row tiles wherever legal (experiments), otherwise never
  // Measured on the step's shapes (profiles/r02_gemm_pairs_tall.log, same box, CUDA events): pairs take 8..9 % off the 3x3 weight
  // gradients and 12 % off the long-K input gradients whose only epilogue input is the ReLU mask, are neutral on the 1x1 weight
  // gradients and up to 1.8x SLOWER on output-heavy 1x1 convolutions (the leader's MMA waits for both CTAs' epilogues); tall tiles
  // are neutral to 6 % slower everywhere (no accumulator double buffering) -- a measured dead end, kept behind the switch.
  static const int cl_env = [] { const char* e = getenv("RB_GEMM_CLUSTER"); return e ? (atoi(e) ? 1 : 0) : 2; }();
  static const int tall_mode = [] { const char* e = getenv("RB_GEMM_TALL"); return (e && atoi(e)) ? 1 : 0; }();
  int cl_mode = cl_env;
  if (cl_env == 2) {
    const int kb = (a->K + BK - 1) / BK;
    if (a->mode == 1) cl_mode = a->taps >= 9 ? 1 : 0;
    else cl_mode = (a->taps == 1 && kb >= 16 && a->mask_src && !a->res && !a->res32 && !a->out32) ? 1 : 0;
  }
  // SMs this launch may use: all of them, or rb_gemm_args.sm_limit for launches that run beside other kernels
  const int nsm = (a->sm_limit > 0 && a->sm_limit < sm_count()) ? (a->sm_limit < 2 ? 2 : a->sm_limit) : sm_count();
  if (a->mode == 1 && a->atomic && a->splits <= 0) {  // splits <= 0: chosen here, together with the tile shape
    int bn_t = 0, sp_t = 1, cl_t = 1;
    pick_tn(a->M, a->N, a->taps, (a->K + BK - 1) / BK, nsm, a->bias_grad != nullptr, a->block_n == 32 ? 64 : a->block_n, cl_mode, &bn_t, &sp_t, &cl_t);
    kp.splits = sp_t;
    auto_bn = bn_t;
    auto_cl = cl_t;
  }
  for (int i = 0; i < 16; ++i) { kp.a_rowoff[i] = a->a_rowoff[i]; kp.b_koff[i] = a->b_koff[i]; }
  kp.out_row_off = a->out_row_off;
  kp.bias = a->bias;
  kp.relu = a->relu; kp.atomic = a->atomic; kp.geom = a->geom;
  kp.kblocks = (a->K + BK - 1) / BK;
  {
    const char* dbg = getenv("RB_GEMM_DEBUG");
    kp.debug = dbg ? atoi(dbg) : 0;
  }
  kp.bias_grad = a->bias_grad;
  kp.row_scale = a->row_scale;
  kp.out_scale = a->out_scale == 0.f ? 1.f : a->out_scale;
  if ((a->bias_grad || a->row_scale || kp.out_scale != 1.f) && !a->atomic) return rb_fail("rb_gemm: bias_grad / row_scale / out_scale need atomic accumulation");
  if (a->bias_grad && a->mode != 1) return rb_fail("rb_gemm: bias_grad needs mode 1 (TN)");
  kp.drop = make_dropk(a->drop);
  kp.drop_gshift = a->drop_gshift;
  kp.mask_scale = a->mask_scale == 0.f ? 1.f : a->mask_scale;
  if (kp.drop.seed) {
    if (a->atomic) return rb_fail("rb_gemm: dropout is not available in atomic mode");
    if (a->relu && (a->res || a->res32)) return rb_fail("rb_gemm: dropout with ReLU after a residual is not defined");
    if ((a->drop_gshift != 0 && (a->drop_gshift < 5 || a->drop_gshift > 16)) || (a->drop_gshift == 0 && (a->N & 1))) return rb_fail("rb_gemm: dropout needs an even N (gshift 0) or gshift >= 5");
    const long long cols = (static_cast<long long>(a->N) + (1LL << a->drop_gshift) - 1) >> a->drop_gshift;
    kp.drop_wpr = static_cast<uint32_t>((cols + 1) >> 1);
    if ((static_cast<long long>(a->M) + a->out_row_off) * kp.drop_wpr >= (1LL << 32)) return rb_fail("rb_gemm: dropout site too large for a 32-bit counter");
  }
  if (gemm_skinny_eligible(a) && !(getenv("RB_GEMM_NO_SKINNY"))) return gemm_skinny_launch(a, kp.drop, static_cast<int>(kp.drop_wpr), st);
  long long tiles_m = (a->M + BM - 1) / BM;
  const long long k_iters = a->mode == 0 ? static_cast<long long>(a->taps) * kp.kblocks : (kp.kblocks + kp.splits - 1) / kp.splits;
  int bn = a->block_n ? a->block_n : (auto_bn ? auto_bn : pick_bn(a->N, tiles_m * (a->mode == 1 ? a->taps * kp.splits : 1), k_iters, nsm));
  // output/residual-dominated problems (short main loop + epilogue inputs): narrower tiles leave room for a deep input ring
  if (!a->block_n && bn == 256 && !a->atomic && (a->res || a->res32 || a->mask_src) && k_iters <= 4) bn = 128;
  if (bn == 32) bn = 64;
  if (a->bias_grad && bn == 256) bn = 128;  // 2 x 16 TMEM columns for the bias-gradient accumulators next to 2 x bn
  if (bn != 64 && bn != 128 && bn != 256) return rb_fail("rb_gemm: unsupported block_n %d", bn);
  // ---- 256-row tiles: long main loops only (the accumulator is not double buffered), never with the bias-gradient MMA --------------
  kp.bm = BM;
  if (auto_bm) {
    kp.bm = auto_bm;
  } else if (a->mode == 0 && tall_mode != 0 && !a->block_n && a->M > BM && k_iters >= 8 && !kp.drop.seed && !a->out32 && !a->res32 && a->out && !a->atomic) {
    // NT: compare the chosen 128-row shape with the tall candidates under the same model
    const bool ein = a->res || a->res32 || a->mask_src;
    double best = launch_clocks(BM, bn, 1, tiles_m * ((a->N + bn - 1) / bn), k_iters, nsm, false);
    if (tall_mode == 1) best = 1e300;
    const int tb[2] = {256, 128};
    for (int b : tb) {
      if (b >= 2 * a->N && b > 64) continue;
      if (ein && b == 256) continue;  // epilogue inputs + 64 KB stages do not leave room for a useful ring
      const long long tl = ((a->M + 255) / 256) * ((a->N + b - 1) / b);
      const double c = launch_clocks(256, b, 1, tl, k_iters, nsm, false);
      if (c < best) { best = c; kp.bm = 256; bn = b; }
    }
  }
  if (kp.bm == 256) tiles_m = (a->M + 255) / 256;
  kp.bn = bn;
  kp.tiles_m = static_cast<int>(tiles_m);
  kp.tiles_n = (a->N + bn - 1) / bn;
  // CTA pairs (cta_group::2): legal for bn >= 128 and at least two row tiles.  Weight gradients: decided by pick_tn's cost model;
  // NT: wherever the main loop is long enough for the operand stream to matter (K per tile >= 256) and the padding of an odd last
  // pair is small
  kp.cl = 1;
  {
    const bool legal = kp.bm == BM && bn >= 128 && tiles_m >= 2 && nsm >= 2;
    if (legal && cl_mode != 0) {
      if (auto_cl) kp.cl = auto_cl;
      else if (cl_mode == 1) kp.cl = 2;
      // (NT, by default: no pairs -- measured neutral to slower on the convolution shapes, profiles/r02_gemm_pairs_tall.log)
    }
  }
  kp.tiles_mp = static_cast<int>((tiles_m + 1) / 2);
  const long long total = (kp.cl == 2 ? kp.tiles_mp : tiles_m) * kp.tiles_n * (a->mode == 1 ? a->taps * kp.splits : 1);
  if (total > 0x7fffffffLL) return rb_fail("rb_gemm: too many tiles");
  kp.tiles_total = static_cast<int>(total);
  if (a->atomic) {
    kp.out32 = a->out32; kp.ldo32 = a->ldo32; kp.out32_z_stride = a->out32_z_stride;
  } else {
    kp.has_out = a->out != nullptr; kp.has_out32 = a->out32 != nullptr;
    kp.has_res = a->res != nullptr; kp.has_res32 = a->res32 != nullptr; kp.has_mask = a->mask_src != nullptr;
  }
  // ---- shared-memory plan --------------------------------------------------------------------------------------
  kp.stage_bytes = (kp.bm / BM) * A_BYTES + (bn / kp.cl) * 128;  // pairs: each CTA stages half of the B tile
  const bool has_ein = kp.has_res || kp.has_res32 || kp.has_mask;
  kp.ein_off_mask = kp.has_res ? 16384 : 0;
  kp.ein_off_res32 = kp.ein_off_mask + (kp.has_mask ? 16384 : 0);
  kp.ein_slot_bytes = kp.ein_off_res32 + (kp.has_res32 ? 32768 : 0);
  kp.out_off_f32 = kp.has_out ? 16384 : 0;
  kp.out_slot_bytes = kp.out_off_f32 + (kp.has_out32 ? 32768 : 0);
  const long long bias_pad = static_cast<long long>(kp.tiles_n) * bn;
  kp.bias_n = (a->bias && !a->atomic && bias_pad <= 2048) ? static_cast<int>(bias_pad) : 0;
  const uint32_t ones_bytes = a->bias_grad ? 8192u + 1024u : 0u;
  uint32_t bars_bytes = 512 + static_cast<uint32_t>(kp.bias_n) * 4 + ones_bytes;
  // short main loops do not need a deep ring: the room goes to the epilogue rings instead (HBM-bound 1x1 convolutions, where
  // the residual / mask stream is as large as the output)
  const int want_stages = k_iters * 2 < 4 ? 4 : static_cast<int>(k_iters * 2 > MAX_STAGES ? MAX_STAGES : k_iters * 2);
  int stages = 0;
  kp.ein_slots = 0;
  kp.out_slots = 2;
  for (int pass = 0; pass < 3 && stages < 2; ++pass) {
    kp.out_slots = (pass == 0 && kp.bm == BM) ? 2 : 1;  // second pass (and tall tiles): one staging slot per team
    if (pass == 2) {                   // third pass: the bias stays in global memory as well
      if (!kp.bias_n) break;
      kp.bias_n = 0;
      bars_bytes = 512 + ones_bytes;
    }
    const long long room = static_cast<long long>(SMEM_LIMIT) - (2LL * kp.out_slots * kp.out_slot_bytes + bars_bytes + 1024 /* alignment slack */);
    if (has_ein) {
      // the epilogue-input ring is what keeps HBM busy when the main loop is short: give it up to MAX_EIN slots after a minimal
      // main-loop ring, then hand what is left back to the main loop
      const int min_stages = want_stages < 3 ? want_stages : (k_iters <= 2 ? 2 : 3);
      long long slots = (room - static_cast<long long>(min_stages) * kp.stage_bytes) / kp.ein_slot_bytes;
      if (slots > MAX_EIN) slots = MAX_EIN;
      if (slots < 2) slots = 2;
      kp.ein_slots = static_cast<int>(slots);
      const long long left = room - slots * kp.ein_slot_bytes;
      stages = left > 0 ? static_cast<int>(left / kp.stage_bytes) : 0;
    } else {
      stages = static_cast<int>(room / kp.stage_bytes);
    }
  }
  if (stages > want_stages) stages = want_stages;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return rb_fail("rb_gemm: shared-memory plan failed (bn=%d)", bn);
  kp.stages = stages;
  kp.off_ein = stages * kp.stage_bytes;
  kp.off_out = kp.off_ein + kp.ein_slots * kp.ein_slot_bytes;
  kp.off_bars = kp.off_out + 2 * kp.out_slots * kp.out_slot_bytes;
  kp.off_bias = kp.off_bars + 512;
  kp.off_ones = (kp.off_bias + static_cast<uint32_t>(kp.bias_n) * 4 + 1023u) & ~1023u;
  const int smem_bytes = static_cast<int>(kp.off_bars + bars_bytes + 1024);
  if (smem_bytes > SMEM_LIMIT) return rb_fail("rb_gemm: shared-memory plan exceeds the limit (%d B)", smem_bytes);

  // ---- tensor maps ---------------------------------------------------------------------------------------------
  CUtensorMap tmA, tmB, tmOut, tmOut32, tmRes, tmRes32, tmMask;
  if (a->mode == 0) {
    if (make_tmap_2d(&tmA, a->A, static_cast<uint64_t>(a->a_cols), static_cast<uint64_t>(a->a_rows), a->lda * 2, 64, BM)) return 1;
    if (make_tmap_2d(&tmB, a->B, static_cast<uint64_t>(a->b_cols), static_cast<uint64_t>(a->b_rows), a->ldb * 2, 64, bn / kp.cl)) return 1;  // pairs: half a tile per CTA
  } else {
    if (make_tmap_2d(&tmA, a->A, static_cast<uint64_t>(a->a_cols), static_cast<uint64_t>(a->a_rows), a->lda * 2, 64, 64)) return 1;
    if (make_tmap_2d(&tmB, a->B, static_cast<uint64_t>(a->b_cols), static_cast<uint64_t>(a->b_rows), a->ldb * 2, 64, 64)) return 1;
  }
  tmOut = tmA; tmOut32 = tmA; tmRes = tmA; tmRes32 = tmA; tmMask = tmA;  // placeholders for unused maps
  const uint64_t M64 = static_cast<uint64_t>(a->M), N64 = static_cast<uint64_t>(a->N);
  if (kp.has_out && make_tmap_2d(&tmOut, static_cast<const rb_t*>(a->out) + a->out_row_off * a->ldo, N64, M64, a->ldo * 2, 64, BM)) return 1;
  if (kp.has_out32 && make_tmap_2d_f32(&tmOut32, a->out32 + a->out_row_off * a->ldo32, N64, M64, a->ldo32 * 4, 32, BM)) return 1;
  if (kp.has_res && make_tmap_2d(&tmRes, static_cast<const rb_t*>(a->res) + a->out_row_off * a->ldres, N64, M64, a->ldres * 2, 64, BM)) return 1;
  if (kp.has_res32 && make_tmap_2d_f32(&tmRes32, a->res32 + a->out_row_off * a->ldres32, N64, M64, a->ldres32 * 4, 32, BM)) return 1;
  if (kp.has_mask && make_tmap_2d(&tmMask, static_cast<const rb_t*>(a->mask_src) + a->out_row_off * a->ldmask, N64, M64, a->ldmask * 2, 64, BM)) return 1;

  const int slots = kp.cl == 2 ? nsm / 2 : nsm;
  const int grid = kp.cl * (kp.tiles_total < slots ? kp.tiles_total : slots);
  const bool lean = !a->atomic && kp.has_out && !kp.has_out32 && !kp.has_res32 && !kp.drop.seed;
  const int epi = a->atomic ? 2 : (lean ? 1 : 0);
  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, GemmKParams);
  static const KernelFn kernels[2][2][3] = {{{umma_gemm_kernel<0, 0, 1, 1>, umma_gemm_kernel<0, 1, 1, 1>, umma_gemm_kernel<0, 2, 1, 1>},
                                             {umma_gemm_kernel<1, 0, 1, 1>, umma_gemm_kernel<1, 1, 1, 1>, umma_gemm_kernel<1, 2, 1, 1>}},
                                            {{umma_gemm_kernel<0, 0, 2, 1>, umma_gemm_kernel<0, 1, 2, 1>, umma_gemm_kernel<0, 2, 2, 1>},
                                             {umma_gemm_kernel<1, 0, 2, 1>, umma_gemm_kernel<1, 1, 2, 1>, umma_gemm_kernel<1, 2, 2, 1>}}};
  static bool configured[2][2][3] = {};
  static bool configured_tall = false;
  KernelFn kern = kernels[kp.cl - 1][a->mode][epi];
  if (kp.bm == 256) {
    if (a->mode != 0 || epi != 1 || kp.cl != 1) return rb_fail("rb_gemm: 256-row tiles exist for the NT convolution path only");
    kern = umma_gemm_kernel<0, 1, 1, 2>;
    if (!configured_tall) {
      RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
      configured_tall = true;
    }
  } else if (!configured[kp.cl - 1][a->mode][epi]) {
    RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    configured[kp.cl - 1][a->mode][epi] = true;
  }
  if (kp.cl == 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    RB_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, tmOut32, tmRes, tmRes32, tmMask, kp));
  } else {
    kern<<<grid, GEMM_THREADS, smem_bytes, st>>>(tmA, tmB, tmOut, tmOut32, tmRes, tmRes32, tmMask, kp);
  }
  RB_CUDA(cudaGetLastError());
  return 0;
}
