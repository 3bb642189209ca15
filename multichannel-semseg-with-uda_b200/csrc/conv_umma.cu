// conv_umma.cu — tcgen05 / TMEM / TMA implicit-GEMM convolution kernels for sm_100a.
//
//   fprop / dgrad ("tap problem" of conv_plan.h):
//     D[128 pixels, BN channels] = sum over taps, 64-channel chunks  A_tap[128 x 64] * W_tap[BN x 64]^T
//     A_tap is ONE 4-D TMA box {64 ch, TW, TH, 1 image} of the NHWC activations, shifted by the tap's
//     (dh, dw); out-of-image coordinates are zero-filled by TMA, which implements the padding
//     (models/drn.py:21-23 conv3x3 padding=dilation).  Strided convs read parity sub-grids of the
//     image through up to 4 tensor maps.  Both operands are K-major, 128B-swizzled.
//   wgrad:
//     dW_tap[128 co, BN ci] = sum over 64-pixel tiles  dY[pix x co]^T * X_tap[pix x ci]
//     both operands MN-major (channels contiguous), split over pixel ranges into a workspace that
//     wgrad_reduce sums into the fp32 OIHW gradient.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
#include <stdlib.h>
#include "common.cuh"
#include "conv_plan.h"
#include "umma_ptx.cuh"

namespace mcd {

using namespace ptx;

constexpr int kThreads = 192;
// epilogue warpgroups of the CTA-pair kernels.  2 (one per TMEM accumulator buffer) was measured SLOWER on B200
// (r01b: forward 25.2 -> 29.1 ms, dgrad 24.1 -> 29.4 ms per iteration): 320 threads cap the epilogue at 168
// registers (270 B of spills) and the extra warps compete with the MMA / TMA warps for issue slots.
#ifndef MCD_SPLIT_STATS
#define MCD_SPLIT_STATS 1
#endif
#ifndef MCD_ACCSTAT
#define MCD_ACCSTAT 0      /* widest tile that keeps per-thread statistics (0 = off, 32 / 64 under test) */
#endif
#ifndef MCD_PAIR_EPI_WG
#define MCD_PAIR_EPI_WG 1
#endif

struct UmmaMaps {
  CUtensorMap a[4];  // activation maps (parity sub-grids for strided problems)
  CUtensorMap b;     // fprop: packed weights; wgrad: dY
};

struct FpropArgs {
  int N, Ht, Wt, TH, TW, tiles_h, tiles_w;
  int tiles_m, tiles_n;   // tiles_m = N * tiles_h * tiles_w pixel tiles, tiles_n = channel tiles of BN
  int kchunks, ntaps, kc_pad;
  int rows;               // produced channels
  int omul, oh0, ow0, Hd, Wd, Cd_s;
  int planar;
  int fmt;                // element format of src / weights / nhwc output: kF16 (forward) or kBF16 (dgrad)
  int packed;             // row-packed thin-channel mode: A = TWp window loads of THp rows each
  int cs_src, smul;       // packed: source channel stride, W stride of the convolution
  const float* bias;
  int relu;               // nhwc output: max(., 0) after bias and addend - the eval-mode unit conv -> folded BatchNorm -> ReLU
  float* stats;
  void* out;
  const __nv_bfloat16* addend;   // optional nhwc tensor (output geometry, element format `fmt`) added in the epilogue:
                                 // the residual gradient (dgrad) or the residual branch of an eval-mode block (forward)
  // dgrad fused with the backward of the BatchNorm+ReLU unit that produced the convolution's input (all nhwc,
  // output geometry): out = mask_src > 0 ? out : 0 and, with bn_y, stats += {sum out, sum out * bn_y}
  const __nv_bfloat16* mask_src;  // bf16 twin of the activation (only its sign is used)
  const __half* bn_y;             // forward tensor: IEEE half
  // stream-K schedule (sk_units > 0): CTA i owns the (tile, k-block) units [i * sk_units, (i+1) * sk_units) of the
  // linearised tile x k-block space, so every CTA does the same amount of MMA work whatever the tile count.  A
  // tile cut by a CTA boundary is finished by the CTA that started it (its LAST segment): the next CTA computes
  // the remaining k-blocks FIRST, parks the fp32 partial accumulator in sk_partial[cta] and raises sk_flags[cta].
  int sk_units;
  float* sk_partial;          // [gridDim.x][BN/32][128][32] fp32
  int* sk_flags;              // [gridDim.x], zero on entry
  // halo-tile mode (HALO kernels): ONE TMA box {64 ch, halo_wb, halo_hb} per (tile, 64-channel chunk) holds every
  // pixel any tap of the 8-wide x 16-tall output tile reads; tap t's A operand is the same buffer entered
  // taps[t].pad0 pixel rows further down (descriptor start + pad0 * 128 B, stride between 8-pixel groups =
  // halo_wb * 128 B).  The 128-byte swizzle is a function of absolute shared-memory address bits
  // (profiles/r01b_exp_swizzle_shift.txt), so a descriptor may start at any row of a tile TMA wrote once.
  int halo_bytes;             // buffer stride (box bytes rounded up to 1024)
  int halo_tx;                // bytes one halo box delivers
  int halo_wb;                // box width in pixels
  int halo_oh, halo_ow;       // box origin relative to the tile origin (most negative tap offset)
  int stages;                 // weight-tile ring depth (runtime in halo mode)
  Tap taps[kMaxTaps];
};

// the (tile, k-block range) segments of one CTA, identical for the producer, MMA and epilogue roles.
// Stream-K order: (1) the tail k-blocks of the tile that starts before this CTA's range (partial, parked for the
// owner), (2) the head k-blocks of the tile that ends after the range (this CTA owns it; the partner parked the
// rest as ITS first segment, so the merge never waits long and overlaps the remaining MMAs), (3) the whole tiles.
struct SegIter {
  int64_t u, u_end;        // data-parallel: tile cursor / tile count; stream-K: whole-tile cursor / end (in units)
  int kblocks, stride, sk;
  int head_tile, head_kb1; // stream-K: pending owned partial tile (kb 0 .. head_kb1), -1 = none
  int tail_tile, tail_kb0; // stream-K: pending foreign partial tile (kb tail_kb0 .. kblocks), -1 = none
  int tile, kb0, kb1;
  __device__ __forceinline__ bool next() {
    if (sk) {
      if (tail_tile >= 0) { tile = tail_tile; kb0 = tail_kb0; kb1 = kblocks; tail_tile = -1; return true; }
      if (head_tile >= 0) { tile = head_tile; kb0 = 0; kb1 = head_kb1; head_tile = -1; return true; }
      if (u >= u_end) return false;
      tile = (int)(u / kblocks); kb0 = 0; kb1 = kblocks;
      u += kblocks;
      return true;
    }
    if (u >= u_end) return false;
    tile = (int)u; kb0 = 0; kb1 = kblocks;
    u += stride;
    return true;
  }
};

__device__ __forceinline__ SegIter make_seg_iter(const FpropArgs& a, int total_tiles, int kblocks) {
  SegIter it;
  it.kblocks = kblocks; it.stride = gridDim.x; it.sk = a.sk_units > 0;
  it.head_tile = it.tail_tile = -1; it.head_kb1 = it.tail_kb0 = 0;
  it.tile = 0; it.kb0 = 0; it.kb1 = 0;
  if (it.sk) {
    const int64_t total = (int64_t)total_tiles * kblocks;
    int64_t b = (int64_t)blockIdx.x * a.sk_units;
    int64_t e = b + a.sk_units < total ? b + a.sk_units : total;
    if (b >= e) { it.u = it.u_end = 0; return it; }
    const int bt = (int)(b / kblocks), bk = (int)(b - (int64_t)bt * kblocks);
    const int et = (int)(e / kblocks), ek = (int)(e - (int64_t)et * kblocks);
    if (bk > 0) { it.tail_tile = bt; it.tail_kb0 = bk; b = (int64_t)(bt + 1) * kblocks; }   // host: sk_units >= kblocks
    if (ek > 0) { it.head_tile = et; it.head_kb1 = ek; e = (int64_t)et * kblocks; }
    it.u = b; it.u_end = e;
  } else {
    it.u = blockIdx.x; it.u_end = total_tiles;
  }
  return it;
}

// CTA pairs walk pair tiles cluster by cluster
__device__ __forceinline__ SegIter make_pair_iter(int total_pair_tiles, int kblocks) {
  SegIter it;
  it.kblocks = kblocks; it.stride = gridDim.x / 2; it.sk = 0;
  it.head_tile = it.tail_tile = -1; it.head_kb1 = it.tail_kb0 = 0;
  it.tile = 0; it.kb0 = 0; it.kb1 = 0;
  it.u = blockIdx.x / 2; it.u_end = total_pair_tiles;
  return it;
}

// PAIR: two CTAs of a cluster (one TPC) run ONE 256-pixel x 256-channel tcgen05.mma.cta_group::2 tile: each CTA
// stages its own 128 pixels (A) and HALF of the weight tile (B), so a k-block costs 32 KB of TMA writes + 32 KB of
// operand reads per SM instead of 48 + 48 - the single-CTA 128x256 tile is shared-memory-bandwidth bound
// (96 KB per 512 MMA cycles at 128 B/clk ~ 68 % of the tensor peak, which is what ncu shows).
// OCC = 2: half the pipeline stages so that two persistent CTAs share an SM (thin layers, BN 64 / 128): their
// per-tile chains (TMA -> MMA -> commit -> epilogue) are latency-bound, a second CTA fills the bubbles.
// EPI2 (r02): TWO epilogue warpgroups in ONE CTA, one per TMEM accumulator buffer (even / odd tiles).  Used by the
// dgrad-epilogue instantiations of the 64- / 128-channel tiles: their epilogue (three extra 16-byte loads per 8 channels,
// ReLU mask, two BatchNorm sums) is bound by the issue latency of ONE warp per scheduler (IPC 0.87 per SM at 246 registers,
// one CTA per SM: profiles/r02_ncu_full_kernels.csv), a second CTA does not fit the register file, eight epilogue warps at
// <= 204 registers do.
#ifndef MCD_THIN_EPI2
#define MCD_THIN_EPI2 1
#endif
constexpr bool fprop_epi2(int BN, bool PAIR, int OCC, bool EXTRAS) {
  return MCD_THIN_EPI2 && EXTRAS && !PAIR && OCC == 1 && (BN == 64 || BN == 128);
}

template <int BN, bool PAIR = false, int OCC = 1, bool EPI2 = false>
struct FpropCfg {
  static constexpr int A_BYTES = 128 * 128;          // 128 pixels x 64 ch bf16
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * 128;   // rows x 64 ch bf16
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = PAIR ? 6 : (OCC == 2 ? (BN == 128 ? 3 : 4)
                                                     : ((BN == 256) ? 4 : (BN == 128 ? 6 : (BN == 64 ? 8 : 4))));
  static constexpr int STAT_ROWS = 1024;             // BatchNorm partial sums staged in smem up to this many channels
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ +
                                    2 * STAT_ROWS * 4;
  static constexpr int ACC_COLS = BN < 32 ? 32 : BN;          // one accumulator
  static constexpr int TMEM_COLS = 2 * ACC_COLS;              // double-buffered (power of two, <= 512)
  static constexpr int MIN_CTAS = (BN <= 32 || OCC == 2) ? 2 : 1;   // thin tiles: two CTAs per SM
  // ACCSTAT: tiles of <= 64 channels have ONE channel tile (tiles_n == 1), so every tile of a persistent CTA covers
  // the same channels: each epilogue thread keeps its BatchNorm partial sums in registers across ALL its tiles and
  // the warp transpose-reduce runs once per kernel instead of once per tile and 32-channel chunk (these layers have
  // 9 - 36 k-blocks of MMA work per tile, the per-tile butterfly made their epilogue the bottleneck)
  static constexpr bool ACCSTAT = !PAIR && BN <= MCD_ACCSTAT;
  static constexpr int ACC_N = BN < 32 ? 32 : BN;
  // CTA pairs: TWO epilogue warpgroups, one per TMEM accumulator buffer (even / odd tiles), so that the epilogue of
  // tile i+1 starts while tile i's is still running (the fused dgrad epilogue of the 256-channel layers is longer
  // than their main loop)
  static constexpr int EPI_WG = ((PAIR && MCD_PAIR_EPI_WG == 2) || EPI2) ? 2 : 1;
  static constexpr int THREADS = 64 + 128 * EPI_WG;
};

// column sums over the 32 lanes (pixels) of a warp for 32 columns: butterfly transpose-reduce, lane j ends up with
// the totals of column j in s1[0] / s2[0]
__device__ __forceinline__ void warp_colsum32(float* s1, float* s2, int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool up = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      float send1 = up ? s1[i] : s1[i + step];
      float keep1 = up ? s1[i + step] : s1[i];
      s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, step);
      float send2 = up ? s2[i] : s2[i + step];
      float keep2 = up ? s2[i + step] : s2[i];
      s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, step);
    }
  }
}

// one statistic at a time (32 instead of 64 live sum registers): used by the dgrad-epilogue instantiations, whose
// register budget is the tightest
__device__ __forceinline__ void warp_colsum32_single(float* s1, int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool up = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      float send1 = up ? s1[i] : s1[i + step];
      float keep1 = up ? s1[i + step] : s1[i];
      s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, step);
    }
  }
}

// Persistent: every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... ; the smem ring runs across tile
// boundaries and the accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of tile i (TMEM ->
// registers -> bf16 / fp32 stores + BatchNorm statistics) overlaps the MMAs of tile i+1.
// EXTRAS = false: the forward instantiation, without the addend / ReLU-mask / BatchNorm-backward inputs of the dgrad
// epilogue (80 fewer live registers: the 168-register two-CTA variants stop spilling)
// the dgrad instantiations of the 16 / 32 channel tiles (layer2 / layer3.0 dgrad with the fused BatchNorm-backward
// epilogue) spill 300 B at the 168 registers two CTAs per SM leave them: MCD_THIN32_DGRAD_OCC1 gives them one CTA
#ifndef MCD_THIN32_DGRAD_OCC1
#define MCD_THIN32_DGRAD_OCC1 0   /* measured: layer2 dgrad 3.9 -> 4.85 ms with one CTA per SM */
#endif
template <int BN, bool PAIR, int OCC = 1, bool HALO = false, bool EXTRAS = true>
__global__ void __launch_bounds__(FpropCfg<BN, PAIR, OCC, fprop_epi2(BN, PAIR, OCC, EXTRAS)>::THREADS,
                                  (EXTRAS && BN <= 32 && MCD_THIN32_DGRAD_OCC1) ? 1 : FpropCfg<BN, PAIR, OCC>::MIN_CTAS)
conv_umma_fprop_kernel(const __grid_constant__ UmmaMaps maps, const __grid_constant__ FpropArgs a) {
  using Cfg = FpropCfg<BN, PAIR, OCC, fprop_epi2(BN, PAIR, OCC, EXTRAS)>;
  const __nv_bfloat16* const x_addend = EXTRAS ? a.addend : nullptr;
  const __nv_bfloat16* const x_mask = EXTRAS ? a.mask_src : nullptr;
  const __half* const x_bny = EXTRAS ? a.bn_y : nullptr;
  static_assert(!PAIR || BN == 256, "CTA pairs run the 256-channel tile only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  // HALO: [halo buffer 0][halo buffer 1][weight-tile ring]; else [A|B stage ring]
  const int nstages = HALO ? a.stages : Cfg::STAGES;
  const int ring_bytes = HALO ? 2 * a.halo_bytes + a.stages * Cfg::B_BYTES : Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + ring_bytes);
  uint64_t* empty_bar = full_bar + nstages;
  uint64_t* tmem_full_bar = empty_bar + nstages;          // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;           // [2]
  uint64_t* halo_full_bar = tmem_empty_bar + 2;           // [2]
  uint64_t* halo_empty_bar = halo_full_bar + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(halo_empty_bar + 2);
  // BatchNorm sum / sum of squares of all tiles of this (persistent) CTA are collected in shared memory and
  // flushed with one global atomic per channel at the end (instead of 2 per channel per warp per tile)
  float* sstat = reinterpret_cast<float*>(smem + ring_bytes + 256);
  const bool stat_sm = a.stats != nullptr && a.rows <= Cfg::STAT_ROWS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;    // PAIR: rank 0 = leader (issues the MMAs, owns full_bar)
  // PAIR: "tiles" are pairs of pixel tiles; CTA `rank` owns pixel tile 2 * (tile / tiles_n) + rank
  const int total_tiles = PAIR ? ((a.tiles_m + 1) / 2) * a.tiles_n : a.tiles_m * a.tiles_n;
  if (stat_sm)
    for (int i = threadIdx.x; i < 2 * a.rows; i += Cfg::THREADS) sstat[i] = 0.f;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.b);
    for (int s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], PAIR ? 8 : 4);
      mbar_init(&halo_full_bar[s], 1); mbar_init(&halo_empty_bar[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kblocks = a.ntaps * a.kchunks;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      SegIter it = PAIR ? make_pair_iter(total_tiles, kblocks) : make_seg_iter(a, total_tiles, kblocks);
      int hb = 0; uint32_t hphase = 0;
      while (it.next()) {
        int mt = PAIR ? 2 * (it.tile / a.tiles_n) + (int)rank : it.tile / a.tiles_n;
        const int n0 = (it.tile % a.tiles_n) * BN;
        const int tw_i = mt % a.tiles_w; mt /= a.tiles_w;
        const int th_i = mt % a.tiles_h; mt /= a.tiles_h;
        const int n_img = mt;                      // PAIR, odd tile count: n_img == N -> TMA zero fill
        const int th0 = th_i * a.TH, tw0 = tw_i * a.TW;
        if (HALO) {
          // per 64-channel chunk: one halo box of activations, then one weight tile per tap through the ring
          for (int kc = 0; kc < a.kchunks; ++kc) {
            mbar_wait(&halo_empty_bar[hb], hphase ^ 1);
            uint8_t* sh = smem + hb * a.halo_bytes;
            if (PAIR) {
              const uint32_t hf = mapa_shared(smem_u32(&halo_full_bar[hb]), 0);
              if (rank == 0) mbar_expect_tx(&halo_full_bar[hb], 2 * a.halo_tx);
              tma_load_4d_pair(sh, &maps.a[0], hf, kc * 64, tw0 + a.halo_ow, th0 + a.halo_oh, n_img);
            } else {
              mbar_expect_tx(&halo_full_bar[hb], a.halo_tx);
              tma_load_4d(sh, &maps.a[0], &halo_full_bar[hb], kc * 64, tw0 + a.halo_ow, th0 + a.halo_oh, n_img);
            }
            for (int t = 0; t < a.ntaps; ++t) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sb = smem + 2 * a.halo_bytes + stage * Cfg::B_BYTES;
              const int wcol = a.taps[t].wk * a.kc_pad + kc * 64;
              if (PAIR) {
                const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
                if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::B_BYTES);
                tma_load_2d_pair(sb, &maps.b, fb, wcol, n0 + (int)rank * (BN / 2));
              } else {
                mbar_expect_tx(&full_bar[stage], Cfg::B_BYTES);
                tma_load_2d(sb, &maps.b, &full_bar[stage], wcol, n0);
              }
              if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
            if (++hb == 2) { hb = 0; hphase ^= 1; }
          }
          continue;
        }
        int t = it.kb0 / a.kchunks, kc = it.kb0 % a.kchunks;
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          const Tap tap = a.taps[t];
          const CUtensorMap* amap = &maps.a[tap.map];
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (PAIR) {
            // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of the pair
            const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            tma_load_4d_pair(sa, amap, fb, kc * 64, tw0 + tap.mdw, th0 + tap.mdh, n_img);
            tma_load_2d_pair(sb, &maps.b, fb, tap.wk * a.kc_pad + kc * 64, n0 + (int)rank * (BN / 2));
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            if (++kc == a.kchunks) { kc = 0; ++t; }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (a.packed) {
            // TW columns x TH rows, column-major in the tile: one {64 elements, TH rows} window box per column
            for (int j = 0; j < a.TW; ++j)
              tma_load_3d(sa + j * a.TH * 128, amap, &full_bar[stage],
                          ((tw0 + j) * a.smul + tap.mdw) * a.cs_src, th0 + tap.mdh, n_img);
          } else {
            tma_load_4d(sa, amap, &full_bar[stage], kc * 64, tw0 + tap.mdw, th0 + tap.mdh, n_img);
          }
          tma_load_2d(sb, &maps.b, &full_bar[stage], tap.wk * a.kc_pad + kc * 64, n0);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          if (++kc == a.kchunks) { kc = 0; ++t; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = instr_desc_16(PAIR ? 256 : 128, BN, 0, 0, (uint32_t)a.fmt);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    SegIter it = PAIR ? make_pair_iter(total_tiles, kblocks) : make_seg_iter(a, total_tiles, kblocks);
    int hb = 0; uint32_t hphase = 0;
    while ((!PAIR || rank == 0) && it.next()) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);      // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS);
      if (HALO) {
        for (int kc = 0; kc < a.kchunks; ++kc) {
          mbar_wait(&halo_full_bar[hb], hphase);
          tc_fence_after();
          const uint32_t sh = smem_u32(smem + hb * a.halo_bytes);
          for (int t = 0; t < a.ntaps; ++t) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (lane == 0) {
              // A: rows (pixels) of the halo buffer starting taps[t].pad0 rows in; 8-pixel groups = tile rows,
              // halo_wb pixels apart.  B: this tap's weight tile.
              const uint64_t adesc = smem_desc_sw128(sh + (uint32_t)a.taps[t].pad0 * 128u, 0,
                                                     (uint32_t)a.halo_wb * 128u);
              const uint64_t bdesc = smem_desc_sw128(smem_u32(smem + 2 * a.halo_bytes + stage * Cfg::B_BYTES), 0, 1024);
              const bool last_t = t == a.ntaps - 1;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t accum = (kc != 0 || t != 0 || k != 0) ? 1u : 0u;
                if (PAIR) umma_bf16_pair(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, accum);
                else umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, accum);
              }
              if (PAIR) {
                umma_commit_pair(&empty_bar[stage], 3);
                if (last_t) umma_commit_pair(&halo_empty_bar[hb], 3);
                if (last_t && kc == a.kchunks - 1) umma_commit_pair(&tmem_full_bar[acc], 3);
              } else {
                umma_commit(&empty_bar[stage]);
                if (last_t) umma_commit(&halo_empty_bar[hb]);
                if (last_t && kc == a.kchunks - 1) umma_commit(&tmem_full_bar[acc]);
              }
            }
            __syncwarp();
            if (++stage == nstages) { stage = 0; phase ^= 1; }
          }
          if (++hb == 2) { hb = 0; hphase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      for (int kb = it.kb0; kb < it.kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = smem_desc_sw128(sa, 0, 1024);
          const uint64_t bdesc = smem_desc_sw128(sb, 0, 1024);
          if (PAIR) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_pair(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                             (kb != it.kb0 || k != 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage], 3);          // frees the stage in BOTH CTAs
            if (kb == it.kb1 - 1) umma_commit_pair(&tmem_full_bar[acc], 3);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)   // 4 x UMMA_K(16) per 64-channel chunk: +32 bytes each
              umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                        (kb != it.kb0 || k != 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (kb == it.kb1 - 1) umma_commit(&tmem_full_bar[acc]);
          }
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ---- epilogue: thread = one output pixel (TMEM lane), loop over 32-channel column chunks ----
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int ncover = a.planar ? a.rows : a.Cd_s;
    const int wg = (warp - 2) >> 2;                // epilogue warpgroup (0 unless EPI_WG == 2)
    int acc = Cfg::EPI_WG == 2 ? wg : 0; uint32_t acc_phase = 0;
    int tile_no = 0;
    float as1[Cfg::ACCSTAT ? Cfg::ACC_N : 1], as2[Cfg::ACCSTAT ? Cfg::ACC_N : 1];
#pragma unroll
    for (int j = 0; j < (Cfg::ACCSTAT ? Cfg::ACC_N : 1); ++j) { as1[j] = 0.f; as2[j] = 0.f; }
    SegIter it = PAIR ? make_pair_iter(total_tiles, kblocks) : make_seg_iter(a, total_tiles, kblocks);
    while (it.next()) {
      if (Cfg::EPI_WG == 2 && ((tile_no++ & 1) != wg)) continue;   // the other warpgroup's accumulator
      const bool sk_dump = it.kb0 > 0;             // stream-K: tail k-blocks of a tile another CTA owns
      const bool sk_merge = it.kb1 < kblocks;      // stream-K: this CTA owns the tile, CTA+1 did the tail
      int mt = PAIR ? 2 * (it.tile / a.tiles_n) + (int)rank : it.tile / a.tiles_n;
      const bool mt_valid = mt < a.tiles_m;
      const int n0 = (it.tile % a.tiles_n) * BN;
      const int tw_i = mt % a.tiles_w; mt /= a.tiles_w;
      const int th_i = mt % a.tiles_h; mt /= a.tiles_h;
      const int n_img = mt;
      const int th0 = th_i * a.TH, tw0 = tw_i * a.TW;
      const int ht = a.packed ? th0 + m % a.TH : th0 + m / a.TW;
      const int wt = a.packed ? tw0 + m / a.TH : tw0 + m % a.TW;
      const bool pvalid = mt_valid && ht < a.Ht && wt < a.Wt;
      const int hd = ht * a.omul + a.oh0, wd = wt * a.omul + a.ow0;
      // epilogue extras (addend / ReLU mask source / BatchNorm input) of one 32-channel chunk: all loads of a chunk
      // are issued together, and those of the first chunk before waiting for the MMAs, so that their latency
      // overlaps the wait instead of adding up load by load
      const bool extras = !a.planar && pvalid && !sk_dump && (x_addend || x_mask || x_bny);
      const int64_t obase = (((int64_t)n_img * a.Hd + hd) * a.Wd + wd) * a.Cd_s + n0;
      uint4 ea[4], em[4], ey[4];
      auto prefetch = [&](int c0) {
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          if (n0 + c0 + j8 * 8 < a.Cd_s) {
            const int64_t off = obase + c0 + j8 * 8;
            if (x_addend) ea[j8] = __ldg(reinterpret_cast<const uint4*>(x_addend + off));
            if (x_mask) em[j8] = __ldg(reinterpret_cast<const uint4*>(x_mask + off));
            if (x_bny) ey[j8] = __ldg(reinterpret_cast<const uint4*>(x_bny + off));
          }
        }
      };
      if (extras) prefetch(0);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS) + ((uint32_t)(q * 32) << 16);
      if (sk_dump) {
        // park the partial accumulator for the owner (CTA blockIdx.x - 1) and raise the flag
        float* dst = a.sk_partial + ((int64_t)blockIdx.x * (BN / 32) * 128 + m) * 32;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (n0 + c0 >= ncover) break;
          float v[32];
          tmem_ld32(tmem_d + (uint32_t)c0, v);
          tmem_ld_wait();
          float4* d4 = reinterpret_cast<float4*>(dst + (int64_t)(c0 / 32) * 128 * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        __threadfence();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");      // one named barrier per epilogue warpgroup
        if (m == 0) asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(a.sk_flags + blockIdx.x), "r"(1) : "memory");
        if (Cfg::EPI_WG == 2) acc_phase ^= 1;
        else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      const float* part = nullptr;
      if (sk_merge) {
        if (m == 0) {
          int ready = 0;
          do {
            asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(ready) : "l"(a.sk_flags + blockIdx.x + 1) : "memory");
            if (!ready) __nanosleep(64);
          } while (!ready);
        }
        __syncwarp();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        part = a.sk_partial + ((int64_t)(blockIdx.x + 1) * (BN / 32) * 128 + m) * 32;
      }
#pragma unroll(Cfg::ACCSTAT ? 2 : 1)
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= ncover) break;   // warp-uniform
        if (extras && c0 > 0) prefetch(c0);
        float v[32];
        tmem_ld32(tmem_d + (uint32_t)c0, v);
        tmem_ld_wait();
        if (part) {
          const float4* p4 = reinterpret_cast<const float4*>(part + (int64_t)(c0 / 32) * 128 * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t4 = __ldcg(p4 + j);
            v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
          }
        }
        if (a.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            int c = n0 + c0 + j;
            v[j] += (c < a.rows) ? __ldg(a.bias + c) : 0.f;
          }
        }
        float vy[32];                       // v * bn_y for the fused BatchNorm-backward sums
        if (pvalid) {
          if (a.planar) {
            float* o = reinterpret_cast<float*>(a.out);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              int c = n0 + c0 + j;
              if (c < a.rows) o[(((int64_t)n_img * a.rows + c) * a.Hd + hd) * a.Wd + wd] = v[j];
            }
          } else {
            const int64_t ooff = (((int64_t)n_img * a.Hd + hd) * a.Wd + wd) * a.Cd_s + n0 + c0;
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + ooff;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              int c = n0 + c0 + j8 * 8;
              if (c < a.Cd_s) {
                float f[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = (c + k < a.rows) ? v[j8 * 8 + k] : 0.f;
                if (x_addend) {
                  float r[8];
                  unpack8r(ea[j8], r, a.fmt);
#pragma unroll
                  for (int k = 0; k < 8; ++k) f[k] += r[k];
                }
                if (a.relu) {
#pragma unroll
                  for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
                }
                if (x_mask) {
                  float r[8];
                  unpack8(em[j8], r);
#pragma unroll
                  for (int k = 0; k < 8; ++k) f[k] = r[k] > 0.f ? f[k] : 0.f;
                }
                if (x_bny) {
                  float r[8];
                  unpack8h(ey[j8], r);
#pragma unroll
                  for (int k = 0; k < 8; ++k) { v[j8 * 8 + k] = f[k]; vy[j8 * 8 + k] = f[k] * r[k]; }
                }
                *reinterpret_cast<uint4*>(o + j8 * 8) = pack8r(f, a.fmt);
              } else if (x_bny) {
#pragma unroll
                for (int k = 0; k < 8; ++k) { v[j8 * 8 + k] = 0.f; vy[j8 * 8 + k] = 0.f; }
              }
            }
          }
        }
        if (Cfg::ACCSTAT) { if (a.stats) {
          // per-thread partial sums, reduced once after the last tile (n0 == 0: one channel tile)
          if (pvalid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (c0 + j < Cfg::ACC_N) {
                as1[c0 + j] += v[j];
                as2[c0 + j] = x_bny ? as2[c0 + j] + vy[j] : fmaf(v[j], v[j], as2[c0 + j]);
              }
            }
          }
        } } else if (a.stats && EXTRAS && MCD_SPLIT_STATS) {
          // dgrad-epilogue instantiations: the two sums one after the other
          const int c = n0 + c0 + lane;
          float* dst = stat_sm ? sstat : a.stats;
          float sA[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) sA[j] = pvalid ? v[j] : 0.f;
          warp_colsum32_single(sA, lane);
          if (c < a.rows) atomicAdd(dst + c, sA[0]);
#pragma unroll
          for (int j = 0; j < 32; ++j) sA[j] = pvalid ? (x_bny ? vy[j] : v[j] * v[j]) : 0.f;
          warp_colsum32_single(sA, lane);
          if (c < a.rows) atomicAdd(dst + a.rows + c, sA[0]);
        } else if (a.stats) {
          // column sums over the 32 pixels of this warp: butterfly transpose-reduce, lane j ends up
          // holding column j.
          float s1[32], s2[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = pvalid ? v[j] : 0.f;
            s1[j] = x;
            s2[j] = x_bny ? (pvalid ? vy[j] : 0.f) : x * x;
          }
          warp_colsum32(s1, s2, lane);
          int c = n0 + c0 + lane;
          if (c < a.rows) {
            float* dst = stat_sm ? sstat : a.stats;
            atomicAdd(dst + c, s1[0]);
            atomicAdd(dst + a.rows + c, s2[0]);
          }
        }
      }
      // all TMEM reads of this warp are complete (tcgen05.wait::ld above): hand the accumulator back
      // (PAIR: to the leader, whose MMAs write both CTAs' TMEM)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && rank != 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty_bar[acc]), 0));
        else mbar_arrive(&tmem_empty_bar[acc]);
      }
      if (Cfg::EPI_WG == 2) acc_phase ^= 1;
      else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (Cfg::ACCSTAT && a.stats) {
#pragma unroll
      for (int cb = 0; cb < Cfg::ACC_N; cb += 32) {
        warp_colsum32(as1 + cb, as2 + cb, lane);
        const int c = cb + lane;
        if (c < a.rows) {
          float* dst = stat_sm ? sstat : a.stats;
          atomicAdd(dst + c, as1[cb]);
          atomicAdd(dst + a.rows + c, as2[cb]);
        }
      }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // PAIR: the peer's smem / barriers stay valid until both are done
  if (warp == 1) {
    if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
  if (stat_sm && blockIdx.x < (PAIR ? 2 * total_tiles : total_tiles))
    for (int i = threadIdx.x; i < 2 * a.rows; i += Cfg::THREADS) atomicAdd(a.stats + i, sstat[i]);
}

// ------------------------------------------------------------------------------------------------
// wgrad
struct WgradArgs {
  int N, tiles_h, tiles_w, TH, TW;   // 64-pixel tiles over the dY grid
  int ntiles;                        // N * tiles_h * tiles_w
  int ksplit;                        // pixel-tile ranges
  int T;                             // taps
  int CoutP, CinP;                   // padded workspace dims (multiples of 128 / BN)
  int packed, cs_src, smul;          // row-packed thin-channel mode (tap = filter row, K-window loads)
  float* ws;                         // [ksplit][T][CoutP][CinP]
  Tap taps[kMaxTaps];
};

template <int BN, int OCC = 1>
struct WgradCfg {
  static constexpr int KPIX = 64;
  static constexpr int A_BYTES = 2 * KPIX * 128;             // two 64-co atoms
  static constexpr int B_BYTES = (BN / 64) * KPIX * 128;     // BN/64 ci atoms
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = OCC == 2 ? (BN == 128 ? 3 : 4) : ((BN == 256) ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 2 * BN;   // double-buffered accumulator
};

// Persistent: work item = (pixel-range split, tap, co tile of 128, ci tile of BN); CTAs walk items round-robin,
// the smem ring runs across items and the TMEM accumulator is double-buffered so that the fp32 partial-tile store
// of item i overlaps the MMAs of item i+1.  Consecutive items share the pixel range (dY / X tiles stay in L2).
template <int BN, int OCC = 1>
__global__ void __launch_bounds__(kThreads, OCC)
conv_umma_wgrad_kernel(const __grid_constant__ UmmaMaps maps, const __grid_constant__ WgradArgs a) {
  using Cfg = WgradCfg<BN, OCC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_co = a.CoutP / 128, n_ci = a.CinP / BN;
  const int per_split = a.T * n_co * n_ci;
  const int total_items = per_split * a.ksplit;
  const int per = (a.ntiles + a.ksplit - 1) / a.ksplit;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.b);
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

#define MCD_WGRAD_DECODE(item)                                              \
  const int split = (item) / per_split;                                     \
  int rem_ = (item) % per_split;                                            \
  const int ci0 = (rem_ % n_ci) * BN; rem_ /= n_ci;                         \
  const int co0 = (rem_ % n_co) * 128; rem_ /= n_co;                        \
  const int t = rem_;                                                       \
  const int tile_beg = split * per;                                         \
  const int ksteps = max(min(tile_beg + per, a.ntiles) - tile_beg, 0);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        MCD_WGRAD_DECODE(item)
        const Tap tap = a.taps[t];
        const CUtensorMap* xmap = &maps.a[tap.map];
        for (int ks = 0; ks < ksteps; ++ks) {
          int tile = tile_beg + ks;
          const int tw_i = tile % a.tiles_w; tile /= a.tiles_w;
          const int th_i = tile % a.tiles_h; tile /= a.tiles_h;
          const int n_img = tile;
          const int th0 = th_i * a.TH, tw0 = tw_i * a.TW;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (a.packed) {
            // pixel order inside the tile is column-major (TW columns of TH rows) for both operands
            for (int j = 0; j < a.TW; ++j) {
#pragma unroll
              for (int i = 0; i < 2; ++i)
                tma_load_4d(sa + i * Cfg::KPIX * 128 + j * a.TH * 128, &maps.b, &full_bar[stage],
                            co0 + i * 64, tw0 + j, th0, n_img);
              tma_load_3d(sb + j * a.TH * 128, xmap, &full_bar[stage],
                          ((tw0 + j) * a.smul + tap.mdw) * a.cs_src, th0 + tap.mdh, n_img);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i)
              tma_load_4d(sa + i * Cfg::KPIX * 128, &maps.b, &full_bar[stage], co0 + i * 64, tw0, th0, n_img);
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_4d(sb + i * Cfg::KPIX * 128, xmap, &full_bar[stage], ci0 + i * 64, tw0 + tap.mdw,
                          th0 + tap.mdh, n_img);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = instr_desc_bf16(128, BN, 1, 1);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      MCD_WGRAD_DECODE(item)
      (void)ci0; (void)co0; (void)t;
      if (ksteps == 0) continue;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          // MN-major, 128B swizzle: LBO = bytes between 64-channel atoms, SBO = 8 pixel rows
          const uint64_t adesc = smem_desc_sw128(sa, Cfg::KPIX * 128, 1024);
          const uint64_t bdesc = smem_desc_sw128(sb, Cfg::KPIX * 128, 1024);
#pragma unroll
          for (int k = 0; k < Cfg::KPIX / 16; ++k)   // 16 pixel rows = 2048 bytes per UMMA_K step
            umma_bf16(tmem_d, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (ks | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (ks == ksteps - 1) umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      MCD_WGRAD_DECODE(item)
      const int co = co0 + q * 32 + lane;
      float* o = a.ws + (((int64_t)split * a.T + t) * a.CoutP + co) * a.CinP + ci0;
      if (ksteps > 0) {
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
      }
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        if (ksteps > 0) {
          tmem_ld32(tmem_d + (uint32_t)c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(o + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      if (ksteps > 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }
#undef MCD_WGRAD_DECODE
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair wgrad for the wide layers (Cin > 128, Cout >= 256): one cta_group::2 MMA tile = 256 co x 256 ci.
// Each CTA stages ITS 128 co of dY (A) and ITS 128 ci of X (B half): 32 KB of TMA writes and operand reads per
// k-step and SM instead of 48 KB for the single-CTA 128 x 256 tile, which is shared-memory-bandwidth bound.
struct WgradPairCfg {
  static constexpr int KPIX = 64;
  static constexpr int A_BYTES = 2 * KPIX * 128;
  static constexpr int B_BYTES = 2 * KPIX * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 6;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 512;      // two 256-column accumulators
};

__global__ void __launch_bounds__(kThreads, 1)
conv_umma_wgrad_pair_kernel(const __grid_constant__ UmmaMaps maps, const __grid_constant__ WgradArgs a) {
  using Cfg = WgradPairCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_co = a.CoutP / 256, n_ci = a.CinP / 256;
  const int per_split = a.T * n_co * n_ci;
  const int total_items = per_split * a.ksplit;
  const int per = (a.ntiles + a.ksplit - 1) / a.ksplit;
  const int first_item = blockIdx.x / 2, item_stride = gridDim.x / 2;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.b);
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

#define MCD_WGRADP_DECODE(item)                                             \
  const int split = (item) / per_split;                                     \
  int rem_ = (item) % per_split;                                            \
  const int ci0 = (rem_ % n_ci) * 256; rem_ /= n_ci;                        \
  const int co0 = (rem_ % n_co) * 256; rem_ /= n_co;                        \
  const int t = rem_;                                                       \
  const int tile_beg = split * per;                                         \
  const int ksteps = max(min(tile_beg + per, a.ntiles) - tile_beg, 0);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = first_item; item < total_items; item += item_stride) {
        MCD_WGRADP_DECODE(item)
        (void)split;
        const Tap tap = a.taps[t];
        const CUtensorMap* xmap = &maps.a[tap.map];
        for (int ks = 0; ks < ksteps; ++ks) {
          int tile = tile_beg + ks;
          const int tw_i = tile % a.tiles_w; tile /= a.tiles_w;
          const int th_i = tile % a.tiles_h; tile /= a.tiles_h;
          const int n_img = tile;
          const int th0 = th_i * a.TH, tw0 = tw_i * a.TW;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
#pragma unroll
          for (int i = 0; i < 2; ++i)
            tma_load_4d_pair(sa + i * Cfg::KPIX * 128, &maps.b, fb, co0 + (int)rank * 128 + i * 64, tw0, th0, n_img);
#pragma unroll
          for (int i = 0; i < 2; ++i)
            tma_load_4d_pair(sb + i * Cfg::KPIX * 128, xmap, fb, ci0 + (int)rank * 128 + i * 64, tw0 + tap.mdw,
                             th0 + tap.mdh, n_img);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = instr_desc_bf16(256, 256, 1, 1);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = first_item; rank == 0 && item < total_items; item += item_stride) {
      MCD_WGRADP_DECODE(item)
      (void)ci0; (void)co0; (void)t; (void)split;
      if (ksteps == 0) continue;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256);
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = smem_desc_sw128(sa, Cfg::KPIX * 128, 1024);
          const uint64_t bdesc = smem_desc_sw128(sb, Cfg::KPIX * 128, 1024);
#pragma unroll
          for (int k = 0; k < Cfg::KPIX / 16; ++k)
            umma_bf16_pair(tmem_d, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (ks | k) != 0);
          umma_commit_pair(&empty_bar[stage], 3);
          if (ks == ksteps - 1) umma_commit_pair(&tmem_full_bar[acc], 3);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = first_item; item < total_items; item += item_stride) {
      MCD_WGRADP_DECODE(item)
      const int co = co0 + (int)rank * 128 + q * 32 + lane;
      float* o = a.ws + (((int64_t)split * a.T + t) * a.CoutP + co) * a.CinP + ci0;
      if (ksteps > 0) {
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
      }
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) {
        float v[32];
        if (ksteps > 0) {
          tmem_ld32(tmem_d + (uint32_t)c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(o + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      if (ksteps > 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank != 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty_bar[acc]), 0));
          else mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }
#undef MCD_WGRADP_DECODE
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// wgrad of the row-packed thin-channel layers (layer0/1/2: Cout <= 64, R <= 7 filter rows).
// One CTA owns a range of 64-pixel tiles and ALL filter rows: per tile it loads the dY tile once (A, one
// 64-channel atom, MN-major) and the R shifted K-windows of X (B_r), and issues R accumulating MMAs
// (M=64 co, N=64 window elements) into R TMEM accumulators.  8 TMA issues + 4R MMAs per tile instead of
// 3 TMA issues per (tile, row) in the generic kernel; the producer lane is no longer the bottleneck.
struct WgradRowsArgs {
  int tiles_h, tiles_w, TH, TW, ntiles, nsplit;
  int R, cs_src, smul, stages, tmem_cols;
  float* ws;                         // [nsplit][R][64 co][64 k]
  Tap taps[8];
};

__global__ void __launch_bounds__(kThreads, 1)
conv_umma_wgrad_rows_kernel(const __grid_constant__ UmmaMaps maps, const __grid_constant__ WgradRowsArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  const int stage_bytes = (1 + a.R) * 8192;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + a.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + a.stages;
  uint64_t* tmem_full_bar = empty_bar + a.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int per = (a.ntiles + a.nsplit - 1) / a.nsplit;
  const int tile_beg = split * per;
  const int ksteps = max(min(tile_beg + per, a.ntiles) - tile_beg, 0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.b);
    for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int tw_i = tile_beg % a.tiles_w, th_i = (tile_beg / a.tiles_w) % a.tiles_h,
          n_img = tile_beg / (a.tiles_w * a.tiles_h);
      for (int ks = 0; ks < ksteps; ++ks) {
        const int th0 = th_i * a.TH, tw0 = tw_i * a.TW;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * stage_bytes;
        mbar_expect_tx(&full_bar[stage], stage_bytes);
        for (int j = 0; j < a.TW; ++j) {
          tma_load_4d(sa + j * a.TH * 128, &maps.b, &full_bar[stage], 0, tw0 + j, th0, n_img);
          for (int r = 0; r < a.R; ++r)
            tma_load_3d(sa + (1 + r) * 8192 + j * a.TH * 128, &maps.a[a.taps[r].map], &full_bar[stage],
                        ((tw0 + j) * a.smul + a.taps[r].mdw) * a.cs_src, th0 + a.taps[r].mdh, n_img);
        }
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
        if (++tw_i == a.tiles_w) { tw_i = 0; if (++th_i == a.tiles_h) { th_i = 0; ++n_img; } }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = instr_desc_bf16(64, 64, 1, 1);
    int stage = 0; uint32_t phase = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint64_t adesc = smem_desc_sw128(sa, 8192, 1024);
        for (int r = 0; r < a.R; ++r) {
          const uint64_t bdesc = smem_desc_sw128(sa + (1 + r) * 8192, 8192, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 64 pixels = 4 x UMMA_K(16) rows of 128 bytes
            umma_bf16(tmem_base + (uint32_t)(r * 64), adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128),
                      idesc, (ks | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        if (ks == ksteps - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == a.stages) { stage = 0; phase ^= 1; }
    }
  } else {
    // M = 64 accumulators live in lanes 0..15 of each 32-lane TMEM sub-partition: row = 16*q + lane
    const int q = warp & 3;
    const int co = q * 16 + lane;
    if (ksteps > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    for (int r = 0; r < a.R; ++r) {
      float* o = a.ws + (((int64_t)split * a.R + r) * 64 + co) * 64;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        float v[32];
        if (ksteps > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(r * 64 + c0), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (lane < 16) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(o + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

// dw[co][ci][r][s] = sum_split ws[split][t][co][ci]
// block = one co x 64 ci: coalesced reads along ci, transpose through smem, contiguous 64*T-float write.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int ksplit, int T, int CoutP,
                    int CinP, int Cout, int Cin, int accumulate) {
  extern __shared__ float tile[];  // [64][T]
  const int co = blockIdx.x, ci0 = blockIdx.y * 64;
  const int nci = min(64, Cin - ci0);
  for (int idx = threadIdx.x; idx < 64 * T; idx += blockDim.x) {
    const int t = idx >> 6, cil = idx & 63;
    float acc = 0.f;
    if (cil < nci) {
      const float* p = ws + ((int64_t)t * CoutP + co) * CinP + ci0 + cil;
      const int64_t kstride = (int64_t)T * CoutP * CinP;
      for (int k = 0; k < ksplit; ++k) acc += p[k * kstride];
    }
    tile[cil * T + t] = acc;
  }
  __syncthreads();
  float* o = dw + ((int64_t)co * Cin + ci0) * T;
  for (int j = threadIdx.x; j < nci * T; j += blockDim.x) o[j] = accumulate ? o[j] + tile[j] : tile[j];
}

// packed: ws[split][r][co][s*Cs + c]  ->  dw[co][c][r][s]
// 8 lanes share one output element (the splits k = lane, lane + 8, ...) and combine with shuffles: the stem layers
// have few weights (4.7 K) but up to 148 splits, one thread per element would serialise 148 dependent loads.
__global__ void wgrad_reduce_packed_kernel(const float* __restrict__ ws, float* __restrict__ dw, int ksplit,
                                           int R, int S, int Cs, int CoutP, int Cout, int Cin, int accumulate) {
  const int64_t total = (int64_t)Cout * Cin * R * S;
  const int sub = threadIdx.x & 7;
  for (int64_t i0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3; i0 < ((total + 31) & ~int64_t(31));
       i0 += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    const int64_t i = i0 < total ? i0 : total - 1;
    int sx = (int)(i % S);
    int r = (int)((i / S) % R);
    int c = (int)((i / ((int64_t)S * R)) % Cin);
    int co = (int)(i / ((int64_t)S * R * Cin));
    float acc = 0.f;
    for (int k = sub; k < ksplit; k += 8) acc += ws[(((int64_t)k * R + r) * CoutP + co) * 64 + sx * Cs + c];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (sub == 0 && i0 < total) dw[i] = accumulate ? dw[i] + acc : acc;
  }
}

// ------------------------------------------------------------------------------------------------
// host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 4-D map over an NHWC bf16 tensor viewed through a (ph,pw) parity sub-grid with step `st`.
static inline CUtensorMapDataType tmap_dtype(int fmt) {
  return fmt == kF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
}

static int encode_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int Cs,
                          int st, int ph, int pw, int box_w, int box_h, int fmt = kBF16) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return MCD_E_CUDA; }
  int Wsub = (W - pw + st - 1) / st, Hsub = (H - ph + st - 1) / st;
  if (Wsub <= 0 || Hsub <= 0) { Wsub = max(Wsub, 1); Hsub = max(Hsub, 1); }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wsub, (cuuint64_t)Hsub, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)st * Cs * 2, (cuuint64_t)st * W * Cs * 2,
                           (cuuint64_t)H * W * Cs * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const char* p = reinterpret_cast<const char*>(base) + ((int64_t)ph * W + pw) * Cs * 2;
  CUresult r = enc(m, tmap_dtype(fmt), 4, const_cast<char*>(p), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(act N=%d H=%d W=%d C=%d Cs=%d st=%d box=%dx%d) failed: %d", N, H,
              W, C, Cs, st, box_w, box_h, (int)r);
    return MCD_E_CUDA;
  }
  return MCD_OK;
}

// 3-D map {W*Cs elements, rows of parity ph (step st), N} for the row-packed window loads
static int encode_rows_map(CUtensorMap* m, const void* base, int N, int H, int W, int Cs, int st, int ph,
                           int box_h, int fmt = kBF16) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return MCD_E_CUDA; }
  int Hsub = (H - ph + st - 1) / st;
  if (Hsub <= 0) Hsub = 1;
  cuuint64_t dims[3] = {(cuuint64_t)W * Cs, (cuuint64_t)Hsub, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)st * W * Cs * 2, (cuuint64_t)H * W * Cs * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_h, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const char* p = reinterpret_cast<const char*>(base) + (int64_t)ph * W * Cs * 2;
  CUresult r = enc(m, tmap_dtype(fmt), 3, const_cast<char*>(p), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(rows N=%d H=%d W=%d Cs=%d st=%d box_h=%d) failed: %d", N, H, W, Cs, st,
              box_h, (int)r);
    return MCD_E_CUDA;
  }
  return MCD_OK;
}

// packed tiles: TW columns of TH rows, TH >= 8 so that every column box starts on a 1024-byte swizzle atom.
// Every column costs one TMA issue per K chunk and the thin layers are TMA-issue-bound, so the tallest tile whose
// padded area is within 10 % of the best one wins.
static void pick_tile_packed(int Ht, int Wt, int npix, int* TH, int* TW) {
  int64_t best = -1;
  for (int th = npix; th >= 8; th >>= 1) {
    int tw = npix / th;
    int64_t area = (int64_t)((Ht + th - 1) / th) * th * ((Wt + tw - 1) / tw) * tw;
    if (best < 0 || area < best) best = area;
  }
  for (int th = npix; th >= 8; th >>= 1) {
    int tw = npix / th;
    int64_t area = (int64_t)((Ht + th - 1) / th) * th * ((Wt + tw - 1) / tw) * tw;
    if (area * 10 <= best * 11) { *TH = th; *TW = tw; return; }
  }
}

static int encode_weight_map(CUtensorMap* m, const void* base, int rows, int64_t kdim, int box_rows, int fmt = kBF16) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return MCD_E_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kdim * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, tmap_dtype(fmt), 2, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(weights rows=%d k=%lld) failed: %d", rows, (long long)kdim, (int)r);
    return MCD_E_CUDA;
  }
  return MCD_OK;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// spatial tile of `npix` pixels minimising padded area; ties -> wider tile
static void pick_tile(int Ht, int Wt, int npix, int* TH, int* TW) {
  int64_t best = -1;
  for (int tw = npix; tw >= 1; tw >>= 1) {
    int th = npix / tw;
    if (tw > 256 || th > 256) continue;
    int64_t area = (int64_t)((Ht + th - 1) / th) * th * ((Wt + tw - 1) / tw) * tw;
    if (best < 0 || area < best) { best = area; *TH = th; *TW = tw; }
  }
}

bool umma_problem_supported(const TapProblem& p) {
  if (p.smul != 1 && p.smul != 2) return false;
  if (p.Cs_src % 8 != 0) return false;
  if (p.ntaps > kMaxTaps) return false;
  return true;
}

static inline bool has_extras(const FpropArgs& a) { return a.addend || a.mask_src || a.bn_y; }

template <int BN, int OCC, bool EXTRAS>
static int launch_fprop_bn_x(const UmmaMaps& maps, const FpropArgs& a, dim3 grid, cudaStream_t st) {
  using Cfg = FpropCfg<BN, false, OCC, fprop_epi2(BN, false, OCC, EXTRAS)>;
  int arc = ensure_dyn_smem<conv_umma_fprop_kernel<BN, false, OCC, false, EXTRAS>>(Cfg::SMEM_BYTES, "conv_umma_fprop");
  if (arc != MCD_OK) return arc;
  conv_umma_fprop_kernel<BN, false, OCC, false, EXTRAS><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(maps, a);
  return check_launch("conv_umma_fprop");
}

template <int BN, int OCC = 1>
static int launch_fprop_bn(const UmmaMaps& maps, const FpropArgs& a, dim3 grid, cudaStream_t st) {
  return has_extras(a) ? launch_fprop_bn_x<BN, OCC, true>(maps, a, grid, st)
                       : launch_fprop_bn_x<BN, OCC, false>(maps, a, grid, st);
}

// MCD_THIN_OCC2=0 keeps one persistent CTA per SM for the 64- / 128-channel tiles (A/B measurements)
static bool thin_occ2() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MCD_THIN_OCC2"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// CTA-pair variant (clusters of 2): grid = 2 * number of persistent pairs
template <bool EXTRAS>
static int launch_fprop_pair_x(const UmmaMaps& maps, const FpropArgs& a, int pairs, cudaStream_t st) {
  using Cfg = FpropCfg<256, true>;
  int arc = ensure_dyn_smem<conv_umma_fprop_kernel<256, true, 1, false, EXTRAS>>(Cfg::SMEM_BYTES, "conv_umma_fprop_pair");
  if (arc != MCD_OK) return arc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_umma_fprop_kernel<256, true, 1, false, EXTRAS>, maps, a);
  if (e != cudaSuccess) { set_error("conv_umma_fprop (CTA pairs): %s", cudaGetErrorString(e)); return MCD_E_CUDA; }
  return check_launch("conv_umma_fprop_pair");
}

static int launch_fprop_pair(const UmmaMaps& maps, const FpropArgs& a, int pairs, cudaStream_t st) {
  return has_extras(a) ? launch_fprop_pair_x<true>(maps, a, pairs, st) : launch_fprop_pair_x<false>(maps, a, pairs, st);
}

// ---- halo-tile launches (HALO = true instantiations; dynamic shared memory sized per problem) -------------------
constexpr int kHaloFixedSmem = 1024 /*align slack*/ + 256 /*barriers*/ + 2 * 1024 * 4 /*BatchNorm staging*/;

template <int BN, bool PAIR, int OCC, bool EXTRAS>
static int launch_fprop_halo_x(const UmmaMaps& maps, const FpropArgs& a, int grid, cudaStream_t st) {
  using Cfg = FpropCfg<BN, PAIR, OCC, fprop_epi2(BN, PAIR, OCC, EXTRAS)>;
  const int smem_bytes = 2 * a.halo_bytes + a.stages * Cfg::B_BYTES + kHaloFixedSmem;
  int arc = ensure_dyn_smem<conv_umma_fprop_kernel<BN, PAIR, OCC, true, EXTRAS>>(smem_bytes, "conv_umma_fprop_halo");
  if (arc != MCD_OK) return arc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = PAIR ? 2 : 1; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_umma_fprop_kernel<BN, PAIR, OCC, true, EXTRAS>, maps, a);
  if (e != cudaSuccess) { set_error("conv_umma_fprop (halo): %s", cudaGetErrorString(e)); return MCD_E_CUDA; }
  return check_launch("conv_umma_fprop_halo");
}

template <int BN, bool PAIR, int OCC>
static int launch_fprop_halo(const UmmaMaps& maps, const FpropArgs& a, int grid, cudaStream_t st) {
  return has_extras(a) ? launch_fprop_halo_x<BN, PAIR, OCC, true>(maps, a, grid, st)
                       : launch_fprop_halo_x<BN, PAIR, OCC, false>(maps, a, grid, st);
}

// MCD_HALO=0 selects the one-box-per-tap (im2col-style) staging for every layer (A/B measurements)
static bool halo_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MCD_HALO"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// Halo-tile plan of a stride-1, multi-tap problem with >= 64 produced channels: 8 x 16 output tiles, box =
// tile + the extent of the tap offsets.  Returns false when the problem keeps the per-tap staging.
// MCD_THIN_OCC2_DGRAD=1: two CTAs per SM also for the dgrad instantiations (fused-epilogue inputs) of the 64 / 128
// channel tiles.  Default 0: they need ~250 registers, at 168 they spill ~400 B per thread in the epilogue loop.
static bool thin_occ2_dgrad() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MCD_THIN_OCC2_DGRAD"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

static bool halo_plan(const TapProblem& p, int BN, bool pair, FpropArgs* a, int* Hb, bool occ2 = true) {
  if (!halo_enabled() || p.packed || p.smul != 1 || p.ntaps < 2 || BN < 64 || (BN == 256 && !pair)) return false;
  int lo_h = 1 << 20, hi_h = -(1 << 20), lo_w = 1 << 20, hi_w = -(1 << 20);
  for (int t = 0; t < p.ntaps; ++t) {
    if (p.taps[t].map != 0) return false;
    lo_h = min(lo_h, (int)p.taps[t].mdh); hi_h = max(hi_h, (int)p.taps[t].mdh);
    lo_w = min(lo_w, (int)p.taps[t].mdw); hi_w = max(hi_w, (int)p.taps[t].mdw);
  }
  const int Wb = 8 + (hi_w - lo_w), hb = 16 + (hi_h - lo_h);
  if (Wb > 256 || hb > 256) return false;
  const int halo_bytes = round_up(Wb * hb * 128, 1024);
  const int b_bytes = (pair ? BN / 2 : BN) * 128;
  const int limit = ((pair || !occ2) ? 224 : 110) * 1024;   // one CTA per SM, or two (64 / 128 tiles, occ2)
  const int stages = min(8, (limit - kHaloFixedSmem - 2 * halo_bytes) / b_bytes);
  if (stages < 3) return false;
  a->halo_bytes = halo_bytes; a->halo_tx = Wb * hb * 128; a->halo_wb = Wb; a->halo_oh = lo_h; a->halo_ow = lo_w;
  a->stages = stages;
  a->TH = 16; a->TW = 8;
  a->tiles_h = (p.Ht + 15) / 16; a->tiles_w = (p.Wt + 7) / 8;
  a->tiles_m = p.N * a->tiles_h * a->tiles_w;
  for (int t = 0; t < p.ntaps; ++t)
    a->taps[t].pad0 = (int16_t)((p.taps[t].mdh - lo_h) * Wb + (p.taps[t].mdw - lo_w));
  *Hb = hb;
  return true;
}

static bool g_pair_enabled = true;   // MCD_CTA_PAIRS=0 selects the single-CTA 128x256 tiles (A/B measurements)
static bool pair_enabled() {
  static bool init = false;
  if (!init) {
    const char* e = getenv("MCD_CTA_PAIRS");
    if (e && e[0] == '0') g_pair_enabled = false;
    init = true;
  }
  return g_pair_enabled;
}

static void fprop_tiling(const TapProblem& p, int planar, int* TH, int* TW, int* tiles_m, int* tiles_n, int* BN) {
  if (p.packed) pick_tile_packed(p.Ht, p.Wt, 128, TH, TW);
  else pick_tile(p.Ht, p.Wt, 128, TH, TW);
  const int ncover = planar ? p.rows : p.Cd_s;
  *BN = ncover > 128 ? 256 : (ncover > 64 ? 128 : (ncover > 32 ? 64 : (ncover > 16 ? 32 : 16)));
  *tiles_m = p.N * ((p.Ht + *TH - 1) / *TH) * ((p.Wt + *TW - 1) / *TW);
  *tiles_n = (ncover + *BN - 1) / *BN;
}

// stream-K pays when the data-parallel schedule would leave the last wave of tiles mostly empty
static bool streamk_plan(int total_tiles, int kblocks, int BN, int packed, int* G, int* units) {
  const int slots = sm_count() * (BN <= 32 ? 2 : 1);
  if (packed || total_tiles <= slots || kblocks < 2) return false;
  const int waves = (total_tiles + slots - 1) / slots;
  if ((double)total_tiles / ((double)waves * slots) > 0.92) return false;
  const int64_t U = (int64_t)total_tiles * kblocks;
  const int64_t W = (U + slots - 1) / slots;
  if (W < kblocks) return false;
  *G = (int)((U + W - 1) / W);
  *units = (int)W;
  return true;
}

// src: activations for this problem; for strided fprop the parity maps are built over (Hs, Ws).
int launch_umma_problem(const void* src, const void* w, const float* bias, void* out, int planar,
                        float* stats, const EpiExtra& ex, const TapProblem& p, int fmt, cudaStream_t st) {
  if (p.ntaps == 0) return MCD_OK;
  if (!umma_problem_supported(p)) { set_error("umma fprop: unsupported problem"); return MCD_E_INVALID; }
  FpropArgs a;
  memset(&a, 0, sizeof(a));
  a.N = p.N; a.Ht = p.Ht; a.Wt = p.Wt;
  a.packed = p.packed; a.cs_src = p.Cs_src; a.smul = p.smul;
  int bn_unused;
  fprop_tiling(p, planar, &a.TH, &a.TW, &a.tiles_m, &a.tiles_n, &bn_unused);
  a.tiles_h = (p.Ht + a.TH - 1) / a.TH; a.tiles_w = (p.Wt + a.TW - 1) / a.TW;
  a.kchunks = (p.Kc + 63) / 64; a.ntaps = p.ntaps; a.kc_pad = p.kc_pad;
  a.rows = p.rows; a.omul = p.omul; a.oh0 = p.oh0; a.ow0 = p.ow0; a.Hd = p.Hd; a.Wd = p.Wd;
  a.Cd_s = p.Cd_s; a.planar = planar; a.fmt = fmt; a.bias = bias; a.stats = stats; a.out = out;
  a.relu = planar ? 0 : ex.relu;
  a.addend = planar ? nullptr : reinterpret_cast<const __nv_bfloat16*>(ex.addend);
  a.mask_src = planar ? nullptr : reinterpret_cast<const __nv_bfloat16*>(ex.mask_src);
  a.bn_y = planar ? nullptr : reinterpret_cast<const __half*>(ex.bn_y);
  for (int t = 0; t < p.ntaps; ++t) a.taps[t] = p.taps[t];

  UmmaMaps maps;
  memset(&maps, 0, sizeof(maps));
  int stp = p.smul;
  bool used[4] = {false, false, false, false};
  for (int t = 0; t < p.ntaps; ++t) used[p.taps[t].map] = true;
  if (p.packed) {
    for (int ph = 0; ph < stp; ++ph) {
      if (!used[ph]) continue;
      int rc = encode_rows_map(&maps.a[ph], src, p.N, p.Hs, p.Ws, p.Cs_src, stp, ph, a.TH, fmt);
      if (rc != MCD_OK) return rc;
    }
  } else {
    for (int ph = 0; ph < stp; ++ph)
      for (int pw = 0; pw < stp; ++pw) {
        int id = ph * stp + pw;
        if (!used[id]) continue;
        int rc = encode_act_map(&maps.a[id], src, p.N, p.Hs, p.Ws, p.Kc, p.Cs_src, stp, ph, pw, a.TW, a.TH, fmt);
        if (rc != MCD_OK) return rc;
      }
  }
  if (!used[0]) maps.a[0] = maps.a[p.taps[0].map];  // keep the prefetch target valid

  int BN, G;
  fprop_tiling(p, planar, &a.TH, &a.TW, &a.tiles_m, &a.tiles_n, &BN);
  const bool want_sk = ex.sk_partial && ex.sk_flags;
  const bool pair = BN == 256 && !p.packed && a.tiles_m >= 2 && !want_sk && pair_enabled();
  int halo_hb = 0;
  // dgrad instantiations (fused-epilogue inputs) of the 64 / 128 tiles keep one CTA per SM: register budget
  const bool occ2_ok = thin_occ2() && (!has_extras(a) || thin_occ2_dgrad());
  if (!want_sk && halo_plan(p, BN, pair, &a, &halo_hb, occ2_ok)) {
    int rc = encode_act_map(&maps.a[0], src, p.N, p.Hs, p.Ws, p.Kc, p.Cs_src, 1, 0, 0, a.halo_wb, halo_hb, fmt);
    if (rc != MCD_OK) return rc;
    rc = encode_weight_map(&maps.b, w, p.rows, (int64_t)p.T_total * p.kc_pad, pair ? BN / 2 : BN, fmt);
    if (rc != MCD_OK) return rc;
    if (pair) {
      const int pair_tiles = ((a.tiles_m + 1) / 2) * a.tiles_n;
      return launch_fprop_halo<256, true, 1>(maps, a, 2 * min(pair_tiles, sm_count() / 2), st);
    }
    const int grid = min(a.tiles_m * a.tiles_n, (occ2_ok ? 2 : 1) * sm_count());
    if (occ2_ok)
      return BN == 128 ? launch_fprop_halo<128, false, 2>(maps, a, grid, st)
                       : launch_fprop_halo<64, false, 2>(maps, a, grid, st);
    return BN == 128 ? launch_fprop_halo<128, false, 1>(maps, a, grid, st)
                     : launch_fprop_halo<64, false, 1>(maps, a, grid, st);
  }
  int rc = encode_weight_map(&maps.b, w, p.rows, (int64_t)p.T_total * p.kc_pad, pair ? BN / 2 : BN, fmt);
  if (rc != MCD_OK) return rc;
  if (pair) {
    const int pair_tiles = ((a.tiles_m + 1) / 2) * a.tiles_n;
    return launch_fprop_pair(maps, a, min(pair_tiles, sm_count() / 2), st);
  }
  const bool occ2 = (BN == 64 || BN == 128) && occ2_ok && !want_sk;
  const bool thin32_two = BN <= 32 && !(has_extras(a) && MCD_THIN32_DGRAD_OCC1);
  const int slots = sm_count() * ((thin32_two || occ2) ? 2 : 1);  // persistent CTAs per SM (see __launch_bounds__)
  G = min(a.tiles_m * a.tiles_n, slots);
  if (ex.sk_partial && ex.sk_flags && BN > MCD_ACCSTAT) {   // per-thread statistics need whole tiles per CTA
    int units = 0, g2 = 0;
    if (streamk_plan(a.tiles_m * a.tiles_n, a.ntaps * a.kchunks, BN, p.packed, &g2, &units)) {
      G = g2; a.sk_units = units;
      a.sk_partial = reinterpret_cast<float*>(ex.sk_partial); a.sk_flags = ex.sk_flags;
    }
  }
  dim3 grid((unsigned)G);
  if (BN == 256) return launch_fprop_bn<256>(maps, a, grid, st);
  if (BN == 128) return occ2 ? launch_fprop_bn<128, 2>(maps, a, grid, st) : launch_fprop_bn<128>(maps, a, grid, st);
  if (BN == 64) return occ2 ? launch_fprop_bn<64, 2>(maps, a, grid, st) : launch_fprop_bn<64>(maps, a, grid, st);
  if (BN == 32) return launch_fprop_bn<32>(maps, a, grid, st);
  return launch_fprop_bn<16>(maps, a, grid, st);
}

bool wgrad_toeplitz_ok(const mcd_conv_geom& g);
size_t wgrad_toeplitz_workspace(const mcd_conv_geom& g);
int wgrad_toeplitz_launch(const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes, const mcd_conv_geom& g,
                          int accumulate, cudaStream_t st);
static int wgrad_bn(const mcd_conv_geom& g);
static bool wgrad_rows_ok(const mcd_conv_geom& g);
static void wgrad_shape(const mcd_conv_geom& g, int* BN, int* CoutP, int* CinP, int* TH, int* TW,
                        int* ntiles, int* ksplit);
static bool wgrad_pair(const mcd_conv_geom& g);
static bool halo_plan(const TapProblem& p, int BN, bool pair, FpropArgs* a, int* Hb, bool occ2);
// which kernel launch_umma_problem() picks for a problem: tile width BN, *pair = CTA-pair (cta_group::2) variant
int umma_problem_tile(const TapProblem& p, int planar, int* pair, int* halo) {
  int TH, TW, tiles_m, tiles_n, BN, hb;
  fprop_tiling(p, planar, &TH, &TW, &tiles_m, &tiles_n, &BN);
  *pair = BN == 256 && !p.packed && tiles_m >= 2 && pair_enabled();
  FpropArgs a;
  *halo = halo_plan(p, BN, *pair != 0, &a, &hb, true) ? 1 : 0;
  return BN;
}

// which kernel umma_wgrad() picks: tile width BN, *rows = all-filter-rows thin-channel kernel
int umma_wgrad_tile(const mcd_conv_geom& g, int* rows) {
  if (wgrad_toeplitz_ok(g)) { *rows = 2; return g.Cout_s; }
  if (!wgrad_rows_ok(g) && wgrad_pair(g)) { *rows = 3; return 256; }
  *rows = wgrad_rows_ok(g);
  if (*rows) return 64;
  return packed_fprop_ok(g) ? 64 : wgrad_bn(g);
}

// layout of the split partial sums the generic wgrad kernel leaves in its workspace: fp32 [ksplit][T][CoutP][CinP];
// false for the layers that use other kernels (stem)
bool umma_wgrad_partial_layout(const mcd_conv_geom& g, int* out4) {
  if (wgrad_toeplitz_ok(g) || (g.stride != 1 && g.stride != 2) || g.R * g.S > kMaxTaps || wgrad_rows_ok(g) || packed_fprop_ok(g)) return false;
  int BN, CoutP, CinP, TH, TW, ntiles, ksplit;
  wgrad_shape(g, &BN, &CoutP, &CinP, &TH, &TW, &ntiles, &ksplit);
  out4[0] = ksplit; out4[1] = g.R * g.S; out4[2] = CoutP; out4[3] = CinP;
  return true;
}

// stream-K workspace of one problem: bytes of fp32 partial tiles and number of int flags (0 / 0: not used)
size_t umma_streamk_workspace(const TapProblem& p, int planar, int* n_flags) {
  *n_flags = 0;
  if (p.ntaps == 0 || !umma_problem_supported(p)) return 0;
  int TH, TW, tiles_m, tiles_n, BN, G = 0, units = 0;
  fprop_tiling(p, planar, &TH, &TW, &tiles_m, &tiles_n, &BN);
  if (!streamk_plan(tiles_m * tiles_n, p.ntaps * ((p.Kc + 63) / 64), BN, p.packed, &G, &units)) return 0;
  *n_flags = G;
  return (size_t)G * BN * 128 * sizeof(float);
}

static int wgrad_bn(const mcd_conv_geom& g) { return g.Cin > 128 ? 256 : (g.Cin > 64 ? 128 : 64); }

static bool pair_enabled();
// CTA-pair tiles (256 co x 256 ci) for the wide layers; MCD_WGRAD_PAIRS=0 keeps the single-CTA 128 x 256 tiles
static bool wgrad_pair(const mcd_conv_geom& g) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MCD_WGRAD_PAIRS"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1 && pair_enabled() && !packed_fprop_ok(g) && g.Cin > 128 && g.Cout >= 256;
}

static void wgrad_shape(const mcd_conv_geom& g, int* BN, int* CoutP, int* CinP, int* TH, int* TW,
                        int* ntiles, int* ksplit) {
  const bool packed = packed_fprop_ok(g);
  *BN = packed ? 64 : wgrad_bn(g);
  if (wgrad_pair(g)) {
    *CoutP = round_up(g.Cout, 256);
    *CinP = round_up(g.Cin, 256);
    pick_tile(g.Ho, g.Wo, 64, TH, TW);
    *ntiles = g.N * ((g.Ho + *TH - 1) / *TH) * ((g.Wo + *TW - 1) / *TW);
    const int base = (*CoutP / 256) * (*CinP / 256) * g.R * g.S;
    int ks = (sm_count() / 2) / base;
    *ksplit = max(1, min(min(ks, *ntiles), 64));
    return;
  }
  *CoutP = round_up(g.Cout, 128);
  *CinP = packed ? 64 : round_up(g.Cin, *BN);
  if (packed) pick_tile_packed(g.Ho, g.Wo, 64, TH, TW);
  else pick_tile(g.Ho, g.Wo, 64, TH, TW);
  *ntiles = g.N * ((g.Ho + *TH - 1) / *TH) * ((g.Wo + *TW - 1) / *TW);
  int base = (*CoutP / 128) * (*CinP / *BN) * g.R * (packed ? 1 : g.S);
  const int slots = sm_count() * ((*BN <= 128 && thin_occ2()) ? 2 : 1);
  int ks = slots / base;                    // one work item per persistent CTA; fewer splits = less partial traffic
  ks = max(1, min(ks, *ntiles));
  ks = min(ks, 64);
  *ksplit = ks;
}

static bool wgrad_rows_ok(const mcd_conv_geom& g) {
  return packed_fprop_ok(g) && g.Cout <= 64 && g.R <= 7;
}

static void wgrad_rows_shape(const mcd_conv_geom& g, int* TH, int* TW, int* ntiles, int* nsplit) {
  pick_tile_packed(g.Ho, g.Wo, 64, TH, TW);
  *ntiles = g.N * ((g.Ho + *TH - 1) / *TH) * ((g.Wo + *TW - 1) / *TW);
  *nsplit = max(1, min(*ntiles, 148));
}

static int umma_wgrad_rows(const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes,
                           const mcd_conv_geom& g, int accumulate, cudaStream_t st) {
  int TH, TW, ntiles, nsplit;
  wgrad_rows_shape(g, &TH, &TW, &ntiles, &nsplit);
  size_t need = sizeof(float) * (size_t)nsplit * g.R * 64 * 64;
  if (ws_bytes < need || !ws) { set_error("umma wgrad: workspace %zu < %zu", ws_bytes, need); return MCD_E_WORKSPACE; }
  TapProblem p;
  plan_fprop_packed(g, p);
  WgradRowsArgs a;
  memset(&a, 0, sizeof(a));
  a.TH = TH; a.TW = TW; a.tiles_h = (g.Ho + TH - 1) / TH; a.tiles_w = (g.Wo + TW - 1) / TW;
  a.ntiles = ntiles; a.nsplit = nsplit; a.R = g.R; a.cs_src = g.Cin_s; a.smul = g.stride;
  const int stage_bytes = (1 + g.R) * 8192;
  a.stages = min(8, (200 * 1024) / stage_bytes);
  a.tmem_cols = g.R * 64 <= 64 ? 64 : (g.R * 64 <= 128 ? 128 : (g.R * 64 <= 256 ? 256 : 512));
  a.ws = reinterpret_cast<float*>(ws);
  for (int r = 0; r < g.R; ++r) a.taps[r] = p.taps[r];
  UmmaMaps maps;
  memset(&maps, 0, sizeof(maps));
  bool used[2] = {false, false};
  for (int r = 0; r < g.R; ++r) used[p.taps[r].map] = true;
  for (int ph = 0; ph < g.stride; ++ph) {
    if (!used[ph]) continue;
    int rc = encode_rows_map(&maps.a[ph], x, g.N, g.H, g.W, g.Cin_s, g.stride, ph, TH);
    if (rc != MCD_OK) return rc;
  }
  if (!used[0]) maps.a[0] = maps.a[p.taps[0].map];
  int rc = encode_act_map(&maps.b, dy, g.N, g.Ho, g.Wo, g.Cout, g.Cout_s, 1, 0, 0, 1, TH);
  if (rc != MCD_OK) return rc;
  const int smem_bytes = a.stages * stage_bytes + 1024 + 256;
  int arc = ensure_dyn_smem<conv_umma_wgrad_rows_kernel>(smem_bytes, "conv_umma_wgrad_rows");
  if (arc != MCD_OK) return arc;
  conv_umma_wgrad_rows_kernel<<<nsplit, kThreads, smem_bytes, st>>>(maps, a);
  rc = check_launch("conv_umma_wgrad_rows");
  if (rc != MCD_OK) return rc;
  int64_t total = (int64_t)g.Cout * g.Cin * g.R * g.S;
  int rgrid = (int)min64((total * 8 + 255) / 256, 148 * 8);
  wgrad_reduce_packed_kernel<<<rgrid, 256, 0, st>>>(a.ws, dw, nsplit, g.R, g.S, g.Cin_s, 64, g.Cout, g.Cin,
                                                    accumulate);
  return check_launch("wgrad_reduce");
}

size_t umma_wgrad_workspace(const mcd_conv_geom& g) {
  if (wgrad_toeplitz_ok(g)) return wgrad_toeplitz_workspace(g);
  if (wgrad_rows_ok(g)) {
    int TH, TW, ntiles, nsplit;
    wgrad_rows_shape(g, &TH, &TW, &ntiles, &nsplit);
    return sizeof(float) * (size_t)nsplit * g.R * 64 * 64;
  }
  int BN, CoutP, CinP, TH, TW, ntiles, ksplit;
  wgrad_shape(g, &BN, &CoutP, &CinP, &TH, &TW, &ntiles, &ksplit);
  return sizeof(float) * (size_t)ksplit * g.R * (packed_fprop_ok(g) ? 1 : g.S) * CoutP * CinP;
}

template <int BN, int OCC = 1>
static int launch_wgrad_bn(const UmmaMaps& maps, const WgradArgs& a, dim3 grid, cudaStream_t st) {
  using Cfg = WgradCfg<BN, OCC>;
  int arc = ensure_dyn_smem<conv_umma_wgrad_kernel<BN, OCC>>(Cfg::SMEM_BYTES, "conv_umma_wgrad");
  if (arc != MCD_OK) return arc;
  conv_umma_wgrad_kernel<BN, OCC><<<grid, kThreads, Cfg::SMEM_BYTES, st>>>(maps, a);
  return check_launch("conv_umma_wgrad");
}

int umma_wgrad(const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes,
               const mcd_conv_geom& g, int accumulate, cudaStream_t st) {
  if (g.stride != 1 && g.stride != 2) { set_error("umma wgrad: stride %d unsupported", g.stride); return MCD_E_INVALID; }
  if (g.R * g.S > kMaxTaps) { set_error("umma wgrad: too many taps"); return MCD_E_INVALID; }
  if (wgrad_toeplitz_ok(g)) {
    if (!dw) { set_error("umma wgrad: partial-sum output is not available for the stem layers"); return MCD_E_INVALID; }
    return wgrad_toeplitz_launch(x, dy, dw, ws, ws_bytes, g, accumulate, st);
  }
  if (wgrad_rows_ok(g)) {
    if (!dw) { set_error("umma wgrad: partial-sum output is not available for the stem layers"); return MCD_E_INVALID; }
    return umma_wgrad_rows(x, dy, dw, ws, ws_bytes, g, accumulate, st);
  }
  int BN, CoutP, CinP, TH, TW, ntiles, ksplit;
  wgrad_shape(g, &BN, &CoutP, &CinP, &TH, &TW, &ntiles, &ksplit);
  const bool packed = packed_fprop_ok(g);
  size_t need = sizeof(float) * (size_t)ksplit * g.R * (packed ? 1 : g.S) * CoutP * CinP;
  if (ws_bytes < need || !ws) { set_error("umma wgrad: workspace %zu < %zu", ws_bytes, need); return MCD_E_WORKSPACE; }

  TapProblem p;
  if (packed) plan_fprop_packed(g, p);
  else plan_fprop(g, p);
  WgradArgs a;
  memset(&a, 0, sizeof(a));
  a.N = g.N; a.TH = TH; a.TW = TW;
  a.tiles_h = (g.Ho + TH - 1) / TH; a.tiles_w = (g.Wo + TW - 1) / TW;
  a.ntiles = ntiles; a.ksplit = ksplit; a.T = p.ntaps; a.CoutP = CoutP; a.CinP = CinP;
  a.packed = packed; a.cs_src = g.Cin_s; a.smul = g.stride;
  a.ws = reinterpret_cast<float*>(ws);
  for (int t = 0; t < p.ntaps; ++t) a.taps[t] = p.taps[t];

  UmmaMaps maps;
  memset(&maps, 0, sizeof(maps));
  bool used[4] = {false, false, false, false};
  for (int t = 0; t < p.ntaps; ++t) used[p.taps[t].map] = true;
  if (packed) {
    for (int ph = 0; ph < g.stride; ++ph) {
      if (!used[ph]) continue;
      int rc = encode_rows_map(&maps.a[ph], x, g.N, g.H, g.W, g.Cin_s, g.stride, ph, TH);
      if (rc != MCD_OK) return rc;
    }
  } else {
    for (int ph = 0; ph < g.stride; ++ph)
      for (int pw = 0; pw < g.stride; ++pw) {
        int id = ph * g.stride + pw;
        if (!used[id]) continue;
        int rc = encode_act_map(&maps.a[id], x, g.N, g.H, g.W, g.Cin, g.Cin_s, g.stride, ph, pw, TW, TH);
        if (rc != MCD_OK) return rc;
      }
  }
  if (!used[0]) maps.a[0] = maps.a[p.taps[0].map];
  // dY map: packed mode loads one column (box width 1) at a time to get column-major pixel order
  int rc = encode_act_map(&maps.b, dy, g.N, g.Ho, g.Wo, g.Cout, g.Cout_s, 1, 0, 0, packed ? 1 : TW, TH);
  if (rc != MCD_OK) return rc;

  if (wgrad_pair(g)) {
    const int pair_items = (CoutP / 256) * (CinP / 256) * a.T * ksplit;
    rc = ensure_dyn_smem<conv_umma_wgrad_pair_kernel>(WgradPairCfg::SMEM_BYTES, "conv_umma_wgrad_pair");
    if (rc != MCD_OK) return rc;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * min(pair_items, sm_count() / 2)));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = WgradPairCfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_umma_wgrad_pair_kernel, maps, a);
    if (e != cudaSuccess) { set_error("conv_umma_wgrad (CTA pairs): %s", cudaGetErrorString(e)); return MCD_E_CUDA; }
    rc = check_launch("conv_umma_wgrad_pair");
    if (rc != MCD_OK) return rc;
    if (!dw) return MCD_OK;
    dim3 rg((unsigned)g.Cout, (unsigned)((g.Cin + 63) / 64));
    wgrad_reduce_kernel<<<rg, 256, sizeof(float) * 64 * a.T, st>>>(a.ws, dw, ksplit, a.T, CoutP, CinP, g.Cout, g.Cin,
                                                                   accumulate);
    return check_launch("wgrad_reduce");
  }
  const int total_items = (CoutP / 128) * (CinP / BN) * a.T * ksplit;
  const bool occ2 = BN <= 128 && thin_occ2();
  dim3 grid((unsigned)min(total_items, sm_count() * (occ2 ? 2 : 1)));
  if (BN == 256) rc = launch_wgrad_bn<256>(maps, a, grid, st);
  else if (BN == 128) rc = occ2 ? launch_wgrad_bn<128, 2>(maps, a, grid, st) : launch_wgrad_bn<128>(maps, a, grid, st);
  else rc = occ2 ? launch_wgrad_bn<64, 2>(maps, a, grid, st) : launch_wgrad_bn<64>(maps, a, grid, st);
  if (rc != MCD_OK) return rc;

  if (!dw) {          // partials only: the caller reduces [ksplit][T][CoutP][CinP] itself (mcd_sgd_pack_multi)
    if (packed) { set_error("umma wgrad: partial-sum output is not available for row-packed layers"); return MCD_E_INVALID; }
    return MCD_OK;
  }
  int64_t total = (int64_t)g.Cout * g.Cin * g.R * g.S;
  int rgrid = (int)min64((total * 8 + 255) / 256, 148 * 8);
  if (packed)
    wgrad_reduce_packed_kernel<<<rgrid, 256, 0, st>>>(a.ws, dw, ksplit, g.R, g.S, g.Cin_s, CoutP, g.Cout, g.Cin,
                                                      accumulate);
  else
  {
    dim3 rg((unsigned)g.Cout, (unsigned)((g.Cin + 63) / 64));
    wgrad_reduce_kernel<<<rg, 256, sizeof(float) * 64 * a.T, st>>>(a.ws, dw, ksplit, a.T, CoutP, CinP, g.Cout,
                                                                   g.Cin, accumulate);
  }
  return check_launch("wgrad_reduce");
}

}  // namespace mcd
