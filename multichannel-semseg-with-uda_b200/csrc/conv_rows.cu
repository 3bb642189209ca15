// conv_rows.cu — "row convolution" tcgen05 kernel for the thin stem layers (channel stride 8 or 16,
// stride 1, dilation 1: layer0 7x7 6->16, layer1 3x3 16->16 and its dgrad; models/drn.py:126-136).
//
// These layers are HBM-bound (AI 50-70 FLOP/B); an im2col-style A operand (one 128-byte K-window per output
// pixel) multiplies the L2->smem traffic by 12-28x and makes them TMA-bound.  Here the image row itself IS the A
// operand: with NHWC and 8 channels per 16-byte unit, the UMMA *no-swizzle* K-major layout
//     element (m, k) at  start + (m%8)*16 + (m/8)*SBO + (k/8)*LBO + (k%8)*2
// with SBO = 128 (8 pixels) and LBO = 16 (ONE pixel) makes k-group j of output pixel m read pixel m+j, i.e.
// A[m][(s,c)] = img[m+s][c] straight from a staged image-row segment: overlapping core matrices, zero expansion.
// One tile = 128 consecutive output pixels of one image row; per filter row r (and per 8-channel half for
// 16-channel tensors) one 2176-byte TMA box {8 ch, 136 px} is staged, and SP/2 MMAs (K=16 = two taps) consume it.
// Persistent CTAs, double-buffered TMEM accumulator, same epilogue contract as conv_umma_fprop_kernel
// (bias, bf16 NHWC store, fused BatchNorm sum / sum of squares, optional fused addend).
#include <stdlib.h>
#include "common.cuh"
#include "conv_plan.h"
#include "umma_ptx.cuh"

namespace mcd {

using namespace ptx;

constexpr int RC_THREADS = 192;
constexpr int RC_SEG = 136;              // staged pixels per row segment: 128 outputs + up to 8 taps
constexpr int RC_PLANE = RC_SEG * 16;    // bytes of one {8 channels x 136 pixels} plane (multiple of 128)

struct RowconvArgs {
  int N, H, W, tiles_w, total_tiles;
  int R, HC, SP, S;         // filter rows, 8-channel halves of the source, padded taps per row (4 or 8), taps
  int roff, woff;           // source row / pixel offset of tap (r=0, s=0) relative to the output pixel
  int NB;                   // GEMM N: produced channels padded to 16
  int rows, Cd_s;           // produced channels, destination channel stride
  int stages, stage_bytes, w_bytes;
  int planar;               // 1: out is planar fp32 [N][rows][H][W]
  int wide16;               // 16-channel source staged as ONE 32-byte-swizzled plane per filter row (see below)
  int fmt;                  // element format of src / weights / nhwc output: kF16 (forward) or kBF16 (dgrad)
  const float* bias;
  float* stats;
  int relu;                        // max(., 0) on the nhwc output (eval-mode unit)
  const __nv_bfloat16* addend;     // epilogue extras, see conv_plan.h EpiExtra
  const __nv_bfloat16* mask_src;   // bf16 twin (sign only)
  const __half* bn_y;              // forward tensor: IEEE half
  void* out;
  const __nv_bfloat16* wpacked;   // [R][HC*SP][NB/8][8][8] bf16, smem-ready no-swizzle K-major B operand
};

// 32-byte swizzle, K-major or MN-major: rows of 32 B (16 bf16), 8-row groups `sbo_bytes` apart, atoms `lbo_bytes` apart
__device__ __forceinline__ uint64_t smem_desc_sw32(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;     // layout type SWIZZLE_32B
  return d;
}

__device__ __forceinline__ uint64_t smem_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;     // descriptor version 1 (sm_100); layout type 0 = SWIZZLE_NONE
  return d;
}

// EXTRAS = false: forward instantiation without the dgrad-epilogue inputs (no spills at the 80 registers that four
// CTAs per SM leave)
template <int NB, bool EXTRAS = true>
__global__ void __launch_bounds__(RC_THREADS, NB == 16 ? 4 : 2)
conv_umma_rowconv_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ RowconvArgs a) {
  const __nv_bfloat16* const x_addend = EXTRAS ? a.addend : nullptr;
  const __nv_bfloat16* const x_mask = EXTRAS ? a.mask_src : nullptr;
  const __half* const x_bny = EXTRAS ? a.bn_y : nullptr;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;                                   // packed weights
  uint8_t* stage0 = smem + a.w_bytes;                    // w_bytes is a multiple of 1024
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage0 + a.stages * a.stage_bytes);
  uint64_t* empty_bar = full_bar + a.stages;
  uint64_t* tmem_full_bar = empty_bar + a.stages;        // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // weights: global -> smem once per (persistent) CTA, then make them visible to the async proxy (UMMA reads)
  for (int i = threadIdx.x; i < a.w_bytes / 16; i += RC_THREADS)
    reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(a.wpacked) + i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&xmap);
    for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * NB);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int planes = a.R * a.HC;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int wt = tile % a.tiles_w;
        const int h = (tile / a.tiles_w) % a.H;
        const int n = tile / (a.tiles_w * a.H);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = stage0 + stage * a.stage_bytes;
        mbar_expect_tx(&full_bar[stage], planes * RC_PLANE);
        if (a.wide16) {
          // 16-channel source: the pixel (32 B) is the swizzle-32B row; a tap shift is a ROW shift of the
          // descriptor start, legal because the swizzle XOR comes from absolute smem address bits.  Half the TMA
          // box rows (136 x 32 B instead of 2 x 136 x 16 B per filter row) and one MMA per tap.
          for (int r = 0; r < a.R; ++r)
            tma_load_4d(st + r * 2 * RC_PLANE, &xmap, &full_bar[stage], 0, wt * 128 + a.woff, h + r + a.roff, n);
        } else {
          for (int r = 0; r < a.R; ++r)
            for (int hf = 0; hf < a.HC; ++hf)
              tma_load_4d(st + (r * a.HC + hf) * RC_PLANE, &xmap, &full_bar[stage], hf * 8, wt * 128 + a.woff,
                          h + r + a.roff, n);
        }
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = instr_desc_16(128, NB, 0, 0, (uint32_t)a.fmt);
    constexpr uint32_t kg_bytes = (NB / 8) * 128;          // one 8-wide k-group of the weights: NB rows x 16 B
    const uint32_t wbase = smem_u32(wsm);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t st = smem_u32(stage0 + stage * a.stage_bytes);
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * NB);
        uint32_t first = 1;
        if (a.wide16) {
          for (int r = 0; r < a.R; ++r) {
            for (int tap = 0; tap < a.S; ++tap) {
              // A: K = the 16 channels of pixel m + tap; B: k-groups (half 0, tap), (half 1, tap) are SP groups apart
              const uint64_t adesc = smem_desc_sw32(st + r * 2 * RC_PLANE + tap * 32, 0, 256);
              const uint64_t bdesc = smem_desc_nosw(wbase + (uint32_t)(r * 2 * a.SP + tap) * kg_bytes,
                                                    (uint32_t)a.SP * kg_bytes, 128);
              umma_bf16(tmem_d, adesc, bdesc, idesc, first ? 0u : 1u);
              first = 0;
            }
          }
        }
        for (int p = 0; p < (a.wide16 ? 0 : planes); ++p) {
          for (int kk = 0; kk < a.SP / 2; ++kk) {
            // A: k-groups = taps 2kk, 2kk+1 -> pixels m+2kk, m+2kk+1 (LBO = one pixel = 16 B, SBO = 8 pixels)
            const uint64_t adesc = smem_desc_nosw(st + p * RC_PLANE + kk * 32, 16, 128);
            const uint64_t bdesc = smem_desc_nosw(wbase + (uint32_t)(p * a.SP + 2 * kk) * kg_bytes, kg_bytes, 128);
            umma_bf16(tmem_d, adesc, bdesc, idesc, first ? 0u : 1u);
            first = 0;
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full_bar[acc]);
      }
      __syncwarp();
      if (++stage == a.stages) { stage = 0; phase ^= 1; }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const bool want_stats = a.stats != nullptr;
    float s1[NB], s2[NB];                                   // per-thread BatchNorm partial sums over all tiles
#pragma unroll
    for (int j = 0; j < NB; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int wt = tile % a.tiles_w;
      const int h = (tile / a.tiles_w) % a.H;
      const int n = tile / (a.tiles_w * a.H);
      const int wpix = wt * 128 + m;
      const bool pvalid = wpix < a.W;
      const int64_t ooff = (((int64_t)n * a.H + h) * a.W + wpix) * a.Cd_s;
      uint4 addv[NB / 8], mskv[NB / 8], yv[NB / 8];         // issue the extra loads before waiting for the MMAs
#pragma unroll
      for (int j8 = 0; j8 < NB / 8; ++j8) {
        if (pvalid && j8 * 8 < a.Cd_s) {
          if (x_addend) addv[j8] = __ldg(reinterpret_cast<const uint4*>(x_addend + ooff + j8 * 8));
          if (x_mask) mskv[j8] = __ldg(reinterpret_cast<const uint4*>(x_mask + ooff + j8 * 8));
          if (x_bny) yv[j8] = __ldg(reinterpret_cast<const uint4*>(x_bny + ooff + j8 * 8));
        }
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      float v[NB];
      if (NB == 16) tmem_ld16(tmem_base + (uint32_t)(acc * NB) + ((uint32_t)(q * 32) << 16), v);
      else tmem_ld32(tmem_base + (uint32_t)(acc * NB) + ((uint32_t)(q * 32) << 16), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);      // registers hold the tile: release the accumulator
      if (a.bias) {
#pragma unroll
        for (int j = 0; j < NB; ++j) v[j] += (j < a.rows) ? __ldg(a.bias + j) : 0.f;
      }
      if (want_stats && pvalid && !x_bny) {
#pragma unroll
        for (int j = 0; j < NB; ++j) { s1[j] += v[j]; s2[j] = fmaf(v[j], v[j], s2[j]); }
      }
      if (a.planar) {
        if (pvalid) {
          float* o = reinterpret_cast<float*>(a.out) + ((int64_t)n * a.rows * a.H + h) * a.W + wpix;
#pragma unroll
          for (int j = 0; j < NB; ++j)
            if (j < a.rows) o[(int64_t)j * a.H * a.W] = v[j];
        }
      } else if (pvalid) {
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out);
#pragma unroll
        for (int j8 = 0; j8 < NB / 8; ++j8) {
          const int c = j8 * 8;
          if (c < a.Cd_s) {
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = (c + k < a.rows) ? v[c + k] : 0.f;
            if (x_addend) {
              float r8[8];
              unpack8r(addv[j8], r8, a.fmt);
#pragma unroll
              for (int k = 0; k < 8; ++k) f[k] += r8[k];
            }
            if (a.relu) {
#pragma unroll
              for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
            }
            if (x_mask) {
              float r8[8];
              unpack8(mskv[j8], r8);
#pragma unroll
              for (int k = 0; k < 8; ++k) f[k] = r8[k] > 0.f ? f[k] : 0.f;
            }
            if (x_bny) {                                     // fused BatchNorm-backward sums: g, g * y
              float r8[8];
              unpack8h(yv[j8], r8);
#pragma unroll
              for (int k = 0; k < 8; ++k) { s1[c + k] += f[k]; s2[c + k] = fmaf(f[k], r8[k], s2[c + k]); }
            }
            *reinterpret_cast<uint4*>(out + ooff + c) = pack8r(f, a.fmt);
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (want_stats) {
      // lanes -> columns: after the butterfly lane l holds the warp total of column l (mod NB)
      if (NB == 16) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
          s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
        }
      }
#pragma unroll
      for (int step = NB / 2; step >= 1; step >>= 1) {
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
      if (lane < NB && lane < a.rows) {
        atomicAdd(a.stats + lane, s1[0]);
        atomicAdd(a.stats + a.rows + lane, s2[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * NB);
}

// ------------------------------------------------------------------------------------------------------------------
// Weight gradient of the same stem layers ("Toeplitz" wgrad, r01b).  dW[co][c][r][s] = sum_p dY[p][co] * x[p + (r,s)][c].
// Per 128-pixel row tile and filter row r the staged image-row segment IS the MN-major A operand:
//     A[m = (s, c8)][k = p] = seg[(p + s) * 16 B + c8 * 2 B]     (no-swizzle MN-major, 8-element groups 16 B apart = ONE
//     pixel, 8-pixel K groups 128 B apart), i.e. overlapping core matrices, zero window expansion - the im2col-style
// kernel (conv_umma_wgrad_rows_kernel) moves 7-8x the bytes through L2 and is L2 -> SM bound.  dY is staged as
// 8-channel planes {8 ch, 128 px} (B operand, N = produced channels).  D_(r,half)[64 = 8 taps x 8 ch][NB] accumulates
// in TMEM over ALL tiles of the persistent CTA; one epilogue per CTA parks it in the workspace
// ws[cta][r * HC + half][64][NB], summed over CTAs by wgrad_toeplitz_reduce_kernel.
struct ToepArgs {
  int N, H, W, tiles_w, total_tiles;
  int R, HC, nacc;          // filter rows, 8-channel halves of x, accumulators = R * HC
  int roff, woff;           // x row / pixel offset of tap (0, 0) relative to the output pixel (-pad)
  int NB, DC;               // GEMM N = dY channel stride (16 / 32), dY planes = NB / 8
  int wide16;               // 16-channel x staged as ONE 32-byte-swizzled plane per filter row: M = 4 taps x 16 ch
  int dy32;                 // NB == 16: dY staged as 32-byte-swizzled pixel rows (one box instead of two planes)
  int xplane_bytes, dy_off; // bytes of one x plane, offset of the dY region inside a stage
  int stages, stage_bytes, tmem_cols;
  float* ws;
};

__global__ void __launch_bounds__(RC_THREADS, 2)
conv_umma_wgrad_toeplitz_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap dymap,
                                const __grid_constant__ ToepArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + a.stages * a.stage_bytes);
  uint64_t* empty_bar = full_bar + a.stages;
  uint64_t* tmem_full_bar = empty_bar + a.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&xmap);
    prefetch_tmap(&dymap);
    for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int dy_off = a.dy_off;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int wt = tile % a.tiles_w;
        const int h = (tile / a.tiles_w) % a.H;
        const int n = tile / (a.tiles_w * a.H);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * a.stage_bytes;
        mbar_expect_tx(&full_bar[stage], a.nacc * a.xplane_bytes + a.DC * 2048);
        if (a.wide16) {
          for (int r = 0; r < a.R; ++r)
            tma_load_4d(st + r * a.xplane_bytes, &xmap, &full_bar[stage], 0, wt * 128 + a.woff, h + r + a.roff, n);
        } else {
          for (int r = 0; r < a.R; ++r)
            for (int hf = 0; hf < a.HC; ++hf)
              tma_load_4d(st + (r * a.HC + hf) * RC_PLANE, &xmap, &full_bar[stage], hf * 8, wt * 128 + a.woff,
                          h + r + a.roff, n);
        }
        if (a.dy32) {
          tma_load_4d(st + dy_off, &dymap, &full_bar[stage], 0, wt * 128, h, n);
        } else {
          for (int d = 0; d < a.DC; ++d)
            tma_load_4d(st + dy_off + d * 2048, &dymap, &full_bar[stage], d * 8, wt * 128, h, n);
        }
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = instr_desc_bf16(64, (uint32_t)a.NB, 1, 1);
    int stage = 0; uint32_t phase = 0;
    uint32_t first = 1;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t st = smem_u32(smem + stage * a.stage_bytes);
        for (int acc = 0; acc < a.nacc; ++acc) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {       // 128 pixels = 8 x UMMA_K(16) pixels
            // A (x): 8-channel plane: 8-element groups ONE pixel (16 B) apart, 8-pixel K groups 128 B apart;
            //        32-byte-swizzled 16-channel rows: 16-element atoms one pixel (32 B) apart, K groups 256 B apart
            const uint64_t adesc = a.wide16 ? smem_desc_sw32(st + acc * a.xplane_bytes + j * 512, 32, 256)
                                            : smem_desc_nosw(st + acc * RC_PLANE + j * 256, 128, 16);
            const uint64_t bdesc = a.dy32 ? smem_desc_sw32(st + dy_off + j * 512, 32, 256)
                                          : smem_desc_nosw(st + dy_off + j * 256, 128, 2048);
            umma_bf16(tmem_base + (uint32_t)(acc * a.NB), adesc, bdesc, idesc, (first && j == 0) ? 0u : 1u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (tile + (int)gridDim.x >= a.total_tiles) umma_commit(tmem_full_bar);
      }
      first = 0;
      __syncwarp();
      if (++stage == a.stages) { stage = 0; phase ^= 1; }
    }
  } else {
    // M = 64 accumulators live in lanes 0..15 of each 32-lane TMEM sub-partition: row m = 16 * q + lane
    const int q = warp & 3;
    const int m = q * 16 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    for (int acc = 0; acc < a.nacc; ++acc) {
      float* o = a.ws + (((int64_t)blockIdx.x * a.nacc + acc) * 64 + m) * a.NB;
      for (int c0 = 0; c0 < a.NB; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * a.NB + c0), v);
        tmem_ld_wait();
        if (lane < 16) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(o + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

// dw[co][c][r][s] = sum_cta ws[cta][r * HC + c / CG][s * CG + c % CG][co]; 8 lanes share one output element
__global__ void wgrad_toeplitz_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int ncta, int nacc,
                                             int HC, int CG, int NB, int R, int S, int Cout, int Cin, int accumulate) {
  const int64_t total = (int64_t)Cout * Cin * R * S;
  const int sub = threadIdx.x & 7;
  for (int64_t i0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3; i0 < ((total + 31) & ~int64_t(31));
       i0 += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    const int64_t i = i0 < total ? i0 : total - 1;
    const int sx = (int)(i % S);
    const int r = (int)((i / S) % R);
    const int c = (int)((i / ((int64_t)S * R)) % Cin);
    const int co = (int)(i / ((int64_t)S * R * Cin));
    const int64_t off = ((int64_t)(r * HC + c / CG) * 64 + sx * CG + c % CG) * NB + co;   // CG channels per M group
    const int64_t cstride = (int64_t)nacc * 64 * NB;
    float acc = 0.f;
    for (int k = sub; k < ncta; k += 8) acc += ws[k * cstride + off];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (sub == 0 && i0 < total) dw[i] = accumulate ? dw[i] + acc : acc;
  }
}

// mode 0 (fprop operand): IEEE half; mode 1 (dgrad operand): bfloat16 - the format of the tensor it multiplies
__global__ void pack_weight_rowconv_kernel(const float* __restrict__ w, uint16_t* __restrict__ dst, int Cout,
                                           int Cin, int R, int S, int HC, int SP, int NB, int mode) {
  const int64_t total = (int64_t)R * HC * SP * NB * 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = f2bits16(rowconv_pack_value(w, i, Cout, Cin, R, S, HC, SP, NB, mode), mode ? kBF16 : kF16);
}

// ---- host -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnR)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFnR get_encode_r() {
  static EncodeTiledFnR fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !p)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFnR>(p);
  }
  return fn;
}

// MCD_ROWCONV_WIDE16=0: stage 16-channel sources as two 8-channel planes (16-byte TMA rows) as in r01
static bool rowconv_wide16_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MCD_ROWCONV_WIDE16"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

bool rowconv_fprop_ok(const mcd_conv_geom& g) {
  return g.stride == 1 && g.dil == 1 && (g.Cin_s == 8 || g.Cin_s == 16) && g.S <= 8 && g.R <= 7 && g.Cout <= 32 &&
         g.Cout_s <= 32 && g.pad <= 7 && g.Ho == g.H && g.Wo == g.W;
}
bool rowconv_dgrad_ok(const mcd_conv_geom& g) {
  return g.stride == 1 && g.dil == 1 && (g.Cout_s == 8 || g.Cout_s == 16) && g.S <= 8 && g.R <= 7 && g.Cin <= 32 &&
         g.Cin_s <= 32 && g.S - 1 - g.pad >= 0 && g.R - 1 - g.pad >= 0 && g.S - 1 - g.pad <= 7 && g.Ho == g.H &&
         g.Wo == g.W;
}

void rowconv_pack_dims(const mcd_conv_geom& g, int mode, int* HC, int* SP, int* NB) {
  *HC = (mode ? g.Cout_s : g.Cin_s) / 8;
  *SP = g.S <= 4 ? 4 : 8;
  *NB = round_up(mode ? g.Cin : g.Cout, 16);
}

int rowconv_pack(const float* w, void* dst, int Cout, int Cin, int R, int S, int Cs, int mode, cudaStream_t st) {
  const int HC = Cs / 8, SP = S <= 4 ? 4 : 8, NB = round_up(mode ? Cin : Cout, 16);
  int64_t total = (int64_t)R * HC * SP * NB * 8;
  int grid = (int)min64((total + 255) / 256, 148 * 4);
  pack_weight_rowconv_kernel<<<grid, 256, 0, st>>>(w, (uint16_t*)dst, Cout, Cin, R, S, HC, SP, NB, mode);
  return check_launch("pack_weight_rowconv");
}

// mode 0: fprop (src = x, produces Cout); mode 1: dgrad (src = dy, produces Cin)
int rowconv_launch(const void* src, const void* wpacked, const float* bias, void* out, int planar, float* stats,
                   const EpiExtra& ex, const mcd_conv_geom& g, int mode, cudaStream_t st) {
  EncodeTiledFnR enc = get_encode_r();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return MCD_E_CUDA; }
  RowconvArgs a;
  memset(&a, 0, sizeof(a));
  a.N = g.N; a.H = g.H; a.W = g.W;
  a.tiles_w = (g.W + 127) / 128;
  a.total_tiles = g.N * g.H * a.tiles_w;
  a.R = g.R; a.S = g.S;
  rowconv_pack_dims(g, mode, &a.HC, &a.SP, &a.NB);
  a.wide16 = (a.HC == 2 && rowconv_wide16_enabled()) ? 1 : 0;
  a.roff = mode ? -(g.R - 1 - g.pad) : -g.pad;
  a.woff = mode ? -(g.S - 1 - g.pad) : -g.pad;
  a.rows = mode ? g.Cin : g.Cout;
  a.Cd_s = mode ? g.Cin_s : g.Cout_s;
  a.stage_bytes = a.R * a.HC * RC_PLANE;
  a.w_bytes = round_up(a.R * a.HC * a.SP * (a.NB / 8) * 128, 1024);
  a.stages = max(2, min(4, (54 * 1024 - a.w_bytes) / a.stage_bytes));
  a.bias = bias; a.stats = stats; a.fmt = mode ? kBF16 : kF16;
  a.relu = planar ? 0 : ex.relu;
  a.addend = planar ? nullptr : (const __nv_bfloat16*)ex.addend;
  a.mask_src = planar ? nullptr : (const __nv_bfloat16*)ex.mask_src;
  a.bn_y = planar ? nullptr : (const __half*)ex.bn_y;
  a.out = out; a.planar = planar; a.wpacked = (const __nv_bfloat16*)wpacked;
  const int srcC = mode ? g.Cout : g.Cin, srcCs = mode ? g.Cout_s : g.Cin_s;
  CUtensorMap xmap;
  cuuint64_t dims[4] = {(cuuint64_t)srcC, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
  cuuint64_t strides[3] = {(cuuint64_t)srcCs * 2, (cuuint64_t)g.W * srcCs * 2, (cuuint64_t)g.H * g.W * srcCs * 2};
  cuuint32_t box[4] = {a.wide16 ? 16u : 8u, RC_SEG, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&xmap, mode ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                   const_cast<void*>(src), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, a.wide16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(rowconv) failed: %d", (int)r); return MCD_E_CUDA; }
  const int smem_bytes = a.w_bytes + a.stages * a.stage_bytes + 1024 + 256;
  const bool extras = a.addend || a.mask_src || a.bn_y;
  auto kern = a.NB == 16 ? (extras ? conv_umma_rowconv_kernel<16, true> : conv_umma_rowconv_kernel<16, false>)
                         : (extras ? conv_umma_rowconv_kernel<32, true> : conv_umma_rowconv_kernel<32, false>);
  int arc = a.NB == 16 ? (extras ? ensure_dyn_smem<conv_umma_rowconv_kernel<16, true>>(smem_bytes, "conv_umma_rowconv")
                                 : ensure_dyn_smem<conv_umma_rowconv_kernel<16, false>>(smem_bytes, "conv_umma_rowconv"))
                       : (extras ? ensure_dyn_smem<conv_umma_rowconv_kernel<32, true>>(smem_bytes, "conv_umma_rowconv")
                                 : ensure_dyn_smem<conv_umma_rowconv_kernel<32, false>>(smem_bytes, "conv_umma_rowconv"));
  if (arc != MCD_OK) return arc;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  int grid = min(a.total_tiles, sms * (a.NB == 16 ? 4 : 2));
  kern<<<grid, RC_THREADS, smem_bytes, st>>>(xmap, a);
  return check_launch("conv_umma_rowconv");
}

// ---- Toeplitz wgrad host side ------------------------------------------------------------------------------------
static int toeplitz_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

bool wgrad_toeplitz_ok(const mcd_conv_geom& g) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("MCD_TOEPLITZ_WGRAD"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled) return false;
  const int HC = g.Cin_s / 8;
  return g.stride == 1 && g.dil == 1 && (g.Cin_s == 8 || g.Cin_s == 16) && (g.Cout_s == 16 || g.Cout_s == 32) &&
         g.S <= 8 && g.R <= 7 && g.pad <= 7 && g.Ho == g.H && g.Wo == g.W && g.R * HC * g.Cout_s <= 512;
}

static int toeplitz_grid(const mcd_conv_geom& g) {
  const int total_tiles = g.N * g.H * ((g.W + 127) / 128);
  return min(total_tiles, 2 * toeplitz_sms());
}

size_t wgrad_toeplitz_workspace(const mcd_conv_geom& g) {
  return sizeof(float) * (size_t)toeplitz_grid(g) * g.R * (g.Cin_s / 8) * 64 * g.Cout_s;
}

int wgrad_toeplitz_launch(const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes, const mcd_conv_geom& g,
                          int accumulate, cudaStream_t st) {
  EncodeTiledFnR enc = get_encode_r();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return MCD_E_CUDA; }
  if (!ws || ws_bytes < wgrad_toeplitz_workspace(g)) {
    set_error("toeplitz wgrad: workspace %zu < %zu", ws_bytes, wgrad_toeplitz_workspace(g));
    return MCD_E_WORKSPACE;
  }
  ToepArgs a;
  memset(&a, 0, sizeof(a));
  a.N = g.N; a.H = g.H; a.W = g.W;
  a.tiles_w = (g.W + 127) / 128;
  a.total_tiles = g.N * g.H * a.tiles_w;
  a.R = g.R; a.HC = g.Cin_s / 8;
  a.wide16 = (a.HC == 2 && g.S <= 4 && rowconv_wide16_enabled()) ? 1 : 0;
  a.nacc = a.wide16 ? a.R : a.R * a.HC;
  a.roff = -g.pad; a.woff = -g.pad;
  a.NB = g.Cout_s; a.DC = g.Cout_s / 8;
  a.dy32 = (a.NB == 16 && rowconv_wide16_enabled()) ? 1 : 0;
  a.xplane_bytes = a.wide16 ? 2 * RC_PLANE : RC_PLANE;
  a.dy_off = round_up(a.nacc * a.xplane_bytes, 256);
  a.stage_bytes = round_up(a.dy_off + a.DC * 2048, 256);
  a.stages = max(2, min(6, (100 * 1024) / a.stage_bytes));
  const int cols = a.nacc * a.NB;
  a.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
  a.ws = reinterpret_cast<float*>(ws);
  CUtensorMap xmap, dymap;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cin_s, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cin_s * 2, (cuuint64_t)g.W * g.Cin_s * 2, (cuuint64_t)g.H * g.W * g.Cin_s * 2};
    cuuint32_t box[4] = {a.wide16 ? 16u : 8u, RC_SEG, 1, 1};
    CUresult r = enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, a.wide16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(toeplitz x) failed: %d", (int)r); return MCD_E_CUDA; }
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cout_s, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cout_s * 2, (cuuint64_t)g.W * g.Cout_s * 2, (cuuint64_t)g.H * g.W * g.Cout_s * 2};
    cuuint32_t box[4] = {a.dy32 ? 16u : 8u, 128, 1, 1};
    CUresult r = enc(&dymap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, a.dy32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(toeplitz dy) failed: %d", (int)r); return MCD_E_CUDA; }
  }
  const int smem_bytes = a.stages * a.stage_bytes + 1024 + 256;
  int arc = ensure_dyn_smem<conv_umma_wgrad_toeplitz_kernel>(smem_bytes, "conv_umma_wgrad_toeplitz");
  if (arc != MCD_OK) return arc;
  const int grid = toeplitz_grid(g);
  conv_umma_wgrad_toeplitz_kernel<<<grid, RC_THREADS, smem_bytes, st>>>(xmap, dymap, a);
  int rc = check_launch("conv_umma_wgrad_toeplitz");
  if (rc != MCD_OK) return rc;
  const int64_t total = (int64_t)g.Cout * g.Cin * g.R * g.S;
  const int rgrid = (int)min64((total * 8 + 255) / 256, 148 * 8);
  wgrad_toeplitz_reduce_kernel<<<rgrid, 256, 0, st>>>(a.ws, dw, grid, a.nacc, a.wide16 ? 1 : a.HC, a.wide16 ? 16 : 8,
                                                      a.NB, g.R, g.S, g.Cout, g.Cin, accumulate);
  return check_launch("wgrad_toeplitz_reduce");
}

}  // namespace mcd
