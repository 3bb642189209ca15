// layout.cu — boundary conversions between the reference's NCHW fp32 / OIHW fp32 tensors and the
// library's NHWC bf16 activations and packed bf16 weights.
#include "common.cuh"

namespace mcd {

// one thread = one pixel x 8-channel group; lanes walk pixels so plane reads are coalesced.
__global__ void nchw_f32_to_nhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst_h,
                                        __nv_bfloat16* __restrict__ dst_b, int N, int C, int HW, int Cs) {
  int groups = Cs >> 3;
  int64_t total = (int64_t)N * HW * groups;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t pix = i % ((int64_t)N * HW);
    int g = (int)(i / ((int64_t)N * HW));
    int n = (int)(pix / HW);
    int hw = (int)(pix % HW);
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int c = g * 8 + k;
      f[k] = c < C ? src[((int64_t)n * C + c) * HW + hw] : 0.f;
    }
    if (dst_h) *reinterpret_cast<uint4*>(dst_h + pix * Cs + g * 8) = pack8h(f);
    if (dst_b) *reinterpret_cast<uint4*>(dst_b + pix * Cs + g * 8) = pack8(f);
  }
}

__global__ void nhwc_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ src, int fmt,
                                        float* __restrict__ dst, int N, int C, int HW, int Cs) {
  int groups = (C + 7) >> 3;
  int64_t total = (int64_t)N * HW * groups;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t pix = i % ((int64_t)N * HW);
    int g = (int)(i / ((int64_t)N * HW));
    int n = (int)(pix / HW);
    int hw = (int)(pix % HW);
    uint4 v = *reinterpret_cast<const uint4*>(src + pix * Cs + g * 8);
    float f[8];
    unpack8r(v, f, fmt);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int c = g * 8 + k;
      if (c < C) dst[((int64_t)n * C + c) * HW + hw] = f[k];
    }
  }
}

// 16-bit -> 16-bit re-encoding of a dense tensor (bf16 <-> IEEE half): creates the missing twin of an activation that
// entered the library in one format only (e.g. a bf16 tensor produced by stock torch code)
__global__ void convert16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int src_fmt, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float f[8];
    unpack8r(src[i], f, src_fmt);
    dst[i] = pack8r(f, src_fmt == kF16 ? kBF16 : kF16);
  }
}

// mode 0: dst[co][r*S+s][kc]            = w[co][kc][r][s]                 (kc < Cin, else 0)
// mode 1: dst[ci][(R-1-r)*S+(S-1-s)][kc] = w[kc][ci][r][s]                 (kc < Cout, else 0)
// packs are stored in the format of the tensor they multiply: mode 0 (fprop) IEEE half, mode 1 (dgrad) bfloat16
__global__ void pack_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ dst,
                                   int Cout, int Cin, int R, int S, int mode, int rows, int kc_pad) {
  int T = R * S;
  int64_t total = (int64_t)rows * T * kc_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int kc = (int)(i % kc_pad);
    int t = (int)((i / kc_pad) % T);
    int row = (int)(i / ((int64_t)kc_pad * T));
    float v = 0.f;
    if (mode == 0) {
      if (kc < Cin) v = w[(((int64_t)row * Cin + kc) * R + t / S) * S + t % S];
    } else {
      if (kc < Cout) {
        int r = R - 1 - t / S, s = S - 1 - t % S;
        v = w[(((int64_t)kc * Cin + row) * R + r) * S + s];
      }
    }
    dst[i] = f2bits16(v, mode ? kBF16 : kF16);
  }
}

// row-packed variants (conv_plan.h): dst[row][r][k], k = s*Cs + c, 64 elements per filter row
// mode 0: row = co, value w[co][c][r][s]            (c < Cin)
// mode 1: row = ci, value w[c][ci][R-1-r][S-1-s]    (c < Cout)   - flipped filter for dgrad
__global__ void pack_weight_rows_kernel(const float* __restrict__ w, uint16_t* __restrict__ dst,
                                        int Cout, int Cin, int R, int S, int Cs, int mode, int rows) {
  int64_t total = (int64_t)rows * R * 64;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int k = (int)(i % 64);
    int r = (int)((i / 64) % R);
    int row = (int)(i / (64 * (int64_t)R));
    int s = k / Cs, c = k % Cs;
    float v = 0.f;
    if (s < S) {
      if (mode == 0) {
        if (c < Cin) v = w[(((int64_t)row * Cin + c) * R + r) * S + s];
      } else {
        if (c < Cout) v = w[(((int64_t)c * Cin + row) * R + (R - 1 - r)) * S + (S - 1 - s)];
      }
    }
    dst[i] = f2bits16(v, mode ? kBF16 : kF16);
  }
}

// ---- multi-tensor re-pack: ONE launch refreshes every bf16 shadow of a model after an optimizer step ------
// item (12 x int64 in device memory): w, dst_fprop, dst_dgrad (0 = absent), Cout, Cin, R, S, kind_f, kind_d,
// cs_f, cs_d, unused.  kind 0 = [rows][taps][kc_pad] layout, 1 = row-packed [rows][R][64] layout, 2 = rowconv
// layout (conv_rows.cu).
// Standard layouts: a block transposes one 32(co) x 32(ci) x T tile through shared memory so that the fp32 OIHW
// reads and BOTH bf16 writes are coalesced (the pad channels of the packs were zeroed when they were created and
// are never touched again).  Row-packed layouts (three tiny stem layers) use the element-wise path.
constexpr int PK_T_MAX = 9;
__global__ void __launch_bounds__(256)
pack_weights_multi_kernel(const int64_t* __restrict__ items) {
  __shared__ float tile[32][32 * PK_T_MAX + 1];
  const int64_t* it = items + (int64_t)blockIdx.x * 12;
  const float* w = reinterpret_cast<const float*>(it[0]);
  uint16_t* dst_f = reinterpret_cast<uint16_t*>(it[1]);     // fprop pack: IEEE half
  uint16_t* dst_d = reinterpret_cast<uint16_t*>(it[2]);     // dgrad pack: bfloat16
  const int Cout = (int)it[3], Cin = (int)it[4], R = (int)it[5], S = (int)it[6];
  const int kind_f = (int)it[7], kind_d = (int)it[8], cs_f = (int)it[9], cs_d = (int)it[10];
  const int T = R * S;
  const bool std_f = dst_f && kind_f == 0, std_d = dst_d && kind_d == 0;
  if ((std_f || std_d) && T <= PK_T_MAX) {
    const int tiles_ci = (Cin + 31) / 32, tiles_co = (Cout + 31) / 32;
    const int kpf = (Cin + 63) / 64 * 64, kpd = (Cout + 63) / 64 * 64;
    for (int tl = blockIdx.y; tl < tiles_ci * tiles_co; tl += gridDim.y) {
      const int co0 = (tl / tiles_ci) * 32, ci0 = (tl % tiles_ci) * 32;
      const int nci = min(32, Cin - ci0), nco = min(32, Cout - co0);
      __syncthreads();
      for (int idx = threadIdx.x; idx < 32 * nci * T; idx += 256) {      // co_l slowest, (ci_l, t) contiguous
        const int co_l = idx / (nci * T), rem = idx % (nci * T);
        if (co_l < nco) tile[co_l][rem] = w[((int64_t)(co0 + co_l) * Cin + ci0) * T + rem];
      }
      __syncthreads();
      if (std_f) {
        for (int idx = threadIdx.x; idx < nco * T * 32; idx += 256) {    // ci_l fastest
          const int ci_l = idx & 31, t = (idx >> 5) % T, co_l = idx / (32 * T);
          if (ci_l < nci)
            dst_f[((int64_t)(co0 + co_l) * T + t) * kpf + ci0 + ci_l] = f2bits16(tile[co_l][ci_l * T + t], kF16);
        }
      }
      if (std_d) {
        for (int idx = threadIdx.x; idx < nci * T * 32; idx += 256) {    // co_l fastest, flipped taps
          const int co_l = idx & 31, t = (idx >> 5) % T, ci_l = idx / (32 * T);
          if (co_l < nco) {
            const int tf = (R - 1 - t / S) * S + (S - 1 - t % S);
            dst_d[((int64_t)(ci0 + ci_l) * T + tf) * kpd + co0 + co_l] = f2bits16(tile[co_l][ci_l * T + t], kBF16);
          }
        }
      }
    }
  }
  // element-wise paths: row-packed layouts, or filters larger than the smem tile allows
  for (int pass = 0; pass < 2; ++pass) {
    uint16_t* dst = pass ? dst_d : dst_f;
    const int fmt = pass ? kBF16 : kF16;
    const int kind = pass ? kind_d : kind_f, Cs = pass ? cs_d : cs_f;
    if (!dst) continue;
    if (kind == 0 && T <= PK_T_MAX) continue;
    const int rows = pass ? Cin : Cout;
    if (kind == 0) {
      const int kc_pad = ((pass ? Cout : Cin) + 63) / 64 * 64;
      const int64_t total = (int64_t)rows * T * kc_pad;
      for (int64_t i = blockIdx.y * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.y * blockDim.x) {
        const int kc = (int)(i % kc_pad), t = (int)((i / kc_pad) % T), row = (int)(i / ((int64_t)kc_pad * T));
        float v = 0.f;
        if (!pass) { if (kc < Cin) v = w[(((int64_t)row * Cin + kc) * R + t / S) * S + t % S]; }
        else if (kc < Cout) v = w[(((int64_t)kc * Cin + row) * R + (R - 1 - t / S)) * S + (S - 1 - t % S)];
        dst[i] = f2bits16(v, fmt);
      }
    } else if (kind == 2) {
      const int HC = Cs / 8, SP = S <= 4 ? 4 : 8, NB = (rows + 15) / 16 * 16;
      const int64_t total = (int64_t)R * HC * SP * NB * 8;
      for (int64_t i = blockIdx.y * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.y * blockDim.x)
        dst[i] = f2bits16(rowconv_pack_value(w, i, Cout, Cin, R, S, HC, SP, NB, pass), fmt);
    } else {
      const int64_t total = (int64_t)rows * R * 64;
      for (int64_t i = blockIdx.y * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.y * blockDim.x) {
        const int k = (int)(i % 64), r = (int)((i / 64) % R), row = (int)(i / (64 * (int64_t)R));
        const int s = k / Cs, c = k % Cs;
        float v = 0.f;
        if (s < S) {
          if (!pass) { if (c < Cin) v = w[(((int64_t)row * Cin + c) * R + r) * S + s]; }
          else if (c < Cout) v = w[(((int64_t)c * Cin + row) * R + (R - 1 - r)) * S + (S - 1 - s)];
        }
        dst[i] = f2bits16(v, fmt);
      }
    }
  }
}

// ---- fused optimizer step: SGD (momentum, weight decay) on every parameter of a model AND the refresh of the packed
// bf16 shadows of its convolution weights, in ONE launch (torch.optim.SGD.step() is ~60 multi-tensor launches and the
// re-pack read all weights a second time).
// item (16 x int64): p, g, buf (0 = no momentum), dst_fprop, dst_dgrad (0 = absent), Cout, Cin, R, S, kind_f, kind_d,
// cs_f, cs_d, numel, 0, 0.  hyper (device, fp32): lr, momentum, weight_decay - read at run time, so a captured CUDA
// graph follows adjust_learning_rate().   d = g + wd*p ; buf = momentum*buf + d ; p -= lr*buf   (torch.optim.SGD with
// dampening 0; buf starts as zeros, which reproduces torch's first-step "buf = d").
__device__ __forceinline__ float sgd_update_g(float* p, float gi, float* buf, int64_t i, float lr, float mom,
                                              float wd) {
  const float d = fmaf(wd, p[i], gi);
  float b = d;
  if (buf) { b = fmaf(mom, buf[i], d); buf[i] = b; }
  const float w = p[i] - lr * b;
  p[i] = w;
  return w;
}

__device__ __forceinline__ float sgd_update(float* p, const float* g, float* buf, int64_t i, float lr, float mom,
                                            float wd) {
  const float d = fmaf(wd, p[i], g[i]);
  float b = d;
  if (buf) { b = fmaf(mom, buf[i], d); buf[i] = b; }
  const float w = p[i] - lr * b;
  p[i] = w;
  return w;
}

__global__ void __launch_bounds__(256)
sgd_pack_multi_kernel(const int64_t* __restrict__ items, const float* __restrict__ hyper) {
  __shared__ float tile[32][32 * PK_T_MAX + 1];
  const int64_t* it = items + (int64_t)blockIdx.x * 16;
  float* w = reinterpret_cast<float*>(it[0]);
  const float* g = reinterpret_cast<const float*>(it[1]);
  float* buf = reinterpret_cast<float*>(it[2]);
  uint16_t* dst_f = reinterpret_cast<uint16_t*>(it[3]);     // fprop pack: IEEE half
  uint16_t* dst_d = reinterpret_cast<uint16_t*>(it[4]);     // dgrad pack: bfloat16
  const int Cout = (int)it[5], Cin = (int)it[6], R = (int)it[7], S = (int)it[8];
  const int kind_f = (int)it[9], kind_d = (int)it[10], cs_f = (int)it[11], cs_d = (int)it[12];
  const int64_t numel = it[13];
  // gradient still in split form (mcd_conv2d_wgrad with dw_oihw == NULL): fp32 [ksplit][T][CoutP][CinP] partial sums
  const float* gws = reinterpret_cast<const float*>(it[14]);
  const int ksplit = (int)(it[15] & 0xFFFF), CoutP = (int)((it[15] >> 16) & 0xFFFF), CinP = (int)((it[15] >> 32) & 0xFFFF);
  const float lr = hyper[0], mom = hyper[1], wd = hyper[2];
  const int T = R * S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool any_pack = dst_f || dst_d;
  if (!any_pack) {                       // BatchNorm affine parameters, biases, ...: element-wise by all blocks
    for (int64_t i = blockIdx.y * (int64_t)blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.y * blockDim.x)
      sgd_update(w, g, buf, i, lr, mom, wd);
    return;
  }
  const bool tiled = T <= PK_T_MAX && (!dst_f || kind_f == 0) && (!dst_d || kind_d == 0);
  if (tiled) {
    // one 32(co) x 32(ci) x T tile per iteration: update while loading, then both packs from shared memory
    const int tiles_ci = (Cin + 31) / 32, tiles_co = (Cout + 31) / 32;
    const int kpf = (Cin + 63) / 64 * 64, kpd = (Cout + 63) / 64 * 64;
    for (int tl = blockIdx.y; tl < tiles_ci * tiles_co; tl += gridDim.y) {
      const int co0 = (tl / tiles_ci) * 32, ci0 = (tl % tiles_ci) * 32;
      const int nci = min(32, Cin - ci0), nco = min(32, Cout - co0);
      __syncthreads();
      if (gws) {
        // reduce the split partial sums of this tile (coalesced along ci) into shared memory, OIHW order
        const int64_t kstride = (int64_t)T * CoutP * CinP;
#pragma unroll 4
        for (int idx = threadIdx.x; idx < T * 32 * 32; idx += 256) {
          const int ci_l = idx & 31, co_l = (idx >> 5) & 31, t = idx >> 10;
          if (ci_l < nci && co_l < nco) {
            const float* q = gws + ((int64_t)t * CoutP + co0 + co_l) * CinP + ci0 + ci_l;
            float acc = 0.f;
#pragma unroll 4
            for (int k = 0; k < ksplit; ++k) acc += __ldcs(q + k * kstride);
            tile[co_l][ci_l * T + t] = acc;
          }
        }
        __syncthreads();
        // (loops below: a warp walks rows, lanes the contiguous dimension - no per-element integer divisions)
        for (int co_l = warp; co_l < nco; co_l += 8) {
          const int64_t base = ((int64_t)(co0 + co_l) * Cin + ci0) * T;
          for (int rem = lane; rem < nci * T; rem += 32)
            tile[co_l][rem] = sgd_update_g(w, tile[co_l][rem], buf, base + rem, lr, mom, wd);
        }
      } else {
        for (int co_l = warp; co_l < nco; co_l += 8) {
          const int64_t base = ((int64_t)(co0 + co_l) * Cin + ci0) * T;
          for (int rem = lane; rem < nci * T; rem += 32)
            tile[co_l][rem] = sgd_update(w, g, buf, base + rem, lr, mom, wd);
        }
      }
      __syncthreads();
      if (dst_f && lane < nci) {          // rows (co_l, t); lane = ci_l
        int co_l = warp / T, t = warp % T;
        for (int row = warp; row < nco * T; row += 8) {
          dst_f[((int64_t)(co0 + co_l) * T + t) * kpf + ci0 + lane] = f2bits16(tile[co_l][lane * T + t], kF16);
          t += 8;
          while (t >= T) { t -= T; ++co_l; }
        }
      }
      if (dst_d && lane < nco) {          // rows (ci_l, t); lane = co_l
        int ci_l = warp / T, t = warp % T;
        for (int row = warp; row < nci * T; row += 8) {
          const int tf = (R - 1 - t / S) * S + (S - 1 - t % S);
          dst_d[((int64_t)(ci0 + ci_l) * T + tf) * kpd + co0 + lane] = f2bits16(tile[lane][ci_l * T + t], kBF16);
          t += 8;
          while (t >= T) { t -= T; ++ci_l; }
        }
      }
    }
    return;
  }
  // thin stem layers (row-packed / row-convolution packs) and large filters: tiny tensors, one block does it all
  if (blockIdx.y != 0) return;
  for (int64_t i = threadIdx.x; i < numel; i += blockDim.x) sgd_update(w, g, buf, i, lr, mom, wd);
  __syncthreads();
  for (int pass = 0; pass < 2; ++pass) {
    uint16_t* dst = pass ? dst_d : dst_f;
    const int fmt = pass ? kBF16 : kF16;
    const int kind = pass ? kind_d : kind_f, Cs = pass ? cs_d : cs_f;
    if (!dst) continue;
    const int rows = pass ? Cin : Cout;
    if (kind == 0) {
      const int kc_pad = ((pass ? Cout : Cin) + 63) / 64 * 64;
      const int64_t total = (int64_t)rows * T * kc_pad;
      for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
        const int kc = (int)(i % kc_pad), t = (int)((i / kc_pad) % T), row = (int)(i / ((int64_t)kc_pad * T));
        float v = 0.f;
        if (!pass) { if (kc < Cin) v = w[(((int64_t)row * Cin + kc) * R + t / S) * S + t % S]; }
        else if (kc < Cout) v = w[(((int64_t)kc * Cin + row) * R + (R - 1 - t / S)) * S + (S - 1 - t % S)];
        dst[i] = f2bits16(v, fmt);
      }
    } else if (kind == 2) {
      const int HC = Cs / 8, SP = S <= 4 ? 4 : 8, NB = (rows + 15) / 16 * 16;
      const int64_t total = (int64_t)R * HC * SP * NB * 8;
      for (int64_t i = threadIdx.x; i < total; i += blockDim.x)
        dst[i] = f2bits16(rowconv_pack_value(w, i, Cout, Cin, R, S, HC, SP, NB, pass), fmt);
    } else {
      const int64_t total = (int64_t)rows * R * 64;
      for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
        const int k = (int)(i % 64), r = (int)((i / 64) % R), row = (int)(i / (64 * (int64_t)R));
        const int s = k / Cs, c = k % Cs;
        float v = 0.f;
        if (s < S) {
          if (!pass) { if (c < Cin) v = w[(((int64_t)row * Cin + c) * R + r) * S + s]; }
          else if (c < Cout) v = w[(((int64_t)c * Cin + row) * R + (R - 1 - r)) * S + (S - 1 - s)];
        }
        dst[i] = f2bits16(v, fmt);
      }
    }
  }
}

}  // namespace mcd

using namespace mcd;

extern "C" {

int mcd_nchw_f32_to_nhwc(const float* src, void* dst_f16, void* dst_bf16, int N, int C, int H, int W, int Cs,
                         int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(src && (dst_f16 || dst_bf16) && N > 0 && C > 0 && H > 0 && W > 0, "nchw->nhwc: bad arguments");
  MCD_REQUIRE(Cs >= C && Cs % 8 == 0, "nchw->nhwc: channel stride %d must be >= C=%d and %% 8 == 0",
              Cs, C);
  int64_t total = (int64_t)N * H * W * (Cs / 8);
  int grid = (int)min64((total + 255) / 256, 148 * 16);
  nchw_f32_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      src, (__nv_bfloat16*)dst_f16, (__nv_bfloat16*)dst_bf16, N, C, H * W, Cs);
  return check_launch("nchw_f32_to_nhwc");
}

int mcd_nhwc_to_nchw_f32(const void* src, int src_fmt, float* dst, int N, int C, int H, int W, int Cs,
                         int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0, "nhwc->nchw: bad arguments");
  MCD_REQUIRE(src_fmt == MCD_FMT_F16 || src_fmt == MCD_FMT_BF16, "nhwc->nchw: bad source format %d", src_fmt);
  MCD_REQUIRE(Cs >= C && Cs % 8 == 0, "nhwc->nchw: channel stride %d must be >= C=%d and %% 8 == 0",
              Cs, C);
  int64_t total = (int64_t)N * H * W * ((C + 7) / 8);
  int grid = (int)min64((total + 255) / 256, 148 * 16);
  nhwc_to_nchw_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)src, src_fmt, dst, N, C, H * W, Cs);
  return check_launch("nhwc_to_nchw_f32");
}

int mcd_convert16(const void* src, int src_fmt, void* dst, int64_t numel, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(src && dst && numel > 0 && numel % 8 == 0, "convert16: bad arguments");
  MCD_REQUIRE(src_fmt == MCD_FMT_F16 || src_fmt == MCD_FMT_BF16, "convert16: bad source format %d", src_fmt);
  const int64_t n8 = numel / 8;
  int grid = (int)min64((n8 + 255) / 256, 148 * 16);
  convert16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, (uint4*)dst, src_fmt, n8);
  return check_launch("convert16");
}

int mcd_pack_weight(const float* w_oihw, void* dst, int Cout, int Cin, int R, int S, int mode,
                    int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(w_oihw && dst && Cout > 0 && Cin > 0 && R > 0 && S > 0, "pack_weight: bad arguments");
  MCD_REQUIRE(mode == 0 || mode == 1, "pack_weight: mode must be 0 (fprop) or 1 (dgrad)");
  int rows = mode ? Cin : Cout;
  int kc_pad = round_up(mode ? Cout : Cin, 64);
  int64_t total = (int64_t)rows * R * S * kc_pad;
  int grid = (int)min64((total + 255) / 256, 148 * 16);
  pack_weight_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w_oihw, (uint16_t*)dst, Cout, Cin,
                                                              R, S, mode, rows, kc_pad);
  return check_launch("pack_weight");
}

int mcd_pack_weight_rows(const float* w_oihw, void* dst, int Cout, int Cin, int R, int S, int Cs,
                         int mode, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(w_oihw && dst && Cout > 0 && Cin > 0 && R > 0 && S > 0, "pack_weight_rows: bad arguments");
  MCD_REQUIRE(mode == 0 || mode == 1, "pack_weight_rows: mode must be 0 (fprop) or 1 (dgrad)");
  MCD_REQUIRE((Cs == 8 || Cs == 16) && S * Cs <= 64 && Cs >= (mode ? Cout : Cin),
              "pack_weight_rows: channel stride %d / S=%d not packable", Cs, S);
  int rows = mode ? Cin : Cout;
  int64_t total = (int64_t)rows * R * 64;
  int grid = (int)min64((total + 255) / 256, 148 * 16);
  pack_weight_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w_oihw, (uint16_t*)dst, Cout, Cin,
                                                                   R, S, Cs, mode, rows);
  return check_launch("pack_weight_rows");
}

int mcd_sgd_pack_multi(const int64_t* items_dev, int n_items, const float* hyper_dev, int blocks_per_item,
                       int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(items_dev && hyper_dev && n_items > 0 && blocks_per_item > 0 && blocks_per_item <= 65535,
              "sgd_pack_multi: bad arguments");
  dim3 grid((unsigned)n_items, (unsigned)blocks_per_item);
  sgd_pack_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(items_dev, hyper_dev);
  return check_launch("sgd_pack_multi");
}

int mcd_pack_weights_multi(const int64_t* items_dev, int n_items, int blocks_per_item, int device,
                           void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(items_dev && n_items > 0 && blocks_per_item > 0 && blocks_per_item <= 65535,
              "pack_weights_multi: bad arguments");
  dim3 grid((unsigned)n_items, (unsigned)blocks_per_item);
  pack_weights_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(items_dev);
  return check_launch("pack_weights_multi");
}

}  // extern "C"
