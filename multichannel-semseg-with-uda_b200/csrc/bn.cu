// bn.cu — BatchNorm2d (+ReLU, +residual) forward / backward on NHWC 16-bit activations.
// Formats (common.cuh): forward tensors y / res / z are IEEE half (z additionally gets a bf16 twin `zb` for the
// weight-gradient GEMM of its consumer and as the autograd-visible handle); gradients dz / dy / dres are bf16; the
// ReLU mask is read from the bf16 twin (only its sign matters).
// Reference semantics: nn.BatchNorm2d(eps=1e-5, momentum=0.1, affine, track_running_stats) as used by
// models/drn.py:34-59 (BasicBlock), :129-131, :199-204 and dilated_fcn.py:632-644 (CBR); eval / --fix_bn
// (models/model_util.py:305-310) uses the running statistics.
// All kernels are single-pass, 16-byte vectorised and HBM-bound.
#include "common.cuh"
#include "umma_ptx.cuh"

namespace mcd {

// Thread layout shared by all kernels: a thread owns ONE 16-byte channel vector (8 channels) and walks over
// pixel rows, so per-channel coefficients live in registers and a warp reads 512 contiguous bytes.
//   vpr = Cs/8 vectors per pixel row, rpi = 256/vpr pixel rows per block iteration.

// ---- per-channel reductions ------------------------------------------------------------------
// MODE 0: stats      : acc0 = sum y, acc1 = sum y^2
// MODE 1: bwd reduce : g = dz * (z > 0 | !relu); acc0 = sum g, acc1 = sum g*xhat(y), acc2 = sum g*xhat(res)
// register budgets: the two-rows-in-flight loops need ~90 (apply) / ~110 (backward reduce) registers; tighter launch
// bounds made them spill inside the streaming loop
template <int MODE>
__global__ void __launch_bounds__(256, MODE == 1 ? 2 : 3)
bn_reduce_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dz,
                 const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ res,
                 const float* __restrict__ mean, const float* __restrict__ rstd,
                 const float* __restrict__ res_mean, const float* __restrict__ res_rstd, int relu,
                 float* __restrict__ out, int64_t P, int C, int Cs) {
  extern __shared__ float sh[];  // 3 * Cs accumulators + 4 * Cs coefficients (mean, rstd, res_mean, res_rstd)
  float* co = sh + 3 * Cs;
  const int vpr = Cs >> 3;
  const int rpi = 256 / vpr;
  const int cv = threadIdx.x % vpr, pr = threadIdx.x / vpr;
  const bool active = pr < rpi;
  const bool dual = MODE == 1 && res_mean != nullptr;
  for (int i = threadIdx.x; i < 3 * Cs; i += 256) sh[i] = 0.f;
  if (MODE == 1) {
    for (int c = threadIdx.x; c < Cs; c += 256) {
      const int cc = min(c, C - 1);
      co[c] = mean[cc]; co[Cs + c] = rstd[cc];
      co[2 * Cs + c] = dual ? res_mean[cc] : 0.f; co[3 * Cs + c] = dual ? res_rstd[cc] : 0.f;
    }
  }
  __syncthreads();
  float a0[8], a1[8], a2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a0[k] = a1[k] = a2[k] = 0.f;
  const int c0 = cv * 8;
  if (active) {
    const int64_t step = (int64_t)gridDim.x * rpi;
    constexpr int U = 2;    // pixel rows per iteration, all loads first
    for (int64_t p0 = (int64_t)blockIdx.x * rpi + pr; p0 < P; p0 += U * step) {
      uint4 vyu[U], vdu[U], vzu[U], vru[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t pu = p0 + u * step;
        vyu[u] = vdu[u] = vzu[u] = vru[u] = make_uint4(0, 0, 0, 0);   // a zero row adds nothing to any sum
        if (pu < P) {
          const int64_t off = pu * Cs + c0;
          vyu[u] = *reinterpret_cast<const uint4*>(y + off);
          if (MODE == 1) {
            vdu[u] = *reinterpret_cast<const uint4*>(dz + off);
            if (relu) vzu[u] = *reinterpret_cast<const uint4*>(z + off);
            if (dual) vru[u] = *reinterpret_cast<const uint4*>(res + off);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
      float fy[8];
      const uint4 vy = vyu[u];
      if (MODE == 0) {
        unpack8h(vy, fy);
#pragma unroll
        for (int k = 0; k < 8; ++k) { a0[k] += fy[k]; a1[k] = fmaf(fy[k], fy[k], a1[k]); }
      } else {
        const uint4 vd = vdu[u], vz = vzu[u], vr = vru[u];
        float fd[8], fz[8];
        unpack8h(vy, fy); unpack8(vd, fd); unpack8(vz, fz);
        float g[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          g[k] = (!relu || fz[k] > 0.f) ? fd[k] : 0.f;
          a0[k] += g[k];
          a1[k] = fmaf(g[k], (fy[k] - co[c0 + k]) * co[Cs + c0 + k], a1[k]);
        }
        if (dual) {
          float fr[8];
          unpack8h(vr, fr);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            a2[k] = fmaf(g[k], (fr[k] - co[2 * Cs + c0 + k]) * co[3 * Cs + c0 + k], a2[k]);
        }
      }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(&sh[c0 + k], a0[k]);
      atomicAdd(&sh[Cs + c0 + k], a1[k]);
      if (dual) atomicAdd(&sh[2 * Cs + c0 + k], a2[k]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    atomicAdd(out + c, sh[c]);
    atomicAdd(out + C + c, sh[Cs + c]);
    if (dual) atomicAdd(out + 2 * C + c, sh[2 * Cs + c]);
  }
}

// dx = mask_src > 0 ? dx : 0 (in place, optional) and sums += {sum dx, sum dx * bn_y} (optional, RAW second moment):
// the un-fused form of the dgrad epilogue extras (conv_plan.h EpiExtra) for paths that cannot fuse them.
__global__ void __launch_bounds__(256, 3)
bn_mask_sums_kernel(__nv_bfloat16* __restrict__ dx, const __nv_bfloat16* __restrict__ mask_src,
                    const __nv_bfloat16* __restrict__ bn_y, float* __restrict__ sums, int64_t P, int C, int Cs) {
  extern __shared__ float sh[];  // 2 * Cs
  const int vpr = Cs >> 3;
  const int rpi = 256 / vpr;
  const int cv = threadIdx.x % vpr, pr = threadIdx.x / vpr;
  for (int i = threadIdx.x; i < 2 * Cs; i += 256) sh[i] = 0.f;
  __syncthreads();
  float a0[8], a1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a0[k] = a1[k] = 0.f;
  const int c0 = cv * 8;
  if (pr < rpi) {
    const int64_t step = (int64_t)gridDim.x * rpi;
    for (int64_t p = (int64_t)blockIdx.x * rpi + pr; p < P; p += step) {
      const int64_t off = p * Cs + c0;
      float g[8];
      unpack8(*reinterpret_cast<const uint4*>(dx + off), g);
      if (mask_src) {
        float m[8];
        unpack8(*reinterpret_cast<const uint4*>(mask_src + off), m);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = m[k] > 0.f ? g[k] : 0.f;
        *reinterpret_cast<uint4*>(dx + off) = pack8(g);
      }
      if (sums) {
        float fy[8];
        unpack8h(*reinterpret_cast<const uint4*>(bn_y + off), fy);
#pragma unroll
        for (int k = 0; k < 8; ++k) { a0[k] += g[k]; a1[k] = fmaf(g[k], fy[k], a1[k]); }
      }
    }
    if (sums) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { atomicAdd(&sh[c0 + k], a0[k]); atomicAdd(&sh[Cs + c0 + k], a1[k]); }
    }
  }
  __syncthreads();
  if (sums)
    for (int c = threadIdx.x; c < C; c += 256) { atomicAdd(sums + c, sh[c]); atomicAdd(sums + C + c, sh[Cs + c]); }
}

__global__ void bn_finalize_kernel(const float* __restrict__ stats, double invP, double unbias,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float momentum, float eps, int training, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ save_mean,
                                   float* __restrict__ save_rstd, int64_t* __restrict__ nbt,
                                   int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (c == 0 && training && nbt) *nbt += training;   // training = number of (identical) batches folded in
  float mean, var;
  if (training) {
    double m = (double)stats[c] * invP;
    double v = (double)stats[C + c] * invP - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m; var = (float)v;
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(v * unbias);
    }
  } else {
    mean = running_mean[c]; var = running_var[c];
  }
  float rstd = rsqrtf(var + eps);
  float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * rstd;
  shift[c] = b - mean * g * rstd;
  save_mean[c] = mean;
  save_rstd[c] = rstd;
}

__global__ void __launch_bounds__(256, 4)
bn_apply_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                const float* __restrict__ shift, const __nv_bfloat16* __restrict__ res,
                const float* __restrict__ rscale, const float* __restrict__ rshift, int relu,
                __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ zb, int64_t P, int Cs) {
  extern __shared__ float co[];  // scale, shift, rscale, rshift : 4 * Cs
  for (int c = threadIdx.x; c < Cs; c += 256) {
    co[c] = scale[c]; co[Cs + c] = shift[c];
    co[2 * Cs + c] = rscale ? rscale[c] : 1.f; co[3 * Cs + c] = rscale ? rshift[c] : 0.f;
  }
  __syncthreads();
  const int vpr = Cs >> 3;
  const int rpi = 256 / vpr;
  const int cv = threadIdx.x % vpr, pr = threadIdx.x / vpr;
  if (pr >= rpi) return;
  const int c0 = cv * 8;
  const int64_t step = (int64_t)gridDim.x * rpi;
  // channel coefficients of this thread's 8 channels live in registers; U pixel rows per iteration with all loads
  // issued before the arithmetic (memory-level parallelism: 2-4 x 16 B in flight per thread)
  float sc[8], sf[8], rc[8], rf[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = co[c0 + k]; sf[k] = co[Cs + c0 + k]; rc[k] = co[2 * Cs + c0 + k]; rf[k] = co[3 * Cs + c0 + k];
  }
  constexpr int U = 2;
  for (int64_t p = (int64_t)blockIdx.x * rpi + pr; p < P; p += U * step) {
    uint4 vy[U], vr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pu = p + u * step;
      vy[u] = make_uint4(0, 0, 0, 0); vr[u] = make_uint4(0, 0, 0, 0);
      if (pu < P) {
        vy[u] = __ldcs(reinterpret_cast<const uint4*>(y + pu * Cs + c0));
        if (res) vr[u] = __ldcs(reinterpret_cast<const uint4*>(res + pu * Cs + c0));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pu = p + u * step;
      if (pu >= P) break;
      float f[8];
      unpack8h(vy[u], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sf[k]);
      if (res) {
        float r[8];
        unpack8h(vr[u], r);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] += fmaf(r[k], rc[k], rf[k]);
      }
      if (relu) {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
      }
      if (z) *reinterpret_cast<uint4*>(z + pu * Cs + c0) = pack8h(f);          // IEEE half: the next fprop's operand
      if (zb) *reinterpret_cast<uint4*>(zb + pu * Cs + c0) = pack8(f);         // bf16 twin: wgrad operand / mask
    }
  }
}

// ---- fused forward: finalize (statistics -> scale/shift, running-stat update) + apply in ONE launch --------
struct BnParams {
  const float* stats;          // [2C] sum / sum of squares (training) or NULL
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; int64_t* nbt;
  float* save;                 // [2C] out: mean, rstd (consumed by the backward kernels)
  float momentum, eps;
  int training;                // 0 eval, k >= 1: k identical batches folded (momentum already adjusted)
};

__device__ __forceinline__ void bn_coeffs(const BnParams& b, int c, int C, double invP, double unbias,
                                          bool writer, float* scale, float* shift) {
  float mean, var;
  if (b.training) {
    double m = (double)b.stats[c] * invP;
    double v = (double)b.stats[C + c] * invP - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m; var = (float)v;
    if (writer && b.running_mean) {
      b.running_mean[c] = (1.f - b.momentum) * b.running_mean[c] + b.momentum * mean;
      b.running_var[c] = (1.f - b.momentum) * b.running_var[c] + b.momentum * (float)(v * unbias);
    }
  } else {
    mean = b.running_mean[c]; var = b.running_var[c];
  }
  const float rstd = rsqrtf(var + b.eps);
  const float g = b.gamma ? b.gamma[c] : 1.f, be = b.beta ? b.beta[c] : 0.f;
  *scale = g * rstd;
  *shift = be - mean * g * rstd;
  if (writer) { b.save[c] = mean; b.save[C + c] = rstd; }
}

// (256, 4): measured - three blocks per SM (no spill) was 21 % slower than four with a 12-byte spill
__global__ void __launch_bounds__(256, 4)
bn_forward_kernel(const __nv_bfloat16* __restrict__ y, BnParams b1, const __nv_bfloat16* __restrict__ res,
                  BnParams b2, int has_res_bn, int relu, __nv_bfloat16* __restrict__ z,
                  __nv_bfloat16* __restrict__ zb, int64_t P, int C, int Cs, double invP, double unbias) {
  extern __shared__ float co[];  // scale, shift, rscale, rshift : 4 * Cs
  const bool writer = blockIdx.x == 0;   // exactly one block updates running stats / saves mean, rstd
  for (int c = threadIdx.x; c < Cs; c += 256) {
    float sc = 0.f, sf = 0.f, rc = 1.f, rf = 0.f;
    if (c < C) {
      bn_coeffs(b1, c, C, invP, unbias, writer, &sc, &sf);
      if (has_res_bn) bn_coeffs(b2, c, C, invP, unbias, writer, &rc, &rf);
    }
    co[c] = sc; co[Cs + c] = sf; co[2 * Cs + c] = rc; co[3 * Cs + c] = rf;
  }
  if (writer && threadIdx.x == 0) {
    if (b1.training && b1.nbt) *b1.nbt += b1.training;
    if (has_res_bn && b2.training && b2.nbt) *b2.nbt += b2.training;
  }
  __syncthreads();
  const int vpr = Cs >> 3;
  const int rpi = 256 / vpr;
  const int cv = threadIdx.x % vpr, pr = threadIdx.x / vpr;
  if (pr >= rpi) return;
  const int c0 = cv * 8;
  const int64_t step = (int64_t)gridDim.x * rpi;
  // channel coefficients of this thread's 8 channels live in registers; U pixel rows per iteration with all loads
  // issued before the arithmetic (memory-level parallelism: 2-4 x 16 B in flight per thread)
  float sc[8], sf[8], rc[8], rf[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = co[c0 + k]; sf[k] = co[Cs + c0 + k]; rc[k] = co[2 * Cs + c0 + k]; rf[k] = co[3 * Cs + c0 + k];
  }
  constexpr int U = 2;
  for (int64_t p = (int64_t)blockIdx.x * rpi + pr; p < P; p += U * step) {
    uint4 vy[U], vr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pu = p + u * step;
      vy[u] = make_uint4(0, 0, 0, 0); vr[u] = make_uint4(0, 0, 0, 0);
      if (pu < P) {
        vy[u] = __ldcs(reinterpret_cast<const uint4*>(y + pu * Cs + c0));
        if (res) vr[u] = __ldcs(reinterpret_cast<const uint4*>(res + pu * Cs + c0));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pu = p + u * step;
      if (pu >= P) break;
      float f[8];
      unpack8h(vy[u], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sf[k]);
      if (res) {
        float r[8];
        unpack8h(vr[u], r);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] += fmaf(r[k], rc[k], rf[k]);
      }
      if (relu) {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
      }
      if (z) *reinterpret_cast<uint4*>(z + pu * Cs + c0) = pack8h(f);          // IEEE half: the next fprop's operand
      if (zb) *reinterpret_cast<uint4*>(zb + pu * Cs + c0) = pack8(f);         // bf16 twin: wgrad operand / mask
    }
  }
}

// ---- bulk-copy staged variant of the forward (r02) --------------------------------------------------------------
// The register-staged kernel above keeps 2 x 16 B per tensor and thread in flight: 32-64 KB per SM, which is what
// bounds it at ~4.8 TB/s (60 % of the DRAM peak in profiles/r02_ncu_full_kernels.csv; a plain copy reaches 6.5).  Here
// the inputs arrive through cp.async.bulk (the 1-D TMA path, no tensor map: a tile of consecutive pixel rows of a dense
// NHWC tensor is one contiguous byte range) into a 3-stage shared-memory ring, so that 2 x 16 KB per tensor and block
// are in flight without costing a register; the arithmetic and the two output stores are unchanged.
constexpr int kBnTileBytes = 16384;      // per tensor and stage
constexpr int kBnStages = 3;

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ptx::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar)) : "memory");
}

template <bool HAS_RES>
__global__ void __launch_bounds__(256, 2)
bn_forward_bulk_kernel(const __nv_bfloat16* __restrict__ y, BnParams b1, const __nv_bfloat16* __restrict__ res,
                       BnParams b2, int has_res_bn, int relu, __nv_bfloat16* __restrict__ z,
                       __nv_bfloat16* __restrict__ zb, int64_t P, int C, int Cs, double invP, double unbias,
                       int rows_per_tile, int ntiles) {
  extern __shared__ __align__(128) uint8_t bsm[];
  constexpr int NT = HAS_RES ? 2 : 1;
  uint8_t* ring = bsm;                                                  // [stage][tensor][kBnTileBytes]
  float* co = reinterpret_cast<float*>(bsm + kBnStages * NT * kBnTileBytes);   // scale, shift, rscale, rshift
  uint64_t* full = reinterpret_cast<uint64_t*>(co + 4 * Cs);
  const int tid = threadIdx.x;
  const bool writer = blockIdx.x == 0;
  if (tid == 0) {
    for (int s = 0; s < kBnStages; ++s) ptx::mbar_init(&full[s], 1);
    ptx::fence_barrier_init();
  }
  for (int c = tid; c < Cs; c += 256) {
    float sc = 0.f, sf = 0.f, rc = 1.f, rf = 0.f;
    if (c < C) {
      bn_coeffs(b1, c, C, invP, unbias, writer, &sc, &sf);
      if (has_res_bn) bn_coeffs(b2, c, C, invP, unbias, writer, &rc, &rf);
    }
    co[c] = sc; co[Cs + c] = sf; co[2 * Cs + c] = rc; co[3 * Cs + c] = rf;
  }
  if (writer && tid == 0) {
    if (b1.training && b1.nbt) *b1.nbt += b1.training;
    if (has_res_bn && b2.training && b2.nbt) *b2.nbt += b2.training;
  }
  __syncthreads();
  const int64_t row_bytes = (int64_t)Cs * 2;
  auto issue = [&](int tile, int stage) {       // thread 0: one bulk copy per tensor
    const int64_t r0 = (int64_t)tile * rows_per_tile;
    const int64_t left = P - r0;
    const uint32_t bytes = (uint32_t)((left < rows_per_tile ? left : (int64_t)rows_per_tile) * row_bytes);
    ptx::mbar_expect_tx(&full[stage], bytes * NT);
    bulk_load_1d(ring + (stage * NT) * kBnTileBytes, reinterpret_cast<const uint8_t*>(y) + r0 * row_bytes, bytes, &full[stage]);
    if (HAS_RES)
      bulk_load_1d(ring + (stage * NT + 1) * kBnTileBytes, reinterpret_cast<const uint8_t*>(res) + r0 * row_bytes, bytes,
                   &full[stage]);
  };
  if (tid == 0)
    for (int s = 0; s < kBnStages; ++s) {
      const int t = blockIdx.x + s * gridDim.x;
      if (t < ntiles) issue(t, s);
    }
  const int vpr = Cs >> 3, rpi = 256 / vpr;        // vpr divides 256 (host check): the channel vector of a thread is fixed
  const int cv = tid % vpr, pr = tid / vpr, c0 = cv * 8;
  float sc[8], sf[8], rc[8], rf[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = co[c0 + k]; sf[k] = co[Cs + c0 + k]; rc[k] = co[2 * Cs + c0 + k]; rf[k] = co[3 * Cs + c0 + k];
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int stage = it % kBnStages;
    const int64_t r0 = (int64_t)tile * rows_per_tile;
    const int nrows = (P - r0 < rows_per_tile) ? (int)(P - r0) : rows_per_tile;
    ptx::mbar_wait(&full[stage], (uint32_t)((it / kBnStages) & 1));
    const uint4* sy = reinterpret_cast<const uint4*>(ring + (stage * NT) * kBnTileBytes);
    const uint4* sr = reinterpret_cast<const uint4*>(ring + (stage * NT + 1) * kBnTileBytes);
#pragma unroll 2
    for (int r = pr; r < nrows; r += rpi) {
      float f[8];
      unpack8h(sy[r * vpr + cv], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sf[k]);
      if (HAS_RES) {
        float q[8];
        unpack8h(sr[r * vpr + cv], q);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] += fmaf(q[k], rc[k], rf[k]);
      }
      if (relu) {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
      }
      const int64_t off = (r0 + r) * Cs + c0;
      if (z) *reinterpret_cast<uint4*>(z + off) = pack8h(f);
      if (zb) *reinterpret_cast<uint4*>(zb + off) = pack8(f);
    }
    __syncthreads();                                // every thread is done with this stage: refill it
    const int nt = tile + kBnStages * gridDim.x;
    if (tid == 0 && nt < ntiles) issue(nt, stage);
  }
}

struct BwdBranch {
  const float* gamma; const float* mean; const float* rstd;
  int training;
};

// second moment of the gradient: sums hold sum(g * xhat) (kind 0) or the raw sum(g * y) (kind 1, produced by the
// fused dgrad epilogue): sum(g * xhat) = rstd * (sum(g * y) - mean * sum(g))
__device__ __forceinline__ float bwd_second(const BwdBranch& b, const float* sums, int sum_off, int c, int raw) {
  const float s = sums[sum_off + c];
  return raw ? b.rstd[c] * (s - b.mean[c] * sums[c]) : s;
}

// dy = A*g + B*y + K with per-channel A = gamma*rstd, B = -A*m2*rstd, K = -A*m1 - B*mean  (training;
// m1 = sum(g)/P, m2 = sum(g*xhat)/P) or B = K = 0 (eval).
__device__ __forceinline__ void bwd_coeffs(const BwdBranch& b, const float* sums, int sum_off, int C, int c,
                                           float invP, int raw, float* A, float* B, float* K) {
  const float a = b.gamma[c] * b.rstd[c];
  *A = a;
  if (b.training) {
    const float m1 = sums[c] * invP, m2 = bwd_second(b, sums, sum_off, c, raw) * invP;
    const float bb = -a * m2 * b.rstd[c];
    *B = bb;
    *K = -a * m1 - bb * b.mean[c];
  } else {
    *B = 0.f; *K = 0.f;
  }
}

__global__ void __launch_bounds__(256, 3)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ z,
                    const __nv_bfloat16* __restrict__ y, BwdBranch b1, const float* __restrict__ sums,
                    int relu, __nv_bfloat16* __restrict__ dy, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, const __nv_bfloat16* __restrict__ res, BwdBranch b2,
                    __nv_bfloat16* __restrict__ dres, float* __restrict__ dres_gamma,
                    float* __restrict__ dres_beta, float invP, int64_t P, int Cs, int C, int raw) {
  extern __shared__ float co[];  // A1, B1, K1, A2, B2, K2 : 6 * Cs
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dgamma) dgamma[c] = bwd_second(b1, sums, C, c, raw);
      if (dbeta) dbeta[c] = sums[c];
      if (dres_gamma) dres_gamma[c] = sums[2 * C + c];
      if (dres_beta) dres_beta[c] = sums[c];
    }
  }
  const bool has_res_bn = dres && b2.gamma;
  for (int c = threadIdx.x; c < Cs; c += 256) {
    float A = 0.f, B = 0.f, K = 0.f, A2 = 0.f, B2 = 0.f, K2 = 0.f;
    if (c < C) {
      bwd_coeffs(b1, sums, C, C, c, invP, raw, &A, &B, &K);
      if (has_res_bn) bwd_coeffs(b2, sums, 2 * C, C, c, invP, 0, &A2, &B2, &K2);
    }
    co[c] = A; co[Cs + c] = B; co[2 * Cs + c] = K;
    co[3 * Cs + c] = A2; co[4 * Cs + c] = B2; co[5 * Cs + c] = K2;
  }
  __syncthreads();
  const int vpr = Cs >> 3;
  const int rpi = 256 / vpr;
  const int cv = threadIdx.x % vpr, pr = threadIdx.x / vpr;
  if (pr >= rpi) return;
  const int c0 = cv * 8;
  const int64_t step = (int64_t)gridDim.x * rpi;
  float cA[8], cB[8], cK[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { cA[k] = co[c0 + k]; cB[k] = co[Cs + c0 + k]; cK[k] = co[2 * Cs + c0 + k]; }
  constexpr int U = 2;      // pixel rows per iteration, all loads first
  for (int64_t p = (int64_t)blockIdx.x * rpi + pr; p < P; p += U * step) {
    uint4 vd[U], vy[U], vz[U], vr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pu = p + u * step;
      vd[u] = vy[u] = vz[u] = vr[u] = make_uint4(0, 0, 0, 0);
      if (pu < P) {
        const int64_t off = pu * Cs + c0;
        vd[u] = __ldcs(reinterpret_cast<const uint4*>(dz + off));
        vy[u] = __ldcs(reinterpret_cast<const uint4*>(y + off));
        if (relu) vz[u] = __ldcs(reinterpret_cast<const uint4*>(z + off));
        if (has_res_bn) vr[u] = __ldcs(reinterpret_cast<const uint4*>(res + off));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pu = p + u * step;
      if (pu >= P) break;
      const int64_t off = pu * Cs + c0;
      float fd[8], fz[8], fy[8], g[8], o[8];
      unpack8(vd[u], fd); unpack8h(vy[u], fy); unpack8(vz[u], fz);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        g[k] = (!relu || fz[k] > 0.f) ? fd[k] : 0.f;
        o[k] = fmaf(cA[k], g[k], fmaf(cB[k], fy[k], cK[k]));
      }
      *reinterpret_cast<uint4*>(dy + off) = pack8(o);
      if (dres) {
        if (has_res_bn) {
          float fr[8];
          unpack8h(vr[u], fr);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            o[k] = fmaf(co[3 * Cs + c0 + k], g[k], fmaf(co[4 * Cs + c0 + k], fr[k], co[5 * Cs + c0 + k]));
          *reinterpret_cast<uint4*>(dres + off) = pack8(o);
        } else {
          *reinterpret_cast<uint4*>(dres + off) = pack8(g);
        }
      }
    }
  }
}

// MCD_BN_BULK=0: the register-staged BatchNorm forward everywhere (A/B measurements).
static bool bn_bulk_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MCD_BN_BULK"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
// (bulk-staged BACKWARD kernels were tried and dropped: faster alone - 512 channels 63 -> 58 us - but slower inside the
// iteration, 137.0 -> 138.2 ms: backward BatchNorm kernels run next to the wgrad kernels of the side stream, and a block
// that needs 50 - 100 KB of shared memory cannot share an SM with a 190 KB wgrad CTA the way these register-staged
// blocks do.)

static inline int rows_grid(int64_t P, int Cs, int rows_per_thread, int max_blocks) {
  int rpi = 256 / (Cs / 8);
  int64_t g = (P + (int64_t)rpi * rows_per_thread - 1) / ((int64_t)rpi * rows_per_thread);
  if (g < 1) g = 1;
  return (int)(g < max_blocks ? g : max_blocks);
}

int bn_stats_launch(const void* y, float* stats, int64_t P, int C, int Cs, cudaStream_t st) {
  int grid = rows_grid(P, Cs, 8, 148 * 3);
  // (Bottleneck trunks reach 2048 channels: 7 * 2048 floats = 56 KB of dynamic shared memory, above the 48 KB default)
  if (ensure_dyn_smem<bn_reduce_kernel<0>>(7 * Cs * (int)sizeof(float), "bn_stats") != MCD_OK) return MCD_E_CUDA;
  bn_reduce_kernel<0><<<grid, 256, 7 * Cs * sizeof(float), st>>>(
      (const __nv_bfloat16*)y, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, stats, P,
      C, Cs);
  return check_launch("bn_stats");
}

int bn_mask_sums_launch(void* dx, const void* mask_src, const void* bn_y, float* sums, int64_t P, int C, int Cs,
                        cudaStream_t st) {
  if (Cs % 8 != 0 || Cs > 2048) { set_error("bn_mask_sums: channel stride %d unsupported", Cs); return MCD_E_INVALID; }
  int grid = rows_grid(P, Cs, 8, 148 * 3);
  bn_mask_sums_kernel<<<grid, 256, 2 * Cs * sizeof(float), st>>>((__nv_bfloat16*)dx, (const __nv_bfloat16*)mask_src,
                                                                  (const __nv_bfloat16*)bn_y, sums, P, C, Cs);
  return check_launch("bn_mask_sums");
}

}  // namespace mcd

using namespace mcd;

extern "C" {

int mcd_bn_stats(const void* y_nhwc, float* stats, int64_t P, int C, int Cs, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(y_nhwc && stats && P > 0 && C > 0, "bn_stats: bad arguments");
  MCD_REQUIRE(Cs >= C && Cs % 8 == 0 && Cs <= 2048, "bn_stats: channel stride %d unsupported", Cs);
  return bn_stats_launch(y_nhwc, stats, P, C, Cs, (cudaStream_t)stream);
}

int mcd_bn_finalize(const float* stats, int64_t P, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps,
                    int training, float* scale, float* shift, float* save_mean, float* save_rstd,
                    int64_t* num_batches_tracked, int C, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(scale && shift && save_mean && save_rstd && C > 0, "bn_finalize: bad arguments");
  MCD_REQUIRE(training ? (stats != nullptr && P > 0) : (running_mean && running_var),
              "bn_finalize: missing statistics for mode training=%d", training);
  double invP = training ? 1.0 / (double)P : 0.0;
  double unbias = (training && P > 1) ? (double)P / (double)(P - 1) : 1.0;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      stats, invP, unbias, gamma, beta, running_mean, running_var, momentum, eps, training, scale,
      shift, save_mean, save_rstd, num_batches_tracked, C);
  return check_launch("bn_finalize");
}

int mcd_bn_forward(const void* y_nhwc, const float* stats, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                   float eps, int training, float* save_mean_rstd, const void* res_nhwc,
                   const float* res_stats, const float* res_gamma, const float* res_beta,
                   float* res_running_mean, float* res_running_var, int64_t* res_num_batches_tracked,
                   float res_momentum, float res_eps, int res_training, float* res_save_mean_rstd, int relu,
                   void* z_nhwc, void* zb_nhwc, int64_t P, int C, int Cs, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(y_nhwc && (z_nhwc || zb_nhwc) && save_mean_rstd && P > 0 && C > 0, "bn_forward: bad arguments");
  MCD_REQUIRE(Cs == C && C % 8 == 0 && Cs <= 2048, "bn_forward: needs dense channels, C %% 8 == 0 (C=%d Cs=%d)", C, Cs);
  MCD_REQUIRE(training ? stats != nullptr : (running_mean && running_var), "bn_forward: missing statistics");
  const int has_res_bn = res_save_mean_rstd != nullptr;
  MCD_REQUIRE(!has_res_bn || (res_nhwc && (res_training ? res_stats != nullptr : (res_running_mean && res_running_var))),
              "bn_forward: incomplete residual BatchNorm");
  BnParams b1{stats, gamma, beta, running_mean, running_var, num_batches_tracked, save_mean_rstd, momentum, eps, training};
  BnParams b2{res_stats, res_gamma, res_beta, res_running_mean, res_running_var, res_num_batches_tracked,
              res_save_mean_rstd, res_momentum, res_eps, res_training};
  double invP = 1.0 / (double)P;
  double unbias = P > 1 ? (double)P / (double)(P - 1) : 1.0;
  // large dense tensors with a power-of-two channel count: inputs staged by bulk copies (bn_forward_bulk_kernel)
  const int vpr = Cs / 8;
  if (bn_bulk_enabled() && vpr >= 2 && vpr <= 256 && (256 % vpr) == 0 && P * (int64_t)Cs * 2 >= (8ll << 20)) {
    const int rows_per_tile = max(1, kBnTileBytes / (Cs * 2));
    const int64_t ntiles = (P + rows_per_tile - 1) / rows_per_tile;
    const int nt = res_nhwc ? 2 : 1;
    const int smem = kBnStages * nt * kBnTileBytes + 4 * Cs * (int)sizeof(float) + 64;
    const int grid = (int)min64(ntiles, 148 * 2);
    if (ntiles < (1ll << 30)) {
      if (res_nhwc) {
        if (ensure_dyn_smem<bn_forward_bulk_kernel<true>>(smem, "bn_forward") != MCD_OK) return MCD_E_CUDA;
        bn_forward_bulk_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)y_nhwc, b1, (const __nv_bfloat16*)res_nhwc, b2, has_res_bn, relu,
            (__nv_bfloat16*)z_nhwc, (__nv_bfloat16*)zb_nhwc, P, C, Cs, invP, unbias, rows_per_tile, (int)ntiles);
      } else {
        if (ensure_dyn_smem<bn_forward_bulk_kernel<false>>(smem, "bn_forward") != MCD_OK) return MCD_E_CUDA;
        bn_forward_bulk_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)y_nhwc, b1, nullptr, b2, has_res_bn, relu, (__nv_bfloat16*)z_nhwc,
            (__nv_bfloat16*)zb_nhwc, P, C, Cs, invP, unbias, rows_per_tile, (int)ntiles);
      }
      return check_launch("bn_forward");
    }
  }
  int grid = rows_grid(P, Cs, 4, 148 * 8);
  bn_forward_kernel<<<grid, 256, 4 * Cs * sizeof(float), (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)y_nhwc, b1, (const __nv_bfloat16*)res_nhwc, b2, has_res_bn, relu,
      (__nv_bfloat16*)z_nhwc, (__nv_bfloat16*)zb_nhwc, P, C, Cs, invP, unbias);
  return check_launch("bn_forward");
}

int mcd_bn_apply(const void* y_nhwc, const float* scale, const float* shift, const void* res_nhwc,
                 const float* rscale, const float* rshift, int relu, void* z_nhwc, void* zb_nhwc, int64_t P,
                 int C, int Cs, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(y_nhwc && scale && shift && (z_nhwc || zb_nhwc) && P > 0, "bn_apply: bad arguments");
  MCD_REQUIRE(Cs == C && C % 8 == 0, "bn_apply: needs dense channels, C %% 8 == 0 (C=%d Cs=%d)", C, Cs);
  MCD_REQUIRE(!rscale || (res_nhwc && rshift), "bn_apply: residual affine without residual");
  MCD_REQUIRE(Cs <= 2048, "bn_apply: channel stride %d unsupported", Cs);
  int grid = rows_grid(P, Cs, 4, 148 * 8);
  bn_apply_kernel<<<grid, 256, 4 * Cs * sizeof(float), (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)y_nhwc, scale, shift, (const __nv_bfloat16*)res_nhwc, rscale, rshift, relu,
      (__nv_bfloat16*)z_nhwc, (__nv_bfloat16*)zb_nhwc, P, Cs);
  return check_launch("bn_apply");
}

int mcd_bn_bwd_reduce(const void* dz_nhwc, const void* z_nhwc, const void* y_nhwc,
                      const float* mean, const float* rstd, const void* res_nhwc,
                      const float* res_mean, const float* res_rstd, int relu, float* sums, int64_t P,
                      int C, int Cs, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(dz_nhwc && y_nhwc && mean && rstd && sums && P > 0, "bn_bwd_reduce: bad arguments");
  MCD_REQUIRE(!relu || z_nhwc, "bn_bwd_reduce: relu needs z");
  MCD_REQUIRE(Cs == C && C % 8 == 0 && Cs <= 2048, "bn_bwd_reduce: needs dense channels (C=%d Cs=%d)", C, Cs);
  MCD_REQUIRE(!res_mean || (res_nhwc && res_rstd), "bn_bwd_reduce: residual stats without residual");
  int grid = rows_grid(P, Cs, 8, 148 * 2);     // two resident blocks per SM (launch bounds): one full wave
  if (ensure_dyn_smem<bn_reduce_kernel<1>>(7 * Cs * (int)sizeof(float), "bn_bwd_reduce") != MCD_OK) return MCD_E_CUDA;
  bn_reduce_kernel<1><<<grid, 256, 7 * Cs * sizeof(float), (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)y_nhwc, (const __nv_bfloat16*)dz_nhwc, (const __nv_bfloat16*)z_nhwc,
      (const __nv_bfloat16*)res_nhwc, mean, rstd, res_mean, res_rstd, relu, sums, P, C, Cs);
  return check_launch("bn_bwd_reduce");
}

int mcd_bn_bwd_apply(const void* dz_nhwc, const void* z_nhwc, const void* y_nhwc, const float* gamma,
                     const float* mean, const float* rstd, const float* sums, int training, int relu,
                     void* dy_nhwc, float* dgamma, float* dbeta, const void* res_nhwc,
                     const float* res_gamma, const float* res_mean, const float* res_rstd,
                     int res_training, void* dres_nhwc, float* dres_gamma, float* dres_beta,
                     int sums_kind, int64_t P, int C, int Cs, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(dz_nhwc && y_nhwc && gamma && mean && rstd && sums && dy_nhwc && P > 0,
              "bn_bwd_apply: bad arguments");
  MCD_REQUIRE(!relu || z_nhwc, "bn_bwd_apply: relu needs z");
  MCD_REQUIRE(Cs == C && C % 8 == 0, "bn_bwd_apply: needs dense channels (C=%d Cs=%d)", C, Cs);
  MCD_REQUIRE(!res_gamma || (res_nhwc && res_mean && res_rstd && dres_nhwc),
              "bn_bwd_apply: incomplete residual branch");
  MCD_REQUIRE(sums_kind == 0 || (sums_kind == 1 && !res_gamma),
              "bn_bwd_apply: raw sums (kind 1) are not defined for a residual BatchNorm branch");
  BwdBranch b1{gamma, mean, rstd, training};
  BwdBranch b2{res_gamma, res_mean, res_rstd, res_training};
  MCD_REQUIRE(Cs <= 2048, "bn_bwd_apply: channel stride %d unsupported", Cs);
  int grid = rows_grid(P, Cs, 4, 148 * 6);
  if (ensure_dyn_smem<bn_bwd_apply_kernel>(6 * Cs * (int)sizeof(float), "bn_bwd_apply") != MCD_OK) return MCD_E_CUDA;
  bn_bwd_apply_kernel<<<grid, 256, 6 * Cs * sizeof(float), (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dz_nhwc, (const __nv_bfloat16*)z_nhwc, (const __nv_bfloat16*)y_nhwc, b1, sums,
      relu, (__nv_bfloat16*)dy_nhwc, dgamma, dbeta, (const __nv_bfloat16*)res_nhwc, b2,
      (__nv_bfloat16*)dres_nhwc, dres_gamma, dres_beta, (float)(1.0 / (double)P), P, Cs, C, sums_kind);
  return check_launch("bn_bwd_apply");
}

}  // extern "C"
