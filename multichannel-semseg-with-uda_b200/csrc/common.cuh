// common.cuh — shared helpers for libmcd_sm100 (error plumbing, launch accounting, bf16 utils).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/mcd_sm100.h"

namespace mcd {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }
inline int64_t max64(int64_t a, int64_t b) { return a > b ? a : b; }

// RAII-free guard: select the device for this call; every entry point starts with it.
inline int enter(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    return MCD_E_CUDA;
  }
  return MCD_OK;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return MCD_E_CUDA;
  }
  count_launch();
  return MCD_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: remember the largest size set
// for (kernel, current device), so that a process that drives several devices (the reference's nn.DataParallel
// pattern, SURVEY 8b "Threading") configures each of them.  A benign race (two threads setting the same value) is
// the worst concurrent outcome.
template <auto Kernel>
inline int ensure_dyn_smem(int bytes, const char* what) {
  constexpr int kMaxDev = 64;
  static int set_bytes[kMaxDev] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  const bool tracked = dev < kMaxDev;
  if (tracked && bytes <= set_bytes[dev]) return MCD_OK;
  cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(max dynamic smem = %d B): %s", what, bytes, cudaGetErrorString(e));
    return MCD_E_CUDA;
  }
  if (tracked) set_bytes[dev] = bytes;
  return MCD_OK;
}

#define MCD_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::mcd::set_error(__VA_ARGS__);      \
      return MCD_E_INVALID;               \
    }                                     \
  } while (0)

#define MCD_ENTER(device)                 \
  do {                                    \
    int _rc = ::mcd::enter(device);       \
    if (_rc != MCD_OK) return _rc;        \
  } while (0)

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }

// element i of the "rowconv" weight pack (conv_rows.cu): dst[r][kg = half*SP + s][n/8][n%8][c%8], n = produced
// channel, c = half*8 + c%8 = reduce channel; mode 0: w[n][c][r][s], mode 1 (dgrad): w[c][n][R-1-r][S-1-s].
__device__ __forceinline__ float rowconv_pack_value(const float* __restrict__ w, int64_t i, int Cout, int Cin, int R,
                                                    int S, int HC, int SP, int NB, int mode) {
  const int ngs = NB / 8;
  const int c8 = (int)(i % 8), n8 = (int)((i / 8) % 8), ng = (int)((i / 64) % ngs);
  const int kg = (int)((i / (64 * (int64_t)ngs)) % (HC * SP));
  const int r = (int)(i / (64 * (int64_t)ngs * HC * SP));
  const int hf = kg / SP, s = kg % SP;
  const int n = ng * 8 + n8, c = hf * 8 + c8;
  if (s >= S) return 0.f;
  if (mode == 0) return (n < Cout && c < Cin) ? w[(((int64_t)n * Cin + c) * R + r) * S + s] : 0.f;
  return (n < Cin && c < Cout) ? w[(((int64_t)c * Cin + n) * R + (R - 1 - r)) * S + (S - 1 - s)] : 0.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- 16-bit storage formats --------------------------------------------------------------------------------
// FORWARD tensors (activations, packed fprop weights) are IEEE half: 11 significant bits keep the ReLU-mask flip
// rate against an fp32 run 8x below bf16 (tests/tools/precision_study.py).  GRADIENT tensors (dy, dx, dlogits, packed
// dgrad weights) and the bf16 TWIN of every activation that feeds a weight-gradient GEMM are bfloat16: fp32 range, no
// loss scaling.  tcgen05.mma kind::f16 requires A and B in the SAME format (profiles/r02_exp_mixed_format.txt:
// mixing is an illegal instruction), hence the twin.  The codes equal the UMMA instruction-descriptor formats.
constexpr int kF16 = 0, kBF16 = 1;

// unpack 8 bf16 (one 16-byte vector) to floats
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
// the same for IEEE half; the pack saturates to +-65504 instead of producing inf
__device__ __forceinline__ void unpack8h(const uint4& v, float* f) {
  const __half2* p = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint32_t f2h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint4 pack8h(const float* f) {
  return make_uint4(f2h2_sat(f[0], f[1]), f2h2_sat(f[2], f[3]), f2h2_sat(f[4], f[5]), f2h2_sat(f[6], f[7]));
}
template <int FMT> __device__ __forceinline__ void unpack8f(const uint4& v, float* f) {
  if (FMT == kF16) unpack8h(v, f); else unpack8(v, f);
}
template <int FMT> __device__ __forceinline__ uint4 pack8f(const float* f) {
  return FMT == kF16 ? pack8h(f) : pack8(f);
}
__device__ __forceinline__ void unpack8r(const uint4& v, float* f, int fmt) {
  if (fmt == kF16) unpack8h(v, f); else unpack8(v, f);
}
__device__ __forceinline__ uint4 pack8r(const float* f, int fmt) { return fmt == kF16 ? pack8h(f) : pack8(f); }
// one element -> 16 storage bits
__device__ __forceinline__ uint16_t f2bits16(float v, int fmt) {
  if (fmt == kF16) return (uint16_t)(f2h2_sat(v, 0.f) & 0xffffu);
  __nv_bfloat16 b = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&b);
}
__device__ __forceinline__ float bits16_2f(uint16_t u, int fmt) {
  if (fmt == kF16) { __half h = *reinterpret_cast<__half*>(&u); return __half2float(h); }
  return __uint_as_float((uint32_t)u << 16);
}

// block-wide sum of `v` (blockDim.x multiple of 32, <= 1024); result valid in thread 0.
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem32[wid] = v;
  __syncthreads();
  float r = 0.f;
  if (wid == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? smem32[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;
}

}  // namespace mcd
