// exp_mixed_format.cu — hardware experiment (sm_100a): does tcgen05.mma kind::f16 accept DIFFERENT element formats
// for A and B (instruction-descriptor a_format != b_format: IEEE half x bfloat16)?  Decides whether the weight
// gradient GEMM can read fp16 forward activations and bf16 gradients directly (no conversion pass).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp_mixed_format scripts/exp_mixed_format.cu && ./exp_mixed_format
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "../multichannel-semseg-with-uda_b200/csrc/umma_ptx.cuh"

using namespace mcd::ptx;

// values with fractional parts that are exact in both formats (multiples of 1/8, |v| < 16)
__device__ __forceinline__ float aval(int r, int k) { return (float)(((r * 7 + k * 3) % 37) - 18) * 0.125f; }
__device__ __forceinline__ float bval(int n, int k) { return (float)(((n * 5 + k * 11) % 29) - 14) * 0.25f; }

__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }
__device__ __forceinline__ uint16_t enc(float v, int bf) {
  if (bf) { __nv_bfloat16 t = __float2bfloat16(v); return *reinterpret_cast<uint16_t*>(&t); }
  __half t = __float2half(v);
  return *reinterpret_cast<uint16_t*>(&t);
}
// kind::f16 instruction descriptor with explicit formats (0 = f16, 1 = bf16), fp32 accumulate
__device__ __forceinline__ uint32_t idesc(uint32_t M, uint32_t N, uint32_t afmt, uint32_t bfmt, uint32_t amn,
                                          uint32_t bmn) {
  return (1u << 4) | (afmt << 7) | (bfmt << 10) | (amn << 15) | (bmn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__global__ void __launch_bounds__(128) exp_kernel(int* results, int only) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 128 rows x 128 B
  uint8_t* sB = smem + 32 * 1024;     // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  uint32_t phase = 0;
  int test = 0;
  for (int major = 0; major < 2; ++major) {          // 0: both K-major, 1: both MN-major (wgrad style)
    for (int af = 0; af < 2; ++af) {
      for (int bfm = 0; bfm < 2; ++bfm) {
        if (only >= 0 && only != test) { ++test; continue; }
        if (major == 0) {
          for (int i = tid; i < 128 * 64; i += 128) {
            int r = i / 64, k = i % 64;
            *reinterpret_cast<uint16_t*>(sA + swz(r * 128 + k * 2)) = enc(aval(r, k), af);
          }
          for (int i = tid; i < 64 * 64; i += 128) {
            int n = i / 64, k = i % 64;
            *reinterpret_cast<uint16_t*>(sB + swz(n * 128 + k * 2)) = enc(bval(n, k), bfm);
          }
        } else {
          // MN-major SW128: A[k = pixel][m = channel], two 64-channel atoms (LBO apart), 64 pixel rows of 128 B
          for (int i = tid; i < 2 * 64 * 64; i += 128) {
            int atom = i / (64 * 64), p = (i / 64) % 64, c = i % 64;
            *reinterpret_cast<uint16_t*>(sA + atom * 8192 + swz(p * 128 + c * 2)) = enc(aval(atom * 64 + c, p), af);
          }
          for (int i = tid; i < 64 * 64; i += 128) {
            int p = i / 64, c = i % 64;
            *reinterpret_cast<uint16_t*>(sB + swz(p * 128 + c * 2)) = enc(bval(c, p), bfm);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
          const uint32_t id = idesc(128, 64, af, bfm, major, major);
          for (int ks = 0; ks < 4; ++ks) {
            uint64_t ad, bd;
            if (major == 0) {
              ad = smem_desc_sw128(smem_u32(sA) + ks * 32, 0, 1024);
              bd = smem_desc_sw128(smem_u32(sB) + ks * 32, 0, 1024);
            } else {
              ad = smem_desc_sw128(smem_u32(sA) + ks * 2048, 8192, 1024);
              bd = smem_desc_sw128(smem_u32(sB) + ks * 2048, 8192, 1024);
            }
            umma_bf16(tmem, ad, bd, id, ks ? 1u : 0u);
          }
          umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        float v[64];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 32, v + 32);
        tmem_ld_wait();
        int bad = 0;
        for (int n = 0; n < 64; ++n) {
          float ref = 0.f;
          for (int k = 0; k < 64; ++k) ref += aval(tid, k) * bval(n, k);
          if (ref != v[n]) ++bad;
        }
        tc_fence_before();
        atomicAdd(&results[test], bad);
        __syncthreads();
        ++test;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  int* d;
  cudaMalloc(&d, 64 * sizeof(int));
  cudaMemset(d, 0, 64 * sizeof(int));
  const int smem = 66 * 1024 + 1024;
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  exp_kernel<<<1, 128, smem>>>(d, only);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  int h[64];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* fm[2] = {"f16", "bf16"};
  int t = 0;
  for (int major = 0; major < 2; ++major)
    for (int af = 0; af < 2; ++af)
      for (int bf = 0; bf < 2; ++bf)
      {
        if (only < 0 || only == t)
          printf("%-9s A=%-4s B=%-4s mismatches (of 8192): %d\n", major ? "MN-major" : "K-major", fm[af], fm[bf], h[t]);
        ++t;
      }
  return 0;
}
