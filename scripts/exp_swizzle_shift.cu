// exp_swizzle_shift.cu — hardware experiment (sm_100a): can a UMMA shared-memory descriptor start at an arbitrary
// ROW of a swizzled tile?  i.e. is the 128B/64B/32B swizzle XOR taken from absolute shared-memory address bits
// (then "start = base + j * row_bytes" reads rows j, j+1, ... of a tile that TMA wrote once) or from
// matrix-relative offsets (then only j % 8 == 0 works unless base_offset compensates).
// This decides whether one halo tile can serve all filter taps of a convolution (no im2col expansion).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp_swizzle_shift scripts/exp_swizzle_shift.cu && ./exp_swizzle_shift
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include "../multichannel-semseg-with-uda_b200/csrc/umma_ptx.cuh"

using namespace mcd::ptx;

__device__ __forceinline__ float aval(int r, int k) { return (float)(((r * 7 + k * 3) % 17) - 8); }
__device__ __forceinline__ float bval(int n, int k) { return (float)(((n * 5 + k * 11) % 13) - 6); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout,
                                              uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)layout << 61;
  return d;
}

// swizzle<B,4,3> on a byte offset relative to a 1024-aligned base
__device__ __forceinline__ uint32_t swz(uint32_t off, int bits) {
  return off ^ (((off >> 7) & ((1u << bits) - 1)) << 4);
}

__global__ void __launch_bounds__(128) exp_kernel(int* results) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // up to 160 rows x 128 B = 20 KB (+ second MN atom)
  uint8_t* sB = smem + 48 * 1024;     // 64 rows x 128 B
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
  // ---------------- K-major, swizzle widths 128 / 64 / 32 bytes ----------------
  for (int wi = 0; wi < 3; ++wi) {
    const int W = 128 >> wi, bits = 3 - wi, layout = wi == 0 ? 2 : (wi == 1 ? 4 : 6);
    const int K = W / 2;
    // fill A_full (160 rows) and B (64 rows), swizzled on absolute offsets
    for (int i = tid; i < 160 * K; i += 128) {
      int r = i / K, k = i % K;
      *reinterpret_cast<__nv_bfloat16*>(sA + swz(r * W + k * 2, bits)) = __float2bfloat16(aval(r, k));
    }
    for (int i = tid; i < 64 * K; i += 128) {
      int n = i / K, k = i % K;
      *reinterpret_cast<__nv_bfloat16*>(sB + swz(n * W + k * 2, bits)) = __float2bfloat16(bval(n, k));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    for (int mode = 0; mode < 2; ++mode) {
      for (int j = 0; j < 10; ++j) {
        if (tid == 0) {
          const uint32_t a0 = smem_u32(sA) + j * W, b0 = smem_u32(sB);
          const uint32_t idesc = instr_desc_bf16(128, 64, 0, 0);
          for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t aa = a0 + ks * 32, bb = b0 + ks * 32;
            const uint64_t ad = make_desc(aa, 0, 8 * W, layout, mode ? ((aa >> 7) & 7) : 0);
            const uint64_t bd = make_desc(bb, 0, 8 * W, layout, 0);
            umma_bf16(tmem, ad, bd, idesc, ks ? 1u : 0u);
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
          for (int k = 0; k < K; ++k) ref += aval(tid + j, k) * bval(n, k);
          if (ref != v[n]) ++bad;
        }
        tc_fence_before();
        atomicAdd(&results[test], bad);
        __syncthreads();
        ++test;
      }
    }
    __syncthreads();
  }
  // ---------------- MN-major SW128 (wgrad-style): A[k = pixel][m = channel], shift along K (pixels) ----------------
  {
    // A_full: 2 atoms (64 channels each) x 96 pixel rows x 128 B, atom stride LBO = 96*128; B: 1 atom x 64 pixel rows
    const int LBO = 96 * 128;
    for (int i = tid; i < 2 * 96 * 64; i += 128) {
      int atom = i / (96 * 64), p = (i / 64) % 96, c = i % 64;
      *reinterpret_cast<__nv_bfloat16*>(sA + atom * LBO + swz(p * 128 + c * 2, 3)) =
          __float2bfloat16(aval(p, atom * 64 + c));
    }
    for (int i = tid; i < 64 * 64; i += 128) {
      int p = i / 64, c = i % 64;
      *reinterpret_cast<__nv_bfloat16*>(sB + swz(p * 128 + c * 2, 3)) = __float2bfloat16(bval(c, p));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    for (int mode = 0; mode < 2; ++mode) {
      for (int j = 0; j < 10; ++j) {
        if (tid == 0) {
          const uint32_t a0 = smem_u32(sA) + j * 128, b0 = smem_u32(sB);
          const uint32_t idesc = instr_desc_bf16(128, 64, 1, 1);
          for (int ks = 0; ks < 4; ++ks) {          // K = 64 pixels, 16 per MMA = 2048 B
            const uint32_t aa = a0 + ks * 2048, bb = b0 + ks * 2048;
            const uint64_t ad = make_desc(aa, LBO, 1024, 2, mode ? ((aa >> 7) & 7) : 0);
            const uint64_t bd = make_desc(bb, 64 * 128, 1024, 2, 0);
            umma_bf16(tmem, ad, bd, idesc, ks ? 1u : 0u);
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
          for (int k = 0; k < 64; ++k) ref += aval(k + j, tid) * bval(n, k);   // D[m][n] = sum_p A[p+j][m] B[p][n]
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

int main() {
  int* d;
  cudaMalloc(&d, 256 * sizeof(int));
  cudaMemset(d, 0, 256 * sizeof(int));
  const int smem = 66 * 1024 + 1024;
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  exp_kernel<<<1, 128, smem>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  int h[256];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[4] = {"K-major SW128", "K-major SW64", "K-major SW32", "MN-major SW128 (K shift)"};
  int t = 0;
  for (int c = 0; c < 4; ++c)
    for (int mode = 0; mode < 2; ++mode) {
      printf("%-26s base_offset=%s  mismatches per row shift j=0..9:", names[c], mode ? "(addr>>7)&7" : "0");
      for (int j = 0; j < 10; ++j) printf(" %5d", h[t++]);
      printf("\n");
    }
  return 0;
}
