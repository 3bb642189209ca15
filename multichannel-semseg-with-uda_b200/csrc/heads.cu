// heads.cu — classifier-head upsampling: the learned depthwise ConvTranspose2d(C,C,16,s8,p4,groups=C)
// of DRNSegPixelClassifier / FusionDRNSegPixelClassifier / ScoreFusionDRNSegPixelClassifier
// (models/dilated_fcn.py:357-366,465-470,479-491) and nn.Upsample(bilinear, align_corners=False) of the
// multitask decoders (models/dilated_fcn.py:676,817-819).  Low-res score maps are planar fp32, the
// full-resolution outputs planar bf16 (or fp32 for regression targets); all kernels are HBM-bound
// on the full-resolution tensor and read / write it exactly once.
#include "common.cuh"

namespace mcd {

// ---- depthwise deconv 16x16 stride 8 pad 4 ------------------------------------------------------
// out[oh][ow] = sum_{a,b in {0,1}} x[ih0-a][iw0-b] * w[kh0+8a][kw0+8b],  ih0 = (oh+4)>>3, kh0 = (oh+4)&7
// "cell row" i (0..h) = the 8 output rows oh = 8i-4 .. 8i+3 that read x[i] (filter rows kh0) and x[i-1]
// (filter rows kh0+8).  One thread owns a cell (i, j): it loads the 6 neighbouring inputs once and emits 8 rows
// of 8 consecutive output columns (one 16-byte store per row) with broadcast LDS.128 weight reads.
// grid: (N*C, ceil((h+1)/cells_per_block)); block 256.
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* f) { *reinterpret_cast<uint4*>(p) = pack8(f); }
__device__ __forceinline__ void store8(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float* f) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
  const float2 p0 = __bfloat1622float2(pv[0]), p1 = __bfloat1622float2(pv[1]);
  f[0] = p0.x; f[1] = p0.y; f[2] = p1.x; f[3] = p1.y;
}
__device__ __forceinline__ void load4(const float* p, float* f) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float* f) {
  unpack8(__ldg(reinterpret_cast<const uint4*>(p)), f);
}
__device__ __forceinline__ void load8(const float* p, float* f) { load4(p, f); load4(p + 4, f + 4); }

// T = type of the full-resolution tensor: bf16 (MCDStep) or fp32 (drop-in default)
template <typename T>
__global__ void __launch_bounds__(256)
deconv16s8_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                      const float* __restrict__ x2, const float* __restrict__ w2,
                      T* __restrict__ out, int C, int h, int wd, int cells_per_block) {
  __shared__ __align__(16) float sw[2][256];
  const int nc = blockIdx.x, c = nc % C;
  const int H = h * 8, W = wd * 8;
  sw[0][threadIdx.x] = w[c * 256 + threadIdx.x];
  sw[1][threadIdx.x] = x2 ? (w2 ? w2 : w)[c * 256 + threadIdx.x] : 0.f;
  __syncthreads();
  const int ninputs = x2 ? 2 : 1;
  T* op = out + (int64_t)nc * H * W;
  const int items = cells_per_block * wd;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int i = blockIdx.y * cells_per_block + it / wd, j = it % wd;
    if (i > h) continue;
    float xa[2][3], xb[2][3];   // [input][column j-1, j, j+1] of rows i (a) and i-1 (b)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float* xp = (q == 0 ? x : x2);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int col = j - 1 + d;
        const bool cok = q < ninputs && col >= 0 && col < wd;
        xa[q][d] = (cok && i < h) ? xp[((int64_t)nc * h + i) * wd + col] : 0.f;
        xb[q][d] = (cok && i >= 1) ? xp[((int64_t)nc * h + i - 1) * wd + col] : 0.f;
      }
    }
#pragma unroll 2
    for (int kh0 = 0; kh0 < 8; ++kh0) {
      const int oh = 8 * i - 4 + kh0;
      if (oh < 0 || oh >= H) continue;
      float f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = 0.f;
      for (int q = 0; q < ninputs; ++q) {
        float wa[16], wb[16];   // filter rows kh0 and kh0+8
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          *reinterpret_cast<float4*>(wa + 4 * v) = *reinterpret_cast<const float4*>(&sw[q][kh0 * 16 + 4 * v]);
          *reinterpret_cast<float4*>(wb + 4 * v) = *reinterpret_cast<const float4*>(&sw[q][(kh0 + 8) * 16 + 4 * v]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {     // ow = 8j+k: iw0 = j, kw0 = k+4
          f[k] = fmaf(xa[q][1], wa[k + 4], f[k]);
          f[k] = fmaf(xa[q][0], wa[k + 12], f[k]);
          f[k] = fmaf(xb[q][1], wb[k + 4], f[k]);
          f[k] = fmaf(xb[q][0], wb[k + 12], f[k]);
        }
#pragma unroll
        for (int k = 4; k < 8; ++k) {     // ow = 8j+k: iw0 = j+1, kw0 = k-4
          f[k] = fmaf(xa[q][2], wa[k - 4], f[k]);
          f[k] = fmaf(xa[q][1], wa[k + 4], f[k]);
          f[k] = fmaf(xb[q][2], wb[k - 4], f[k]);
          f[k] = fmaf(xb[q][1], wb[k + 4], f[k]);
        }
      }
      store8(op + (int64_t)oh * W + j * 8, f);
    }
  }
}

// dx[ih][iw] = sum_{kh,kw} dout[8ih-4+kh][8iw-4+kw] * w[kh][kw]
// grid: (N*C, h); block 128: thread t handles iw = t, t+128, ...
template <typename T>
__global__ void __launch_bounds__(128)
deconv16s8_bwd_dx_kernel(const T* __restrict__ dout, const float* __restrict__ w,
                         float* __restrict__ dx, int C, int h, int wd) {
  __shared__ float sw[256];
  const int nc = blockIdx.x, c = nc % C, ih = blockIdx.y;
  const int H = h * 8, W = wd * 8;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sw[i] = w[c * 256 + i];
  __syncthreads();
  const T* dp = dout + (int64_t)nc * H * W;
  for (int iw = threadIdx.x; iw < wd; iw += blockDim.x) {
    float acc = 0.f;
    for (int kh = 0; kh < 16; ++kh) {
      const int oh = 8 * ih - 4 + kh;
      if (oh < 0 || oh >= H) continue;
      const int ow0 = 8 * iw - 4;  // multiple of 4 -> 8-byte aligned groups of 4
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const int ow = ow0 + 4 * k4;
        if (ow < 0 || ow + 3 >= W) continue;  // W % 8 == 0 and ow % 4 == 0: group fully in or out
        float d4[4];
        load4(dp + (int64_t)oh * W + ow, d4);
        const float* ww = &sw[kh * 16 + 4 * k4];
        acc = fmaf(d4[0], ww[0], acc); acc = fmaf(d4[1], ww[1], acc);
        acc = fmaf(d4[2], ww[2], acc); acc = fmaf(d4[3], ww[3], acc);
      }
    }
    dx[(int64_t)nc * h * wd + ih * wd + iw] = acc;
  }
}

// dw[c][kh][kw] = sum_{n,ih,iw} x[n,c,ih,iw] * dout[n,c,8ih-4+kh,8iw-4+kw]
// Every dout element (oh, ow) feeds exactly four bins: with ih0 = (oh+4)>>3, kh0 = (oh+4)&7 and iw0 = (ow+4)>>3,
// kw0 = (ow+4)&7 these are (ih0|ih0-1, kh0|kh0+8) x (iw0|iw0-1, kw0|kw0+8).  A block owns one (c, kh0): all rows
// with that residue, i.e. both kh bins, so dout is read exactly ONCE, as 16-byte chunks (8 ow = one iw step).
// grid: (C*8 [c,kh0], nsplit chunks of (n,ih0) rows); block 256 = 8 warps, one row per warp-iteration; dw pre-zeroed.
template <typename T>
__global__ void __launch_bounds__(256)
deconv16s8_bwd_dw_kernel(const T* __restrict__ dout, const float* __restrict__ x,
                         float* __restrict__ dw, int N, int C, int h, int wd) {
  __shared__ float red[8][32];
  const int c = blockIdx.x >> 3, kh0 = blockIdx.x & 7;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = h * 8, W = wd * 8;
  const int rows = N * (h + 1);
  const int per = (rows + gridDim.y - 1) / gridDim.y;
  const int r_beg = blockIdx.y * per, r_end = min(r_beg + per, rows);
  float accA[16], accB[16];      // A: (ih0, kh0) bins kw 0..15;  B: (ih0-1, kh0+8)
#pragma unroll
  for (int k = 0; k < 16; ++k) { accA[k] = 0.f; accB[k] = 0.f; }
  for (int rr = r_beg + warp; rr < r_end; rr += 8) {
    const int n = rr / (h + 1), ih0 = rr % (h + 1);
    const int oh = 8 * ih0 + kh0 - 4;
    if (oh < 0 || oh >= H) continue;
    const T* dp = dout + (((int64_t)n * C + c) * H + oh) * W;
    const float* xa = x + (((int64_t)n * C + c) * h + ih0) * wd;     // valid if ih0 < h
    const float* xb = xa - wd;                                       // valid if ih0 >= 1
    const bool va = ih0 < h, vb = ih0 >= 1;
    for (int j0 = lane; j0 < wd; j0 += 96) {
      float raw[3][8];
#pragma unroll
      for (int q = 0; q < 3; ++q) {          // three independent vector loads per lane in flight
        const int j = j0 + 32 * q;
        if (j < wd) load8(dp + j * 8, raw[q]);
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int j = j0 + 32 * q;
        if (j >= wd) continue;
        const float* d = raw[q];
        const float a0 = va ? xa[j] : 0.f, am = (va && j >= 1) ? xa[j - 1] : 0.f, ap = (va && j + 1 < wd) ? xa[j + 1] : 0.f;
        const float b0 = vb ? xb[j] : 0.f, bm = (vb && j >= 1) ? xb[j - 1] : 0.f, bp = (vb && j + 1 < wd) ? xb[j + 1] : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {      // ow = 8j+k: iw0 = j, kw0 = k+4
          accA[k + 4] = fmaf(a0, d[k], accA[k + 4]);  accA[k + 12] = fmaf(am, d[k], accA[k + 12]);
          accB[k + 4] = fmaf(b0, d[k], accB[k + 4]);  accB[k + 12] = fmaf(bm, d[k], accB[k + 12]);
        }
#pragma unroll
        for (int k = 4; k < 8; ++k) {      // ow = 8j+k: iw0 = j+1, kw0 = k-4
          accA[k - 4] = fmaf(ap, d[k], accA[k - 4]);  accA[k + 4] = fmaf(a0, d[k], accA[k + 4]);
          accB[k - 4] = fmaf(bp, d[k], accB[k - 4]);  accB[k + 4] = fmaf(b0, d[k], accB[k + 4]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    float va_ = accA[k], vb_ = accB[k];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      va_ += __shfl_xor_sync(0xffffffffu, va_, o);
      vb_ += __shfl_xor_sync(0xffffffffu, vb_, o);
    }
    if (lane == 0) { red[warp][k] = va_; red[warp][16 + k] = vb_; }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += red[k][threadIdx.x];
    const int kh = threadIdx.x < 16 ? kh0 : kh0 + 8, kw = threadIdx.x & 15;
    atomicAdd(dw + c * 256 + kh * 16 + kw, sum);
  }
}

// ---- bilinear upsample, align_corners=False ------------------------------------------------------
__device__ __forceinline__ void bil_src(int o, int s, int in, int* i0, int* i1, float* lam) {
  float src = (o + 0.5f) / (float)s - 0.5f;
  if (src < 0.f) src = 0.f;
  int a = (int)src;
  if (a > in - 1) a = in - 1;
  *i0 = a;
  *i1 = a + (a < in - 1 ? 1 : 0);
  *lam = src - (float)a;
}

template <bool OUT_F32>
__global__ void __launch_bounds__(256)
bilinear_fwd_kernel(const float* __restrict__ x, void* __restrict__ out, int h, int wd, int s,
                    int64_t total8) {
  const int H = h * s, W = wd * s;
  const int w8 = W >> 3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total8;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % w8);
    const int oh = (int)((i / w8) % H);
    const int64_t nc = i / ((int64_t)w8 * H);
    int h0, h1; float lh;
    bil_src(oh, s, h, &h0, &h1, &lh);
    const float* r0 = x + (nc * h + h0) * wd;
    const float* r1 = x + (nc * h + h1) * wd;
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int w0, w1; float lw;
      bil_src(j * 8 + k, s, wd, &w0, &w1, &lw);
      const float top = r0[w0] + lw * (r0[w1] - r0[w0]);
      const float bot = r1[w0] + lw * (r1[w1] - r1[w0]);
      f[k] = top + lh * (bot - top);
    }
    const int64_t o = (nc * H + oh) * (int64_t)W + j * 8;
    if (OUT_F32) {
      float* op = reinterpret_cast<float*>(out) + o;
      *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(op + 4) = make_float4(f[4], f[5], f[6], f[7]);
    } else {
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + o) = pack8(f);
    }
  }
}

// separable gather: block = one (n*c, ih) low-res row.
//   col[ow] = sum_{oh} wh(oh, ih) * dout[oh][ow]     (coalesced along ow)
//   dx[iw]  = sum_{ow} ww(ow, iw) * col[ow]
template <bool IN_F32>
__global__ void __launch_bounds__(256)
bilinear_bwd_kernel(const void* __restrict__ dout, float* __restrict__ dx, int h, int wd, int s) {
  extern __shared__ float col[];  // W floats
  const int H = h * s, W = wd * s;
  const int64_t nc = blockIdx.x;
  const int ih = blockIdx.y;
  const int oh_lo = max(0, s * (ih - 1)), oh_hi = min(H - 1, s * (ih + 2) - 1);
  for (int ow = threadIdx.x; ow < W; ow += blockDim.x) {
    float acc = 0.f;
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      int h0, h1; float lh;
      bil_src(oh, s, h, &h0, &h1, &lh);
      float wgt = (h0 == ih ? 1.f - lh : 0.f) + (h1 == ih ? lh : 0.f);
      if (wgt == 0.f) continue;
      const int64_t o = (nc * H + oh) * (int64_t)W + ow;
      float v = IN_F32 ? reinterpret_cast<const float*>(dout)[o]
                       : bf2f(reinterpret_cast<const __nv_bfloat16*>(dout)[o]);
      acc = fmaf(wgt, v, acc);
    }
    col[ow] = acc;
  }
  __syncthreads();
  for (int iw = threadIdx.x; iw < wd; iw += blockDim.x) {
    const int ow_lo = max(0, s * (iw - 1)), ow_hi = min(W - 1, s * (iw + 2) - 1);
    float acc = 0.f;
    for (int ow = ow_lo; ow <= ow_hi; ++ow) {
      int w0, w1; float lw;
      bil_src(ow, s, wd, &w0, &w1, &lw);
      float wgt = (w0 == iw ? 1.f - lw : 0.f) + (w1 == iw ? lw : 0.f);
      acc = fmaf(wgt, col[ow], acc);
    }
    dx[(nc * h + ih) * wd + iw] = acc;
  }
}

}  // namespace mcd

using namespace mcd;

extern "C" {

int mcd_deconv16s8_fwd(const float* x, const float* w, const float* x2, const float* w2, void* out, int out_f32,
                       int N, int C, int h, int w_, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(x && w && out && N > 0 && C > 0 && h > 0 && w_ > 0, "deconv16s8_fwd: bad arguments");
  int cells = w_ >= 256 ? 1 : 256 / w_;
  dim3 grid((unsigned)(N * C), (unsigned)((h + 1 + cells - 1) / cells));
  if (out_f32)
    deconv16s8_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, x2, w2, (float*)out, C, h, w_, cells);
  else
    deconv16s8_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, x2, w2, (__nv_bfloat16*)out,
                                                                                 C, h, w_, cells);
  return check_launch("deconv16s8_fwd");
}

int mcd_deconv16s8_bwd(const void* dout, int dout_f32, const float* x, const float* w, float* dx, float* dw,
                       int N, int C, int h, int w_, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(dout && w && N > 0 && C > 0 && h > 0 && w_ > 0, "deconv16s8_bwd: bad arguments");
  MCD_REQUIRE(!dw || x, "deconv16s8_bwd: dw needs x");
  if (dx) {
    dim3 grid((unsigned)(N * C), (unsigned)h);
    if (dout_f32)
      deconv16s8_bwd_dx_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>((const float*)dout, w, dx, C, h, w_);
    else
      deconv16s8_bwd_dx_kernel<__nv_bfloat16><<<grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dout, w,
                                                                                      dx, C, h, w_);
    int rc = check_launch("deconv16s8_bwd_dx");
    if (rc != MCD_OK) return rc;
  }
  if (dw) {
    cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)C * 256, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("deconv dw memset: %s", cudaGetErrorString(e)); return MCD_E_CUDA; }
    int nsplit = min(N * (h + 1), 32);
    dim3 grid((unsigned)(C * 8), (unsigned)nsplit);
    if (dout_f32)
      deconv16s8_bwd_dw_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)dout, x, dw, N, C, h, w_);
    else
      deconv16s8_bwd_dw_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dout, x,
                                                                                      dw, N, C, h, w_);
    return check_launch("deconv16s8_bwd_dw");
  }
  return MCD_OK;
}

int mcd_bilinear_up_fwd(const float* x, void* out, int out_f32, int N, int C, int h, int w_, int s,
                        int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(x && out && N > 0 && C > 0 && h > 0 && w_ > 0, "bilinear_up_fwd: bad arguments");
  MCD_REQUIRE(s == 2 || s == 4 || s == 8, "bilinear_up_fwd: scale %d unsupported (2/4/8)", s);
  MCD_REQUIRE((w_ * s) % 8 == 0, "bilinear_up_fwd: output width must be a multiple of 8");
  int64_t total8 = (int64_t)N * C * h * s * (w_ * s / 8);
  int grid = (int)min64((total8 + 255) / 256, 148 * 16);
  if (out_f32)
    bilinear_fwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, h, w_, s, total8);
  else
    bilinear_fwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, h, w_, s, total8);
  return check_launch("bilinear_up_fwd");
}

int mcd_bilinear_up_bwd(const void* dout, int dout_f32, float* dx, int N, int C, int h, int w_,
                        int s, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(dout && dx && N > 0 && C > 0 && h > 0 && w_ > 0, "bilinear_up_bwd: bad arguments");
  MCD_REQUIRE(s == 2 || s == 4 || s == 8, "bilinear_up_bwd: scale %d unsupported (2/4/8)", s);
  dim3 grid((unsigned)(N * C), (unsigned)h);
  size_t smem = sizeof(float) * (size_t)w_ * s;
  if (dout_f32)
    bilinear_bwd_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(dout, dx, h, w_, s);
  else
    bilinear_bwd_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(dout, dx, h, w_, s);
  return check_launch("bilinear_up_bwd");
}

}  // extern "C"
