// conv_direct.cu — smem-tiled CUDA-core convolution kernels (fp32 accumulate, bf16 storage).
// They accept any geometry, serve as the on-device cross-check of the tcgen05 kernels and run the
// layers the tcgen05 path does not take.  Reference semantics: nn.Conv2d (models/drn.py:21-23).
#include "common.cuh"
#include "conv_plan.h"

namespace mcd {

constexpr int DT = 64;  // tile: 64 pixels x 64 produced channels, 16 reduce channels per step

template <bool PLANAR>
__global__ void __launch_bounds__(256)
conv_direct_kernel(const __nv_bfloat16* __restrict__ src, const __nv_bfloat16* __restrict__ w,
                   const float* __restrict__ bias, void* __restrict__ out,
                   const __nv_bfloat16* __restrict__ addend, const TapProblem p, const int fmt) {
  __shared__ float As[16][DT + 4];
  __shared__ float Bs[16][DT + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t npix = (int64_t)p.N * p.Ht * p.Wt;
  const int64_t pix0 = (int64_t)blockIdx.x * DT;
  const int row0 = blockIdx.y * DT;

  // loader roles
  const bool loadA = tid < 128;
  const int lidx = (tid & 127) >> 1, lhalf = tid & 1;
  int ln = 0, lht = 0, lwt = 0;
  bool lvalid = false;
  if (loadA) {
    int64_t P = pix0 + lidx;
    lvalid = P < npix;
    if (lvalid) {
      lwt = (int)(P % p.Wt);
      lht = (int)((P / p.Wt) % p.Ht);
      ln = (int)(P / ((int64_t)p.Wt * p.Ht));
    }
  }
  const int lrow = row0 + lidx;
  const int64_t wrow_stride = (int64_t)p.T_total * p.kc_pad;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < p.ntaps; ++t) {
    const Tap tap = p.taps[t];
    const __nv_bfloat16* aptr = nullptr;
    if (loadA && lvalid) {
      int hs = lht * p.smul + tap.dh, ws = lwt * p.smul + tap.dw;
      if (hs >= 0 && hs < p.Hs && ws >= 0 && ws < p.Ws)
        aptr = src + (((int64_t)ln * p.Hs + hs) * p.Ws + ws) * p.Cs_src;
    }
    const __nv_bfloat16* bptr =
        (!loadA && lrow < p.rows) ? w + lrow * wrow_stride + (int64_t)tap.wk * p.kc_pad : nullptr;
    for (int kc0 = 0; kc0 < p.Kc; kc0 += 16) {
      int kc = kc0 + lhalf * 8;
      float f[8];
      uint4 v = make_uint4(0, 0, 0, 0);
      if (loadA) {
        if (aptr && kc < p.Kc) v = *reinterpret_cast<const uint4*>(aptr + kc);
      } else {
        if (bptr && kc < p.kc_pad) v = *reinterpret_cast<const uint4*>(bptr + kc);
      }
      unpack8r(v, f, fmt);     // src, w and the nhwc output share the format (kF16 forward / kBF16 dgrad)
      if (loadA && kc + 8 > p.Kc) {  // partial vector: channels past Kc must not contribute
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (kc + k >= p.Kc) f[k] = 0.f;
      }
      __syncthreads();
      if (loadA) {
#pragma unroll
        for (int k = 0; k < 8; ++k) As[lhalf * 8 + k][lidx] = f[k];
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) Bs[lhalf * 8 + k][lidx] = f[k];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t P = pix0 + ty * 4 + i;
    if (P >= npix) continue;
    int wt = (int)(P % p.Wt), ht = (int)((P / p.Wt) % p.Ht), n = (int)(P / ((int64_t)p.Wt * p.Ht));
    int hd = ht * p.omul + p.oh0, wd = wt * p.omul + p.ow0;
    if (hd >= p.Hd || wd >= p.Wd) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int row = row0 + tx * 4 + j;
      if (PLANAR) {
        if (row < p.rows) {
          float v = acc[i][j] + (bias ? bias[row] : 0.f);
          reinterpret_cast<float*>(out)[(((int64_t)n * p.rows + row) * p.Hd + hd) * p.Wd + wd] = v;
        }
      } else {
        if (row < p.Cd_s) {
          float v = row < p.rows ? acc[i][j] + (bias ? bias[row] : 0.f) : 0.f;
          const int64_t oidx = (((int64_t)n * p.Hd + hd) * p.Wd + wd) * p.Cd_s + row;
          if (addend) v += bf2f(addend[oidx]);
          reinterpret_cast<uint16_t*>(out)[oidx] = f2bits16(v, fmt);
        }
      }
    }
  }
}

// wgrad: dw[co][ci][r][s] += sum_{pixels of this block} dy[p][co] * x[src(p,t)][ci]
// grid: x = pixel chunks, y = co tiles * ci tiles, z = taps
constexpr int WG_PCH = 1024;
__global__ void __launch_bounds__(256)
wgrad_direct_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                    float* __restrict__ dw, const mcd_conv_geom g) {
  __shared__ float Ds[16][DT + 4];  // [pixel][co]
  __shared__ float Xs[16][DT + 4];  // [pixel][ci]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // tx: ci group, ty: co group
  const int ci_tiles = (g.Cin + DT - 1) / DT;
  const int co0 = (blockIdx.y / ci_tiles) * DT, ci0 = (blockIdx.y % ci_tiles) * DT;
  const int r = blockIdx.z / g.S, s = blockIdx.z % g.S;
  const int64_t npix = (int64_t)g.N * g.Ho * g.Wo;
  const int64_t pbeg = (int64_t)blockIdx.x * WG_PCH;
  const int64_t pend = min(pbeg + WG_PCH, npix);
  const bool loadD = tid < 128;
  const int lp = (tid & 127) >> 3, lv = tid & 7;  // pixel within chunk of 16, vector of 8 channels

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t pc = pbeg; pc < pend; pc += 16) {
    int64_t P = pc + lp;
    uint4 v = make_uint4(0, 0, 0, 0);
    int cbase = 0, climit = 0;
    if (P < pend) {
      int wo = (int)(P % g.Wo), ho = (int)((P / g.Wo) % g.Ho), n = (int)(P / ((int64_t)g.Wo * g.Ho));
      if (loadD) {
        cbase = co0 + lv * 8; climit = g.Cout;
        if (cbase < g.Cout) v = *reinterpret_cast<const uint4*>(dy + P * g.Cout_s + cbase);
      } else {
        cbase = ci0 + lv * 8; climit = g.Cin;
        int hs = ho * g.stride - g.pad + r * g.dil, ws = wo * g.stride - g.pad + s * g.dil;
        if (cbase < g.Cin && hs >= 0 && hs < g.H && ws >= 0 && ws < g.W)
          v = *reinterpret_cast<const uint4*>(x + (((int64_t)n * g.H + hs) * g.W + ws) * g.Cin_s + cbase);
      }
    }
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (cbase + k >= climit) f[k] = 0.f;
    __syncthreads();
    if (loadD) {
#pragma unroll
      for (int k = 0; k < 8; ++k) Ds[lp][lv * 8 + k] = f[k];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) Xs[lp][lv * 8 + k] = f[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Ds[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Xs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= g.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int ci = ci0 + tx * 4 + j;
      if (ci >= g.Cin) continue;
      atomicAdd(dw + (((int64_t)co * g.Cin + ci) * g.R + r) * g.S + s, acc[i][j]);
    }
  }
}

// per-channel column sum of an nhwc tensor: out[c] += sum_p t[p][c]  (bias gradients)
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ t, float* __restrict__ out, int64_t P,
                              int C, int Cs) {
  // block = 256 threads = 32 channel lanes x 8 pixel lanes
  __shared__ float red[8][33];
  int c = blockIdx.y * 32 + (threadIdx.x & 31);
  int pl = threadIdx.x >> 5;
  float s = 0.f;
  if (c < C)
    for (int64_t p = (int64_t)blockIdx.x * 8 + pl; p < P; p += (int64_t)gridDim.x * 8)
      s += bf2f(t[p * Cs + c]);
  red[pl][threadIdx.x & 31] = s;
  __syncthreads();
  if (pl == 0 && c < C) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += red[k][threadIdx.x];
    atomicAdd(out + c, r);
  }
}

int launch_direct_problem(const void* src, const void* w, const float* bias, void* out, int planar,
                          const void* addend, const TapProblem& p, int fmt, cudaStream_t st) {
  if (p.ntaps == 0) return MCD_OK;
  int64_t npix = (int64_t)p.N * p.Ht * p.Wt;
  int rows_cover = planar ? p.rows : p.Cd_s;
  dim3 grid((unsigned)((npix + DT - 1) / DT), (unsigned)((rows_cover + DT - 1) / DT));
  if (planar)
    conv_direct_kernel<true><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src,
                                                   (const __nv_bfloat16*)w, bias, out, nullptr, p, fmt);
  else
    conv_direct_kernel<false><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, (const __nv_bfloat16*)w, bias, out,
                                                    (const __nv_bfloat16*)addend, p, fmt);
  return check_launch("conv_direct");
}

int wgrad_direct(const void* x, const void* dy, float* dw, const mcd_conv_geom& g, int accumulate,
                 cudaStream_t st) {
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)g.Cout * g.Cin * g.R * g.S, st);
    if (e != cudaSuccess) { set_error("wgrad memset: %s", cudaGetErrorString(e)); return MCD_E_CUDA; }
  }
  int64_t npix = (int64_t)g.N * g.Ho * g.Wo;
  dim3 grid((unsigned)((npix + WG_PCH - 1) / WG_PCH),
            (unsigned)(((g.Cout + DT - 1) / DT) * ((g.Cin + DT - 1) / DT)), (unsigned)(g.R * g.S));
  wgrad_direct_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw, g);
  return check_launch("wgrad_direct");
}

int colsum(const void* t, float* out, int64_t P, int C, int Cs, int accumulate, cudaStream_t st) {
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C, st);
    if (e != cudaSuccess) { set_error("colsum memset: %s", cudaGetErrorString(e)); return MCD_E_CUDA; }
  }
  dim3 grid((unsigned)min64((P + 7) / 8, 148 * 4), (unsigned)((C + 31) / 32));
  colsum_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)t, out, P, C, Cs);
  return check_launch("colsum");
}

}  // namespace mcd
