// variants.cu — the small memory-bound kernels behind the option surface around the MCD hot path (SURVEY.md
// section 8 row f4): the Gate / Concat fusion heads (models/fusion.py:6-50), FuseDRNSegBase's per-stage adds
// (models/dilated_fcn.py:294-331), nn.UpsamplingBilinear2d (align_corners=True; `use_torch_up`, :354-355,443-444),
// F.softmax over channels (ScoreGateFusion, fusion.py:13-15), F.sigmoid of the seg -> boundary convolution
// (dilated_fcn.py:965-966) and ProbCrossEntropyLoss2d (loss.py:16-30, the Gate fusions' criterion).  Element-wise kernels
// have 16-byte vector bodies with scalar tails; measured bandwidths: profiles/r02f_bench_variants.txt.
#include "common.cuh"

namespace mcd {

// ---- z = a + b on nhwc IEEE-half activations, written as IEEE half and (optionally) the bf16 twin ----------------
__global__ void __launch_bounds__(256)
add_nhwc_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out16,
                uint4* __restrict__ twin, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float fa[8], fb[8];
    unpack8h(a[i], fa);
    unpack8h(b[i], fb);
#pragma unroll
    for (int k = 0; k < 8; ++k) fa[k] += fb[k];
    out16[i] = pack8h(fa);
    if (twin) twin[i] = pack8(fa);
  }
}

// ---- out = x1 * sigmoid(a) + x2 * (1 - sigmoid(a))   (GateFusion, models/fusion.py:17-21) ---------------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

// element-wise kernels: a float4 body over the first 4 * n4 elements (n4 = 0 when a pointer is not 16-byte aligned)
// and a scalar tail; 16 bytes per load keep enough bytes in flight for the HBM latency at 2048 threads per SM
__device__ __forceinline__ float4 ld4(const float* p, int64_t i) { return reinterpret_cast<const float4*>(p)[i]; }
__device__ __forceinline__ void st4(float* p, int64_t i, float4 v) { reinterpret_cast<float4*>(p)[i] = v; }
__device__ __forceinline__ float gate1(float x1, float x2, float a) {
  const float g = sigmoidf_(a);
  return x1 * g + x2 * (1.f - g);
}

__global__ void __launch_bounds__(256)
gate_fwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ a,
                float* __restrict__ out, int64_t n, int64_t n4) {
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = t0; i < n4; i += stride) {
    const float4 u = ld4(x1, i), v = ld4(x2, i), w = ld4(a, i);
    st4(out, i, make_float4(gate1(u.x, v.x, w.x), gate1(u.y, v.y, w.y), gate1(u.z, v.z, w.z), gate1(u.w, v.w, w.w)));
  }
  for (int64_t i = 4 * n4 + t0; i < n; i += stride) out[i] = gate1(x1[i], x2[i], a[i]);
}

__device__ __forceinline__ void gate_bwd1(float x1, float x2, float a, float d, float* g1, float* g2, float* ga) {
  const float g = sigmoidf_(a);
  *g1 = d * g;
  *g2 = d * (1.f - g);
  *ga = d * (x1 - x2) * g * (1.f - g);
}

__global__ void __launch_bounds__(256)
gate_bwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ a,
                const float* __restrict__ dout, float* __restrict__ dx1, float* __restrict__ dx2,
                float* __restrict__ da, int64_t n, int64_t n4) {
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = t0; i < n4; i += stride) {
    const float4 u = ld4(x1, i), v = ld4(x2, i), w = ld4(a, i), d = ld4(dout, i);
    float4 r1, r2, ra;
    gate_bwd1(u.x, v.x, w.x, d.x, &r1.x, &r2.x, &ra.x);
    gate_bwd1(u.y, v.y, w.y, d.y, &r1.y, &r2.y, &ra.y);
    gate_bwd1(u.z, v.z, w.z, d.z, &r1.z, &r2.z, &ra.z);
    gate_bwd1(u.w, v.w, w.w, d.w, &r1.w, &r2.w, &ra.w);
    if (dx1) st4(dx1, i, r1);
    if (dx2) st4(dx2, i, r2);
    if (da) st4(da, i, ra);
  }
  for (int64_t i = 4 * n4 + t0; i < n; i += stride) {
    float r1, r2, ra;
    gate_bwd1(x1[i], x2[i], a[i], dout[i], &r1, &r2, &ra);
    if (dx1) dx1[i] = r1;
    if (dx2) dx2[i] = r2;
    if (da) da[i] = ra;
  }
}

// ---- softmax over the channel axis of a planar fp32 tensor (one thread per pixel, coalesced across pixels) -------
// C <= SM_CMAX (= 44): the channel vector of the pixel lives in registers - ONE pass over the input, all C loads of a thread
// independent and in flight together; larger C: three passes (the re-reads hit L2).
constexpr int SM_CMAX = 44;          // 41 classes padded to a multiple of 4

template <bool REG>
__global__ void __launch_bounds__(256)
softmax_ch_fwd_kernel(const float* __restrict__ x, float* __restrict__ p, int C, int64_t HW, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW, px = i % HW;
    const float* xp = x + n * C * HW + px;
    float* pp = p + n * C * HW + px;
    if (REG) {
      float v[SM_CMAX];
#pragma unroll
      for (int c = 0; c < SM_CMAX; ++c) v[c] = c < C ? xp[c * HW] : -INFINITY;
      float m = -INFINITY;
#pragma unroll
      for (int c = 0; c < SM_CMAX; ++c) m = fmaxf(m, v[c]);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < SM_CMAX; ++c) { v[c] = __expf(v[c] - m); s += v[c]; }      // exp(-inf) = 0 for the padding
      const float inv = 1.f / s;
#pragma unroll
      for (int c = 0; c < SM_CMAX; ++c)
        if (c < C) pp[c * HW] = v[c] * inv;
    } else {
      float m = -INFINITY;
      for (int c = 0; c < C; ++c) m = fmaxf(m, xp[c * HW]);
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += __expf(xp[c * HW] - m);
      const float inv = 1.f / s;
      for (int c = 0; c < C; ++c) pp[c * HW] = __expf(xp[c * HW] - m) * inv;
    }
  }
}

// dx = p * (dp - sum_c dp * p).  Two passes over p and dp (the second hits L1 / L2): measured 4.3 TB/s of algorithmic
// traffic; register-resident variants were slower (both vectors: 138 registers, one block per SM, 3.7 TB/s; p only with
// dp re-read: 2.6 TB/s) - scripts/bench_variants.py
__global__ void __launch_bounds__(256)
softmax_ch_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp, float* __restrict__ dx, int C,
                      int64_t HW, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW, px = i % HW;
    const int64_t base = n * C * HW + px;
    float dot = 0.f;
    for (int c = 0; c < C; ++c) dot = fmaf(dp[base + c * HW], p[base + c * HW], dot);
    for (int c = 0; c < C; ++c) dx[base + c * HW] = p[base + c * HW] * (dp[base + c * HW] - dot);
  }
}

// ---- torch.cat([a, b], 1) and its backward (split) on planar fp32 tensors ------------------------------------------
// cat == 1: ab[n][0:Ca] = a[n], ab[n][Ca:] = b[n];  cat == 0: the reverse copy (a / b may be null).
// T = float4 when the per-image element counts are multiples of 4 and the pointers 16-byte aligned (ea / eb / total
// are then counted in float4 units), else float.
template <typename T>
__global__ void __launch_bounds__(256)
cat2_kernel(T* __restrict__ a, T* __restrict__ b, T* __restrict__ ab, int64_t ea, int64_t eb, int64_t total, int cat) {
  const int64_t e = ea + eb;   // elements per image
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / e, r = i % e;
    T* part = r < ea ? (a ? a + n * ea + r : nullptr) : (b ? b + n * eb + (r - ea) : nullptr);
    if (!part) continue;
    if (cat) ab[i] = *part; else *part = ab[i];
  }
}

// ---- sigmoid ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int64_t n4) {
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = t0; i < n4; i += stride) {
    const float4 v = ld4(x, i);
    st4(y, i, make_float4(sigmoidf_(v.x), sigmoidf_(v.y), sigmoidf_(v.z), sigmoidf_(v.w)));
  }
  for (int64_t i = 4 * n4 + t0; i < n; i += stride) y[i] = sigmoidf_(x[i]);
}
__global__ void __launch_bounds__(256)
sigmoid_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, int64_t n,
                   int64_t n4) {
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = t0; i < n4; i += stride) {
    const float4 v = ld4(y, i), d = ld4(dy, i);
    st4(dx, i, make_float4(d.x * v.x * (1.f - v.x), d.y * v.y * (1.f - v.y), d.z * v.z * (1.f - v.z),
                           d.w * v.w * (1.f - v.w)));
  }
  for (int64_t i = 4 * n4 + t0; i < n; i += stride) dx[i] = dy[i] * y[i] * (1.f - y[i]);
}

// ---- out = a + b (+ c) on fp32 tensors (the shortcut decoders' h1 + h2 + h3, dilated_fcn.py:875,884,904) -----------
__global__ void __launch_bounds__(256)
add3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
            float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a[i] + b[i] + (c ? c[i] : 0.f);
}

// ---- bilinear upsample, align_corners=True (nn.UpsamplingBilinear2d) -------------------------------------------------
// ATen's area_pixel_compute_scale: scale = (in - 1) / (out - 1) for out > 1, else 0; src = scale * o, in float.
__device__ __forceinline__ void bil_src_ac(int o, float scale, int in, int* i0, int* i1, float* lam) {
  const float src = scale * (float)o;
  int a = (int)src;
  if (a > in - 1) a = in - 1;
  *i0 = a;
  *i1 = a + (a < in - 1 ? 1 : 0);
  *lam = src - (float)a;
}

// V = 8: a thread produces 8 adjacent output pixels of a row (W % 8 == 0) and stores them as 2 x float4 / one 16-byte
// bf16 vector; V = 1: any width.  The gathers hit L1 / L2 (the source is s^2 times smaller than the output).
template <bool OUT_F32, int V>
__global__ void __launch_bounds__(256)
bilinear_ac_fwd_kernel(const float* __restrict__ x, void* __restrict__ out, int h, int wd, int H, int W, float sh,
                       float sw, int64_t total) {
  const int wv = W / V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % wv);
    const int oh = (int)((i / wv) % H);
    const int64_t nc = i / ((int64_t)wv * H);
    int h0, h1;
    float lh;
    bil_src_ac(oh, sh, h, &h0, &h1, &lh);
    const float* r0 = x + (nc * h + h0) * wd;
    const float* r1 = x + (nc * h + h1) * wd;
    const float l0h = 1.f - lh;
    float f[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      int w0, w1;
      float lw;
      bil_src_ac(j * V + k, sw, wd, &w0, &w1, &lw);
      // ATen's weighting order: h0lambda * (w0lambda * p00 + w1lambda * p01) + h1lambda * (w0lambda * p10 + w1lambda * p11)
      const float l0w = 1.f - lw;
      f[k] = l0h * (l0w * r0[w0] + lw * r0[w1]) + lh * (l0w * r1[w0] + lw * r1[w1]);
    }
    const int64_t o = (nc * H + oh) * (int64_t)W + (int64_t)j * V;
    if (V == 8) {
      if (OUT_F32) {
        float* op = reinterpret_cast<float*>(out) + o;
        *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(op + 4) = make_float4(f[4 % V], f[5 % V], f[6 % V], f[7 % V]);
      } else {
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + o) = pack8(f);
      }
    } else {
      if (OUT_F32) reinterpret_cast<float*>(out)[o] = f[0];
      else reinterpret_cast<__nv_bfloat16*>(out)[o] = f2bf(f[0]);
    }
  }
}

// block = one (n*c, ih) low-resolution row, separable gather like bilinear_bwd_kernel (heads.cu)
template <bool IN_F32>
__global__ void __launch_bounds__(256)
bilinear_ac_bwd_kernel(const void* __restrict__ dout, float* __restrict__ dx, int h, int wd, int H, int W, float sh,
                       float sw) {
  extern __shared__ float col[];  // W floats
  const int64_t nc = blockIdx.x;
  const int ih = blockIdx.y;
  int oh_lo = 0, oh_hi = H - 1;
  if (sh > 0.f) {
    oh_lo = max(0, (int)floorf((float)(ih - 1) / sh) - 1);
    oh_hi = min(H - 1, (int)ceilf((float)(ih + 1) / sh) + 1);
  }
  for (int ow = threadIdx.x; ow < W; ow += blockDim.x) {
    float acc = 0.f;
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      int h0, h1; float lh;
      bil_src_ac(oh, sh, h, &h0, &h1, &lh);
      const float wgt = (h0 == ih ? 1.f - lh : 0.f) + (h1 == ih ? lh : 0.f);
      if (wgt == 0.f) continue;
      const int64_t o = (nc * H + oh) * (int64_t)W + ow;
      const float v = IN_F32 ? reinterpret_cast<const float*>(dout)[o]
                             : bf2f(reinterpret_cast<const __nv_bfloat16*>(dout)[o]);
      acc = fmaf(wgt, v, acc);
    }
    col[ow] = acc;
  }
  __syncthreads();
  for (int iw = threadIdx.x; iw < wd; iw += blockDim.x) {
    int ow_lo = 0, ow_hi = W - 1;
    if (sw > 0.f) {
      ow_lo = max(0, (int)floorf((float)(iw - 1) / sw) - 1);
      ow_hi = min(W - 1, (int)ceilf((float)(iw + 1) / sw) + 1);
    }
    float acc = 0.f;
    for (int ow = ow_lo; ow <= ow_hi; ++ow) {
      int w0, w1; float lw;
      bil_src_ac(ow, sw, wd, &w0, &w1, &lw);
      const float wgt = (w0 == iw ? 1.f - lw : 0.f) + (w1 == iw ? lw : 0.f);
      acc = fmaf(wgt, col[ow], acc);
    }
    dx[(nc * h + ih) * wd + iw] = acc;
  }
}

// ---- ProbCrossEntropyLoss2d (loss.py:16-30): NLLLoss2d(weight, mean)(log(p), target) on a probability map -----------
// acc[0] += sum w[y] * -log p[y], acc[1] += sum w[y], acc[2] += labels outside [0, C) and != ignore_index
__global__ void __launch_bounds__(256)
prob_ce_fwd_kernel(const float* __restrict__ p, const int64_t* __restrict__ target, const float* __restrict__ weight,
                   int64_t ignore_index, float* __restrict__ acc, int C, int64_t HW, int64_t total) {
  __shared__ float red[32];
  float s_l = 0.f, s_w = 0.f, s_bad = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t y = target[i];
    if (y == ignore_index) continue;
    if (y < 0 || y >= C) { s_bad += 1.f; continue; }
    const int64_t n = i / HW, px = i % HW;
    const float w = weight ? weight[y] : 1.f;
    s_l -= w * logf(p[(n * C + y) * HW + px]);
    s_w += w;
  }
  float t = block_sum(s_l, red);
  if (threadIdx.x == 0 && t != 0.f) atomicAdd(acc + 0, t);
  t = block_sum(s_w, red);
  if (threadIdx.x == 0 && t != 0.f) atomicAdd(acc + 1, t);
  t = block_sum(s_bad, red);
  if (threadIdx.x == 0 && t != 0.f) atomicAdd(acc + 2, t);
}

// dp[n,c,px] = c == y ? -gscale * w[y] / (p[y] * acc[1]) : 0   (acc[1] = 1 for size_average=False)
// VEC = 4: a thread owns 4 adjacent pixels (HW % 4 == 0, 16-byte aligned dp) and writes one float4 per channel
template <int VEC>
__global__ void __launch_bounds__(256)
prob_ce_bwd_kernel(const float* __restrict__ p, const int64_t* __restrict__ target, const float* __restrict__ weight,
                   int64_t ignore_index, const float* __restrict__ acc, const float* __restrict__ gscale,
                   float* __restrict__ dp, int C, int64_t HW, int64_t total) {
  const float scale = gscale[0] / acc[1];
  const int64_t groups = total / VEC;
  for (int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; gi < groups; gi += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = gi * VEC;
    const int64_t n = i / HW, px = i % HW;
    int yy[VEC];
    float gy[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int64_t y = target[i + k];
      const bool valid = y != ignore_index && y >= 0 && y < C;
      yy[k] = valid ? (int)y : -1;
      gy[k] = valid ? -scale * (weight ? weight[y] : 1.f) / p[(n * C + y) * HW + px + k] : 0.f;
    }
    for (int c = 0; c < C; ++c) {
      float* o = dp + (n * C + c) * HW + px;
      if (VEC == 4) {
        *reinterpret_cast<float4*>(o) = make_float4(yy[0] == c ? gy[0] : 0.f, yy[1] == c ? gy[1] : 0.f,
                                                    yy[2] == c ? gy[2] : 0.f, yy[3] == c ? gy[3] : 0.f);
      } else {
        o[0] = yy[0] == c ? gy[0] : 0.f;
      }
    }
  }
}

static inline int grid_for(int64_t n) { return (int)max64(1, min64((n + 255) / 256, 148 * 16)); }
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace mcd

using namespace mcd;

extern "C" {

int mcd_add_nhwc(const void* a_f16, const void* b_f16, void* out_f16, void* out_bf16, int64_t numel, int device,
                 void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(a_f16 && b_f16 && out_f16 && numel > 0 && numel % 8 == 0, "add_nhwc: bad arguments");
  add_nhwc_kernel<<<grid_for(numel / 8), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)a_f16, (const uint4*)b_f16, (uint4*)out_f16, (uint4*)out_bf16, numel / 8);
  return check_launch("add_nhwc");
}

int mcd_gate_fuse_fwd(const float* x1, const float* x2, const float* gate_logits, float* out, int64_t numel,
                      int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(x1 && x2 && gate_logits && out && numel > 0, "gate_fuse_fwd: bad arguments");
  const int64_t n4 = (al16(x1) && al16(x2) && al16(gate_logits) && al16(out)) ? numel / 4 : 0;
  gate_fwd_kernel<<<grid_for(n4 ? n4 : numel), 256, 0, (cudaStream_t)stream>>>(x1, x2, gate_logits, out, numel, n4);
  return check_launch("gate_fuse_fwd");
}

int mcd_gate_fuse_bwd(const float* x1, const float* x2, const float* gate_logits, const float* dout, float* dx1,
                      float* dx2, float* dgate_logits, int64_t numel, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(x1 && x2 && gate_logits && dout && numel > 0, "gate_fuse_bwd: bad arguments");
  const int64_t n4 = (al16(x1) && al16(x2) && al16(gate_logits) && al16(dout) && al16(dx1) && al16(dx2) &&
                      al16(dgate_logits)) ? numel / 4 : 0;
  gate_bwd_kernel<<<grid_for(n4 ? n4 : numel), 256, 0, (cudaStream_t)stream>>>(x1, x2, gate_logits, dout, dx1, dx2,
                                                                               dgate_logits, numel, n4);
  return check_launch("gate_fuse_bwd");
}

int mcd_softmax_ch_fwd(const float* x, float* p, int N, int C, int64_t HW, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(x && p && N > 0 && C > 0 && HW > 0, "softmax_ch_fwd: bad arguments");
  const int64_t total = (int64_t)N * HW;
  if (C <= SM_CMAX)
    softmax_ch_fwd_kernel<true><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, p, C, HW, total);
  else
    softmax_ch_fwd_kernel<false><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, p, C, HW, total);
  return check_launch("softmax_ch_fwd");
}

int mcd_softmax_ch_bwd(const float* p, const float* dp, float* dx, int N, int C, int64_t HW, int device,
                       void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(p && dp && dx && N > 0 && C > 0 && HW > 0, "softmax_ch_bwd: bad arguments");
  const int64_t total = (int64_t)N * HW;
  softmax_ch_bwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(p, dp, dx, C, HW, total);
  return check_launch("softmax_ch_bwd");
}

int mcd_cat2_f32(const float* a, int Ca, const float* b, int Cb, float* out, int N, int64_t HW, int device,
                 void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(a && b && out && N > 0 && Ca > 0 && Cb > 0 && HW > 0, "cat2_f32: bad arguments");
  const int64_t total = (int64_t)N * (Ca + Cb) * HW, ea = Ca * HW, eb = Cb * HW;
  float *pa = const_cast<float*>(a), *pb = const_cast<float*>(b);
  if (ea % 4 == 0 && eb % 4 == 0 && al16(a) && al16(b) && al16(out))
    cat2_kernel<float4><<<grid_for(total / 4), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(pa), reinterpret_cast<float4*>(pb), reinterpret_cast<float4*>(out), ea / 4, eb / 4,
        total / 4, 1);
  else
    cat2_kernel<float><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(pa, pb, out, ea, eb, total, 1);
  return check_launch("cat2_f32");
}

int mcd_split2_f32(const float* src, float* a, int Ca, float* b, int Cb, int N, int64_t HW, int device,
                   void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(src && (a || b) && N > 0 && Ca > 0 && Cb > 0 && HW > 0, "split2_f32: bad arguments");
  const int64_t total = (int64_t)N * (Ca + Cb) * HW, ea = Ca * HW, eb = Cb * HW;
  float* ps = const_cast<float*>(src);
  if (ea % 4 == 0 && eb % 4 == 0 && al16(a) && al16(b) && al16(src))
    cat2_kernel<float4><<<grid_for(total / 4), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(a), reinterpret_cast<float4*>(b), reinterpret_cast<float4*>(ps), ea / 4, eb / 4,
        total / 4, 0);
  else
    cat2_kernel<float><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(a, b, ps, ea, eb, total, 0);
  return check_launch("split2_f32");
}

int mcd_sigmoid_fwd(const float* x, float* y, int64_t numel, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(x && y && numel > 0, "sigmoid_fwd: bad arguments");
  const int64_t n4 = (al16(x) && al16(y)) ? numel / 4 : 0;
  sigmoid_fwd_kernel<<<grid_for(n4 ? n4 : numel), 256, 0, (cudaStream_t)stream>>>(x, y, numel, n4);
  return check_launch("sigmoid_fwd");
}

int mcd_sigmoid_bwd(const float* y, const float* dy, float* dx, int64_t numel, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(y && dy && dx && numel > 0, "sigmoid_bwd: bad arguments");
  const int64_t n4 = (al16(y) && al16(dy) && al16(dx)) ? numel / 4 : 0;
  sigmoid_bwd_kernel<<<grid_for(n4 ? n4 : numel), 256, 0, (cudaStream_t)stream>>>(y, dy, dx, numel, n4);
  return check_launch("sigmoid_bwd");
}

int mcd_add3_f32(const float* a, const float* b, const float* c, float* out, int64_t numel, int device,
                 void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(a && b && out && numel > 0, "add3_f32: bad arguments");
  add3_kernel<<<grid_for(numel), 256, 0, (cudaStream_t)stream>>>(a, b, c, out, numel);
  return check_launch("add3_f32");
}

int mcd_prob_ce2d_fwd(const float* p, const int64_t* target, const float* weight, int64_t ignore_index, float* acc,
                      int N, int C, int H, int W, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(p && target && acc && N > 0 && C > 0 && H > 0 && W > 0, "prob_ce2d_fwd: bad arguments");
  const int64_t HW = (int64_t)H * W, total = (int64_t)N * HW;
  prob_ce_fwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(p, target, weight, ignore_index, acc, C, HW,
                                                                        total);
  return check_launch("prob_ce2d_fwd");
}

int mcd_prob_ce2d_bwd(const float* p, const int64_t* target, const float* weight, int64_t ignore_index,
                      const float* acc, const float* gscale, float* dp, int N, int C, int H, int W, int device,
                      void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(p && target && acc && gscale && dp && N > 0 && C > 0 && H > 0 && W > 0, "prob_ce2d_bwd: bad arguments");
  const int64_t HW = (int64_t)H * W, total = (int64_t)N * HW;
  if (HW % 4 == 0 && al16(dp))
    prob_ce_bwd_kernel<4><<<grid_for(total / 4), 256, 0, (cudaStream_t)stream>>>(p, target, weight, ignore_index, acc,
                                                                                 gscale, dp, C, HW, total);
  else
    prob_ce_bwd_kernel<1><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(p, target, weight, ignore_index, acc,
                                                                             gscale, dp, C, HW, total);
  return check_launch("prob_ce2d_bwd");
}

int mcd_bilinear_ac_up_fwd(const float* x, void* out, int out_f32, int N, int C, int h, int w_, int s, int device,
                           void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(x && out && N > 0 && C > 0 && h > 0 && w_ > 0 && s >= 1, "bilinear_ac_up_fwd: bad arguments");
  const int H = h * s, W = w_ * s;
  const float sh = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sw = W > 1 ? (float)(w_ - 1) / (float)(W - 1) : 0.f;
  const int64_t total = (int64_t)N * C * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  if (W % 8 == 0 && al16(out)) {
    if (out_f32) bilinear_ac_fwd_kernel<true, 8><<<grid_for(total / 8), 256, 0, st>>>(x, out, h, w_, H, W, sh, sw, total / 8);
    else bilinear_ac_fwd_kernel<false, 8><<<grid_for(total / 8), 256, 0, st>>>(x, out, h, w_, H, W, sh, sw, total / 8);
  } else {
    if (out_f32) bilinear_ac_fwd_kernel<true, 1><<<grid_for(total), 256, 0, st>>>(x, out, h, w_, H, W, sh, sw, total);
    else bilinear_ac_fwd_kernel<false, 1><<<grid_for(total), 256, 0, st>>>(x, out, h, w_, H, W, sh, sw, total);
  }
  return check_launch("bilinear_ac_up_fwd");
}

int mcd_bilinear_ac_up_bwd(const void* dout, int dout_f32, float* dx, int N, int C, int h, int w_, int s, int device,
                           void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(dout && dx && N > 0 && C > 0 && h > 0 && w_ > 0 && s >= 1, "bilinear_ac_up_bwd: bad arguments");
  const int H = h * s, W = w_ * s;
  MCD_REQUIRE((size_t)W * sizeof(float) <= 48 * 1024, "bilinear_ac_up_bwd: output rows wider than 12288 unsupported");
  const float sh = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sw = W > 1 ? (float)(w_ - 1) / (float)(W - 1) : 0.f;
  dim3 grid((unsigned)(N * C), (unsigned)h);
  const size_t smem = sizeof(float) * (size_t)W;
  if (dout_f32)
    bilinear_ac_bwd_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(dout, dx, h, w_, H, W, sh, sw);
  else
    bilinear_ac_bwd_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(dout, dx, h, w_, H, W, sh, sw);
  return check_launch("bilinear_ac_up_bwd");
}

}  // extern "C"
