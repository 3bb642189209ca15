// loss.cu — fused per-pixel losses on planar (NCHW) full-resolution logits.  Logits (and their gradients, which
// autograd requires to have the same dtype) are either bf16 - MCDStep's choice, half the bytes - or fp32, the
// drop-in default that keeps `outputs.data.cpu().numpy()` of the reference testers working (f32 flag of every entry
// point; the register-resident fast paths exist for bf16 only).
//   CrossEntropyLoss2d = log_softmax(dim=1) + NLLLoss2d(weight, mean)        loss.py:7-13
//   Diff2d             = mean |softmax(a) - softmax(b)|                      loss.py:93-100
//   F.mse_loss (HHA regression)                              models/dilated_fcn.py:712,958
//   sigmoid-average boundary head + bce2d       models/dilated_fcn.py:913-923, loss.py:131-138
//   tester argmax + calc_entropy                      adapt_tester.py:104-124, util.py:44-48
// One thread owns two horizontally adjacent pixels (4-byte bf16x2 loads, 128 B per warp and channel
// plane); channel loops re-read the logits from L1/L2, so DRAM sees each logit once per kernel.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"

namespace mcd {

__device__ __forceinline__ float2 ld2(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ void st2(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return bf2f(*p); }
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = f2bf(v); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }

// max and sum-exp over channels for a pixel pair
template <typename T>
__device__ __forceinline__ void softmax_stats(const T* base, int C, int64_t HW, float2* mx,
                                              float2* se) {
  float2 m = make_float2(-INFINITY, -INFINITY);
  for (int c = 0; c < C; ++c) {
    float2 v = ld2(base + c * HW);
    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y);
  }
  float2 s = make_float2(0.f, 0.f);
  for (int c = 0; c < C; ++c) {
    float2 v = ld2(base + c * HW);
    s.x += __expf(v.x - m.x); s.y += __expf(v.y - m.y);
  }
  *mx = m; *se = s;
}

struct LabelInfo { float w; int y; bool bad; };
__device__ __forceinline__ LabelInfo read_label(const int64_t* target, int64_t idx, const float* weight,
                                                int64_t ignore_index, int C) {
  LabelInfo li;
  int64_t y = target[idx];
  li.bad = false; li.w = 0.f; li.y = -1;
  if (y == ignore_index) return li;
  if (y < 0 || y >= C) { li.bad = true; return li; }
  li.y = (int)y;
  li.w = weight ? weight[y] : 1.f;
  return li;
}

// grid-stride over pixel pairs; npairs = N*H*W/2 (W even)
template <typename T>
__global__ void __launch_bounds__(256)
ce2d_fwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ target,
                const float* __restrict__ weight, int64_t ignore_index, float* __restrict__ acc, int C,
                int64_t HW, int64_t npairs) {
  __shared__ float red[32];
  float lsum = 0.f, wsum = 0.f, bad = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npairs;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i * 2;
    const int64_t n = pix / HW, hw = pix % HW;
    const T* base = logits + n * C * HW + hw;
    float2 m, s;
    softmax_stats(base, C, HW, &m, &s);
    LabelInfo l0 = read_label(target, pix, weight, ignore_index, C);
    LabelInfo l1 = read_label(target, pix + 1, weight, ignore_index, C);
    if (l0.y >= 0) { float xy = ld1(base + l0.y * HW); lsum += l0.w * (m.x + __logf(s.x) - xy); wsum += l0.w; }
    if (l1.y >= 0) { float xy = ld1(base + l1.y * HW + 1); lsum += l1.w * (m.y + __logf(s.y) - xy); wsum += l1.w; }
    bad += (l0.bad ? 1.f : 0.f) + (l1.bad ? 1.f : 0.f);
  }
  float r0 = block_sum(lsum, red), r1 = block_sum(wsum, red), r2 = block_sum(bad, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc + 0, r0); atomicAdd(acc + 1, r1);
    if (r2 != 0.f) atomicAdd(acc + 2, r2);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
ce2d_bwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ target,
                const float* __restrict__ weight, int64_t ignore_index, const float* __restrict__ acc,
                const float* __restrict__ gscale, T* __restrict__ dlogits, int C,
                int64_t HW, int64_t npairs) {
  const float wtot = acc[1];
  const float coef = wtot > 0.f ? gscale[0] / wtot : 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npairs;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i * 2;
    const int64_t n = pix / HW, hw = pix % HW;
    const T* base = logits + n * C * HW + hw;
    T* dbase = dlogits + n * C * HW + hw;
    float2 m, s;
    softmax_stats(base, C, HW, &m, &s);
    LabelInfo l0 = read_label(target, pix, weight, ignore_index, C);
    LabelInfo l1 = read_label(target, pix + 1, weight, ignore_index, C);
    const float k0 = coef * l0.w / s.x, k1 = coef * l1.w / s.y;
    for (int c = 0; c < C; ++c) {
      float2 v = ld2(base + c * HW);
      float g0 = k0 * __expf(v.x - m.x) - (c == l0.y ? coef * l0.w : 0.f);
      float g1 = k1 * __expf(v.y - m.y) - (c == l1.y ? coef * l1.w : 0.f);
      st2(dbase + c * HW, g0, g1);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
diff2d_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b,
                  float* __restrict__ acc, float4* __restrict__ stats, int C, int64_t HW, int64_t npairs) {
  __shared__ float red[32];
  float lsum = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npairs;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i * 2;
    const int64_t n = pix / HW, hw = pix % HW;
    const T* pa = a + n * C * HW + hw;
    const T* pb = b + n * C * HW + hw;
    float2 ma, sa, mb, sb;
    softmax_stats(pa, C, HW, &ma, &sa);
    softmax_stats(pb, C, HW, &mb, &sb);
    const float ia0 = 1.f / sa.x, ia1 = 1.f / sa.y, ib0 = 1.f / sb.x, ib1 = 1.f / sb.y;
    if (stats) {   // per-pixel (max_a, 1/sum_a, max_b, 1/sum_b): the backward kernel skips its two statistic passes
      stats[pix] = make_float4(ma.x, ia0, mb.x, ib0);
      stats[pix + 1] = make_float4(ma.y, ia1, mb.y, ib1);
    }
    for (int c = 0; c < C; ++c) {
      float2 va = ld2(pa + c * HW), vb = ld2(pb + c * HW);
      lsum += fabsf(__expf(va.x - ma.x) * ia0 - __expf(vb.x - mb.x) * ib0);
      lsum += fabsf(__expf(va.y - ma.y) * ia1 - __expf(vb.y - mb.y) * ib1);
    }
  }
  float r = block_sum(lsum, red);
  if (threadIdx.x == 0) atomicAdd(acc, r);
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

template <typename T>
__global__ void __launch_bounds__(256)
diff2d_bwd_kernel(const T* __restrict__ a, const T* __restrict__ b,
                  const float* __restrict__ gscale, const float4* __restrict__ stats,
                  T* __restrict__ da, T* __restrict__ db, float inv_numel, int C,
                  int64_t HW, int64_t npairs) {
  const float coef = gscale[0] * inv_numel;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npairs;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i * 2;
    const int64_t n = pix / HW, hw = pix % HW;
    const int64_t off = n * C * HW + hw;
    const T* pa = a + off;
    const T* pb = b + off;
    float2 ma, mb;
    float ia0, ia1, ib0, ib1;
    if (stats) {
      const float4 s0 = stats[pix], s1 = stats[pix + 1];
      ma = make_float2(s0.x, s1.x); mb = make_float2(s0.z, s1.z);
      ia0 = s0.y; ia1 = s1.y; ib0 = s0.w; ib1 = s1.w;
    } else {
      float2 sa, sb;
      softmax_stats(pa, C, HW, &ma, &sa);
      softmax_stats(pb, C, HW, &mb, &sb);
      ia0 = 1.f / sa.x; ia1 = 1.f / sa.y; ib0 = 1.f / sb.x; ib1 = 1.f / sb.y;
    }
    // dot products  sum_c sign_c * p_c  for both distributions
    float da0 = 0.f, da1 = 0.f, db0 = 0.f, db1 = 0.f;
    for (int c = 0; c < C; ++c) {
      float2 va = ld2(pa + c * HW), vb = ld2(pb + c * HW);
      float pa0 = __expf(va.x - ma.x) * ia0, pb0 = __expf(vb.x - mb.x) * ib0;
      float pa1 = __expf(va.y - ma.y) * ia1, pb1 = __expf(vb.y - mb.y) * ib1;
      float s0 = sgn(pa0 - pb0), s1 = sgn(pa1 - pb1);
      da0 += s0 * pa0; db0 += s0 * pb0; da1 += s1 * pa1; db1 += s1 * pb1;
    }
    for (int c = 0; c < C; ++c) {
      float2 va = ld2(pa + c * HW), vb = ld2(pb + c * HW);
      float pa0 = __expf(va.x - ma.x) * ia0, pb0 = __expf(vb.x - mb.x) * ib0;
      float pa1 = __expf(va.y - ma.y) * ia1, pb1 = __expf(vb.y - mb.y) * ib1;
      float s0 = sgn(pa0 - pb0), s1 = sgn(pa1 - pb1);
      st2(da + off + c * HW, coef * pa0 * (s0 - da0), coef * pa1 * (s1 - da1));
      st2(db + off + c * HW, coef * pb0 * (db0 - s0), coef * pb1 * (db1 - s1));
    }
  }
}

// ---- the other discrepancy criteria of loss.py:68-171 (get_prob_distance_criterion names other than 'diff'): all are
// per-pixel functions of the two softmax distributions averaged over N*C*H*W elements.  With pa = softmax(a), la = log pa:
//   MODE 1  symkl / nmlsymkl (Symkl2d :103-117) and mysymkl (MySymkl2d :141-151):  0.5 * sum_c (pa - pb)(la - lb)
//           (the view(-1, n_target_ch) of Symkl2d regroups elements of an element-wise mean: no effect)
//   MODE 2  jsd (JSD :79-90):  0.5 * sum_c [ pa (la - lm) + pb (lb - lm) ],  lm = log_softmax((a + b) / 2)
//   MODE 3  mis_symkl (MisSymKLD :68-76) and spatial_jsd (SpatialJSD2d :154-170): F.kl_div fed with PROBABILITIES where it
//           expects log-probabilities:  0.5 * sum_c [ pb (lb - pa) + pa (la - pb) ]
// Gradients are the exact derivatives through both arguments (torch >= 1.x differentiates F.kl_div w.r.t. its target):
//   1: da_c = 0.5 [ pa_c (la_c - lb_c) + pa_c - pb_c - pa_c KL(a||b) ]                    (db: a <-> b)
//   2: da_c = 0.5 pa_c [ (la_c - lm_c) - KL(a||m) ] - 0.25 (pa_c + pb_c - 2 pm_c)        (db: a <-> b)
//   3: da_c = 0.5 pa_c [ (la_c + 1 - 2 pb_c) - (sum_k pa_k la_k + 1 - 2 <pa, pb>) ]       (db: a <-> b)
// Multi-pass over the channel planes of a pixel pair (the re-reads hit L1 / L2); these criteria are options of the
// trainers (--d_loss), not the default path.
struct PixSoft { float m, ls; };        // max and log(sum exp): log p_c = v_c - m - ls

template <typename T, int MODE>
__global__ void __launch_bounds__(256)
pairdist_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ acc, const float* __restrict__ gscale,
                T* __restrict__ da, T* __restrict__ db, float inv_numel, int C, int64_t HW, int64_t npairs) {
  __shared__ float red[32];
  const bool bwd = da != nullptr;
  const float coef = bwd ? gscale[0] * inv_numel : 0.f;
  float lsum = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npairs; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i * 2;
    const int64_t n = pix / HW, hw = pix % HW, off = n * C * HW + hw;
    const T* pa = a + off;
    const T* pb = b + off;
    // pass 1: softmax statistics of a, b (and of m = (a + b) / 2 for the Jensen-Shannon form)
    float ma[2] = {-INFINITY, -INFINITY}, mb[2] = {-INFINITY, -INFINITY}, mm[2] = {-INFINITY, -INFINITY};
    for (int c = 0; c < C; ++c) {
      const float2 va = ld2(pa + c * HW), vb = ld2(pb + c * HW);
      ma[0] = fmaxf(ma[0], va.x); ma[1] = fmaxf(ma[1], va.y);
      mb[0] = fmaxf(mb[0], vb.x); mb[1] = fmaxf(mb[1], vb.y);
      if (MODE == 2) { mm[0] = fmaxf(mm[0], 0.5f * (va.x + vb.x)); mm[1] = fmaxf(mm[1], 0.5f * (va.y + vb.y)); }
    }
    float sa[2] = {0.f, 0.f}, sb[2] = {0.f, 0.f}, sm[2] = {0.f, 0.f};
    for (int c = 0; c < C; ++c) {
      const float2 va = ld2(pa + c * HW), vb = ld2(pb + c * HW);
      sa[0] += __expf(va.x - ma[0]); sa[1] += __expf(va.y - ma[1]);
      sb[0] += __expf(vb.x - mb[0]); sb[1] += __expf(vb.y - mb[1]);
      if (MODE == 2) { sm[0] += __expf(0.5f * (va.x + vb.x) - mm[0]); sm[1] += __expf(0.5f * (va.y + vb.y) - mm[1]); }
    }
    float lsa[2], lsb[2], lsm[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) { lsa[h] = __logf(sa[h]); lsb[h] = __logf(sb[h]); lsm[h] = MODE == 2 ? __logf(sm[h]) : 0.f; }
    // pass 2: the per-pixel scalars (loss term and the inner products the gradients need)
    float t0[2] = {0.f, 0.f}, t1[2] = {0.f, 0.f}, t2[2] = {0.f, 0.f};
    for (int c = 0; c < C; ++c) {
      const float2 va2 = ld2(pa + c * HW), vb2 = ld2(pb + c * HW);
      const float va[2] = {va2.x, va2.y}, vb[2] = {vb2.x, vb2.y};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float la = va[h] - ma[h] - lsa[h], lb = vb[h] - mb[h] - lsb[h];
        const float qa = __expf(la), qb = __expf(lb);
        if (MODE == 1) { t0[h] += qa * (la - lb); t1[h] += qb * (lb - la); }            // KL(a||b), KL(b||a)
        if (MODE == 2) {
          const float lm = 0.5f * (va[h] + vb[h]) - mm[h] - lsm[h];
          t0[h] += qa * (la - lm); t1[h] += qb * (lb - lm);                              // KL(a||m), KL(b||m)
        }
        if (MODE == 3) { t0[h] += qa * la; t1[h] += qb * lb; t2[h] += qa * qb; }         // -H(a), -H(b), <pa, pb>
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (MODE == 1 || MODE == 2) lsum += 0.5f * (t0[h] + t1[h]);
      if (MODE == 3) lsum += 0.5f * (t0[h] + t1[h] - 2.f * t2[h]);
    }
    if (!bwd) continue;
    // pass 3: gradients
    for (int c = 0; c < C; ++c) {
      const float2 va2 = ld2(pa + c * HW), vb2 = ld2(pb + c * HW);
      const float va[2] = {va2.x, va2.y}, vb[2] = {vb2.x, vb2.y};
      float ga[2], gb[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float la = va[h] - ma[h] - lsa[h], lb = vb[h] - mb[h] - lsb[h];
        const float qa = __expf(la), qb = __expf(lb);
        if (MODE == 1) {
          ga[h] = 0.5f * (qa * (la - lb) + qa - qb - qa * t0[h]);
          gb[h] = 0.5f * (qb * (lb - la) + qb - qa - qb * t1[h]);
        } else if (MODE == 2) {
          const float lm = 0.5f * (va[h] + vb[h]) - mm[h] - lsm[h];
          const float mix = 0.25f * (qa + qb - 2.f * __expf(lm));
          ga[h] = 0.5f * qa * ((la - lm) - t0[h]) - mix;
          gb[h] = 0.5f * qb * ((lb - lm) - t1[h]) - mix;
        } else {
          ga[h] = 0.5f * qa * ((la + 1.f - 2.f * qb) - (t0[h] + 1.f - 2.f * t2[h]));
          gb[h] = 0.5f * qb * ((lb + 1.f - 2.f * qa) - (t1[h] + 1.f - 2.f * t2[h]));
        }
      }
      st2(da + off + c * HW, coef * ga[0], coef * ga[1]);
      st2(db + off + c * HW, coef * gb[0], coef * gb[1]);
    }
  }
  if (acc) {
    const float r = block_sum(lsum, red);
    if (threadIdx.x == 0) atomicAdd(acc, r);
  }
}

// ---- register-resident variants (C <= kRegC): a thread loads all channels of its pixel pair ONCE (C independent
// 4-byte loads in flight per tensor), then max / sum-exp / loss / gradients come from registers: one pass over the
// logits with no dependent re-reads.  Used for the MCD heads (C = 41); wider heads take the multi-pass kernels.
constexpr int kRegC = 42;   // >= the 41 classes of the MCD heads; wider heads take the multi-pass kernels
constexpr uint32_t kNegInfPair = 0xFF80FF80u;   // two bf16 -inf: padding channels vanish in max and sum-exp

__device__ __forceinline__ float lo_f(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float hi_f(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ void load_pair(const __nv_bfloat16* base, int C, int64_t HW, uint32_t (&r)[kRegC]) {
#pragma unroll
  for (int c = 0; c < kRegC; ++c)
    r[c] = (c < C) ? __ldg(reinterpret_cast<const unsigned int*>(base + c * HW)) : kNegInfPair;
}
__device__ __forceinline__ void reg_softmax_stats(const uint32_t (&r)[kRegC], float2* mx, float2* se) {
  float2 m = make_float2(-INFINITY, -INFINITY);
#pragma unroll
  for (int c = 0; c < kRegC; ++c) { m.x = fmaxf(m.x, lo_f(r[c])); m.y = fmaxf(m.y, hi_f(r[c])); }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < kRegC; ++c) { s.x += __expf(lo_f(r[c]) - m.x); s.y += __expf(hi_f(r[c]) - m.y); }
  *mx = m; *se = s;
}

// in place: r[c] <- half2(exp(v_lo - m.x), exp(v_hi - m.y)); returns the two sums.  One MUFU per logit - the
// probabilities are then re-used from registers (fp16 keeps 11 bits: 5e-4 relative, below the bf16 gradients'
// own rounding and averaged out in the loss sums) instead of being recomputed in every later pass.
__device__ __forceinline__ float2 reg_exp_inplace(uint32_t (&r)[kRegC], float2 m) {
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < kRegC; ++c) {
    const float e0 = __expf(lo_f(r[c]) - m.x), e1 = __expf(hi_f(r[c]) - m.y);
    s.x += e0; s.y += e1;
    const __half2 h = __floats2half2_rn(e0, e1);
    r[c] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return s;
}
__device__ __forceinline__ float2 reg_max(const uint32_t (&r)[kRegC]) {
  float2 m = make_float2(-INFINITY, -INFINITY);
#pragma unroll
  for (int c = 0; c < kRegC; ++c) { m.x = fmaxf(m.x, lo_f(r[c])); m.y = fmaxf(m.y, hi_f(r[c])); }
  return m;
}
__device__ __forceinline__ float2 h2f(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }

__global__ void __launch_bounds__(128, 3)
ce2d_fwd_reg_kernel(const __nv_bfloat16* __restrict__ logits, const int64_t* __restrict__ target,
                    const float* __restrict__ weight, int64_t ignore_index, float* __restrict__ acc, int C,
                    int64_t HW, int64_t npairs) {
  __shared__ float red[32];
  float lsum = 0.f, wsum = 0.f, bad = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npairs;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i * 2;
    const int64_t n = pix / HW, hw = pix % HW;
    const __nv_bfloat16* base = logits + n * C * HW + hw;
    uint32_t r[kRegC];
    load_pair(base, C, HW, r);
    LabelInfo l0 = read_label(target, pix, weight, ignore_index, C);
    LabelInfo l1 = read_label(target, pix + 1, weight, ignore_index, C);
    float2 m, s;
    reg_softmax_stats(r, &m, &s);
    float x0 = 0.f, x1 = 0.f;
#pragma unroll
    for (int c = 0; c < kRegC; ++c) {
      x0 = (c == l0.y) ? lo_f(r[c]) : x0;
      x1 = (c == l1.y) ? hi_f(r[c]) : x1;
    }
    if (l0.y >= 0) { lsum += l0.w * (m.x + __logf(s.x) - x0); wsum += l0.w; }
    if (l1.y >= 0) { lsum += l1.w * (m.y + __logf(s.y) - x1); wsum += l1.w; }
    bad += (l0.bad ? 1.f : 0.f) + (l1.bad ? 1.f : 0.f);
  }
  float r0 = block_sum(lsum, red), r1 = block_sum(wsum, red), r2 = block_sum(bad, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc + 0, r0); atomicAdd(acc + 1, r1);
    if (r2 != 0.f) atomicAdd(acc + 2, r2);
  }
}

__global__ void __launch_bounds__(128, 2)
diff2d_fwd_reg_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                      float* __restrict__ acc, float4* __restrict__ stats, int C, int64_t HW, int64_t npairs) {
  __shared__ float red[32];
  float lsum = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npairs;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i * 2;
    const int64_t n = pix / HW, hw = pix % HW;
    uint32_t ra[kRegC], rb[kRegC];
    load_pair(a + n * C * HW + hw, C, HW, ra);
    load_pair(b + n * C * HW + hw, C, HW, rb);
    const float2 ma = reg_max(ra), mb = reg_max(rb);
    const float2 sa = reg_exp_inplace(ra, ma), sb = reg_exp_inplace(rb, mb);
    const float ia0 = 1.f / sa.x, ia1 = 1.f / sa.y, ib0 = 1.f / sb.x, ib1 = 1.f / sb.y;
    if (stats) {
      stats[pix] = make_float4(ma.x, ia0, mb.x, ib0);
      stats[pix + 1] = make_float4(ma.y, ia1, mb.y, ib1);
    }
#pragma unroll
    for (int c = 0; c < kRegC; ++c) {
      const float2 ea = h2f(ra[c]), eb = h2f(rb[c]);
      lsum += fabsf(ea.x * ia0 - eb.x * ib0) + fabsf(ea.y * ia1 - eb.y * ib1);
    }
  }
  float r = block_sum(lsum, red);
  if (threadIdx.x == 0) atomicAdd(acc, r);
}

// register-resident backward of Diff2d (C <= kRegC): both heads' logits of a pixel pair are loaded ONCE, the
// probabilities exp(v - max) are computed ONCE (kept as half2 like the forward kernel) and both passes - the dot
// products sum_c sign_c p_c and the gradients - run from registers.  The multi-pass kernel above reads every logit
// twice and evaluates 4 exp per logit.  Opt-in (MCD_DIFF2D_BWD_REG=1): measured slower than the multi-pass kernel.
__global__ void __launch_bounds__(128, 2)
diff2d_bwd_reg_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                      const float* __restrict__ gscale, __nv_bfloat16* __restrict__ da,
                      __nv_bfloat16* __restrict__ db, float inv_numel, int C, int64_t HW, int64_t npairs) {
  const float coef = gscale[0] * inv_numel;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npairs;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i * 2;
    const int64_t n = pix / HW, hw = pix % HW;
    const int64_t off = n * C * HW + hw;
    uint32_t ra[kRegC], rb[kRegC];
    load_pair(a + off, C, HW, ra);
    load_pair(b + off, C, HW, rb);
    const float2 ma = reg_max(ra), mb = reg_max(rb);
    const float2 sa = reg_exp_inplace(ra, ma), sb = reg_exp_inplace(rb, mb);
    const float ia0 = 1.f / sa.x, ia1 = 1.f / sa.y, ib0 = 1.f / sb.x, ib1 = 1.f / sb.y;
    float da0 = 0.f, da1 = 0.f, db0 = 0.f, db1 = 0.f;
#pragma unroll
    for (int c = 0; c < kRegC; ++c) {     // padding channels hold exp(-inf) = 0: sign(0 - 0) = 0, no contribution
      const float2 ea = h2f(ra[c]), eb = h2f(rb[c]);
      const float pa0 = ea.x * ia0, pb0 = eb.x * ib0, pa1 = ea.y * ia1, pb1 = eb.y * ib1;
      const float s0 = sgn(pa0 - pb0), s1 = sgn(pa1 - pb1);
      da0 = fmaf(s0, pa0, da0); db0 = fmaf(s0, pb0, db0);
      da1 = fmaf(s1, pa1, da1); db1 = fmaf(s1, pb1, db1);
    }
#pragma unroll
    for (int c = 0; c < kRegC; ++c) {
      if (c < C) {
        const float2 ea = h2f(ra[c]), eb = h2f(rb[c]);
        const float pa0 = ea.x * ia0, pb0 = eb.x * ib0, pa1 = ea.y * ia1, pb1 = eb.y * ib1;
        const float s0 = sgn(pa0 - pb0), s1 = sgn(pa1 - pb1);
        st2(da + off + c * HW, coef * pa0 * (s0 - da0), coef * pa1 * (s1 - da1));
        st2(db + off + c * HW, coef * pb0 * (db0 - s0), coef * pb1 * (db1 - s1));
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
mse_fwd_kernel(const T* __restrict__ pred, const float* __restrict__ target,
               float* __restrict__ acc, int64_t numel) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel;
       i += (int64_t)gridDim.x * blockDim.x) {
    float d = ld1(pred + i) - target[i];
    s = fmaf(d, d, s);
  }
  float r = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(acc, r);
}

template <typename T>
__global__ void __launch_bounds__(256)
mse_bwd_kernel(const T* __restrict__ pred, const float* __restrict__ target,
               const float* __restrict__ gscale, T* __restrict__ dpred, float inv_numel,
               int64_t numel) {
  const float coef = 2.f * gscale[0] * inv_numel;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel;
       i += (int64_t)gridDim.x * blockDim.x)
    st1(dpred + i, coef * (ld1(pred + i) - target[i]));
}

__global__ void __launch_bounds__(256)
sum_f32_kernel(const float* __restrict__ x, float* __restrict__ acc, int64_t numel) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel;
       i += (int64_t)gridDim.x * blockDim.x)
    s += x[i];
  float r = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(acc, r);
}

__device__ __forceinline__ float sigmoidf(float v) { return 1.f / (1.f + __expf(-v)); }

template <typename T>
__global__ void __launch_bounds__(256)
sigmoid3_bce_fwd_kernel(const T* __restrict__ h1, const T* __restrict__ h2,
                        const T* __restrict__ h3, const float* __restrict__ target,
                        const float* __restrict__ tsum, float* __restrict__ acc,
                        T* __restrict__ p_out, float inv_global, int64_t numel) {
  __shared__ float red[32];
  // beta = 1 - mean(target) over the GLOBAL batch (loss.py:133; tsum is all-reduced under data parallelism)
  const float beta = tsum ? 1.f - tsum[0] * inv_global : 0.f;
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel;
       i += (int64_t)gridDim.x * blockDim.x) {
    float p = (sigmoidf(ld1(h1 + i)) + sigmoidf(ld1(h2 + i)) + sigmoidf(ld1(h3 + i))) * (1.f / 3.f);
    if (p_out) st1(p_out + i, p);
    if (target) {
      float t = target[i];
      float w = 1.f - beta + (2.f * beta - 1.f) * t;
      float lp = fmaxf(__logf(p), -100.f), lq = fmaxf(__logf(1.f - p), -100.f);
      s -= w * (t * lp + (1.f - t) * lq);
    }
  }
  if (acc) {
    float r = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(acc, r);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
sigmoid3_bce_bwd_kernel(const T* __restrict__ h1, const T* __restrict__ h2,
                        const T* __restrict__ h3, const float* __restrict__ target,
                        const float* __restrict__ tsum, const float* __restrict__ gscale,
                        T* __restrict__ dh1, T* __restrict__ dh2,
                        T* __restrict__ dh3, float inv_global, int64_t numel) {
  const float beta = 1.f - tsum[0] * inv_global;
  const float coef = gscale[0] / (float)numel;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel;
       i += (int64_t)gridDim.x * blockDim.x) {
    float s1 = sigmoidf(ld1(h1 + i)), s2 = sigmoidf(ld1(h2 + i)), s3 = sigmoidf(ld1(h3 + i));
    float p = (s1 + s2 + s3) * (1.f / 3.f);
    float t = target[i];
    float w = 1.f - beta + (2.f * beta - 1.f) * t;
    // F.binary_cross_entropy backward: w * (p - t) / max(p*(1-p), 1e-12)
    float dp = coef * w * (p - t) / fmaxf(p * (1.f - p), 1e-12f) * (1.f / 3.f);
    st1(dh1 + i, dp * s1 * (1.f - s1));
    st1(dh2 + i, dp * s2 * (1.f - s2));
    st1(dh3 + i, dp * s3 * (1.f - s3));
  }
}

// bce2d on a probability map (loss.py:130-138): beta = 1 - mean(t); w = 1 - beta + (2 beta - 1) t;
// F.binary_cross_entropy(p, t, w) with torch's log clamp at -100.  acc[0] += sum w * bce
__global__ void __launch_bounds__(256)
bce2d_fwd_kernel(const float* __restrict__ p, const float* __restrict__ target, const float* __restrict__ tsum,
                 float* __restrict__ acc, float inv_global, int64_t numel) {
  __shared__ float red[32];
  const float beta = 1.f - tsum[0] * inv_global;
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
    const float pi = p[i], t = target[i];
    const float w = 1.f - beta + (2.f * beta - 1.f) * t;
    const float lp = fmaxf(logf(pi), -100.f), lq = fmaxf(logf(1.f - pi), -100.f);
    s -= w * (t * lp + (1.f - t) * lq);
  }
  float r = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(acc, r);
}

// dp = gscale / numel * w * (p - t) / max(p (1 - p), 1e-12)   (F.binary_cross_entropy backward)
__global__ void __launch_bounds__(256)
bce2d_bwd_kernel(const float* __restrict__ p, const float* __restrict__ target, const float* __restrict__ tsum,
                 const float* __restrict__ gscale, float* __restrict__ dp, float inv_global, int64_t numel) {
  const float beta = 1.f - tsum[0] * inv_global;
  const float coef = gscale[0] / (float)numel;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
    const float pi = p[i], t = target[i];
    const float w = 1.f - beta + (2.f * beta - 1.f) * t;
    dp[i] = coef * w * (pi - t) / fmaxf(pi * (1.f - pi), 1e-12f);
  }
}

// morphological boundary of a label / value map (models/dilated_fcn.py:769-773 get_boundary): out = 1 where the 3x3
// maximum differs from the 3x3 minimum (max_pool2d of x and of -x, padding excluded), else 0.  T = int64 labels or
// fp32 values; comparisons are exact in either type.
template <typename T>
__global__ void __launch_bounds__(256)
label_boundary_kernel(const T* __restrict__ x, float* __restrict__ out, int H, int W, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W), h = (int)((i / W) % H);
    const T* base = x + (i - (int64_t)h * W - w);
    T mx = base[(int64_t)h * W + w], mn = mx;
#pragma unroll
    for (int dh = -1; dh <= 1; ++dh) {
      const int hh = h + dh;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int dw = -1; dw <= 1; ++dw) {
        const int ww = w + dw;
        if (ww < 0 || ww >= W) continue;
        const T v = base[(int64_t)hh * W + ww];
        mx = v > mx ? v : mx;
        mn = v < mn ? v : mn;
      }
    }
    out[i] = mx != mn ? 1.f : 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
argmax_entropy_kernel(const T* __restrict__ logits, int64_t* __restrict__ labels,
                      float* __restrict__ acc, int C, int C_arg, int64_t HW, int64_t npairs) {
  __shared__ float red[32];
  float esum = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npairs;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i * 2;
    const int64_t n = pix / HW, hw = pix % HW;
    const T* base = logits + n * C * HW + hw;
    float2 m, s;
    softmax_stats(base, C, HW, &m, &s);
    float b0 = -INFINITY, b1 = -INFINITY;
    int a0 = 0, a1 = 0;
    const float i0 = 1.f / s.x, i1 = 1.f / s.y;
    for (int c = 0; c < C; ++c) {
      float2 v = ld2(base + c * HW);
      if (c < C_arg) {
        if (v.x > b0) { b0 = v.x; a0 = c; }
        if (v.y > b1) { b1 = v.y; a1 = c; }
      }
      float p0 = __expf(v.x - m.x) * i0, p1 = __expf(v.y - m.y) * i1;
      esum += p0 * __logf(p0 + 1e-6f) + p1 * __logf(p1 + 1e-6f);
    }
    if (labels) { labels[pix] = a0; labels[pix + 1] = a1; }
  }
  if (acc) {
    float r = block_sum(esum, red);
    if (threadIdx.x == 0) atomicAdd(acc, r);
  }
}

__global__ void __launch_bounds__(256)
sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                int64_t numel, float lr, float momentum, float wd, int first) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel;
       i += (int64_t)gridDim.x * blockDim.x) {
    float d = fmaf(wd, p[i], g[i]);
    float b = d;
    if (momentum != 0.f) {
      b = first ? d : fmaf(momentum, buf[i], d);
      buf[i] = b;
    }
    p[i] -= lr * b;
  }
}

static inline int grid_for(int64_t items) {
  return (int)max64(1, min64((items + 255) / 256, 148 * 16));
}
static inline int grid_reg(int64_t items) {   // 128-thread blocks, 3 per SM, a few waves
  return (int)max64(1, min64((items + 127) / 128, 148 * 3 * 8));
}

}  // namespace mcd

using namespace mcd;

#define MCD_CHECK_PLANAR(name)                                                                   \
  MCD_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, name ": bad sizes");                             \
  MCD_REQUIRE(W % 2 == 0, name ": W must be even (pixel-pair vectorisation), got %d", W)

typedef __nv_bfloat16 bf16_t;

template <typename T>
static int launch_pairdist(int mode, const T* a, const T* b, float* acc, const float* gscale, T* da, T* db, float inv_numel,
                           int C, int64_t HW, int64_t npairs, cudaStream_t st) {
  const int grid = grid_for(npairs);
  if (mode == 1) pairdist_kernel<T, 1><<<grid, 256, 0, st>>>(a, b, acc, gscale, da, db, inv_numel, C, HW, npairs);
  else if (mode == 2) pairdist_kernel<T, 2><<<grid, 256, 0, st>>>(a, b, acc, gscale, da, db, inv_numel, C, HW, npairs);
  else pairdist_kernel<T, 3><<<grid, 256, 0, st>>>(a, b, acc, gscale, da, db, inv_numel, C, HW, npairs);
  return check_launch("pairdist");
}


extern "C" {

int mcd_ce2d_fwd(const void* logits, int f32, const int64_t* target, const float* weight,
                 int64_t ignore_index, float* acc, int N, int C, int H, int W, int device,
                 void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(logits && target && acc, "ce2d_fwd: null pointer");
  MCD_CHECK_PLANAR("ce2d_fwd");
  int64_t npairs = (int64_t)N * H * W / 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (f32)
    ce2d_fwd_kernel<float><<<grid_for(npairs), 256, 0, st>>>((const float*)logits, target, weight, ignore_index, acc,
                                                             C, (int64_t)H * W, npairs);
  else if (C <= kRegC)
    ce2d_fwd_reg_kernel<<<grid_reg(npairs), 128, 0, st>>>((const bf16_t*)logits, target, weight, ignore_index, acc, C,
                                                          (int64_t)H * W, npairs);
  else
    ce2d_fwd_kernel<bf16_t><<<grid_for(npairs), 256, 0, st>>>((const bf16_t*)logits, target, weight, ignore_index,
                                                              acc, C, (int64_t)H * W, npairs);
  return check_launch("ce2d_fwd");
}

int mcd_ce2d_bwd(const void* logits, int f32, const int64_t* target, const float* weight,
                 int64_t ignore_index, const float* acc, const float* gscale, void* dlogits, int N,
                 int C, int H, int W, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(logits && target && acc && gscale && dlogits, "ce2d_bwd: null pointer");
  MCD_CHECK_PLANAR("ce2d_bwd");
  int64_t npairs = (int64_t)N * H * W / 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (f32)
    ce2d_bwd_kernel<float><<<grid_for(npairs), 256, 0, st>>>((const float*)logits, target, weight, ignore_index, acc,
                                                             gscale, (float*)dlogits, C, (int64_t)H * W, npairs);
  else
    ce2d_bwd_kernel<bf16_t><<<grid_for(npairs), 256, 0, st>>>((const bf16_t*)logits, target, weight, ignore_index,
                                                              acc, gscale, (bf16_t*)dlogits, C, (int64_t)H * W,
                                                              npairs);
  return check_launch("ce2d_bwd");
}

int mcd_diff2d_fwd(const void* a, const void* b, int f32, float* acc, float* stats, int N, int C, int H, int W,
                   int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(a && b && acc, "diff2d_fwd: null pointer");
  MCD_CHECK_PLANAR("diff2d_fwd");
  int64_t npairs = (int64_t)N * H * W / 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (f32)
    diff2d_fwd_kernel<float><<<grid_for(npairs), 256, 0, st>>>((const float*)a, (const float*)b, acc, (float4*)stats,
                                                               C, (int64_t)H * W, npairs);
  else if (C <= kRegC)
    diff2d_fwd_reg_kernel<<<grid_reg(npairs), 128, 0, st>>>((const bf16_t*)a, (const bf16_t*)b, acc, (float4*)stats,
                                                            C, (int64_t)H * W, npairs);
  else
    diff2d_fwd_kernel<bf16_t><<<grid_for(npairs), 256, 0, st>>>((const bf16_t*)a, (const bf16_t*)b, acc,
                                                                (float4*)stats, C, (int64_t)H * W, npairs);
  return check_launch("diff2d_fwd");
}

int mcd_diff2d_bwd(const void* a, const void* b, int f32, const float* gscale, const float* stats, void* da,
                   void* db, int N, int C, int H, int W, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(a && b && gscale && da && db, "diff2d_bwd: null pointer");
  MCD_CHECK_PLANAR("diff2d_bwd");
  int64_t npairs = (int64_t)N * H * W / 2;
  float inv = (float)(1.0 / ((double)N * C * H * W));
  cudaStream_t st = (cudaStream_t)stream;
  if (f32) {
    diff2d_bwd_kernel<float><<<grid_for(npairs), 256, 0, st>>>((const float*)a, (const float*)b, gscale,
                                                               (const float4*)stats, (float*)da, (float*)db, inv, C,
                                                               (int64_t)H * W, npairs);
    return check_launch("diff2d_bwd");
  }
  static int use_reg = -1;
  // measured (r01b, 22 pairs): register-resident 836 us vs multi-pass 760 us per launch - 84 live packed logits
  // + the two passes spill at 255 registers; opt-in only
  if (use_reg < 0) { const char* e = getenv("MCD_DIFF2D_BWD_REG"); use_reg = (e && e[0] == '1') ? 1 : 0; }
  if (use_reg && C <= kRegC) {
    diff2d_bwd_reg_kernel<<<grid_reg(npairs), 128, 0, st>>>((const bf16_t*)a, (const bf16_t*)b, gscale, (bf16_t*)da,
                                                            (bf16_t*)db, inv, C, (int64_t)H * W, npairs);
    return check_launch("diff2d_bwd");
  }
  diff2d_bwd_kernel<bf16_t><<<grid_for(npairs), 256, 0, st>>>((const bf16_t*)a, (const bf16_t*)b, gscale,
                                                              (const float4*)stats, (bf16_t*)da, (bf16_t*)db, inv, C,
                                                              (int64_t)H * W, npairs);
  return check_launch("diff2d_bwd");
}

int mcd_pairdist(int mode, const void* a, const void* b, int f32, float* acc, const float* gscale, void* da, void* db,
                 float inv_numel, int N, int C, int H, int W, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(a && b && (acc || da), "pairdist: null pointer");
  MCD_REQUIRE(mode >= 1 && mode <= 3, "pairdist: mode %d (1 symmetric KL, 2 Jensen-Shannon, 3 kl_div on probabilities)", mode);
  MCD_REQUIRE((da != nullptr) == (db != nullptr) && (!da || gscale), "pairdist: da, db and gscale go together");
  MCD_CHECK_PLANAR("pairdist");
  const int64_t HW = (int64_t)H * W, npairs = (int64_t)N * HW / 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (f32)
    return launch_pairdist<float>(mode, (const float*)a, (const float*)b, acc, gscale, (float*)da, (float*)db, inv_numel, C,
                                  HW, npairs, st);
  return launch_pairdist<bf16_t>(mode, (const bf16_t*)a, (const bf16_t*)b, acc, gscale, (bf16_t*)da, (bf16_t*)db, inv_numel,
                                 C, HW, npairs, st);
}

int mcd_mse_fwd(const void* pred, int f32, const float* target, float* acc, int64_t numel, int device,
                void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(pred && target && acc && numel > 0, "mse_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (f32) mse_fwd_kernel<float><<<grid_for(numel), 256, 0, st>>>((const float*)pred, target, acc, numel);
  else mse_fwd_kernel<bf16_t><<<grid_for(numel), 256, 0, st>>>((const bf16_t*)pred, target, acc, numel);
  return check_launch("mse_fwd");
}

int mcd_mse_bwd(const void* pred, int f32, const float* target, const float* gscale, void* dpred,
                int64_t numel, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(pred && target && gscale && dpred && numel > 0, "mse_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const float inv = (float)(1.0 / (double)numel);
  if (f32) mse_bwd_kernel<float><<<grid_for(numel), 256, 0, st>>>((const float*)pred, target, gscale, (float*)dpred,
                                                                  inv, numel);
  else mse_bwd_kernel<bf16_t><<<grid_for(numel), 256, 0, st>>>((const bf16_t*)pred, target, gscale, (bf16_t*)dpred,
                                                               inv, numel);
  return check_launch("mse_bwd");
}

int mcd_sum_f32(const float* x, float* acc, int64_t numel, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(x && acc && numel > 0, "sum_f32: bad arguments");
  sum_f32_kernel<<<grid_for(numel), 256, 0, (cudaStream_t)stream>>>(x, acc, numel);
  return check_launch("sum_f32");
}

int mcd_sigmoid3_bce_fwd(const void* h1, const void* h2, const void* h3, int f32, const float* target,
                         const float* tsum, float* acc, void* p_out, int64_t numel, int64_t numel_global,
                         int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(h1 && h2 && h3 && numel > 0, "sigmoid3_bce_fwd: bad arguments");
  MCD_REQUIRE(!target || (tsum && acc), "sigmoid3_bce_fwd: target needs tsum and acc");
  MCD_REQUIRE(target || p_out, "sigmoid3_bce_fwd: nothing to compute");
  cudaStream_t st = (cudaStream_t)stream;
  const float inv_global = (float)(1.0 / (double)(numel_global > 0 ? numel_global : numel));
  if (f32)
    sigmoid3_bce_fwd_kernel<float><<<grid_for(numel), 256, 0, st>>>(
        (const float*)h1, (const float*)h2, (const float*)h3, target, tsum, target ? acc : nullptr, (float*)p_out,
        inv_global, numel);
  else
    sigmoid3_bce_fwd_kernel<bf16_t><<<grid_for(numel), 256, 0, st>>>(
        (const bf16_t*)h1, (const bf16_t*)h2, (const bf16_t*)h3, target, tsum, target ? acc : nullptr,
        (bf16_t*)p_out, inv_global, numel);
  return check_launch("sigmoid3_bce_fwd");
}

int mcd_sigmoid3_bce_bwd(const void* h1, const void* h2, const void* h3, int f32, const float* target,
                         const float* tsum, const float* gscale, void* dh1, void* dh2, void* dh3,
                         int64_t numel, int64_t numel_global, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(h1 && h2 && h3 && target && tsum && gscale && dh1 && dh2 && dh3 && numel > 0,
              "sigmoid3_bce_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const float inv_global = (float)(1.0 / (double)(numel_global > 0 ? numel_global : numel));
  if (f32)
    sigmoid3_bce_bwd_kernel<float><<<grid_for(numel), 256, 0, st>>>(
        (const float*)h1, (const float*)h2, (const float*)h3, target, tsum, gscale, (float*)dh1, (float*)dh2,
        (float*)dh3, inv_global, numel);
  else
    sigmoid3_bce_bwd_kernel<bf16_t><<<grid_for(numel), 256, 0, st>>>(
        (const bf16_t*)h1, (const bf16_t*)h2, (const bf16_t*)h3, target, tsum, gscale, (bf16_t*)dh1, (bf16_t*)dh2,
        (bf16_t*)dh3, inv_global, numel);
  return check_launch("sigmoid3_bce_bwd");
}

int mcd_bce2d_fwd(const float* p, const float* target, const float* tsum, float* acc, int64_t numel,
                  int64_t numel_global, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(p && target && tsum && acc && numel > 0, "bce2d_fwd: bad arguments");
  const float inv_global = (float)(1.0 / (double)(numel_global > 0 ? numel_global : numel));
  bce2d_fwd_kernel<<<grid_for(numel), 256, 0, (cudaStream_t)stream>>>(p, target, tsum, acc, inv_global, numel);
  return check_launch("bce2d_fwd");
}

int mcd_bce2d_bwd(const float* p, const float* target, const float* tsum, const float* gscale, float* dp,
                  int64_t numel, int64_t numel_global, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(p && target && tsum && gscale && dp && numel > 0, "bce2d_bwd: bad arguments");
  const float inv_global = (float)(1.0 / (double)(numel_global > 0 ? numel_global : numel));
  bce2d_bwd_kernel<<<grid_for(numel), 256, 0, (cudaStream_t)stream>>>(p, target, tsum, gscale, dp, inv_global, numel);
  return check_launch("bce2d_bwd");
}

int mcd_label_boundary(const void* x, int is_int64, float* out, int N, int H, int W, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(x && out && N > 0 && H > 0 && W > 0, "label_boundary: bad arguments");
  const int64_t total = (int64_t)N * H * W;
  if (is_int64)
    label_boundary_kernel<int64_t><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const int64_t*)x, out, H, W, total);
  else
    label_boundary_kernel<float><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const float*)x, out, H, W, total);
  return check_launch("label_boundary");
}

int mcd_argmax_entropy(const void* logits, int f32, int64_t* labels, float* acc, int N, int C, int C_arg,
                       int H, int W, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(logits && (labels || acc), "argmax_entropy: null pointer");
  MCD_CHECK_PLANAR("argmax_entropy");
  MCD_REQUIRE(C_arg >= 1 && C_arg <= C, "argmax_entropy: C_arg=%d out of range", C_arg);
  int64_t npairs = (int64_t)N * H * W / 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (f32)
    argmax_entropy_kernel<float><<<grid_for(npairs), 256, 0, st>>>((const float*)logits, labels, acc, C, C_arg,
                                                                   (int64_t)H * W, npairs);
  else
    argmax_entropy_kernel<bf16_t><<<grid_for(npairs), 256, 0, st>>>((const bf16_t*)logits, labels, acc, C, C_arg,
                                                                    (int64_t)H * W, npairs);
  return check_launch("argmax_entropy");
}

int mcd_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t numel, float lr,
                 float momentum, float weight_decay, int first_step, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(param && grad && numel > 0, "sgd_step: bad arguments");
  MCD_REQUIRE(momentum == 0.f || momentum_buf, "sgd_step: momentum needs a buffer");
  sgd_step_kernel<<<grid_for(numel), 256, 0, (cudaStream_t)stream>>>(param, grad, momentum_buf, numel,
                                                                      lr, momentum, weight_decay,
                                                                      first_step);
  return check_launch("sgd_step");
}

}  // extern "C"
