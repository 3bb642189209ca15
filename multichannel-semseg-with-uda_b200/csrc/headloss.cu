// headloss.cu — classifier head (x8 upsampling) FUSED with its per-pixel loss, forward and backward in ONE kernel:
// the full-resolution logits (41 x 480 x 640 per image and head: the largest tensors of an MCD iteration) and their
// gradients are never materialised.  Replaces, per loss evaluation, the chain
//     deconv16s8_fwd -> {ce2d | diff2d}_fwd -> {ce2d | diff2d}_bwd -> deconv16s8_bwd_dx (-> deconv16s8_bwd_dw)
// (5.5 GB of HBM traffic per phase at 22 images) by a kernel whose only traffic is the 60x80 score maps, the labels
// and the 41 x 16 x 16 filters.
//   heads  : learned depthwise ConvTranspose2d(C,C,16,s8,p4,groups=C) of DRNSegPixelClassifier /
//            FusionDRNSegPixelClassifier / ScoreFusionDRNSegPixelClassifier (models/dilated_fcn.py:357-366,465-491:
//            one or two (input, filter) pairs per head), or nn.Upsample(x8, bilinear, align_corners=False) of the
//            multitask decoders (:676,817-819) - "filter == NULL"
//   losses : CrossEntropyLoss2d (loss.py:7-13; one head) or Diff2d (loss.py:93-100; two heads)
// Output pixel (oh, ow) reads the 2x2 inputs x[ih0-a][iw0-b], ih0 = (oh+4)>>3, with filter taps (kh0+8a, kw0+8b),
// kh0 = (oh+4)&7: a "cell" (ci, cj) = the 8x8 output pixels with ih0 = ci, iw0 = cj shares its four inputs.  A block
// of 256 threads works on 4 horizontally adjacent cells at a time (thread = one output pixel):
//   phase 1  logits of all channels in registers (filters and inputs staged in shared memory as float4 per (channel,
//            position) / (channel, cell)), softmax, loss term, d(loss)/d(logit) -> shared memory (bf16)
//   phase 2  the transposed products that turn the logit gradients into score-map gradients (work unit = channel x
//            quarter of the 64 positions, 4 cells x 4 taps partial sums -> a 2 x 5 tile per channel in shared memory ->
//            one global atomicAdd per score-map element of the item) and filter gradients (thread t owns bin (position
//            t / 4, tap t % 4) of every channel in REGISTERS over all items of the persistent block; one atomicAdd per
//            filter element and block at the end).
// Gradients are produced for an upstream gradient of 1; the autograd Function scales them (mcd_b200/headloss.py).
#include "common.cuh"

namespace mcd {

constexpr int HL_THREADS = 256;
constexpr int HL_CELLS = 4;          // cells per block iteration
constexpr int HL_MAXC = 44;          // register arrays: classes padded to a multiple of 4 (41 -> 44)
constexpr int HL_WS = 260;           // floats per channel of a staged filter: [64 pos][4 taps] + 4 (bank spreading)
constexpr int HL_DL = 260;           // bf16 per channel of the staged logit gradients: [64 pos][4 cells] + 4

struct HeadIn {
  const float* x;     // [N, C, h, w] score map (NULL: input absent)
  const float* w;     // [C, 256] filter (NULL: bilinear upsampling)
  float* dx;          // [N, C, h, w] += d loss / d x      (NULL: not wanted)
  float* dw;          // [C, 256]     += d loss / d w      (NULL: not wanted)
};

struct HeadLossArgs {
  int N, C, h, w;
  int mode;                    // 0: cross entropy on head 0;  1: Diff2d between head 0 and head 1
  int nheads, nin;             // heads (1 / 2), inputs per head (1 / 2)
  HeadIn in[2][2];             // [head][input]
  // cross entropy
  const int64_t* target;       // [N, 8h, 8w]
  const float* cls_weight;     // [C] or NULL
  int64_t ignore_index;
  const float* wsum;           // device scalar: the (global) normaliser sum_i w[y_i]; NULL: 1 (size_average = False)
  // Diff2d
  float inv_numel;             // 1 / (N * C * H * W * world)
  float* acc;                  // [0] += loss numerator, [2] += number of bad labels (CE)
  int want_grad;               // any dx / dw wanted
  int total_items, groups_per_row;
  // shared memory carve-up (in floats), filled by the host
  int off_w, off_x, off_xt, off_dl, off_dx, smem_bytes;
};

// shared-memory fp32 atomic add (native RED.shared)
__device__ __forceinline__ void sm_add(float* p, float v) { atomicAdd(p, v); }

// DWN: number of (head, input) pairs whose filter gradient is accumulated in registers (0 = none wanted)
// CT / NIT: class count and inputs per head as COMPILE-TIME constants (0 = take them from the arguments).  The kernel is
// issue-bound and, with run-time C and nin, two thirds of the instructions of its inner loops were integer address
// arithmetic and bounds predicates (scripts/gpu_hl_exp.sh: 41 x 2 x (2 loads + 4 FMAs) cost 1400 instructions per thread);
// with CT = 41, NIT = 1 (every MCD head of the reference) the shared-memory offsets become immediates.
template <int MODE, bool BILINEAR, int DWN, int CT, int NIT>
__global__ void __launch_bounds__(HL_THREADS, 1)
head_loss_kernel(const __grid_constant__ HeadLossArgs a) {
  constexpr bool DW = DWN > 0;
  extern __shared__ __align__(16) uint8_t hl_smem[];
  float* ws = reinterpret_cast<float*>(hl_smem) + a.off_w;      // [head][in][C][64 pos][4 taps] (+4 per channel)
  float* xs = reinterpret_cast<float*>(hl_smem) + a.off_x;      // [head][in][C][cell][tap]
  float* xt = reinterpret_cast<float*>(hl_smem) + a.off_xt;     // [head][in][C][tap][cell]      (filter gradients)
  __nv_bfloat16* dls = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<float*>(hl_smem) + a.off_dl);   // [head][C][pos][cell]
  float* dxs = reinterpret_cast<float*>(hl_smem) + a.off_dx;    // [head][in][C][2 rows][5 cols] score-map gradients of this item
  __shared__ float red[32];
  const int C = CT > 0 ? CT : a.C, h = a.h, w = a.w, H = 8 * a.h, W = 8 * a.w;
  const int tid = threadIdx.x, q = tid >> 6, pos = tid & 63, kh0 = pos >> 3, kw0 = pos & 7;
  constexpr int NH = MODE >= 1 ? 2 : 1;     // MODE 1: Diff2d(a, b); MODE 2: CE(a, y) + CE(b, y), one label read
  const int NI = NIT > 0 ? NIT : a.nin;

  // ---- filters -> shared memory [c][pos][tap], tap = 2a + b <-> (kh0 + 8a, kw0 + 8b)
  if (!BILINEAR) {
    for (int hd = 0; hd < NH; ++hd)
      for (int k = 0; k < NI; ++k) {
        const float* wg = a.in[hd][k].w;
        float* wd = ws + ((hd * NI + k) * C) * HL_WS;
        for (int i = tid; i < C * 256; i += HL_THREADS) {
          const int c = i >> 8, p = (i >> 2) & 63, t = i & 3;
          wd[c * HL_WS + (i & 255)] = wg[c * 256 + ((p >> 3) + 8 * (t >> 1)) * 16 + (p & 7) + 8 * (t & 1)];
        }
      }
  }
  // bilinear tap weights of this thread's position: rows (ci, ci-1) <-> (lam, 1-lam), lam = (k0 + .5) / 8
  const float lh = (kh0 + 0.5f) * 0.125f, lw = (kw0 + 0.5f) * 0.125f;
  const float bw4[4] = {lh * lw, lh * (1.f - lw), (1.f - lh) * lw, (1.f - lh) * (1.f - lw)};
  // filter gradients: thread t owns bin (pos = t / 4, tap = t % 4) of EVERY channel, in registers over the whole kernel
  float dwacc[DW ? DWN : 1][DW ? HL_MAXC : 1];
#pragma unroll
  for (int i = 0; i < (DW ? DWN : 1); ++i)
#pragma unroll
    for (int c = 0; c < (DW ? HL_MAXC : 1); ++c) dwacc[i][c] = 0.f;

  float loss_acc = 0.f, bad_acc = 0.f;
  const float wsum = (MODE != 1 && a.wsum) ? a.wsum[0] : 1.f;
  const float ce_coef = wsum > 0.f ? 1.f / wsum : 0.f;

  for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
    const int g = item % a.groups_per_row;
    const int ci = (item / a.groups_per_row) % (h + 1);
    const int n = item / (a.groups_per_row * (h + 1));
    const int cj0 = g * HL_CELLS;
    __syncthreads();        // the previous item's phase 2 has finished with xs / dls / dxs
    // ---- stage the inputs of the 4 cells: xs[hd][k][c][cell][tap] = x[n][c][row(ci - a)][col(cj - b)]
    for (int hd = 0; hd < NH; ++hd)
      for (int k = 0; k < NI; ++k) {
        const float* xg = a.in[hd][k].x;
        float* xd = xs + ((hd * NI + k) * C) * (HL_CELLS * 4);
        float* xtd = xt + ((hd * NI + k) * C) * (HL_CELLS * 4);
        for (int i = tid; i < C * HL_CELLS * 4; i += HL_THREADS) {
          const int c = i / (HL_CELLS * 4), cell = (i >> 2) % HL_CELLS, t = i & 3;
          int r = ci - (t >> 1), col = cj0 + cell - (t & 1);
          float v = 0.f;
          if (BILINEAR) {       // clamped source index (align_corners = False border behaviour)
            r = min(max(r, 0), h - 1); col = min(max(col, 0), w - 1);
            v = xg[(((int64_t)n * C + c) * h + r) * w + col];
          } else if (r >= 0 && r < h && col >= 0 && col < w) {
            v = xg[(((int64_t)n * C + c) * h + r) * w + col];
          }
          xd[i] = v;
          if (DW) xtd[(c * 4 + t) * HL_CELLS + cell] = v;
        }
      }
    if (a.want_grad)
      for (int i = tid; i < NH * NI * C * 10; i += HL_THREADS) dxs[i] = 0.f;
    __syncthreads();
    // ---- phase 1: this thread's output pixel
    const int cj = cj0 + q;
    const int oh = 8 * ci - 4 + kh0, ow = 8 * cj - 4 + kw0;
    const bool pvalid = oh >= 0 && oh < H && ow >= 0 && ow < W && cj <= w;
    float v0[HL_MAXC], v1[MODE >= 1 ? HL_MAXC : 1];
#pragma unroll
    for (int c = 0; c < HL_MAXC; ++c) {
      float s0 = -INFINITY, s1 = -INFINITY;
      if (c < C) {
        s0 = 0.f; s1 = 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (k < NI) {
            const float4 x0 = *reinterpret_cast<const float4*>(xs + (((0 * NI + k) * C + c) * HL_CELLS + q) * 4);
            if (BILINEAR) {
              s0 += x0.x * bw4[0] + x0.y * bw4[1] + x0.z * bw4[2] + x0.w * bw4[3];
            } else {
              const float4 w0 = *reinterpret_cast<const float4*>(ws + ((0 * NI + k) * C + c) * HL_WS + pos * 4);
              s0 += x0.x * w0.x + x0.y * w0.y + x0.z * w0.z + x0.w * w0.w;
            }
            if (MODE >= 1) {
              const float4 x1 = *reinterpret_cast<const float4*>(xs + (((1 * NI + k) * C + c) * HL_CELLS + q) * 4);
              if (BILINEAR) {
                s1 += x1.x * bw4[0] + x1.y * bw4[1] + x1.z * bw4[2] + x1.w * bw4[3];
              } else {
                const float4 w1 = *reinterpret_cast<const float4*>(ws + ((1 * NI + k) * C + c) * HL_WS + pos * 4);
                s1 += x1.x * w1.x + x1.y * w1.y + x1.z * w1.z + x1.w * w1.w;
              }
            }
          }
        }
      }
      v0[c] = s0;
      if (MODE >= 1) v1[c] = s1;
    }
#if defined(HL_EXP) && HL_EXP == 1
    { float t = 0.f;
#pragma unroll
      for (int c = 0; c < HL_MAXC; ++c) t += (c < C ? v0[c] : 0.f) + (MODE >= 1 && c < C ? v1[c] : 0.f);
      loss_acc += t; continue; }
#endif
    // cross entropy: the label and its raw logit (before the logits are overwritten by their exponentials)
    float wy = 0.f, xy = 0.f, xy1 = 0.f;
    int y = -1;
    if (MODE != 1) {
      if (pvalid) {
        const int64_t yy = __ldcs(a.target + ((int64_t)n * H + oh) * W + ow);
        if (yy != a.ignore_index) {
          if (yy < 0 || yy >= C) bad_acc += 1.f;
          else { y = (int)yy; wy = a.cls_weight ? a.cls_weight[y] : 1.f; }
        }
      }
#pragma unroll
      for (int c = 0; c < HL_MAXC; ++c) {
        xy = (c == y) ? v0[c] : xy;
        if (MODE == 2) xy1 = (c == y) ? v1[c] : xy1;
      }
    }
    // softmax statistics
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int c = 0; c < HL_MAXC; ++c) { m0 = fmaxf(m0, v0[c]); if (MODE >= 1) m1 = fmaxf(m1, v1[c]); }
    float se0 = 0.f, se1 = 0.f;
#pragma unroll
    for (int c = 0; c < HL_MAXC; ++c) {
      v0[c] = __expf(v0[c] - m0); se0 += v0[c];          // padding channels: exp(-inf) = 0
      if (MODE >= 1) { v1[c] = __expf(v1[c] - m1); se1 += v1[c]; }
    }
    if (MODE != 1) {
      // nll = max + log(sum exp) - x_y;  d logit_c = w_y / wsum * (p_c - [c == y])
      if (y >= 0) loss_acc += wy * (m0 + __logf(se0) - xy);
      const float kk = wy * ce_coef / se0;
#pragma unroll
      for (int c = 0; c < HL_MAXC; ++c) v0[c] = (y >= 0) ? (kk * v0[c] - (c == y ? wy * ce_coef : 0.f)) : 0.f;
      if (MODE == 2) {      // the second classifier on the same labels
        if (y >= 0) loss_acc += wy * (m1 + __logf(se1) - xy1);
        const float k1 = wy * ce_coef / se1;
#pragma unroll
        for (int c = 0; c < HL_MAXC; ++c) v1[c] = (y >= 0) ? (k1 * v1[c] - (c == y ? wy * ce_coef : 0.f)) : 0.f;
      }
    } else {
      // Diff2d: loss += sum_c |pa - pb|; d/d logit_a = k pa (s_c - sum s pa), d/d logit_b = k pb (sum s pb - s_c)
      const float i0 = 1.f / se0, i1 = 1.f / se1;
      float da = 0.f, db = 0.f, l = 0.f;
#pragma unroll
      for (int c = 0; c < HL_MAXC; ++c) {
        const float pa = v0[c] * i0, pb = v1[c] * i1, d = pa - pb;
        const float s = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        l += fabsf(d);
        da = fmaf(s, pa, da); db = fmaf(s, pb, db);
        v0[c] = pa; v1[c] = pb;
      }
      if (pvalid) loss_acc += l;
      const float kk = pvalid ? a.inv_numel : 0.f;
#pragma unroll
      for (int c = 0; c < HL_MAXC; ++c) {
        const float pa = v0[c], pb = v1[c], d = pa - pb;
        const float s = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        v0[c] = kk * pa * (s - da);
        v1[c] = kk * pb * (db - s);
      }
    }
#if defined(HL_EXP) && HL_EXP == 2
    { float t = 0.f;
#pragma unroll
      for (int c = 0; c < HL_MAXC; ++c) t += v0[c] + (MODE >= 1 ? v1[c] : 0.f);
      loss_acc += t * 1e-30f; continue; }
#endif
    if (!a.want_grad) continue;
    // logit gradients -> shared memory (bf16, [head][c][pos][cell]); invalid pixels contribute zeros
#pragma unroll
    for (int c = 0; c < HL_MAXC; ++c) {
      if (c < C) {
        // (pixels outside the image already hold zeros: no label / k = 0 above)
        dls[(0 * C + c) * HL_DL + pos * 4 + q] = __float2bfloat16_rn(v0[c]);
        if (MODE >= 1) dls[(1 * C + c) * HL_DL + pos * 4 + q] = __float2bfloat16_rn(v1[c]);
      }
    }
    __syncthreads();
#if defined(HL_EXP) && HL_EXP == 3
    continue;
#endif
    // ---- phase 2a: score-map gradients.  Work unit (channel c, quarter of the 64 positions): the 4 cells x 4 taps
    //      partial sums over its 16 positions (one 8-byte load = the logit gradient of a position in all 4 cells, one
    //      16-byte load = its 4 filter taps), added to the item's 2 x 5 score-map gradient tile: tap (a, b) of cell
    //      `cell` belongs to input row ci - a, input column cj0 + cell - b  ->  tile column cell + 1 - b.
    for (int k = 0; k < NI; ++k) {
      for (int hd = 0; hd < NH; ++hd) {
        if (!a.in[hd][k].dx) continue;
        const float* wd = ws + ((hd * NI + k) * C) * HL_WS;
        for (int u = tid; u < C * 4; u += HL_THREADS) {
          const int c = u % C, qt = u / C;
          float sacc[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sacc[i][j] = 0.f;
#pragma unroll 4
          for (int pi = 0; pi < 16; ++pi) {
            const int p = qt * 16 + pi;
            const uint2 draw = *reinterpret_cast<const uint2*>(dls + (hd * C + c) * HL_DL + p * 4);
            const float d0 = __uint_as_float(draw.x << 16), d1 = __uint_as_float(draw.x & 0xffff0000u);
            const float d2 = __uint_as_float(draw.y << 16), d3 = __uint_as_float(draw.y & 0xffff0000u);
            float4 wv;
            if (BILINEAR) {
              const float wh = ((p >> 3) + 0.5f) * 0.125f, ww = ((p & 7) + 0.5f) * 0.125f;
              wv = make_float4(wh * ww, wh * (1.f - ww), (1.f - wh) * ww, (1.f - wh) * (1.f - ww));
            } else {
              wv = *reinterpret_cast<const float4*>(wd + c * HL_WS + p * 4);
            }
            const float dd[4] = {d0, d1, d2, d3};
            const float wt[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int cell = 0; cell < 4; ++cell)
#pragma unroll
              for (int t = 0; t < 4; ++t) sacc[cell][t] = fmaf(dd[cell], wt[t], sacc[cell][t]);
          }
          float* dxc = dxs + ((hd * NI + k) * C + c) * 10;
#pragma unroll
          for (int cell = 0; cell < 4; ++cell)
#pragma unroll
            for (int t = 0; t < 4; ++t) sm_add(dxc + (t >> 1) * 5 + cell + 1 - (t & 1), sacc[cell][t]);
        }
      }
    }
    // ---- phase 2b: filter gradients in registers: bin (pos = tid / 4, tap = tid % 4) of every channel
    if (DW) {
      const int p2 = tid >> 2, t2 = tid & 3;
#pragma unroll
      for (int hd = 0; hd < 2; ++hd)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (hd >= NH || k >= NI || hd * NI + k >= DWN || !a.in[hd][k].dw) continue;
          const float* xtd = xt + ((hd * NI + k) * C) * (HL_CELLS * 4);
#pragma unroll
          for (int c = 0; c < HL_MAXC; ++c) {
            if (c < C) {
              const uint2 draw = *reinterpret_cast<const uint2*>(dls + (hd * C + c) * HL_DL + p2 * 4);
              const float4 xv = *reinterpret_cast<const float4*>(xtd + (c * 4 + t2) * HL_CELLS);
              float s = __uint_as_float(draw.x << 16) * xv.x;
              s = fmaf(__uint_as_float(draw.x & 0xffff0000u), xv.y, s);
              s = fmaf(__uint_as_float(draw.y << 16), xv.z, s);
              s = fmaf(__uint_as_float(draw.y & 0xffff0000u), xv.w, s);
              dwacc[(hd * NI + k) < DWN ? (hd * NI + k) : 0][c] += s;
            }
          }
        }
    }
    __syncthreads();
    // ---- the item's score-map gradient tiles -> global (one atomic per element; two heads that read the same score
    //      map, i.e. share the dx pointer, are summed first)
    for (int k = 0; k < NI; ++k)
      for (int hd = 0; hd < NH; ++hd) {
        float* dxg = a.in[hd][k].dx;
        if (!dxg) continue;
        const bool merged = NH > 1 && a.in[0][k].dx == a.in[1][k].dx;
        if (merged && hd == 1) continue;
        for (int i = tid; i < C * 10; i += HL_THREADS) {
          const int c = i / 10, ar = (i / 5) & 1, jj = i % 5;
          float v = dxs[((hd * NI + k) * C + c) * 10 + ar * 5 + jj];
          if (merged) v += dxs[((1 * NI + k) * C + c) * 10 + ar * 5 + jj];
          int r = ci - ar, col = cj0 - 1 + jj;
          if (BILINEAR) { r = min(max(r, 0), h - 1); col = min(max(col, 0), w - 1); }
          else if (r < 0 || r >= h || col < 0 || col >= w) continue;
          if (v != 0.f) atomicAdd(dxg + (((int64_t)n * C + c) * h + r) * w + col, v);
        }
      }
  }
  // ---- block results
  const float r0 = block_sum(loss_acc, red);
  if (tid == 0 && r0 != 0.f) atomicAdd(a.acc + 0, r0);
  if (MODE != 1) {
    const float r2 = block_sum(bad_acc, red);
    if (tid == 0 && r2 != 0.f) atomicAdd(a.acc + 2, r2);
  }
  if (DW) {
    const int p2 = tid >> 2, t2 = tid & 3;
    const int bin = ((p2 >> 3) + 8 * (t2 >> 1)) * 16 + (p2 & 7) + 8 * (t2 & 1);
#pragma unroll
    for (int hd = 0; hd < 2; ++hd)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (hd >= NH || k >= NI || hd * NI + k >= DWN || !a.in[hd][k].dw) continue;
#pragma unroll
        for (int c = 0; c < HL_MAXC; ++c) {
          const float v = dwacc[(hd * NI + k) < DWN ? (hd * NI + k) : 0][c];
          if (c < C && v != 0.f) atomicAdd(a.in[hd][k].dw + c * 256 + bin, v);
        }
      }
  }
}

// per-pixel class weights summed over a label map: the normaliser of CrossEntropyLoss2d (loss.py:12-13, NLLLoss2d
// weighted mean); acc[0] += sum w[y], acc[1] += number of labels outside [0, C) that are not ignore_index
__global__ void __launch_bounds__(256)
label_weight_sum_kernel(const int64_t* __restrict__ target, const float* __restrict__ weight, int64_t ignore_index,
                        int C, float* __restrict__ acc, int64_t total) {
  __shared__ float red[32];
  float s = 0.f, bad = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t y = target[i];
    if (y == ignore_index) continue;
    if (y < 0 || y >= C) { bad += 1.f; continue; }
    s += weight ? weight[y] : 1.f;
  }
  const float r0 = block_sum(s, red), r1 = block_sum(bad, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc, r0);
    if (r1 != 0.f) atomicAdd(acc + 1, r1);
  }
}

static int sm_count_hl() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int MODE, bool BIL, int DWN, int CT, int NIT>
static int launch_head_loss_x(HeadLossArgs& a, cudaStream_t st) {
  int rc = ensure_dyn_smem<head_loss_kernel<MODE, BIL, DWN, CT, NIT>>(a.smem_bytes, "head_loss");
  if (rc != MCD_OK) return rc;
  const int grid = min(a.total_items, sm_count_hl());
  head_loss_kernel<MODE, BIL, DWN, CT, NIT><<<grid, HL_THREADS, a.smem_bytes, st>>>(a);
  return check_launch("head_loss");
}

template <int MODE, bool BIL, int DWN>
static int launch_head_loss(HeadLossArgs& a, cudaStream_t st) {
  // the 41-class, one-input heads of the reference's trainers get the constant-folded instantiation
  if constexpr (!(MODE == 0 && DWN == 2)) {       // (one head, two filter gradients) implies two inputs
    if (a.C == 41 && a.nin == 1) return launch_head_loss_x<MODE, BIL, DWN, 41, 1>(a, st);
  }
  return launch_head_loss_x<MODE, BIL, DWN, 0, 0>(a, st);
}

}  // namespace mcd

using namespace mcd;

extern "C" {

int mcd_label_weight_sum(const int64_t* target, const float* weight, int64_t ignore_index, int C, float* acc2,
                         int64_t numel, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(target && acc2 && numel > 0 && C > 0, "label_weight_sum: bad arguments");
  const int grid = (int)max64(1, min64((numel + 255) / 256, 148 * 8));
  label_weight_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(target, weight, ignore_index, C, acc2, numel);
  return check_launch("label_weight_sum");
}

int mcd_head_loss(int mode, int nheads, int nin, const float* const* x, const float* const* w, float* const* dx,
                  float* const* dw, const int64_t* target, const float* cls_weight, int64_t ignore_index,
                  const float* wsum, float inv_numel, float* acc, int N, int C, int h, int w_, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(mode >= 0 && mode <= 2, "head_loss: mode must be 0 (cross entropy), 1 (Diff2d) or 2 (cross entropy of two heads)");
  MCD_REQUIRE(nheads == (mode ? 2 : 1) && (nin == 1 || nin == 2), "head_loss: %d heads / %d inputs unsupported", nheads, nin);
  MCD_REQUIRE(x && w && dx && dw && acc && N > 0 && C > 0 && h > 0 && w_ > 0, "head_loss: bad arguments");
  MCD_REQUIRE(C <= HL_MAXC, "head_loss: at most %d classes (got %d)", HL_MAXC, C);
  MCD_REQUIRE(mode == 1 || target, "head_loss: cross entropy needs labels");
  HeadLossArgs a;
  memset(&a, 0, sizeof(a));
  a.N = N; a.C = C; a.h = h; a.w = w_; a.mode = mode; a.nheads = nheads; a.nin = nin;
  bool bil = true, any_w = false, any_dw = false;
  for (int hd = 0; hd < nheads; ++hd)
    for (int k = 0; k < nin; ++k) {
      const int i = hd * nin + k;
      MCD_REQUIRE(x[i], "head_loss: input %d of head %d is NULL", k, hd);
      a.in[hd][k] = HeadIn{x[i], w[i], dx[i], dw[i]};
      any_w |= w[i] != nullptr;
      bil &= w[i] == nullptr;
      any_dw |= dw[i] != nullptr;
      a.want_grad |= (dx[i] != nullptr) || (dw[i] != nullptr);
      MCD_REQUIRE(!dw[i] || w[i], "head_loss: a filter gradient needs the filter");
    }
  MCD_REQUIRE(bil || !(!any_w), "head_loss: internal");
  for (int hd = 0; hd < nheads; ++hd)
    for (int k = 0; k < nin; ++k)
      MCD_REQUIRE(bil == (w[hd * nin + k] == nullptr), "head_loss: learned and bilinear heads cannot be mixed");
  a.target = target; a.cls_weight = cls_weight; a.ignore_index = ignore_index; a.wsum = wsum;
  a.inv_numel = inv_numel; a.acc = acc;
  a.groups_per_row = (w_ + 1 + HL_CELLS - 1) / HL_CELLS;
  a.total_items = N * (h + 1) * a.groups_per_row;
  // the two heads of a Diff2d either share a score map (same x AND same dx: the classifiers of one generator) or not
  if (nheads == 2)
    for (int k = 0; k < nin; ++k) {
      const bool same_x = x[k] == x[nin + k];
      MCD_REQUIRE(!dx[k] || !dx[nin + k] || (dx[k] == dx[nin + k]) == same_x,
                  "head_loss: heads that read the same score map must share its gradient buffer (and only those)");
      MCD_REQUIRE(!same_x || !dx[k] == !dx[nin + k] || true, "head_loss: internal");
    }
  // shared memory: filters, staged inputs (two layouts), logit gradients (bf16), the item's score-map gradient tile
  const int nhi = nheads * nin;
  int off = 0;
  a.off_w = off; off += bil ? 0 : nhi * C * HL_WS;
  a.off_x = off; off += nhi * C * HL_CELLS * 4;
  a.off_xt = off; off += any_dw ? nhi * C * HL_CELLS * 4 : 0;
  a.off_dl = off; off += (nheads * C * HL_DL) / 2;      // bf16
  a.off_dx = off; off += nhi * C * 10;
  a.smem_bytes = off * 4 + 16;
  MCD_REQUIRE(a.smem_bytes <= 227 * 1024, "head_loss: %d bytes of shared memory needed", a.smem_bytes);
  // filter gradients live in registers: at most 2 (head, input) pairs (one head with two inputs, or two heads with one)
  MCD_REQUIRE(!any_dw || nhi <= 2, "head_loss: filter gradients of %d head inputs are not supported in one launch "
              "(two heads x two inputs: train the filters through the un-fused head)", nhi);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) {
    if (bil) return launch_head_loss<0, true, 0>(a, st);
    if (!any_dw) return launch_head_loss<0, false, 0>(a, st);
    return nhi == 1 ? launch_head_loss<0, false, 1>(a, st) : launch_head_loss<0, false, 2>(a, st);
  }
  if (mode == 2) {
    if (bil) return launch_head_loss<2, true, 0>(a, st);
    return any_dw ? launch_head_loss<2, false, 2>(a, st) : launch_head_loss<2, false, 0>(a, st);
  }
  if (bil) return launch_head_loss<1, true, 0>(a, st);
  return any_dw ? launch_head_loss<1, false, 2>(a, st) : launch_head_loss<1, false, 0>(a, st);
}

}  // extern "C"
