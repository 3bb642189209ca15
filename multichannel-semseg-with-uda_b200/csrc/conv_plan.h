// conv_plan.h — host-side lowering of one nn.Conv2d (fprop / dgrad) to "tap problems":
// a dense gather-GEMM over a tile-space pixel grid, shared by the direct and the tcgen05 kernels.
//
//   out[n, ht*omul+oh0, wt*omul+ow0, row] = sum_t sum_kc src[n, ht*smul+dh_t, wt*smul+dw_t, kc]
//                                                      * w_packed[row][wk_t][kc]
//
// fprop  stride s : one problem, smul = s, omul = 1, dh = r*dil - pad, wk = r*S+s.
// dgrad  stride 1 : one problem over dy with the flipped-tap pack (mcd_pack_weight mode 1).
// dgrad  stride 2 : four problems, one per output parity class (ph,pw); only the taps whose source
//                   row (h + pad - r*dil)/2 is an integer contribute (models/drn.py:175-180 strided
//                   convs: layer2, layer3.0, layer4.0 and their 1x1 downsample branches).
#pragma once
#include <stdint.h>

namespace mcd {

constexpr int kMaxTaps = 49;

// optional epilogue inputs of a problem (all nhwc bf16 in the OUTPUT geometry, or null)
struct EpiExtra {
  const void* addend = nullptr;    // out += addend                       (identity-shortcut gradient)
  const void* mask_src = nullptr;  // out  = mask_src > 0 ? out : 0       (ReLU backward of the producing unit)
  const void* bn_y = nullptr;      // stats += {sum out, sum out * bn_y}  (BatchNorm backward sums)
  int relu = 0;                    // out  = max(out, 0)                  (forward of an eval-mode conv-BN-ReLU unit)
  void* sk_partial = nullptr;      // stream-K workspace (tcgen05 path): fp32 partial tiles ...
  int* sk_flags = nullptr;         // ... and ready flags (zeroed by the caller); see mcd_conv2d_streamk_workspace
};

struct Tap {
  int16_t dh, dw;   // source offset in source-grid pixels (after smul scaling of the tile coord)
  int16_t wk;       // tap slot inside the packed weight row
  int16_t map;      // tcgen05 path: which parity tensor map (0..3); direct path: unused
  int16_t mdh, mdw; // tcgen05 path: offset inside the parity sub-grid
  int16_t pad0, pad1;
};

struct TapProblem {
  int N;
  int Hs, Ws;          // source geometry
  int Cs_src;          // source channel stride
  int Kc;              // reduce channels actually present in the source (<= Cs_src)
  int kc_pad;          // padded reduce channels per tap in the packed weights
  int T_total;         // taps per packed weight row
  int Ht, Wt;          // tile-space grid
  int smul;            // source multiplier (direct kernel)
  int omul, oh0, ow0;  // output placement
  int Hd, Wd;          // destination geometry
  int Cd_s;            // destination channel stride (nhwc) / channel count (planar)
  int rows;            // GEMM-N: produced channels
  int packed;          // 1: row-packed thin-channel problem (one "tap" = one filter row, see plan_*_packed)
  int ntaps;
  Tap taps[kMaxTaps];
};

inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int posmod(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

// fprop: returns 1 problem
inline void plan_fprop(const mcd_conv_geom& g, TapProblem& p) {
  p.N = g.N; p.Hs = g.H; p.Ws = g.W; p.Cs_src = g.Cin_s; p.Kc = g.Cin;
  p.kc_pad = (g.Cin + 63) / 64 * 64; p.T_total = g.R * g.S;
  p.Ht = g.Ho; p.Wt = g.Wo; p.smul = g.stride; p.omul = 1; p.oh0 = 0; p.ow0 = 0;
  p.Hd = g.Ho; p.Wd = g.Wo; p.Cd_s = g.Cout_s; p.rows = g.Cout; p.ntaps = 0; p.packed = 0;
  for (int r = 0; r < g.R; ++r)
    for (int s = 0; s < g.S; ++s) {
      Tap& t = p.taps[p.ntaps++];
      t.dh = (int16_t)(r * g.dil - g.pad); t.dw = (int16_t)(s * g.dil - g.pad);
      t.wk = (int16_t)(r * g.S + s);
      // parity decomposition for stride 2: source row = stride*ht + dh = stride*(ht + floor(dh/stride)) + (dh mod stride)
      int ph = posmod(t.dh, g.stride), pw = posmod(t.dw, g.stride);
      t.map = (int16_t)(ph * g.stride + pw);
      t.mdh = (int16_t)floordiv(t.dh, g.stride); t.mdw = (int16_t)floordiv(t.dw, g.stride);
      t.pad0 = t.pad1 = 0;
    }
}

// dgrad: returns the number of problems written to p[] (1 for stride 1, stride^2 for stride 2);
// problems with ntaps == 0 mean "that parity class of dx is identically zero".
inline int plan_dgrad(const mcd_conv_geom& g, TapProblem* p) {
  int st = g.stride, np = 0;
  for (int ph = 0; ph < st; ++ph)
    for (int pw = 0; pw < st; ++pw) {
      TapProblem& q = p[np++];
      q.N = g.N; q.Hs = g.Ho; q.Ws = g.Wo; q.Cs_src = g.Cout_s; q.Kc = g.Cout;
      q.kc_pad = (g.Cout + 63) / 64 * 64; q.T_total = g.R * g.S;
      q.Ht = (g.H - ph + st - 1) / st; q.Wt = (g.W - pw + st - 1) / st;
      q.smul = 1; q.omul = st; q.oh0 = ph; q.ow0 = pw;
      q.Hd = g.H; q.Wd = g.W; q.Cd_s = g.Cin_s; q.rows = g.Cin; q.ntaps = 0; q.packed = 0;
      for (int r = 0; r < g.R; ++r)
        for (int s = 0; s < g.S; ++s) {
          int nh = ph + g.pad - r * g.dil, nw = pw + g.pad - s * g.dil;
          if (posmod(nh, st) != 0 || posmod(nw, st) != 0) continue;
          Tap& t = q.taps[q.ntaps++];
          t.dh = (int16_t)floordiv(nh, st); t.dw = (int16_t)floordiv(nw, st);
          // mode-1 pack stores tap (r,s) at slot (R-1-r)*S + (S-1-s)
          t.wk = (int16_t)((g.R - 1 - r) * g.S + (g.S - 1 - s));
          t.map = 0; t.mdh = t.dh; t.mdw = t.dw; t.pad0 = t.pad1 = 0;
        }
    }
  return np;
}

// ---- row-packed thin-channel problems ---------------------------------------------------------------
// For Cs_src in {8,16} and dilation 1 the S taps of one filter row read S*Cs_src CONTIGUOUS bf16 of the NHWC
// image (pixels w*stride-pad .. +S-1), at most 64 elements.  One 64-element window per output pixel is then one
// K=64 chunk: layer0 (7x7, 6->16 ch: 7 chunks instead of 49), layer1 (3x3, 16->16: 3 instead of 9), layer2
// (3x3 stride 2, 16->32) and the dgrad of layer1 (models/drn.py:126-136).  The window is fetched by a 3-D TMA
// map {W*Cs elements, rows, N} whose inner coordinate is an ELEMENT offset, so image-border windows are
// zero-filled by TMA exactly like the padded convolution.
//   tap t = filter row r:  dh/mdh/map as usual (row parity for stride 2), mdw = window start in PIXELS
//   relative to wt*smul, wk = r.  Packed weights: [rows][R][64] with k = s*Cs_src + c.
inline bool packed_fprop_ok(const mcd_conv_geom& g) {
  return g.dil == 1 && (g.Cin_s == 8 || g.Cin_s == 16) && g.S * g.Cin_s <= 64 && g.stride <= 2 && g.R <= kMaxTaps;
}
inline void plan_fprop_packed(const mcd_conv_geom& g, TapProblem& p) {
  p.N = g.N; p.Hs = g.H; p.Ws = g.W; p.Cs_src = g.Cin_s; p.Kc = 64; p.kc_pad = 64; p.T_total = g.R;
  p.Ht = g.Ho; p.Wt = g.Wo; p.smul = g.stride; p.omul = 1; p.oh0 = 0; p.ow0 = 0;
  p.Hd = g.Ho; p.Wd = g.Wo; p.Cd_s = g.Cout_s; p.rows = g.Cout; p.ntaps = 0; p.packed = 1;
  for (int r = 0; r < g.R; ++r) {
    Tap& t = p.taps[p.ntaps++];
    t.dh = (int16_t)(r - g.pad); t.dw = (int16_t)(-g.pad); t.wk = (int16_t)r;
    t.map = (int16_t)posmod(t.dh, g.stride);
    t.mdh = (int16_t)floordiv(t.dh, g.stride); t.mdw = t.dw; t.pad0 = t.pad1 = 0;
  }
}
// dgrad of a stride-1 conv whose dy is thin: windows over dy, flipped filter rows.
inline bool packed_dgrad_ok(const mcd_conv_geom& g) {
  return g.dil == 1 && g.stride == 1 && (g.Cout_s == 8 || g.Cout_s == 16) && g.S * g.Cout_s <= 64;
}
inline void plan_dgrad_packed(const mcd_conv_geom& g, TapProblem& q) {
  q.N = g.N; q.Hs = g.Ho; q.Ws = g.Wo; q.Cs_src = g.Cout_s; q.Kc = 64; q.kc_pad = 64; q.T_total = g.R;
  q.Ht = g.H; q.Wt = g.W; q.smul = 1; q.omul = 1; q.oh0 = 0; q.ow0 = 0;
  q.Hd = g.H; q.Wd = g.W; q.Cd_s = g.Cin_s; q.rows = g.Cin; q.ntaps = 0; q.packed = 1;
  for (int rp = 0; rp < g.R; ++rp) {          // rp = R-1-r
    Tap& t = q.taps[q.ntaps++];
    t.dh = (int16_t)(rp - (g.R - 1 - g.pad)); t.dw = (int16_t)(-(g.S - 1 - g.pad)); t.wk = (int16_t)rp;
    t.map = 0; t.mdh = t.dh; t.mdw = t.dw; t.pad0 = t.pad1 = 0;
  }
}

}  // namespace mcd
