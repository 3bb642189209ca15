// conv_api.cu — C-ABI entry points for convolution: validation, lowering to tap problems
// (conv_plan.h) and dispatch to the tcgen05 (conv_umma.cu) or CUDA-core (conv_direct.cu) kernels.
#include "common.cuh"
#include "conv_plan.h"

namespace mcd {
int launch_direct_problem(const void* src, const void* w, const float* bias, void* out, int planar,
                          const void* addend, const TapProblem& p, int fmt, cudaStream_t st);
int wgrad_direct(const void* x, const void* dy, float* dw, const mcd_conv_geom& g, int accumulate,
                 cudaStream_t st);
int colsum(const void* t, float* out, int64_t P, int C, int Cs, int accumulate, cudaStream_t st);
bool umma_problem_supported(const TapProblem& p);
int launch_umma_problem(const void* src, const void* w, const float* bias, void* out, int planar,
                        float* stats, const EpiExtra& ex, const TapProblem& p, int fmt, cudaStream_t st);
size_t umma_wgrad_workspace(const mcd_conv_geom& g);
size_t umma_streamk_workspace(const TapProblem& p, int planar, int* n_flags);
int umma_wgrad(const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes,
               const mcd_conv_geom& g, int accumulate, cudaStream_t st);
int umma_problem_tile(const TapProblem& p, int planar, int* pair, int* halo);
int umma_wgrad_tile(const mcd_conv_geom& g, int* rows);
bool umma_wgrad_partial_layout(const mcd_conv_geom& g, int* out4);
bool rowconv_fprop_ok(const mcd_conv_geom& g);
bool rowconv_dgrad_ok(const mcd_conv_geom& g);
int rowconv_pack(const float* w, void* dst, int Cout, int Cin, int R, int S, int Cs, int mode, cudaStream_t st);
int rowconv_launch(const void* src, const void* wpacked, const float* bias, void* out, int planar, float* stats,
                   const EpiExtra& ex, const mcd_conv_geom& g, int mode, cudaStream_t st);
int bn_mask_sums_launch(void* dx, const void* mask_src, const void* bn_y, float* sums, int64_t P, int C, int Cs,
                        cudaStream_t st);
int bn_stats_launch(const void* y, float* stats, int64_t P, int C, int Cs, cudaStream_t st);

static int validate(const mcd_conv_geom* g) {
  MCD_REQUIRE(g, "conv: null geometry");
  MCD_REQUIRE(g->N > 0 && g->H > 0 && g->W > 0 && g->Cin > 0 && g->Cout > 0, "conv: bad sizes");
  MCD_REQUIRE(g->R > 0 && g->S > 0 && g->R * g->S <= kMaxTaps, "conv: filter %dx%d unsupported", g->R, g->S);
  MCD_REQUIRE(g->stride >= 1 && g->dil >= 1 && g->pad >= 0, "conv: bad stride/dil/pad");
  MCD_REQUIRE(g->Cin_s >= g->Cin && g->Cin_s % 8 == 0, "conv: Cin_s=%d must be >= Cin=%d, %% 8 == 0", g->Cin_s, g->Cin);
  MCD_REQUIRE(g->Cout_s >= g->Cout && g->Cout_s % 8 == 0, "conv: Cout_s=%d must be >= Cout=%d, %% 8 == 0", g->Cout_s, g->Cout);
  int Ho = (g->H + 2 * g->pad - g->dil * (g->R - 1) - 1) / g->stride + 1;
  int Wo = (g->W + 2 * g->pad - g->dil * (g->S - 1) - 1) / g->stride + 1;
  MCD_REQUIRE(Ho == g->Ho && Wo == g->Wo, "conv: Ho/Wo (%d,%d) inconsistent with geometry (%d,%d)",
              g->Ho, g->Wo, Ho, Wo);
  return MCD_OK;
}

static bool use_umma(int algo, bool supported, int* rc) {
  *rc = MCD_OK;
  if (algo == MCD_ALGO_DIRECT) return false;
  if (algo == MCD_ALGO_UMMA) {
    if (!supported) { set_error("conv: MCD_ALGO_UMMA requested for an unsupported shape"); *rc = MCD_E_INVALID; }
    return supported;
  }
  return supported;
}

}  // namespace mcd

using namespace mcd;

extern "C" {

int mcd_conv2d_fprop(const void* x_nhwc, const void* w_packed, const float* bias, void* y,
                     int y_layout, float* stats, void* sk_partial, int* sk_flags, const mcd_conv_geom* g,
                     int algo, int device, void* stream) {
  MCD_ENTER(device);
  int rc = validate(g);
  if (rc != MCD_OK) return rc;
  MCD_REQUIRE(x_nhwc && w_packed && y, "conv fprop: null pointer");
  MCD_REQUIRE(y_layout == MCD_OUT_NHWC_BF16 || y_layout == MCD_OUT_PLANAR_F32, "conv fprop: bad y_layout");
  cudaStream_t st = (cudaStream_t)stream;
  TapProblem p;
  plan_fprop(*g, p);
  int planar = y_layout == MCD_OUT_PLANAR_F32;
  bool umma = use_umma(algo, umma_problem_supported(p), &rc);
  if (rc != MCD_OK) return rc;
  if (umma) {
    if (rowconv_fprop_ok(*g))                            // w_packed: mcd_pack_weight_rowconv mode 0
      return rowconv_launch(x_nhwc, w_packed, bias, y, planar, stats, EpiExtra(), *g, 0, st);
    if (packed_fprop_ok(*g)) plan_fprop_packed(*g, p);   // w_packed is then the mcd_pack_weight_rows layout
    EpiExtra ex;
    ex.sk_partial = sk_partial; ex.sk_flags = sk_flags;
    return launch_umma_problem(x_nhwc, w_packed, bias, y, planar, stats, ex, p, kF16, st);
  }
  rc = launch_direct_problem(x_nhwc, w_packed, bias, y, planar, nullptr, p, kF16, st);
  if (rc != MCD_OK) return rc;
  if (stats) {
    MCD_REQUIRE(!planar, "conv fprop: fused BN statistics need the nhwc output layout");
    return bn_stats_launch(y, stats, (int64_t)g->N * g->Ho * g->Wo, g->Cout, g->Cout_s, st);
  }
  return MCD_OK;
}

int mcd_conv2d_fprop_act_supported(const mcd_conv_geom* g) {
  if (!g || validate(g) != MCD_OK) return 0;
  TapProblem p;
  plan_fprop(*g, p);
  return umma_problem_supported(p) ? 1 : 0;
}

int mcd_conv2d_fprop_act(const void* x_nhwc, const void* w_packed, const float* bias, const void* res_nhwc, int relu,
                         void* y_nhwc, void* sk_partial, int* sk_flags, const mcd_conv_geom* g, int algo, int device,
                         void* stream) {
  MCD_ENTER(device);
  int rc = validate(g);
  if (rc != MCD_OK) return rc;
  MCD_REQUIRE(x_nhwc && w_packed && y_nhwc, "conv fprop_act: null pointer");
  MCD_REQUIRE(algo != MCD_ALGO_DIRECT, "conv fprop_act: tcgen05 paths only");
  cudaStream_t st = (cudaStream_t)stream;
  TapProblem p;
  plan_fprop(*g, p);
  MCD_REQUIRE(umma_problem_supported(p), "conv fprop_act: shape not supported by the tcgen05 path");
  EpiExtra ex;
  ex.addend = res_nhwc; ex.relu = relu != 0;
  if (rowconv_fprop_ok(*g)) return rowconv_launch(x_nhwc, w_packed, bias, y_nhwc, 0, nullptr, ex, *g, 0, st);
  if (packed_fprop_ok(*g)) plan_fprop_packed(*g, p);
  ex.sk_partial = sk_partial; ex.sk_flags = sk_flags;
  return launch_umma_problem(x_nhwc, w_packed, bias, y_nhwc, 0, nullptr, ex, p, kF16, st);
}

int mcd_conv2d_dgrad(const void* dy_nhwc, const void* w_packed_dgrad, void* dx_nhwc, const void* add_nhwc,
                     const void* relu_src_nhwc, const void* bn_y_nhwc, float* bn_sums, void* sk_partial,
                     int* sk_flags, const mcd_conv_geom* g, int algo, int device, void* stream) {
  MCD_ENTER(device);
  int rc = validate(g);
  if (rc != MCD_OK) return rc;
  MCD_REQUIRE(dy_nhwc && w_packed_dgrad && dx_nhwc, "conv dgrad: null pointer");
  MCD_REQUIRE(g->stride <= 2, "conv dgrad: stride %d unsupported", g->stride);
  MCD_REQUIRE((bn_y_nhwc != nullptr) == (bn_sums != nullptr), "conv dgrad: bn_y and bn_sums go together");
  MCD_REQUIRE(!bn_sums || g->Cin_s == g->Cin, "conv dgrad: fused BatchNorm sums need dense channels");
  cudaStream_t st = (cudaStream_t)stream;
  EpiExtra ex;
  ex.addend = add_nhwc; ex.mask_src = relu_src_nhwc; ex.bn_y = bn_y_nhwc;
  const bool post_wanted = relu_src_nhwc || bn_sums;
  if (algo != MCD_ALGO_DIRECT && rowconv_dgrad_ok(*g))   // w_packed_dgrad: mcd_pack_weight_rowconv mode 1
    return rowconv_launch(dy_nhwc, w_packed_dgrad, nullptr, dx_nhwc, 0, bn_sums, ex, *g, 1, st);
  TapProblem p[4];
  int np = plan_dgrad(*g, p);
  if (algo != MCD_ALGO_DIRECT && packed_dgrad_ok(*g) && umma_problem_supported(p[0])) {
    plan_dgrad_packed(*g, p[0]);                          // w_packed_dgrad: mcd_pack_weight_rows mode 1
    return launch_umma_problem(dy_nhwc, w_packed_dgrad, nullptr, dx_nhwc, 0, bn_sums, ex, p[0], kBF16, st);
  }
  // the ReLU mask / BatchNorm sums are fused into the epilogue when every pixel of dx is produced by a tcgen05
  // problem; otherwise one extra pass over dx applies them
  bool fused = post_wanted && algo != MCD_ALGO_DIRECT;
  bool any_empty = false;
  for (int i = 0; i < np; ++i) {
    any_empty |= (p[i].ntaps == 0);
    if (p[i].ntaps && !umma_problem_supported(p[i])) fused = false;
  }
  if (any_empty) fused = false;
  if (!fused) { ex.mask_src = nullptr; ex.bn_y = nullptr; }
  if (np == 1) { ex.sk_partial = sk_partial; ex.sk_flags = sk_flags; }
  if (any_empty) {   // parity classes without taps: dx = 0 (+ addend) there
    const size_t bytes = (size_t)g->N * g->H * g->W * g->Cin_s * 2;
    cudaError_t e = add_nhwc ? cudaMemcpyAsync(dx_nhwc, add_nhwc, bytes, cudaMemcpyDeviceToDevice, st)
                             : cudaMemsetAsync(dx_nhwc, 0, bytes, st);
    if (e != cudaSuccess) { set_error("dgrad fill: %s", cudaGetErrorString(e)); return MCD_E_CUDA; }
  }
  for (int i = 0; i < np; ++i) {
    if (p[i].ntaps == 0 || p[i].Ht <= 0 || p[i].Wt <= 0) continue;
    bool umma = use_umma(algo, umma_problem_supported(p[i]), &rc);
    if (rc != MCD_OK) return rc;
    rc = umma ? launch_umma_problem(dy_nhwc, w_packed_dgrad, nullptr, dx_nhwc, 0, fused ? bn_sums : nullptr, ex,
                                    p[i], kBF16, st)
              : launch_direct_problem(dy_nhwc, w_packed_dgrad, nullptr, dx_nhwc, 0, add_nhwc, p[i], kBF16, st);
    if (rc != MCD_OK) return rc;
  }
  if (post_wanted && !fused)
    return bn_mask_sums_launch(dx_nhwc, relu_src_nhwc, bn_y_nhwc, bn_sums, (int64_t)g->N * g->H * g->W, g->Cin,
                               g->Cin_s, st);
  return MCD_OK;
}

int mcd_conv2d_pack_kind(const mcd_conv_geom* g, int pass, int algo) {
  if (!g || algo == MCD_ALGO_DIRECT) return 0;
  if (g->stride > 2) return 0;
  if (pass == 0 ? rowconv_fprop_ok(*g) : rowconv_dgrad_ok(*g)) return 2;
  return pass == 0 ? (packed_fprop_ok(*g) ? 1 : 0) : (packed_dgrad_ok(*g) ? 1 : 0);
}

int mcd_conv2d_kernel_id(const mcd_conv_geom* g, int pass, int y_layout, int algo) {
  if (!g || validate(g) != MCD_OK) return -1;
  if (algo == MCD_ALGO_DIRECT || g->stride > 2) return 9000;
  if (pass == 2) {
    int rows = 0;
    const int bn = umma_wgrad_tile(*g, &rows);
    return (rows == 3 ? 10000 : (rows == 2 ? 8000 : (rows ? 5000 : 4000))) + bn;
  }
  if (pass == 0 ? rowconv_fprop_ok(*g) : rowconv_dgrad_ok(*g)) return 3000 + 16;
  TapProblem p[4];
  int pair = 0, halo = 0;
  if (pass == 0) {
    plan_fprop(*g, p[0]);
    if (packed_fprop_ok(*g)) plan_fprop_packed(*g, p[0]);
  } else {
    plan_dgrad(*g, p);
    if (packed_dgrad_ok(*g)) plan_dgrad_packed(*g, p[0]);
  }
  const int bn = umma_problem_tile(p[0], pass == 0 && y_layout == MCD_OUT_PLANAR_F32, &pair, &halo);
  return (p[0].packed ? 2000 : (halo ? (pair ? 7000 : 6000) : (pair ? 1000 : 0))) + bn;
}

int mcd_pack_weight_rowconv(const float* w_oihw, void* dst, int Cout, int Cin, int R, int S, int Cs, int mode,
                            int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(w_oihw && dst && Cout > 0 && Cin > 0 && R > 0 && S > 0, "pack_weight_rowconv: bad arguments");
  MCD_REQUIRE(mode == 0 || mode == 1, "pack_weight_rowconv: mode must be 0 (fprop) or 1 (dgrad)");
  MCD_REQUIRE((Cs == 8 || Cs == 16) && S <= 8 && R <= 7 && Cs >= (mode ? Cout : Cin) && (mode ? Cin : Cout) <= 32,
              "pack_weight_rowconv: channel stride %d / filter %dx%d not packable", Cs, R, S);
  return rowconv_pack(w_oihw, dst, Cout, Cin, R, S, Cs, mode, (cudaStream_t)stream);
}

size_t mcd_conv2d_streamk_workspace(const mcd_conv_geom* g, int pass, int y_layout, int algo, int* n_flags) {
  int nf = 0;
  size_t bytes = 0;
  if (g && algo != MCD_ALGO_DIRECT && validate(g) == MCD_OK) {
    if (pass == 0) {
      if (!rowconv_fprop_ok(*g) && !packed_fprop_ok(*g)) {
        TapProblem p;
        plan_fprop(*g, p);
        bytes = umma_streamk_workspace(p, y_layout == MCD_OUT_PLANAR_F32, &nf);
      }
    } else if (g->stride == 1 && !rowconv_dgrad_ok(*g) && !packed_dgrad_ok(*g)) {
      TapProblem p[4];
      if (plan_dgrad(*g, p) == 1) bytes = umma_streamk_workspace(p[0], 0, &nf);
    }
  }
  if (n_flags) *n_flags = nf;
  return bytes;
}

size_t mcd_conv2d_wgrad_workspace(const mcd_conv_geom* g, int algo) {
  if (!g || algo == MCD_ALGO_DIRECT) return 0;
  return umma_wgrad_workspace(*g);
}

int mcd_conv2d_wgrad_partials(const mcd_conv_geom* g, int algo, int32_t* layout4) {
  if (!g || !layout4 || algo == MCD_ALGO_DIRECT || validate(g) != MCD_OK) return 0;
  int out[4];
  if (!umma_wgrad_partial_layout(*g, out)) return 0;
  for (int i = 0; i < 4; ++i) layout4[i] = out[i];
  return 1;
}

int mcd_conv2d_wgrad(const void* x_nhwc, const void* dy_nhwc, float* dw_oihw, float* dbias,
                     void* workspace, size_t workspace_bytes, const mcd_conv_geom* g, int accumulate,
                     int algo, int device, void* stream) {
  MCD_ENTER(device);
  int rc = validate(g);
  if (rc != MCD_OK) return rc;
  MCD_REQUIRE(x_nhwc && dy_nhwc, "conv wgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  bool supported = (g->stride == 1 || g->stride == 2);
  bool umma = use_umma(algo, supported, &rc);
  if (rc != MCD_OK) return rc;
  MCD_REQUIRE(dw_oihw || umma, "conv wgrad: dw_oihw == NULL (partial sums only) needs the tcgen05 path");
  rc = umma ? umma_wgrad(x_nhwc, dy_nhwc, dw_oihw, workspace, workspace_bytes, *g, accumulate, st)
            : wgrad_direct(x_nhwc, dy_nhwc, dw_oihw, *g, accumulate, st);
  if (rc != MCD_OK) return rc;
  if (dbias) return colsum(dy_nhwc, dbias, (int64_t)g->N * g->Ho * g->Wo, g->Cout, g->Cout_s, accumulate, st);
  return MCD_OK;
}

}  // extern "C"
