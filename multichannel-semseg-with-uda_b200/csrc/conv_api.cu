// conv_api.cu — C-ABI entry points for convolution: validation, lowering to tap problems
// (conv_plan.h) and dispatch to the tcgen05 (conv_umma.cu) or CUDA-core (conv_direct.cu) kernels.
#include "common.cuh"
#include "conv_plan.h"

namespace mcd {
int launch_direct_problem(const void* src, const void* w, const float* bias, void* out, int planar,
                          const void* addend, const TapProblem& p, cudaStream_t st);
int wgrad_direct(const void* x, const void* dy, float* dw, const mcd_conv_geom& g, int accumulate,
                 cudaStream_t st);
int colsum(const void* t, float* out, int64_t P, int C, int Cs, int accumulate, cudaStream_t st);
bool umma_problem_supported(const TapProblem& p);
int launch_umma_problem(const void* src, const void* w, const float* bias, void* out, int planar,
                        float* stats, const void* addend, const TapProblem& p, cudaStream_t st);
size_t umma_wgrad_workspace(const mcd_conv_geom& g);
int umma_wgrad(const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes,
               const mcd_conv_geom& g, int accumulate, cudaStream_t st);
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
                     int y_layout, float* stats, const mcd_conv_geom* g, int algo, int device,
                     void* stream) {
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
    if (packed_fprop_ok(*g)) plan_fprop_packed(*g, p);   // w_packed is then the mcd_pack_weight_rows layout
    return launch_umma_problem(x_nhwc, w_packed, bias, y, planar, stats, nullptr, p, st);
  }
  rc = launch_direct_problem(x_nhwc, w_packed, bias, y, planar, nullptr, p, st);
  if (rc != MCD_OK) return rc;
  if (stats) {
    MCD_REQUIRE(!planar, "conv fprop: fused BN statistics need the nhwc output layout");
    return bn_stats_launch(y, stats, (int64_t)g->N * g->Ho * g->Wo, g->Cout, g->Cout_s, st);
  }
  return MCD_OK;
}

int mcd_conv2d_dgrad(const void* dy_nhwc, const void* w_packed_dgrad, void* dx_nhwc, const void* add_nhwc,
                     const mcd_conv_geom* g, int algo, int device, void* stream) {
  MCD_ENTER(device);
  int rc = validate(g);
  if (rc != MCD_OK) return rc;
  MCD_REQUIRE(dy_nhwc && w_packed_dgrad && dx_nhwc, "conv dgrad: null pointer");
  MCD_REQUIRE(g->stride <= 2, "conv dgrad: stride %d unsupported", g->stride);
  cudaStream_t st = (cudaStream_t)stream;
  TapProblem p[4];
  int np = plan_dgrad(*g, p);
  if (algo != MCD_ALGO_DIRECT && packed_dgrad_ok(*g) && umma_problem_supported(p[0])) {
    plan_dgrad_packed(*g, p[0]);                          // w_packed_dgrad: mcd_pack_weight_rows mode 1
    return launch_umma_problem(dy_nhwc, w_packed_dgrad, nullptr, dx_nhwc, 0, nullptr, add_nhwc, p[0], st);
  }
  bool any_empty = false;
  for (int i = 0; i < np; ++i) any_empty |= (p[i].ntaps == 0);
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
    rc = umma ? launch_umma_problem(dy_nhwc, w_packed_dgrad, nullptr, dx_nhwc, 0, nullptr, add_nhwc, p[i], st)
              : launch_direct_problem(dy_nhwc, w_packed_dgrad, nullptr, dx_nhwc, 0, add_nhwc, p[i], st);
    if (rc != MCD_OK) return rc;
  }
  return MCD_OK;
}

int mcd_conv2d_pack_kind(const mcd_conv_geom* g, int pass, int algo) {
  if (!g || algo == MCD_ALGO_DIRECT) return 0;
  if (g->stride > 2) return 0;
  return pass == 0 ? (packed_fprop_ok(*g) ? 1 : 0) : (packed_dgrad_ok(*g) ? 1 : 0);
}

size_t mcd_conv2d_wgrad_workspace(const mcd_conv_geom* g, int algo) {
  if (!g || algo == MCD_ALGO_DIRECT) return 0;
  return umma_wgrad_workspace(*g);
}

int mcd_conv2d_wgrad(const void* x_nhwc, const void* dy_nhwc, float* dw_oihw, float* dbias,
                     void* workspace, size_t workspace_bytes, const mcd_conv_geom* g, int accumulate,
                     int algo, int device, void* stream) {
  MCD_ENTER(device);
  int rc = validate(g);
  if (rc != MCD_OK) return rc;
  MCD_REQUIRE(x_nhwc && dy_nhwc && dw_oihw, "conv wgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  bool supported = (g->stride == 1 || g->stride == 2);
  bool umma = use_umma(algo, supported, &rc);
  if (rc != MCD_OK) return rc;
  rc = umma ? umma_wgrad(x_nhwc, dy_nhwc, dw_oihw, workspace, workspace_bytes, *g, accumulate, st)
            : wgrad_direct(x_nhwc, dy_nhwc, dw_oihw, *g, accumulate, st);
  if (rc != MCD_OK) return rc;
  if (dbias) return colsum(dy_nhwc, dbias, (int64_t)g->N * g->Ho * g->Wo, g->Cout, g->Cout_s, accumulate, st);
  return MCD_OK;
}

}  // extern "C"
