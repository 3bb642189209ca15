/*
 * mcd_sm100.h — C-ABI of libmcd_sm100.so, the B200 (sm_100a) kernel library behind the
 * Maximum-Classifier-Discrepancy (MCD) training / inference step of
 * LittleWat/multichannel-semseg-with-uda.
 *
 * The reference has no FFI: every op below replaces a stock PyTorch call made by the reference's
 * nn.Modules (file:line cited per entry point, paths relative to the reference checkout).  The
 * host-side mirror of those modules (multichannel-semseg-with-uda_b200/{models,loss.py}) binds
 * these symbols with ctypes (mcd_b200/abi.py); INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every entry point returns 0 on success, a negative MCD_E_* code on failure; the message is
 *     available from mcd_last_error() (thread-local).  Nothing throws across the ABI.
 *   - all pointers are DEVICE pointers unless the name ends in _host; the library allocates no
 *     persistent device memory: outputs and workspaces are supplied by the caller.
 *   - `stream` is a cudaStream_t passed as void*; `device` is the CUDA ordinal the pointers live on
 *     (the call is re-entrant and may come from PyTorch's autograd worker thread).
 *   - "nhwc" tensors are 16-bit, channel count C must be a multiple of 8 (16-byte TMA stride rule);
 *     "planar" tensors are NCHW-contiguous.  Image geometry is (N, H, W).
 *   - 16-bit formats: FORWARD tensors (x, y, z, residuals: "f16" below) are IEEE half - its 11 significant bits
 *     keep per-layer activations 8x closer to an fp32 run than bf16 and the ReLU-mask flip rate 8x lower;
 *     GRADIENT tensors (dy, dx, dz, "bf16" below) are bfloat16 - fp32 range, no loss scaling.  tcgen05.mma needs
 *     both GEMM operands in ONE format, so every activation that feeds a weight gradient also exists as a bf16
 *     TWIN (written by the kernel that produces it: mcd_bn_forward / mcd_nchw_f32_to_nhwc); the twin is what
 *     mcd_conv2d_wgrad reads, and ReLU masks are taken from it (only the sign matters).
 *   - conv weights are consumed in a packed 16-bit form produced by mcd_pack_weight():
 *         [rows][taps][kc_pad]   kc_pad = round_up(reduce-channels, 64)
 *     fprop (IEEE half) : rows = Cout, taps in (r,s) order, reduce-channels = Cin
 *     dgrad (bfloat16)  : rows = Cin,  taps flipped,        reduce-channels = Cout
 */
#ifndef MCD_SM100_H
#define MCD_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCD_ABI_VERSION 21

enum {
  MCD_OK = 0,
  MCD_E_INVALID = -1,   /* bad argument / unsupported shape            */
  MCD_E_CUDA = -2,      /* a CUDA runtime / driver call failed          */
  MCD_E_WORKSPACE = -3, /* caller workspace too small                   */
  MCD_E_ARCH = -4       /* device is not sm_100                         */
};

/* conv algorithm selector */
enum {
  MCD_ALGO_AUTO = 0,   /* tcgen05 implicit GEMM when the shape allows it, else direct */
  MCD_ALGO_DIRECT = 1, /* smem-tiled CUDA-core kernel (any shape; cross-check + thin layers) */
  MCD_ALGO_UMMA = 2    /* tcgen05/TMEM/TMA implicit GEMM; MCD_E_INVALID if shape unsupported */
};

/* 16-bit element formats (= the tcgen05 instruction-descriptor codes) */
enum {
  MCD_FMT_F16 = 0, /* IEEE half: forward tensors */
  MCD_FMT_BF16 = 1 /* bfloat16: gradients, wgrad twins */
};

/* output layout of mcd_conv2d_fprop */
enum {
  MCD_OUT_NHWC_BF16 = 0, /* [N,Ho,Wo,Cout] 16-bit nhwc (f16 from fprop), Cout % 8 == 0; name kept from ABI 13 */
  MCD_OUT_PLANAR_F32 = 1 /* [N,Cout,Ho,Wo] fp32 (score maps: seg / decoder heads, any Cout) */
};

/* Geometry of one convolution (nn.Conv2d semantics: models/drn.py:21-23,126-128,195-205). */
typedef struct mcd_conv_geom {
  int32_t N, H, W;     /* input image geometry                       */
  int32_t Cin, Cout;   /* logical channel counts                      */
  int32_t Cin_s;       /* channel stride of the nhwc input  (>= Cin, % 8 == 0) */
  int32_t Cout_s;      /* channel stride of the nhwc output / dy (>= Cout, % 8 == 0) */
  int32_t R, S;        /* filter size                                 */
  int32_t stride, dil, pad;
  int32_t Ho, Wo;      /* output geometry                             */
} mcd_conv_geom;

const char* mcd_last_error(void);
int mcd_version(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
int64_t mcd_launch_count(void);
/* 0 if `device` is an sm_100 part, MCD_E_ARCH otherwise. */
int mcd_check_device(int device);

/* ---- layout ----------------------------------------------------------------------------- */
/* NCHW fp32 -> NHWC 16-bit with channel stride Cs (zero fill of channels >= C), as IEEE half (dst_f16) and / or
 * bfloat16 (dst_bf16): either may be NULL.  Replaces the implicit layout of `Variable(...).cuda()` inputs
 * (adapt_trainer.py:156-160); gradients entering from fp32 score maps use the bf16 output only. */
int mcd_nchw_f32_to_nhwc(const float* src, void* dst_f16, void* dst_bf16, int N, int C, int H, int W, int Cs,
                         int device, void* stream);
/* NHWC 16-bit (format src_fmt, channel stride Cs) -> NCHW fp32 (first C channels). */
int mcd_nhwc_to_nchw_f32(const void* src, int src_fmt, float* dst, int N, int C, int H, int W, int Cs,
                         int device, void* stream);
/* Re-encode a dense 16-bit tensor in the OTHER format (numel % 8 == 0): creates a missing twin. */
int mcd_convert16(const void* src, int src_fmt, void* dst, int64_t numel, int device, void* stream);
/* fp32 OIHW nn.Conv2d weight -> packed 16-bit.  mode 0 = fprop (IEEE half), 1 = dgrad (bfloat16), see header comment.
 * dst holds rows*R*S*kc_pad bf16 where rows = (mode ? Cin : Cout), kc_pad = round_up(mode ? Cout : Cin, 64). */
int mcd_pack_weight(const float* w_oihw, void* dst, int Cout, int Cin, int R, int S, int mode,
                    int device, void* stream);

/* Row-packed pack for thin-channel convolutions (channel stride Cs in {8,16}, S*Cs <= 64, dilation 1):
 * dst[rows][R][64] with k = s*Cs + c.  mode 0 = fprop (rows = Cout, Cs = Cin_s), 1 = dgrad (rows = Cin,
 * Cs = Cout_s, flipped filter).  Which pack a convolution wants: mcd_conv2d_pack_kind(). */
int mcd_pack_weight_rows(const float* w_oihw, void* dst, int Cout, int Cin, int R, int S, int Cs,
                         int mode, int device, void* stream);
/* "Row convolution" pack for the stride-1, dilation-1 stem layers (models/drn.py:126-136: 7x7 6->16, 3x3 16->16;
 * channel stride Cs in {8,16}, S <= 8, R <= 7, <= 32 produced channels): the smem-ready no-swizzle K-major UMMA
 * operand dst[R][(Cs/8)*SP][NB/8][8][8] bf16 with SP = (S <= 4 ? 4 : 8) taps per row and NB = round_up(rows, 16),
 * rows = (mode ? Cin : Cout); i.e. R*(Cs/8)*SP*NB*8 elements.  mode as for mcd_pack_weight_rows. */
int mcd_pack_weight_rowconv(const float* w_oihw, void* dst, int Cout, int Cin, int R, int S, int Cs,
                            int mode, int device, void* stream);
/* Multi-tensor re-pack: one launch for all convolutions of a model (after optimizer.step()).  items_dev:
 * device array of n_items x 12 int64 {w, dst_fprop, dst_dgrad (0 = absent), Cout, Cin, R, S, kind_fprop,
 * kind_dgrad (mcd_conv2d_pack_kind), Cs_fprop, Cs_dgrad, 0}.  The destination buffers must have been produced
 * once by mcd_pack_weight / mcd_pack_weight_rows (their zero padding is kept). */
int mcd_pack_weights_multi(const int64_t* items_dev, int n_items, int blocks_per_item, int device,
                           void* stream);
/* Fused optimizer step of a whole model: torch.optim.SGD semantics (solvers' optimizer_g.step(), reference
 * models/model_util.py:289-302: momentum, weight decay, dampening 0, no Nesterov) on every parameter plus the refresh
 * of the packed bf16 shadows of the convolution weights, one launch.  items_dev: n_items x 16 int64 {param, grad,
 * momentum_buf (0 = none; zero-initialised before the first step), dst_fprop, dst_dgrad (0 = absent / not a
 * convolution weight), Cout, Cin, R, S, kind_fprop, kind_dgrad, Cs_fprop, Cs_dgrad, numel, grad_partials (0 = none; else the
 * workspace of mcd_conv2d_wgrad(dw_oihw = NULL), `grad` is then ignored), ksplit | CoutP << 16 | CinP << 32};
 * hyper_dev: device fp32 {lr, momentum, weight_decay}, read when the kernel RUNS (CUDA-graph replays follow
 * adjust_learning_rate()). */
int mcd_sgd_pack_multi(const int64_t* items_dev, int n_items, const float* hyper_dev, int blocks_per_item,
                       int device, void* stream);
/* 0 = mcd_pack_weight() layout, 1 = mcd_pack_weight_rows(), 2 = mcd_pack_weight_rowconv() for (geometry, pass, algo);
 * pass: 0 = fprop, 1 = dgrad. */
int mcd_conv2d_pack_kind(const mcd_conv_geom* g, int pass, int algo);

/* Which kernel the library launches for (geometry, pass, layout, algo); pass: 0 = fprop, 1 = dgrad, 2 = wgrad.
 * Returns 1000 * kind + tile width BN; kind: 0 = conv_umma_fprop_kernel<BN> (one CTA per 128-pixel tile),
 * 1 = conv_umma_fprop_kernel<256, pair> (cta_group::2, 256 x 256 tile), 2 = row-packed thin-channel mode of kind 0,
 * 3 = conv_umma_rowconv_kernel, 4 = conv_umma_wgrad_kernel<BN>, 5 = conv_umma_wgrad_rows_kernel,
 * 6 = kind 0 with halo-tile staging (one activation box per tile and 64-channel chunk serves all taps), 7 = kind 1
 * with halo-tile staging, 8 = conv_umma_wgrad_toeplitz_kernel (stem layers), 9 = CUDA-core direct kernels,
 * 10 = conv_umma_wgrad_pair_kernel (cta_group::2, 256 co x 256 ci tile); -1 = invalid geometry.  Used by bench.py to attribute measured time to kernels. */
int mcd_conv2d_kernel_id(const mcd_conv_geom* g, int pass, int y_layout, int algo);

/* ---- convolution (nn.Conv2d: models/drn.py:21-23,126-131,171-205; dilated_fcn.py:226-232,632-658,821-823) */
/* y = conv(x, w) (+ bias); x and the nhwc y are f16, w_packed the mode-0 pack.  If `stats` != NULL (fp32 [2*Cout], caller-zeroed) the kernel also
 * accumulates per-channel sum and sum of squares of y for train-mode BatchNorm. */
int mcd_conv2d_fprop(const void* x_nhwc, const void* w_packed, const float* bias, void* y,
                     int y_layout, float* stats, void* sk_partial, int* sk_flags, const mcd_conv_geom* g,
                     int algo, int device, void* stream);
/* Optional stream-K workspace of fprop (pass 0) / dgrad (pass 1): when the pixel-tile count of a layer does not
 * fill the last wave of persistent CTAs (8 x 60x80 pixels = 300 tiles on 148 SMs), the tcgen05 kernels split the
 * (tile, k-block) space evenly over the CTAs instead; tiles cut by a CTA boundary exchange one fp32 partial tile
 * through `sk_partial` (returned size in bytes, contents arbitrary) and `sk_flags` (*n_flags ints, ZEROED by the
 * caller before every call).  Returns 0 / *n_flags = 0 when the geometry does not use it; passing NULL for either
 * pointer selects the plain tile-per-CTA schedule.  Results do not depend on the schedule beyond fp32 summation
 * order. */
size_t mcd_conv2d_streamk_workspace(const mcd_conv_geom* g, int pass, int y_layout, int algo, int* n_flags);
/* Inference form of the unit conv -> BatchNorm (eval: running statistics) -> (+ residual) -> ReLU of
 * models/drn.py:43-59,126-131,195-205 with the BatchNorm folded into the operands by the caller (w_packed =
 * pack(gamma / sqrt(var + eps) * W), bias = beta - mean * gamma / sqrt(var + eps) [+ scaled conv bias]):
 *   y_nhwc (IEEE half) = [relu]( conv(x, w_packed) + bias + res_nhwc )      res_nhwc: IEEE half, y's geometry, or NULL
 * one kernel, no pre-BatchNorm tensor.  tcgen05 paths only: mcd_conv2d_fprop_act_supported(g) says whether geometry
 * g has one (else run mcd_conv2d_fprop + mcd_bn_apply). */
int mcd_conv2d_fprop_act_supported(const mcd_conv_geom* g);
int mcd_conv2d_fprop_act(const void* x_nhwc, const void* w_packed, const float* bias, const void* res_nhwc, int relu,
                         void* y_nhwc, void* sk_partial, int* sk_flags, const mcd_conv_geom* g, int algo, int device,
                         void* stream);
/* dx = conv_transpose(dy, w) (+ add_nhwc): gradient wrt the nhwc input; dy, dx, add are bf16, w_packed is the
 * mode-1 (bf16) pack, relu_src the bf16 twin of the input, bn_y the f16 pre-BatchNorm tensor.
 * add_nhwc (may be NULL): tensor of dx's geometry added in the epilogue - the gradient that reaches the same
 * activation through the identity shortcut of a BasicBlock (models/drn.py:53-58), saving a separate add pass.
 * The backward of the BatchNorm+ReLU unit that PRODUCED the convolution's input (models/drn.py:47-49,126-131)
 * can start in the same epilogue (each may be NULL; all nhwc tensors of dx's geometry):
 *   relu_src_nhwc : the convolution's input x = relu(...):  dx = x > 0 ? dx : 0
 *   bn_y_nhwc, bn_sums (fp32 [2*Cin], caller-zeroed): bn_sums += {sum dx, sum dx * bn_y} per channel, bn_y = the
 *   input of that BatchNorm - the RAW sums mcd_bn_bwd_apply() accepts with sums_kind = 1. */
int mcd_conv2d_dgrad(const void* dy_nhwc, const void* w_packed_dgrad, void* dx_nhwc, const void* add_nhwc,
                     const void* relu_src_nhwc, const void* bn_y_nhwc, float* bn_sums, void* sk_partial,
                     int* sk_flags, const mcd_conv_geom* g, int algo, int device, void* stream);
/* dw (fp32 OIHW) = sum_pixels dy (x) x ; x_nhwc is the bf16 TWIN of the convolution's input, dy bf16;
 * dbias (fp32 [Cout], may be NULL).  accumulate = 0 overwrites,
 * 1 adds to the existing contents (gradient accumulation straight into param.grad / all-reduce buckets).
 * workspace: mcd_conv2d_wgrad_workspace() bytes.
 * dw_oihw == NULL ("partials only"): the tcgen05 kernel leaves its split partial sums in `workspace` as fp32
 * [ksplit][R*S][CoutP][CinP] (mcd_conv2d_wgrad_partials() gives the four numbers; returns 0 for layers that have
 * no such form) and the reduction + OIHW transpose is done by the consumer - mcd_sgd_pack_multi() reads that form
 * directly, which removes one reduction kernel per layer and one write + read of every weight gradient. */
size_t mcd_conv2d_wgrad_workspace(const mcd_conv_geom* g, int algo);
int mcd_conv2d_wgrad_partials(const mcd_conv_geom* g, int algo, int32_t* layout4);
int mcd_conv2d_wgrad(const void* x_nhwc, const void* dy_nhwc, float* dw_oihw, float* dbias,
                     void* workspace, size_t workspace_bytes, const mcd_conv_geom* g, int accumulate,
                     int algo, int device, void* stream);

/* ---- BatchNorm2d (+ReLU, +residual)  (nn.BatchNorm2d defaults eps 1e-5 momentum 0.1:
 *      models/drn.py:34-59,129-131,167-169,199-204; --fix_bn: models/model_util.py:305-310) -------- */
/* per-channel sum / sum-of-squares of an f16 nhwc tensor into caller-zeroed stats[2*C]. */
int mcd_bn_stats(const void* y_nhwc, float* stats, int64_t P, int C, int Cs, int device, void* stream);
/* training: mean/var from stats (count = P), writes scale/shift (fp32 [C] each), save_mean,
 * save_rstd, and updates running_mean / running_var (unbiased) with `momentum`.
 * eval (training == 0): scale/shift from running stats, save_* = running mean / rstd.
 * num_batches_tracked (int64 scalar, may be NULL) is incremented by `training` in training mode
 * (training = k > 1 folds k identical batches: pass momentum' = 1-(1-momentum)^k). */
int mcd_bn_finalize(const float* stats, int64_t P, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps,
                    int training, float* scale, float* shift, float* save_mean, float* save_rstd,
                    int64_t* num_batches_tracked, int C, int device, void* stream);
/* Fused forward: mcd_bn_finalize (for the main and, when res_save_mean_rstd != NULL, the residual/downsample
 * BatchNorm) + mcd_bn_apply in ONE launch.  save_mean_rstd: fp32 [2*C] out (mean, rstd) for the backward.
 * res_nhwc with res_save_mean_rstd == NULL is an identity residual.  y / res are f16; the result is written as f16
 * (z_nhwc) and / or as its bf16 twin (zb_nhwc) - either may be NULL. */
int mcd_bn_forward(const void* y_nhwc, const float* stats, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                   float eps, int training, float* save_mean_rstd, const void* res_nhwc,
                   const float* res_stats, const float* res_gamma, const float* res_beta,
                   float* res_running_mean, float* res_running_var, int64_t* res_num_batches_tracked,
                   float res_momentum, float res_eps, int res_training, float* res_save_mean_rstd, int relu,
                   void* z_nhwc, void* zb_nhwc, int64_t P, int C, int Cs, int device, void* stream);
/* z = act(scale*y + shift + residual'), residual' = res (identity) or rscale*res + rshift
 * (downsample branch, models/drn.py:53-56); res / rscale may be NULL; relu = 0/1. */
int mcd_bn_apply(const void* y_nhwc, const float* scale, const float* shift, const void* res_nhwc,
                 const float* rscale, const float* rshift, int relu, void* z_nhwc, void* zb_nhwc, int64_t P,
                 int C, int Cs, int device, void* stream);
/* backward (dz, dy, dres bf16; z = the bf16 twin, used as the ReLU mask; y, res f16).
 * reductions: g = dz * (relu ? z > 0 : 1);
 *   sums[0:C]  = sum g, sums[C:2C] = sum g * xhat(y)   [, sums[2C:3C] = sum g * xhat(res) if res_mean]
 * sums is caller-zeroed fp32 [3*C]. */
int mcd_bn_bwd_reduce(const void* dz_nhwc, const void* z_nhwc, const void* y_nhwc,
                      const float* mean, const float* rstd, const void* res_nhwc,
                      const float* res_mean, const float* res_rstd, int relu, float* sums, int64_t P,
                      int C, int Cs, int device, void* stream);
/* dy = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat))  (training) or gamma*rstd*g (eval);
 * optional second branch dres with its own gamma/mean/rstd (downsample BN) or, when
 * res_gamma == NULL and dres != NULL, the identity residual gradient dres = g.
 * dgamma/dbeta (fp32 [C]) are written from sums.  sums_kind 0: sums as produced by mcd_bn_bwd_reduce;
 * 1: sums[C:2C] holds the raw sum g*y (mcd_conv2d_dgrad's fused epilogue; no residual BatchNorm branch). */
int mcd_bn_bwd_apply(const void* dz_nhwc, const void* z_nhwc, const void* y_nhwc, const float* gamma,
                     const float* mean, const float* rstd, const float* sums, int training, int relu,
                     void* dy_nhwc, float* dgamma, float* dbeta, const void* res_nhwc,
                     const float* res_gamma, const float* res_mean, const float* res_rstd,
                     int res_training, void* dres_nhwc, float* dres_gamma, float* dres_beta,
                     int sums_kind, int64_t P, int C, int Cs, int device, void* stream);

/* ---- classifier heads --------------------------------------------------------------------- */
/* Depthwise ConvTranspose2d(C,C,16,stride 8,pad 4,groups C,bias=False)
 * (dilated_fcn.py:357-366,465-470,479-491).  x,x2: planar fp32 [N,C,h,w]; w,w2: fp32 [C,1,16,16];
 * out: planar [N,C,8h,8w], bf16 (out_f32 == 0) or fp32.  out = up_w(x) (+ up_w2(x2) when x2 != NULL; if
 * w2 == NULL the same weight is used, i.e. AddFusion up(x1+x2)).
 * Full-resolution tensors ("logits") and their gradients share ONE dtype (autograd's rule): bf16 - half the bytes,
 * what MCDStep uses - or fp32, the drop-in default (`outputs.data.cpu().numpy()` of adapt_tester.py:114-118). */
int mcd_deconv16s8_fwd(const float* x, const float* w, const float* x2, const float* w2, void* out, int out_f32,
                       int N, int C, int h, int w_, int device, void* stream);
/* dx (planar fp32 [N,C,h,w], overwritten) and dw (fp32 [C,256], overwritten) from dout (planar bf16 / fp32). */
int mcd_deconv16s8_bwd(const void* dout, int dout_f32, const float* x, const float* w, float* dx, float* dw,
                       int N, int C, int h, int w_, int device, void* stream);
/* ---- classifier head FUSED with its loss: the full-resolution logits are never materialised -----------------
 * One kernel = x8 upsampling head (learned 16x16/s8 depthwise deconv, models/dilated_fcn.py:357-366,465-491, or
 * nn.Upsample(x8, bilinear) of the multitask decoders :676,817-819 when the filter pointers are NULL) + softmax +
 * loss + the gradients of the loss w.r.t. the head's inputs and filters, for an upstream gradient of 1.
 *   mode 0: CrossEntropyLoss2d (loss.py:7-13) of ONE head;  mode 1: Diff2d (loss.py:93-100) between TWO heads;
 *   mode 2: CrossEntropyLoss2d of TWO heads against the same labels, summed (adapt_trainer.py:171-175:
 *           criterion(outputs1, lbls) + criterion(outputs2, lbls)), labels read once.
 * x / w / dx / dw: host arrays of nheads * nin DEVICE pointers, index head * nin + input; a head's logits are the sum
 * over its `nin` (input, filter) pairs (ScoreAddFusion: up1(x1) + up2(x2); AddFusion: pass x1 + x2 as one input).
 * dx[i] (fp32 [N,C,h,w]) and dw[i] (fp32 [C,256]) are ACCUMULATED INTO (caller-zeroed; two heads may share one dx
 * buffer when they read the same score map); NULL = gradient not wanted.
 * target / cls_weight / ignore_index / wsum: cross entropy only; wsum = device scalar, the (global, all-reduced)
 * normaliser sum_i w[y_i] from mcd_label_weight_sum (NULL: 1).  inv_numel: Diff2d only, 1 / (N*C*H*W*world).
 * acc (fp32[4], caller-zeroed): acc[0] += loss numerator (sum w*nll, or sum |pa - pb|), acc[2] += bad labels. */
int mcd_head_loss(int mode, int nheads, int nin, const float* const* x, const float* const* w, float* const* dx,
                  float* const* dw, const int64_t* target, const float* cls_weight, int64_t ignore_index,
                  const float* wsum, float inv_numel, float* acc, int N, int C, int h, int w_, int device,
                  void* stream);
/* acc2[0] += sum_i w[y_i] over a label map (ignore_index skipped), acc2[1] += labels outside [0, C). */
int mcd_label_weight_sum(const int64_t* target, const float* weight, int64_t ignore_index, int C, float* acc2,
                         int64_t numel, int device, void* stream);
/* nn.Upsample(scale_factor=s, mode='bilinear'), align_corners=False (dilated_fcn.py:676,817-819).
 * x planar fp32 [N,C,h,w] -> out planar bf16 (out_f32 == 0) or fp32 [N,C,s*h,s*w]. */
int mcd_bilinear_up_fwd(const float* x, void* out, int out_f32, int N, int C, int h, int w_, int s,
                        int device, void* stream);
int mcd_bilinear_up_bwd(const void* dout, int dout_f32, float* dx, int N, int C, int h, int w_,
                        int s, int device, void* stream);

/* ---- per-pixel losses (loss.py:7-13,93-100,131-138; dilated_fcn.py:712,958; util.py:44-48) - */
/* CrossEntropyLoss2d: log_softmax(dim=1) + NLLLoss2d(weight, mean, ignore_index).
 * logits planar [N,C,H,W] bf16 (f32 == 0) or fp32; target int64 [N,H,W]; weight fp32 [C] or NULL.
 * acc (fp32 [4], caller-zeroed): acc[0] += sum w*nll, acc[1] += sum w, acc[2] += #bad labels. */
int mcd_ce2d_fwd(const void* logits, int f32, const int64_t* target, const float* weight,
                 int64_t ignore_index, float* acc, int N, int C, int H, int W, int device,
                 void* stream);
/* dlogits (planar, dtype of logits) = gscale[0] * w[y]*(softmax - onehot) / acc[1]; gscale is a device fp32 scalar
 * (the upstream gradient). */
int mcd_ce2d_bwd(const void* logits, int f32, const int64_t* target, const float* weight,
                 int64_t ignore_index, const float* acc, const float* gscale, void* dlogits, int N,
                 int C, int H, int W, int device, void* stream);
/* Diff2d: mean |softmax(a) - softmax(b)| over N*C*H*W.  acc[0] += sum |.|
 * stats (fp32 [N*H*W*4], may be NULL): per-pixel (max_a, 1/sumexp_a, max_b, 1/sumexp_b) written by the forward
 * and consumed by the backward, which then skips its two statistic passes. */
int mcd_diff2d_fwd(const void* a, const void* b, int f32, float* acc, float* stats, int N, int C, int H, int W,
                   int device, void* stream);
int mcd_diff2d_bwd(const void* a, const void* b, int f32, const float* gscale, const float* stats, void* da,
                   void* db, int N, int C, int H, int W, int device, void* stream);
/* The other discrepancy criteria of loss.py:68-171 (get_prob_distance_criterion: every name but 'diff'), forward and
 * backward in one entry point.  mode 1: symkl / nmlsymkl / mysymkl = mean 0.5 (pa - pb)(log pa - log pb);
 * mode 2: jsd = mean 0.5 [pa (log pa - log pm) + pb (log pb - log pm)], pm = softmax((a + b) / 2);
 * mode 3: mis_symkl / spatial_jsd = mean 0.5 [pb (log pb - pa) + pa (log pa - pb)] (kl_div fed with probabilities).
 * acc (may be NULL): acc[0] += sum over all elements (divide by N*C*H*W); da / db (may be NULL, dtype of a / b,
 * overwritten) = gscale[0] * inv_numel * d(sum)/d(a|b), exact derivatives through both arguments. */
int mcd_pairdist(int mode, const void* a, const void* b, int f32, float* acc, const float* gscale, void* da, void* db,
                 float inv_numel, int N, int C, int H, int W, int device, void* stream);
/* F.mse_loss(pred, target) with pred planar bf16 / fp32, target planar fp32: acc[0] += sum (p-t)^2 */
int mcd_mse_fwd(const void* pred, int f32, const float* target, float* acc, int64_t numel, int device,
                void* stream);
int mcd_mse_bwd(const void* pred, int f32, const float* target, const float* gscale, void* dpred,
                int64_t numel, int device, void* stream);
/* boundary head: p = (sigmoid(h1)+sigmoid(h2)+sigmoid(h3))/3 (dilated_fcn.py:913-923) followed by
 * bce2d (loss.py:131-138).  h* planar bf16 / fp32 [numel]; target fp32 in {0,1}.
 * tsum (fp32[1]) = sum(target) from mcd_sum_f32, all-reduced by the caller under data parallelism: beta =
 * 1 - tsum / numel_global is the batch-GLOBAL class balance of loss.py:133 (numel_global <= 0: numel). */
int mcd_sum_f32(const float* x, float* acc, int64_t numel, int device, void* stream);
int mcd_sigmoid3_bce_fwd(const void* h1, const void* h2, const void* h3, int f32, const float* target,
                         const float* tsum, float* acc, void* p_out, int64_t numel, int64_t numel_global,
                         int device, void* stream);
int mcd_sigmoid3_bce_bwd(const void* h1, const void* h2, const void* h3, int f32, const float* target,
                         const float* tsum, const float* gscale, void* dh1, void* dh2, void* dh3,
                         int64_t numel, int64_t numel_global, int device, void* stream);
/* bce2d(input, target) on a probability map (loss.py:130-138): p, target fp32 [numel]; acc[0] += sum w * bce with
 * w = 1 - beta + (2 beta - 1) t and torch's log clamp at -100; dp = gscale / numel * w (p - t) / max(p (1-p), 1e-12). */
int mcd_bce2d_fwd(const float* p, const float* target, const float* tsum, float* acc, int64_t numel,
                  int64_t numel_global, int device, void* stream);
int mcd_bce2d_bwd(const float* p, const float* target, const float* tsum, const float* gscale, float* dp,
                  int64_t numel, int64_t numel_global, int device, void* stream);
/* get_boundary of models/dilated_fcn.py:769-773: out (fp32 [N,H,W]) = 1 where the 3x3 max of x differs from the
 * 3x3 min (max_pool2d(x) != -max_pool2d(-x), stride 1, padding 1), else 0.  x: int64 labels (is_int64) or fp32. */
int mcd_label_boundary(const void* x, int is_int64, float* out, int N, int H, int W, int device, void* stream);
/* Testers: argmax over channels [0, C_arg) (first max wins, like torch.max) + entropy partial
 * acc[0] += sum_c p*log(p+1e-6) over all C channels (adapt_tester.py:104-124, util.py:44-48). */
int mcd_argmax_entropy(const void* logits, int f32, int64_t* labels, float* acc, int N, int C, int C_arg,
                       int H, int W, int device, void* stream);

/* ---- input pipeline and evaluation counts ----------------------------- */
/* transform.py:302-314 get_img_transform (ToTensor: uint8 / 255; Normalize: (x - mean) / std) and the channel
 * concatenation of datasets.py:667-695, over a batch, in one pass.  Plane s is [N*H*W][src_stride[s]] uint8 (HWC
 * bytes as PIL decodes them); channels [src_first[s], src_first[s] + src_count[s]) are taken.  src_raw[s] != 0 marks a
 * label plane (the boundary map of the 7-channel input): value relabel_from becomes relabel_to (ReLabel(255, 1)),
 * converted with .float(), neither scaled nor normalised.  mean / stdv: HOST arrays with one entry per OUTPUT
 * channel (NULL = 0 / 1).  Outputs (any subset): NCHW fp32 [N,C,H,W] - bit-identical to torchvision's fp32
 * arithmetic; NHWC IEEE half and its bfloat16 twin [N,H,W,CP=8] (zero padded) - the stem convolution's operand. */
int mcd_input_transform(const void* const* src, const int* src_stride, const int* src_first, const int* src_count,
                        const int* src_raw, int n_src, const float* mean, const float* stdv, int relabel_from,
                        int relabel_to, float* out_nchw, void* out_nhwc_f16, void* out_nhwc_bf16, int CP, int N, int H,
                        int W, int device, void* stream);
/* transform.py:21-48,317-324 ToLabel + ReLabel(background_id, n_class - 1): uint8 label map -> int64. */
int mcd_relabel_u8(const void* src, int64_t* dst, int from, int to, int64_t numel, int device, void* stream);
/* adapt_tester.py:124-126 Image.resize(test_img_shape, NEAREST) of uint8 label maps [N,H,W] -> [N,OH,OW];
 * ytab / xtab: DEVICE int32 source indices per output row / column (mcd_b200.pipeline.pil_nearest_table). */
int mcd_resize_nearest_u8(const void* src, void* dst, const int* ytab_dev, const int* xtab_dev, int N, int H, int W,
                          int OH, int OW, int device, void* stream);
/* eval.py:21-23 fast_hist(a = ground truth, b = prediction, n): hist[n*a + b] += 1 for 0 <= a < n (int64 counts,
 * accumulated into hist[0 .. n*n)); hist[n*n] counts predictions outside [0, n), for which numpy's bincount would
 * leave the n x n table.  Labels: uint8 or int64.  n <= 104. */
int mcd_fast_hist(const void* gt, int gt_is_int64, const void* pred, int pred_is_int64, int n, int64_t numel,
                  int64_t* hist, int device, void* stream);
/* transform.py:285-294 unnormalize: uint8((x * std + mean) * 255) in float64, x fp32 [N,3,H,W] -> uint8 [N,H,W,3]. */
int mcd_unnormalize_u8(const float* x, void* dst, const double* mean3, const double* std3, int N, int H, int W,
                       int device, void* stream);

/* ---- option surface around the hot path: fusion heads, FuseDRNSegBase, torch_up ----------- */
/* FuseDRNSegBase.forward `x = torch.add(x, x_dK)` (models/dilated_fcn.py:308-329) and AddFusion on ver2 trunk
 * features (models/fusion.py:24-29): z = a + b on dense IEEE-half nhwc activations, written as IEEE half and, when
 * out_bf16 != NULL, as the bfloat16 twin.  numel % 8 == 0. */
int mcd_add_nhwc(const void* a_f16, const void* b_f16, void* out_f16, void* out_bf16, int64_t numel, int device,
                 void* stream);
/* GateFusion (models/fusion.py:17-21): out = x1 * sigmoid(a) + x2 * (1 - sigmoid(a)), a = conv(cat(x1, x2));
 * all planar fp32 of one shape.  Backward: any of dx1 / dx2 / dgate_logits may be NULL. */
int mcd_gate_fuse_fwd(const float* x1, const float* x2, const float* gate_logits, float* out, int64_t numel,
                      int device, void* stream);
int mcd_gate_fuse_bwd(const float* x1, const float* x2, const float* gate_logits, const float* dout, float* dx1,
                      float* dx2, float* dgate_logits, int64_t numel, int device, void* stream);
/* F.softmax over dim 1 of a planar fp32 [N,C,HW] tensor (ScoreGateFusion, models/fusion.py:13-15) and its backward
 * dx = p * (dp - sum_c dp * p). */
int mcd_softmax_ch_fwd(const float* x, float* p, int N, int C, int64_t HW, int device, void* stream);
int mcd_softmax_ch_bwd(const float* p, const float* dp, float* dx, int N, int C, int64_t HW, int device,
                       void* stream);
/* torch.cat([a, b], 1) of planar fp32 tensors [N,Ca,HW], [N,Cb,HW] (models/fusion.py:16,37,47) and its backward
 * (either output of the split may be NULL). */
int mcd_cat2_f32(const float* a, int Ca, const float* b, int Cb, float* out, int N, int64_t HW, int device,
                 void* stream);
int mcd_split2_f32(const float* src, float* a, int Ca, float* b, int Cb, int N, int64_t HW, int device,
                   void* stream);
/* F.sigmoid on fp32 (seg2bd boundary maps, models/dilated_fcn.py:965-966); backward from the OUTPUT y. */
int mcd_sigmoid_fwd(const float* x, float* y, int64_t numel, int device, void* stream);
int mcd_sigmoid_bwd(const float* y, const float* dy, float* dx, int64_t numel, int device, void* stream);
/* out = a + b (+ c, may be NULL) on fp32: the shortcut decoders' h1 + h2 + h3 (models/dilated_fcn.py:875,884,904). */
int mcd_add3_f32(const float* a, const float* b, const float* c, float* out, int64_t numel, int device,
                 void* stream);
/* ProbCrossEntropyLoss2d (loss.py:16-30; the criterion adapt_mfnet_trainer.py:149 selects for the Gate fusions):
 * NLLLoss2d(weight, mean, ignore_index)(log(p), target) on a planar fp32 probability map p [N,C,H,W].
 * acc (fp32 [4], caller-zeroed) as in mcd_ce2d_fwd: acc[0] += sum w * -log p[y], acc[1] += sum w, acc[2] += #bad labels.
 * Backward: dp = gscale[0] * (c == y ? -w[y] / (p[y] * acc[1]) : 0), dense over all channels. */
int mcd_prob_ce2d_fwd(const float* p, const int64_t* target, const float* weight, int64_t ignore_index, float* acc,
                      int N, int C, int H, int W, int device, void* stream);
int mcd_prob_ce2d_bwd(const float* p, const int64_t* target, const float* weight, int64_t ignore_index,
                      const float* acc, const float* gscale, float* dp, int N, int C, int H, int W, int device,
                      void* stream);
/* nn.UpsamplingBilinear2d(scale_factor=s) = bilinear with align_corners=True (`use_torch_up`,
 * models/dilated_fcn.py:354-355,443-444).  Same tensor conventions as mcd_bilinear_up_fwd / _bwd; any s >= 1. */
int mcd_bilinear_ac_up_fwd(const float* x, void* out, int out_f32, int N, int C, int h, int w_, int s, int device,
                           void* stream);
int mcd_bilinear_ac_up_bwd(const void* dout, int dout_f32, float* dx, int N, int C, int h, int w_, int s, int device,
                           void* stream);

/* ---- optimiser (models/model_util.py:289-302: SGD momentum / weight decay, torch semantics) - */
int mcd_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t numel, float lr,
                 float momentum, float weight_decay, int first_step, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCD_SM100_H */
