// pipeline.cu — the byte-side neighbours of the MCD step: loader output -> network input, and
// network output -> evaluation counts.  All HBM-bound integer / byte work, one pass each.
//
//   mcd_input_transform   uint8 HWC image planes -> ToTensor (/255) -> Normalize (mean, std) -> channel concat
//                         (transform.py:302-314, datasets.py:667-695) -> NHWC IEEE half + bf16 twin (what the stem
//                         convolution reads) and / or NCHW fp32 (what the reference's loader hands to the trainer)
//   mcd_relabel_u8        ToLabel + ReLabel(255, n_class - 1) (transform.py:21-48,317-324): uint8 -> int64
//   mcd_resize_nearest_u8 Image.resize(size, NEAREST) of the predicted label map (adapt_tester.py:124-126), index
//                         tables supplied by the host
//   mcd_fast_hist         eval.py:21-23 confusion matrix counts
//   mcd_unnormalize_u8    transform.py:285-294 (x * std + mean) * 255 -> uint8 HWC
#include "common.cuh"

namespace mcd {

constexpr int kMaxSrc = 3;
constexpr int kMaxCh = 8;
constexpr int kPixPerBlock = 256;

struct TransformArgs {
  const uint8_t* src[kMaxSrc];   // [npix][stride] bytes
  int stride[kMaxSrc];           // bytes per pixel of the plane
  int first[kMaxSrc];            // first channel taken
  int count[kMaxSrc];            // channels taken
  int raw[kMaxSrc];              // 1: label plane - no /255, no normalisation, relabel_from -> relabel_to
  int n_src, C, CP;
  int relabel_from, relabel_to;
  float mean[kMaxCh], stdv[kMaxCh];
  float* out_nchw;               // [N][C][H*W] or NULL
  uint16_t* out_f16;             // [npix][CP] or NULL
  uint16_t* out_bf16;            // [npix][CP] or NULL
  int64_t npix, hw;
};

// One block = 256 consecutive pixels.  The source bytes of the block are contiguous in every plane: they are staged in
// shared memory with 4-byte loads (256 * stride bytes is a multiple of 4, bases are checked by the host), each thread
// then owns one pixel.  Arithmetic follows torchvision bit for bit: float(v) / 255 (IEEE division), (x - mean) / std.
__global__ void __launch_bounds__(kPixPerBlock)
input_transform_kernel(const __grid_constant__ TransformArgs a) {
  extern __shared__ __align__(16) uint8_t stage[];
  const int tid = threadIdx.x;
  const int64_t p0 = (int64_t)blockIdx.x * kPixPerBlock;
  const int64_t left = a.npix - p0;
  const int npx = left < kPixPerBlock ? (int)left : kPixPerBlock;
  int off[kMaxSrc];
  int o = 0;
#pragma unroll
  for (int s = 0; s < kMaxSrc; ++s) {
    off[s] = o;
    if (s < a.n_src) o += kPixPerBlock * a.stride[s];
  }
#pragma unroll
  for (int s = 0; s < kMaxSrc; ++s) {
    if (s >= a.n_src) continue;
    const uint8_t* g = a.src[s] + p0 * a.stride[s];
    const int nbytes = npx * a.stride[s];
    if (npx == kPixPerBlock) {
      const uint32_t* g4 = reinterpret_cast<const uint32_t*>(g);
      uint32_t* s4 = reinterpret_cast<uint32_t*>(stage + off[s]);
      for (int i = tid; i < nbytes / 4; i += kPixPerBlock) s4[i] = __ldcs(g4 + i);
    } else {
      for (int i = tid; i < nbytes; i += kPixPerBlock) stage[off[s] + i] = g[i];
    }
  }
  __syncthreads();
  if (tid >= npx) return;
  float v[kMaxCh];
#pragma unroll
  for (int c = 0; c < kMaxCh; ++c) v[c] = 0.f;
  int c = 0;
#pragma unroll
  for (int s = 0; s < kMaxSrc; ++s) {
    if (s >= a.n_src) continue;
    const uint8_t* px = stage + off[s] + tid * a.stride[s] + a.first[s];
    for (int k = 0; k < a.count[s]; ++k, ++c) {
      int b = px[k];
      float x;
      if (a.raw[s]) {
        if (b == a.relabel_from) b = a.relabel_to;
        x = (float)b;
      } else {
        x = __fdiv_rn((float)b, 255.f);
        x = __fdiv_rn(__fsub_rn(x, a.mean[c]), a.stdv[c]);
      }
#pragma unroll
      for (int j = 0; j < kMaxCh; ++j)       // register array: no dynamic indexing
        if (j == c) v[j] = x;
    }
  }
  const int64_t p = p0 + tid;
  if (a.out_nchw) {
    const int64_t n = p / a.hw, r = p % a.hw;
#pragma unroll
    for (int j = 0; j < kMaxCh; ++j)
      if (j < a.C) __stcs(a.out_nchw + (n * a.C + j) * a.hw + r, v[j]);
  }
  if (a.out_f16) {
    *reinterpret_cast<uint4*>(a.out_f16 + p * 8) = pack8h(v);
  }
  if (a.out_bf16) {
    *reinterpret_cast<uint4*>(a.out_bf16 + p * 8) = pack8(v);
  }
}

__global__ void __launch_bounds__(256)
relabel_u8_kernel(const uint8_t* __restrict__ src, int64_t* __restrict__ dst, int from, int to, int64_t n) {
  // 4 labels per thread: one 4-byte load, four 8-byte stores (two 16-byte vectors)
  const int64_t n4 = n / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t w = __ldcs(reinterpret_cast<const uint32_t*>(src) + i);
    int64_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = (w >> (8 * k)) & 0xff;
      o[k] = b == from ? to : b;
    }
    longlong2* d = reinterpret_cast<longlong2*>(dst + 4 * i);
    d[0] = make_longlong2(o[0], o[1]);
    d[1] = make_longlong2(o[2], o[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = n4 * 4 + threadIdx.x;
    const int b = src[i];
    dst[i] = b == from ? to : b;
  }
}

__global__ void __launch_bounds__(256)
resize_nearest_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int* __restrict__ ytab,
                         const int* __restrict__ xtab, int N, int H, int W, int OH, int OW) {
  const int64_t total = (int64_t)N * OH * OW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
    const int64_t n = i / ((int64_t)OW * OH);
    dst[i] = src[(n * H + ytab[oy]) * W + xtab[ox]];
  }
}

// eval.py:21-23:  k = (a >= 0) & (a < n);  bincount(n * a[k] + b[k], minlength = n^2).  Counts in shared memory per
// block (n <= 104), one 64-bit global atomic per non-empty bin per block.  A prediction outside [0, n) would make
// numpy's bincount overflow the n x n table (the reference then fails in reshape): it is counted in hist[n * n].
template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
fast_hist_kernel(const TA* __restrict__ gt, const TB* __restrict__ pred, int n, int64_t numel,
                 unsigned long long* __restrict__ hist) {
  extern __shared__ uint32_t bins[];
  const int nb = n * n + 1;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) bins[i] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = (int64_t)gt[i], b = (int64_t)pred[i];
    if (a >= 0 && a < n) atomicAdd(bins + ((b >= 0 && b < n) ? (int)(a * n + b) : n * n), 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x)
    if (bins[i]) atomicAdd(hist + i, (unsigned long long)bins[i]);
}

// transform.py:285-294: np.uint8((x * std + mean) * 255) on an HWC float64 array; x fp32 NCHW here.  The float -> uint8
// conversion truncates toward zero and wraps modulo 256 (numpy's C cast through a wider integer).
__global__ void __launch_bounds__(256)
unnormalize_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ dst, double m0, double m1, double m2, double s0,
                      double s1, double s2, int64_t npix, int64_t hw) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = p / hw, r = p % hw;
    const double mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double v = __dmul_rn(__dadd_rn(__dmul_rn((double)x[(n * 3 + c) * hw + r], sd[c]), mean[c]), 255.0);
      dst[p * 3 + c] = (uint8_t)((long long)v & 0xff);
    }
  }
}

static inline int grid_of(int64_t items) { return (int)max64(1, min64((items + 255) / 256, 148 * 16)); }

}  // namespace mcd

using namespace mcd;

extern "C" {

int mcd_input_transform(const void* const* src, const int* src_stride, const int* src_first, const int* src_count,
                        const int* src_raw, int n_src, const float* mean, const float* stdv, int relabel_from,
                        int relabel_to, float* out_nchw, void* out_nhwc_f16, void* out_nhwc_bf16, int CP, int N, int H,
                        int W, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(src && src_stride && src_first && src_count && src_raw && n_src >= 1 && n_src <= kMaxSrc,
              "input_transform: 1..%d source planes", kMaxSrc);
  MCD_REQUIRE(N > 0 && H > 0 && W > 0 && (out_nchw || out_nhwc_f16 || out_nhwc_bf16), "input_transform: bad arguments");
  TransformArgs a{};
  int C = 0, smem = 0;
  for (int s = 0; s < n_src; ++s) {
    MCD_REQUIRE(src[s] && src_stride[s] >= 1 && src_stride[s] <= 16 && src_first[s] >= 0 && src_count[s] >= 1 &&
                    src_first[s] + src_count[s] <= src_stride[s],
                "input_transform: plane %d: channels [%d, %d) of %d", s, src_first[s], src_first[s] + src_count[s],
                src_stride[s]);
    MCD_REQUIRE(((uintptr_t)src[s] & 3) == 0, "input_transform: plane %d is not 4-byte aligned", s);
    a.src[s] = (const uint8_t*)src[s];
    a.stride[s] = src_stride[s]; a.first[s] = src_first[s]; a.count[s] = src_count[s]; a.raw[s] = src_raw[s] != 0;
    C += src_count[s];
    smem += kPixPerBlock * src_stride[s];
  }
  MCD_REQUIRE(C <= kMaxCh, "input_transform: %d channels (max %d)", C, kMaxCh);
  MCD_REQUIRE((!out_nhwc_f16 && !out_nhwc_bf16) || CP == 8, "input_transform: NHWC outputs are padded to 8 channels");
  for (int c = 0; c < C; ++c) {
    a.mean[c] = mean ? mean[c] : 0.f;
    a.stdv[c] = stdv ? stdv[c] : 1.f;
  }
  a.n_src = n_src; a.C = C; a.CP = CP;
  a.relabel_from = relabel_from; a.relabel_to = relabel_to;
  a.out_nchw = out_nchw; a.out_f16 = (uint16_t*)out_nhwc_f16; a.out_bf16 = (uint16_t*)out_nhwc_bf16;
  a.hw = (int64_t)H * W; a.npix = a.hw * N;
  const int64_t blocks = (a.npix + kPixPerBlock - 1) / kPixPerBlock;
  MCD_REQUIRE(blocks < (1ll << 31), "input_transform: too many pixels");
  input_transform_kernel<<<(int)blocks, kPixPerBlock, smem, (cudaStream_t)stream>>>(a);
  return check_launch("input_transform");
}

int mcd_relabel_u8(const void* src, int64_t* dst, int from, int to, int64_t numel, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(src && dst && numel > 0, "relabel_u8: bad arguments");
  MCD_REQUIRE(((uintptr_t)src & 3) == 0 && ((uintptr_t)dst & 15) == 0, "relabel_u8: unaligned buffers");
  relabel_u8_kernel<<<grid_of(numel / 4 + 1), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)src, dst, from, to, numel);
  return check_launch("relabel_u8");
}

int mcd_resize_nearest_u8(const void* src, void* dst, const int* ytab_dev, const int* xtab_dev, int N, int H, int W,
                          int OH, int OW, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(src && dst && ytab_dev && xtab_dev && N > 0 && H > 0 && W > 0 && OH > 0 && OW > 0,
              "resize_nearest_u8: bad arguments");
  resize_nearest_u8_kernel<<<grid_of((int64_t)N * OH * OW), 256, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)src, (uint8_t*)dst, ytab_dev, xtab_dev, N, H, W, OH, OW);
  return check_launch("resize_nearest_u8");
}

int mcd_fast_hist(const void* gt, int gt_is_int64, const void* pred, int pred_is_int64, int n, int64_t numel,
                  int64_t* hist, int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(gt && pred && hist && numel > 0, "fast_hist: bad arguments");
  MCD_REQUIRE(n >= 1 && n <= 104, "fast_hist: n=%d (1..104: the counts of a block live in shared memory)", n);
  const int smem = (n * n + 1) * 4;
  const int grid = (int)max64(1, min64((numel + 256 * 16 - 1) / (256 * 16), 148 * 8));
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* h = reinterpret_cast<unsigned long long*>(hist);
  if (gt_is_int64 && pred_is_int64)
    fast_hist_kernel<int64_t, int64_t><<<grid, 256, smem, st>>>((const int64_t*)gt, (const int64_t*)pred, n, numel, h);
  else if (gt_is_int64)
    fast_hist_kernel<int64_t, uint8_t><<<grid, 256, smem, st>>>((const int64_t*)gt, (const uint8_t*)pred, n, numel, h);
  else if (pred_is_int64)
    fast_hist_kernel<uint8_t, int64_t><<<grid, 256, smem, st>>>((const uint8_t*)gt, (const int64_t*)pred, n, numel, h);
  else
    fast_hist_kernel<uint8_t, uint8_t><<<grid, 256, smem, st>>>((const uint8_t*)gt, (const uint8_t*)pred, n, numel, h);
  return check_launch("fast_hist");
}

int mcd_unnormalize_u8(const float* x, void* dst, const double* mean3, const double* std3, int N, int H, int W,
                       int device, void* stream) {
  MCD_ENTER(device);
  MCD_REQUIRE(x && dst && mean3 && std3 && N > 0 && H > 0 && W > 0, "unnormalize_u8: bad arguments");
  const int64_t hw = (int64_t)H * W;
  unnormalize_u8_kernel<<<grid_of(hw * N), 256, 0, (cudaStream_t)stream>>>(x, (uint8_t*)dst, mean3[0], mean3[1], mean3[2],
                                                                          std3[0], std3[1], std3[2], hw * N, hw);
  return check_launch("unnormalize_u8");
}

}  // extern "C"
