"""Loader output -> network input and network output -> evaluation counts on the GPU (csrc/pipeline.cu; SURVEY 8f rows 2
and 3).  Everything here is byte / integer work and bit-exact against the reference's CPU code:

    transform_images   transform.py:302-314 (ToTensor + Normalize) + datasets.py:667-695 (channel concatenation)
    relabel            transform.py:21-48,317-324 (ToLabel + ReLabel(255, n_class - 1))
    resize_nearest     adapt_tester.py:124-126 (PIL Image.resize(size, NEAREST) of the predicted labels)
    fast_hist          eval.py:21-23
    unnormalize        transform.py:285-294

The PIL decode and the BILINEAR / NEAREST `Scale` to img_shape stay in the loader (CPU, per file); this module starts at
the decoded uint8 HWC arrays, batched: uint8 is what crosses PCIe (6 + 1 bytes per pixel instead of 24 + 8).
"""
import ctypes

import torch

from . import abi, ops

U8, I64, F32 = torch.uint8, torch.int64, torch.float32
IMAGENET_MEAN = (.485, .456, .406, .485, .485, .485)      # transform.py:307 (6 entries; zip() uses the first c)
IMAGENET_STD = (.229, .224, .225, .229, .229, .229)
CITY_MEAN = (0.290101, 0.328081, 0.286964)                # transform.py:311
CITY_STD = (0.182954, 0.186566, 0.184475)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _dev(t):
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _plane(t):
    """[N,H,W,c] (or [N,H,W]) contiguous uint8 on the GPU"""
    assert t.is_cuda and t.dtype == U8 and t.is_contiguous() and t.dim() in (3, 4), \
        "image planes are contiguous uint8 [N,H,W,c] tensors on the GPU"
    return t if t.dim() == 4 else t.unsqueeze(-1)


def normalisation(normalize_way, n_ch):
    """per-plane (mean, std) exactly as transform.py:306-315 applies them to an image with n_ch channels"""
    if normalize_way == "imagenet":
        return IMAGENET_MEAN[:n_ch], IMAGENET_STD[:n_ch]
    if normalize_way == "city":
        return CITY_MEAN[:n_ch], CITY_STD[:n_ch]
    return (0.0,) * n_ch, (1.0,) * n_ch            # "No normalization..."


def transform_images(planes, normalize_way="imagenet", label_planes=(), channels=None, out="nhwc", twin=None):
    """`planes`: uint8 [N,H,W,c] tensors (one per decoded file: rgb, hha, ...), each normalised like a separate image
    by the reference's img_transform and concatenated along channels (datasets.py:667-680); `label_planes`: uint8
    [N,H,W] maps appended as float channels after ReLabel(255, 1) (the boundary channel, datasets.py:688-693).
    `channels`: optional (first, count) per plane to take a sub-range (the MFNet streams).

    out = "nchw": fp32 [N,C,H,W], bit-identical to the loader's tensor.
    out = "nhwc": the stem convolution's operand (16-bit channels_last, C padded to 8) - IEEE half, and while autograd
                  records (or twin=True) the bfloat16 twin carrying the half tensor as `_mcd_h16` (ops.to_nhwc)."""
    planes = [_plane(p) for p in planes]
    labels = [_plane(p) for p in label_planes]
    allp = planes + labels
    n, h, w, _ = allp[0].shape
    for p in allp:
        assert tuple(p.shape[:3]) == (n, h, w) and p.device == allp[0].device
    if channels is None:
        channels = [(0, p.shape[3]) for p in planes]
    channels = list(channels) + [(0, 1)] * len(labels)
    assert len(allp) <= 3, "at most three planes per call"
    mean, std = [], []
    for p, (first, cnt) in zip(planes, channels):
        m, s = normalisation(normalize_way, p.shape[3])
        mean += list(m[first:first + cnt])
        std += list(s[first:first + cnt])
    mean += [0.0] * len(labels)
    std += [1.0] * len(labels)
    c = len(mean)
    k = len(allp)
    src = (ctypes.c_void_p * k)(*[p.data_ptr() for p in allp])
    stride = (ctypes.c_int * k)(*[p.shape[3] for p in allp])
    first = (ctypes.c_int * k)(*[f for f, _ in channels])
    count = (ctypes.c_int * k)(*[cnt for _, cnt in channels])
    raw = (ctypes.c_int * k)(*([0] * len(planes) + [1] * len(labels)))
    mean_a, std_a = (ctypes.c_float * c)(*mean), (ctypes.c_float * c)(*std)
    dev = allp[0].device
    o32 = o16 = ob = None
    if out == "nchw":
        o32 = torch.empty((n, c, h, w), dtype=F32, device=dev)
    else:
        assert out == "nhwc"
        twin = ops.want_twin() if twin is None else twin
        o16 = ops.nhwc_empty(n, 8, h, w, dev, ops.F16)
        ob = ops.nhwc_empty(n, 8, h, w, dev, ops.BF16) if twin else None
    abi.check(abi.lib().mcd_input_transform(src, stride, first, count, raw, k, mean_a, std_a, 255, 1, _p(o32), _p(o16),
                                            _p(ob), 8, n, h, w, _dev(allp[0]), _stream(allp[0])), "input_transform")
    if o32 is not None:
        return o32
    if ob is None:
        return o16
    ob._mcd_h16 = o16
    return ob


def relabel(lbl_u8, n_class, background_id=255):
    """get_lbl_transform(...)(label image) over a batch: int64 labels with background_id -> n_class - 1."""
    assert lbl_u8.is_cuda and lbl_u8.dtype == U8 and lbl_u8.is_contiguous()
    out = torch.empty(lbl_u8.shape, dtype=I64, device=lbl_u8.device)
    abi.check(abi.lib().mcd_relabel_u8(_p(lbl_u8), _p(out), int(background_id), int(n_class) - 1, lbl_u8.numel(),
                                       _dev(lbl_u8), _stream(lbl_u8)), "relabel_u8")
    return out


def pil_nearest_table(n_in, n_out):
    """source index of every output coordinate of PIL's Image.resize(NEAREST): the affine scale path accumulates
    x_in = (0.5 * scale) + x_out * scale in double precision step by step and truncates."""
    scale = float(n_in) / float(n_out)
    tab, xin = [], 0.5 * scale
    for _ in range(n_out):
        tab.append(min(max(int(xin), 0), n_in - 1))
        xin += scale
    return tab


_tables = {}


def resize_nearest(lbl_u8, size):
    """uint8 label maps [N,H,W] -> [N,size[1],size[0]]  (PIL sizes are (width, height): adapt_tester.py:125)."""
    assert lbl_u8.is_cuda and lbl_u8.dtype == U8 and lbl_u8.is_contiguous() and lbl_u8.dim() == 3
    n, h, w = lbl_u8.shape
    ow, oh = int(size[0]), int(size[1])
    key = (h, w, oh, ow, lbl_u8.device)
    if key not in _tables:
        _tables[key] = (torch.tensor(pil_nearest_table(h, oh), dtype=torch.int32, device=lbl_u8.device),
                        torch.tensor(pil_nearest_table(w, ow), dtype=torch.int32, device=lbl_u8.device))
    yt, xt = _tables[key]
    out = torch.empty((n, oh, ow), dtype=U8, device=lbl_u8.device)
    abi.check(abi.lib().mcd_resize_nearest_u8(_p(lbl_u8), _p(out), _p(yt), _p(xt), n, h, w, oh, ow, _dev(lbl_u8),
                                              _stream(lbl_u8)), "resize_nearest_u8")
    return out


def fast_hist(a, b, n, hist=None):
    """eval.py:21-23 with a = ground truth, b = prediction (uint8 or int64 GPU tensors of equal size): returns / adds
    into the int64 [n, n] count matrix `hist` (device tensor of n*n + 1 counters: the last one counts predictions
    outside [0, n), which make the reference's np.bincount(...).reshape(n, n) fail)."""
    assert a.is_cuda and b.is_cuda and a.numel() == b.numel()
    a, b = a.contiguous(), b.contiguous()
    for t in (a, b):
        assert t.dtype in (U8, I64), "labels are uint8 (decoded PNG) or int64 (argmax)"
    if hist is None:
        hist = torch.zeros(n * n + 1, dtype=I64, device=a.device)
    assert hist.dtype == I64 and hist.numel() == n * n + 1 and hist.is_contiguous()
    abi.check(abi.lib().mcd_fast_hist(_p(a), int(a.dtype == I64), _p(b), int(b.dtype == I64), n, a.numel(), _p(hist),
                                      _dev(a), _stream(a)), "fast_hist")
    return hist


def hist_matrix(hist, n):
    """device counters -> numpy [n, n] int64 (raises like the reference when a prediction was out of range)"""
    h = hist.cpu().numpy()
    if h[n * n] != 0:
        raise ValueError("cannot reshape array of size > %d into shape (%d,%d)" % (n * n, n, n))
    return h[:n * n].reshape(n, n)


def unnormalize(x, normalize_way="imagenet"):
    """transform.py:285-294 for a batch: fp32 [N,3,H,W] -> uint8 [N,H,W,3]."""
    if normalize_way != "imagenet":
        raise NotImplementedError()
    assert x.is_cuda and x.dtype == F32 and x.dim() == 4 and x.shape[1] == 3
    x = x.contiguous()
    n, _, h, w = x.shape
    out = torch.empty((n, h, w, 3), dtype=U8, device=x.device)
    mean = (ctypes.c_double * 3)(.485, .456, .406)
    std = (ctypes.c_double * 3)(.229, .224, .225)
    abi.check(abi.lib().mcd_unnormalize_u8(_p(x), _p(out), mean, std, n, h, w, _dev(x), _stream(x)), "unnormalize_u8")
    return out


class Batch:
    """A pre-transformed batch for MCDStep: `streams` = the generators' NHWC inputs (one for early fusion and the
    multitask encoders, two for MFNet), `aux` = fp32 NCHW regression / boundary targets of the multitask trainers
    (src_imgs[:, 3:] in adapt_triple_multitask_trainer.py:194-196)."""

    def __init__(self, streams, aux=None):
        self.streams, self.aux = tuple(streams), aux
        self.device = self.streams[0].device


class InputPipeline:
    """uint8 batches as the loader decodes them -> what MCDStep consumes, inside the captured iteration.

        mode "early":     (rgb, hha) -> one 6-channel stream                      adapt_trainer.py:156-160
        mode "mfnet":     (rgb, hha) -> RGB stream, HHA stream                    adapt_mfnet_trainer.py:186-187
        mode "multitask": (rgb, hha[, boundary]) -> RGB stream + fp32 targets     adapt_triple_multitask_trainer.py:194-196
    """

    def __init__(self, mode="early", n_class=41, normalize_way="imagenet", background_id=255):
        assert mode in ("early", "mfnet", "multitask")
        self.mode, self.n_class, self.way, self.bg = mode, n_class, normalize_way, background_id

    def images(self, planes, boundary=None):
        if self.mode == "early":
            return Batch([transform_images(planes, self.way)])
        if self.mode == "mfnet":
            return Batch([transform_images([p], self.way) for p in planes])
        rgb = transform_images(planes[:1], self.way)
        aux = transform_images(planes[1:], self.way, label_planes=() if boundary is None else (boundary,), out="nchw")
        return Batch([rgb], aux)

    def labels(self, lbl_u8):
        return relabel(lbl_u8, self.n_class, self.bg)
