"""Classifier head fused with its per-pixel loss (csrc/headloss.cu): upsampling x8, softmax, loss and all gradients in
ONE kernel - the full-resolution logits (41 x 480 x 640 per image and head) and their gradients are never written.

    head_ce2d(inputs, filters, target, criterion)            CrossEntropyLoss2d(head(inputs), target)    loss.py:7-13
    head_diff2d(inputs_a, filters_a, inputs_b, filters_b)    Diff2d(head_a(.), head_b(.))                loss.py:93-100
    head_ce2d_pair(inputs, filters_a, filters_b, target, ...) CE(head_a(.), t) + CE(head_b(.), t)   adapt_trainer.py:171-175

`inputs` / `filters`: lists of one or two fp32 score maps [N,C,h,w] and depthwise ConvTranspose2d filters [C,1,16,16]
(DRNSegPixelClassifier: one pair; ScoreFusionDRNSegPixelClassifier: up1(x1) + up2(x2), two pairs); filters = None is
nn.Upsample(x8, bilinear, align_corners=False) of the multitask decoders.  Same values as the un-fused module path up to
fp32 rounding (the logits stay in fp32 registers instead of a bf16 / fp32 tensor).  The kernel computes the gradients for
an upstream gradient of 1 during the forward pass; backward() scales them.
"""
import ctypes

import torch

from . import abi, ops

F32 = torch.float32
_enabled = True


def set_enabled(flag):
    """A/B switch: False makes every caller take the un-fused head + criterion path."""
    global _enabled
    prev, _enabled = _enabled, bool(flag)
    return prev


def enabled():
    return _enabled


def fits(n_heads, n_in, n_class, want_dw, bilinear):
    """can one launch do it: the shared-memory working set (filters, staged inputs, logit gradients) fits and the filter
    gradients (kept in registers) are those of at most two (head, input) pairs"""
    nhi = n_heads * n_in
    if want_dw and not bilinear and nhi > 2:
        return False
    floats = 0 if bilinear else nhi * n_class * 260
    floats += nhi * n_class * 16 * (2 if want_dw else 1) + (n_heads * n_class * 260) // 2 + nhi * n_class * 10
    return n_class <= 44 and floats * 4 + 16 <= 227 * 1024


def _ptrs(ts):
    arr = (ctypes.c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def _f32c(t):
    if t is None:
        return None
    t = t if t.dtype == F32 else t.float()
    return t.contiguous()


def _launch(mode, n_heads, n_in, xs, ws, dxs, dws, target, cls_weight, ignore_index, wsum, inv_numel, acc):
    x0 = xs[0]
    n, c, h, w = x0.shape
    for t in xs:
        assert t.is_cuda and t.dtype == F32 and t.is_contiguous() and tuple(t.shape) == (n, c, h, w)
    for t in ws:
        assert t is None or (t.dtype == F32 and t.is_contiguous() and t.numel() == c * 256)
    dev = x0.device.index if x0.device.index is not None else torch.cuda.current_device()
    abi.check(abi.lib().mcd_head_loss(
        mode, n_heads, n_in, _ptrs(xs), _ptrs(ws), _ptrs(dxs), _ptrs(dws),
        ctypes.c_void_p(target.data_ptr()) if target is not None else None,
        ctypes.c_void_p(cls_weight.data_ptr()) if cls_weight is not None else None, int(ignore_index),
        ctypes.c_void_p(wsum.data_ptr()) if wsum is not None else None, float(inv_numel),
        ctypes.c_void_p(acc.data_ptr()), n, c, h, w, dev,
        ctypes.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream)), "head_loss")


def _dp():
    import loss as _loss
    return _loss._dp_group, _loss.dp_world()


class _HeadCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, target, cls_weight, ignore_index, size_average, n_in, *xw):
        xs, ws = list(xw[:n_in]), list(xw[n_in:])
        need = ctx.needs_input_grad[5:]
        dxs = [torch.zeros_like(x) if need[i] else None for i, x in enumerate(xs)]
        dws = [torch.zeros_like(w) if (w is not None and need[n_in + i]) else None for i, w in enumerate(ws)]
        wsum = None
        if size_average:
            acc2 = ops.zeros_f32(2, target.device)
            dev = target.device.index if target.device.index is not None else torch.cuda.current_device()
            abi.check(abi.lib().mcd_label_weight_sum(
                ctypes.c_void_p(target.data_ptr()), ctypes.c_void_p(cls_weight.data_ptr()) if cls_weight is not None else None,
                int(ignore_index), xs[0].shape[1], ctypes.c_void_p(acc2.data_ptr()), target.numel(), dev,
                ctypes.c_void_p(torch.cuda.current_stream(target.device).cuda_stream)), "label_weight_sum")
            wsum = acc2[0:1]
            group, world = _dp()
            if world > 1:       # nn.DataParallel computes the criterion on the gathered outputs: GLOBAL normaliser
                torch.distributed.all_reduce(wsum, op=torch.distributed.ReduceOp.SUM, group=group)
        acc = ops.zeros_f32(4, xs[0].device)
        _launch(0, 1, n_in, xs, ws, dxs, dws, target, cls_weight, ignore_index, wsum, 0.0, acc)
        ctx.n_in = n_in
        ctx.save_for_backward(*[t for t in dxs + dws if t is not None])
        ctx.present = [t is not None for t in dxs + dws]
        return acc[0] / wsum[0] if size_average else acc[0].clone()

    @staticmethod
    def backward(ctx, go):
        saved = list(ctx.saved_tensors)
        grads = [saved.pop(0) * go if p else None for p in ctx.present]
        return (None, None, None, None, None) + tuple(grads)


class _HeadCEPairFn(torch.autograd.Function):
    """criterion(F1(feat), lbls) + criterion(F2(feat), lbls) in one launch (mode 2): both classifiers read the same score
    maps and the same labels; the score-map gradient is their sum, accumulated into one buffer."""

    @staticmethod
    def forward(ctx, target, cls_weight, ignore_index, size_average, n_in, *xw):
        xs, wa, wb = list(xw[:n_in]), list(xw[n_in:2 * n_in]), list(xw[2 * n_in:])
        need = ctx.needs_input_grad[5:]
        dxs = [torch.zeros_like(x) if need[i] else None for i, x in enumerate(xs)]
        dwa = [torch.zeros_like(w) if need[n_in + i] else None for i, w in enumerate(wa)]
        dwb = [torch.zeros_like(w) if need[2 * n_in + i] else None for i, w in enumerate(wb)]
        wsum = None
        if size_average:
            acc2 = ops.zeros_f32(2, target.device)
            dev = target.device.index if target.device.index is not None else torch.cuda.current_device()
            abi.check(abi.lib().mcd_label_weight_sum(
                ctypes.c_void_p(target.data_ptr()), ctypes.c_void_p(cls_weight.data_ptr()) if cls_weight is not None else None,
                int(ignore_index), xs[0].shape[1], ctypes.c_void_p(acc2.data_ptr()), target.numel(), dev,
                ctypes.c_void_p(torch.cuda.current_stream(target.device).cuda_stream)), "label_weight_sum")
            wsum = acc2[0:1]
            group, world = _dp()
            if world > 1:
                torch.distributed.all_reduce(wsum, op=torch.distributed.ReduceOp.SUM, group=group)
        acc = ops.zeros_f32(4, xs[0].device)
        _launch(2, 2, n_in, xs + xs, wa + wb, dxs + dxs, dwa + dwb, target, cls_weight, ignore_index, wsum, 0.0, acc)
        out = dxs + dwa + dwb
        ctx.present = [t is not None for t in out]
        ctx.save_for_backward(*[t for t in out if t is not None])
        return acc[0] / wsum[0] if size_average else acc[0].clone()

    @staticmethod
    def backward(ctx, go):
        saved = list(ctx.saved_tensors)
        grads = [saved.pop(0) * go if p else None for p in ctx.present]
        return (None, None, None, None, None) + tuple(grads)


class _HeadDiffFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, n_in, *xw):
        xa, wa = list(xw[:n_in]), list(xw[n_in:2 * n_in])
        xb, wb = list(xw[2 * n_in:3 * n_in]), list(xw[3 * n_in:])
        need = ctx.needs_input_grad[1:]
        n_xa, n_wa, n_xb, n_wb = need[:n_in], need[n_in:2 * n_in], need[2 * n_in:3 * n_in], need[3 * n_in:]
        dxa, dxb, dwa, dwb = [], [], [], []
        for k in range(n_in):
            shared = xb[k].data_ptr() == xa[k].data_ptr()       # both classifiers read the same score map
            da = torch.zeros_like(xa[k]) if (n_xa[k] or (shared and n_xb[k])) else None
            dxa.append(da)
            dxb.append(da if shared else (torch.zeros_like(xb[k]) if n_xb[k] else None))
            dwa.append(torch.zeros_like(wa[k]) if (wa[k] is not None and n_wa[k]) else None)
            dwb.append(torch.zeros_like(wb[k]) if (wb[k] is not None and n_wb[k]) else None)
        group, world = _dp()
        numel = float(xa[0].shape[0] * xa[0].shape[1] * xa[0].shape[2] * xa[0].shape[3] * 64) * world
        acc = ops.zeros_f32(4, xa[0].device)
        _launch(1, 2, n_in, xa + xb, wa + wb, dxa + dxb, dwa + dwb, None, None, -100, None, 1.0 / numel, acc)
        out = []                                   # gradients in input order; a shared score map reports its sum once
        for k in range(n_in):
            out.append(dxa[k])
        out += dwa
        for k in range(n_in):
            out.append(None if (dxb[k] is dxa[k]) else dxb[k])
        out += dwb
        ctx.present = [t is not None for t in out]
        ctx.save_for_backward(*[t for t in out if t is not None])
        return acc[0] / numel

    @staticmethod
    def backward(ctx, go):
        saved = list(ctx.saved_tensors)
        grads = [saved.pop(0) * go if p else None for p in ctx.present]
        return (None,) + tuple(grads)


def _check_target(inputs, target):
    n, _, h, w = inputs[0].shape
    if tuple(target.shape) != (n, 8 * h, 8 * w):
        raise ValueError("CrossEntropyLoss2d: input and target batch or spatial sizes don't match: target %s, input %s"
                         % (list(target.shape), [n, inputs[0].shape[1], 8 * h, 8 * w]))


def head_ce2d(inputs, filters, target, weight=None, ignore_index=-100, size_average=True):
    """CrossEntropyLoss2d(weight, size_average, ignore_index)(head(inputs), target) without the logits tensor."""
    _check_target(inputs, target)
    n_in = len(inputs)
    filters = [None] * n_in if filters is None else list(filters)
    xs = [_f32c(x) for x in inputs]
    ws = [_f32c(w) for w in filters]
    if target.dtype != torch.int64:
        target = target.long()
    w = None if weight is None else weight.to(device=xs[0].device, dtype=F32).contiguous()
    return _HeadCEFn.apply(target.contiguous(), w, ignore_index, size_average, n_in, *xs, *ws)


def head_ce2d_pair(inputs, filters_a, filters_b, target, weight=None, ignore_index=-100, size_average=True):
    """CrossEntropyLoss2d(head_a(inputs), target) + CrossEntropyLoss2d(head_b(inputs), target): the supervised term of
    every MCD phase (adapt_trainer.py:171-175,191-194), one launch for both classifiers."""
    _check_target(inputs, target)
    n_in = len(inputs)
    xs = [_f32c(x) for x in inputs]
    if target.dtype != torch.int64:
        target = target.long()
    w = None if weight is None else weight.to(device=xs[0].device, dtype=F32).contiguous()
    return _HeadCEPairFn.apply(target.contiguous(), w, ignore_index, size_average, n_in, *xs,
                               *[_f32c(t) for t in filters_a], *[_f32c(t) for t in filters_b])


def head_diff2d(inputs_a, filters_a, inputs_b, filters_b):
    """Diff2d()(head_a(inputs_a), head_b(inputs_b)) = mean |softmax - softmax| without the logits tensors."""
    n_in = len(inputs_a)
    fa = [None] * n_in if filters_a is None else list(filters_a)
    fb = [None] * n_in if filters_b is None else list(filters_b)
    xa = [_f32c(x) for x in inputs_a]
    xb = [xa[k] if inputs_b[k] is inputs_a[k] else _f32c(inputs_b[k]) for k in range(n_in)]
    return _HeadDiffFn.apply(n_in, *xa, *[_f32c(w) for w in fa], *xb, *[_f32c(w) for w in fb])


def classifier_inputs(head, feats):
    """(inputs, filters) of one of the reference's pixel classifiers applied to `feats`, or None when the module is not
    one of the plain learned-deconvolution heads (models/dilated_fcn.py:340-366,431-491)."""
    from models.dilated_fcn import (DRNSegPixelClassifier, FusionDRNSegPixelClassifier,
                                    ScoreFusionDRNSegPixelClassifier)
    from models.fusion import AddFusion
    from .nn import DepthwiseDeconv16s8

    def score(h):      # ver2: the 1x1 `seg` convolution sits in the head (models/dilated_fcn.py:362-365,467-468)
        return head.seg(h) if getattr(head, "ver", "ver1") == "ver2" else h
    if type(head) is DRNSegPixelClassifier and len(feats) == 1 and type(head.up) is DepthwiseDeconv16s8:
        return [score(feats[0])], [head.up.weight]
    if type(head) is FusionDRNSegPixelClassifier and len(feats) == 2 and type(head.up) is DepthwiseDeconv16s8:
        # up([seg](fusion(x1, x2))): the fusion (Add: a 60x80 torch add; Gate / ConcatConv: library kernels) runs as a
        # module, the upsampling + criterion as the fused kernel
        return [score(head.fusion(feats[0], feats[1]))], [head.up.weight]
    if type(head) is ScoreFusionDRNSegPixelClassifier and isinstance(head.fusion, AddFusion) and len(feats) == 2:
        return [feats[0], feats[1]], [head.up1.weight, head.up2.weight]
    return None
