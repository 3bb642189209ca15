"""autograd Functions and nn.Module building blocks over libmcd_sm100.

The modules subclass the stock torch containers (`nn.Conv2d`, `nn.BatchNorm2d`, `nn.ConvTranspose2d`)
only so that parameters, buffers, initialisation and state_dict keys are byte-compatible with the
reference's checkpoints (adapt_trainer.py:232-245, adapt_tester.py:79-83); their `forward` never calls
a stock torch kernel - it launches the library's kernels through `ops`.

Numerics (mcd_b200/ops.py): forward activations are IEEE half, gradients bfloat16.  The unit of execution is
conv -> BatchNorm (+ residual) -> ReLU as ONE autograd Function (`_UnitFn`): the IEEE-half pre-BatchNorm tensor never
crosses an autograd edge, what does is the unit's output as a bfloat16 twin (autograd requires gradients in the dtype
of the tensor they belong to, and gradients are bfloat16) that carries the IEEE-half original as `_mcd_h16`.
"""
import os

import torch
import torch.nn as nn

from . import ops

BF16, F16, F32 = torch.bfloat16, torch.float16, torch.float32

# number of identical forward passes one BatchNorm forward stands for (MCDStep re-uses the phase-B target
# forward for the first phase-C step): the running statistics then take that many momentum updates at once.
_bn_repeat = 1


class bn_update_repeat:
    def __init__(self, k):
        self.k = int(k)

    def __enter__(self):
        global _bn_repeat
        self.prev, _bn_repeat = _bn_repeat, self.k

    def __exit__(self, *a):
        global _bn_repeat
        _bn_repeat = self.prev


# dtype of the full-resolution predictions the heads / decoders return: fp32 is the drop-in default (the reference
# testers call `.data.cpu().numpy()` on them, adapt_tester.py:114-118); MCDStep switches to bfloat16, which halves
# the traffic of the head and loss kernels - the largest tensors of the step.
_logits_dtype = F32


class logits_dtype:
    """`with logits_dtype(torch.bfloat16):` or `logits_dtype.set(...)`."""

    def __init__(self, dtype):
        assert dtype in (BF16, F32)
        self.dtype = dtype

    def __enter__(self):
        global _logits_dtype
        self.prev, _logits_dtype = _logits_dtype, self.dtype

    def __exit__(self, *a):
        global _logits_dtype
        _logits_dtype = self.prev

    @staticmethod
    def set(dtype):
        global _logits_dtype
        assert dtype in (BF16, F32)
        prev, _logits_dtype = _logits_dtype, dtype
        return prev

    @staticmethod
    def get():
        return _logits_dtype


_overlap_wgrad = True
_side_streams = {}


def _side_stream(device):
    s = _side_streams.get(device)
    if s is None:
        s = _side_streams[device] = torch.cuda.Stream(device)
    return s


class DirectGrads:
    """Context of MCDStep's backward passes: convolution weight gradients are written by the wgrad kernels
    directly into param.grad on the side stream (see _conv_backward) instead of travelling through autograd.
    `join()` makes the main stream wait for them and releases the operands that were kept alive."""

    def __init__(self, defer=False):
        self.defer = defer   # weight gradients may stay in split form for FusedSGD (MCDStep with fused_sgd)
        self.keep = []
        self.stash = {}      # data_ptr of a block input -> identity-shortcut gradient awaiting conv1's dgrad
        self.bnsums = {}     # data_ptr of a dgrad output -> (raw BatchNorm-backward sums, data_ptr of the BN input)

    def __enter__(self):
        global _direct
        self.prev, _direct = _direct, self
        return self

    def __exit__(self, *a):
        global _direct
        _direct = self.prev

    def join(self, device):
        side = _side_streams.get(device)
        if side is not None:
            torch.cuda.current_stream(device).wait_stream(side)
        self.keep.clear()
        self.bnsums.clear()
        if self.stash:
            self.stash.clear()
            raise RuntimeError("mcd_b200: an identity-shortcut gradient was stashed but never consumed")


_direct = None
# Opt-in (A/B switch, OFF by default): fuse the ReLU mask and the BatchNorm-backward sums {sum g, sum g * y} of the
# producing unit into the dgrad epilogue of its sole-consumer convolution (direct-gradient mode only), which removes that
# unit's reduction pass.  With the round-2 kernels the epilogue costs more than the pass it replaces: +47...+95 us per
# launch (scripts/bench_thin_dgrad.py) against a memory-bound reduction that overlaps the tensor-bound wgrad stream;
# whole iteration at 22 pairs: 137.4 ms fused everywhere, 135.4 ms for >= 128 / 256 channels only, 133.7 ms for >= 512
# only, 133.5 ms never (profiles/r02e_ab_fuse_bn_bwd.txt).  MCD_FUSE_BN_BWD=1 [MCD_FUSE_BN_BWD_MIN_C=c] switches it on
# (for convolutions whose input has >= c channels).
_fuse_bn_bwd = os.environ.get("MCD_FUSE_BN_BWD", "0") == "1"
_fuse_bn_bwd_min_c = int(os.environ.get("MCD_FUSE_BN_BWD_MIN_C", "0"))


def set_fuse_bn_bwd(flag, min_channels=0):
    """see above; off = separate reduction pass (the default)."""
    global _fuse_bn_bwd, _fuse_bn_bwd_min_c
    _fuse_bn_bwd, _fuse_bn_bwd_min_c = bool(flag), int(min_channels)


def set_overlap_wgrad(flag):
    global _overlap_wgrad
    prev, _overlap_wgrad = _overlap_wgrad, bool(flag)
    return prev


def _as_nhwc_grad(dy):
    """Gradients arriving from autograd for an nhwc activation: make them nhwc bfloat16 again."""
    if ops.is_nhwc(dy) and dy.dtype == BF16:
        return dy
    if dy.dtype in (BF16, F16) and dy.shape[1] % 8 == 0:
        return ops.to_nhwc(dy.contiguous(memory_format=torch.channels_last), grad=True)
    return ops.to_nhwc(dy, grad=True)


# ---------------------------------------------------------------------------------------------
def _weight_tag(w):
    return (w._version, w.data_ptr(), getattr(w, "_mcd_step", 0))


def _conv_backward(mod, g, x, dy, need_dx, need_dw, want_db, bn_y, w_tag=None):
    """dgrad + wgrad of one convolution.  x: the input handle (its bf16 twin feeds wgrad and the ReLU mask of the fused
    dgrad epilogue), dy: bf16 nhwc gradient of the output, bn_y: see _UnitFn.  Returns (dx, dw, db); in
    direct-gradient mode dw / db are written in place and None is returned for them."""
    dx = dw = db = None
    only_db = want_db and not need_dw         # frozen weight, trainable bias: dbias comes with the wgrad call
    need_dw = need_dw or want_db
    if w_tag is not None and need_dx and w_tag != _weight_tag(mod.weight):
        # dgrad re-reads the packed weights at backward time: they must still be the ones the forward used
        raise RuntimeError("mcd_b200: the weights of a convolution were modified between its forward and its "
                           "backward (optimizer step inside a forward/backward pair?)")
    if _direct is not None and need_dw:
        # direct-gradient mode (MCDStep): wgrad is enqueued on the side stream and writes straight into
        # param.grad (or its all-reduce bucket view); nothing is returned to autograd for the weights, so no
        # accumulation kernel runs and the main stream goes on with dgrad / BatchNorm backward while the
        # tensor-bound wgrad kernels trail behind.  MCDStep joins the side stream before optimizer.step().
        main = torch.cuda.current_stream(dy.device)
        # set_overlap_wgrad(False): everything on the main stream (per-kernel timing in bench.py)
        side = _side_stream(dy.device) if _overlap_wgrad else main
        ready = torch.cuda.Event()
        ready.record(main)                    # dy (and everything before it) is complete here
        w, b = mod.weight, mod.bias
        acc = getattr(w, "_mcd_written", False) and w.grad is not None
        defer = getattr(mod, "_mcd_defer", None) if (_direct.defer and not only_db) else None
        if defer is not None and defer[0] != g.key():
            defer = None
        if defer is not None and getattr(w, "_mcd_deferred", False):
            raise RuntimeError("mcd_b200: a deferred weight gradient was produced twice before optimizer_g.step()")
        if defer is None and w.grad is None and not only_db:
            w.grad = torch.empty_like(w)
        if want_db and b.grad is None:
            b.grad = torch.empty_like(b)
    # dgrad is enqueued FIRST: it heads the critical path (dgrad -> BatchNorm backward -> next dgrad) and
    # cannot share an SM with the persistent wgrad CTAs (both want ~190 KB of shared memory); the wgrad
    # that follows on the side stream then overlaps the memory-bound BatchNorm kernels of the next unit.
    add = _direct.stash.pop(x.data_ptr(), None) if _direct is not None else None   # identity-shortcut gradient
    if need_dx and bn_y is not None and _fuse_bn_bwd and _direct is not None and x.shape[1] >= _fuse_bn_bwd_min_c:
        dx, sums = ops.conv_dgrad(dy, mod.packed(1, g), g, add=add, relu_src=x, bn_y=bn_y)
        _direct.bnsums[dx.data_ptr()] = (sums, bn_y.data_ptr(), dx)
    elif need_dx:
        dx = ops.conv_dgrad(dy, mod.packed(1, g), g, add=add)
    elif add is not None:
        dx = add
    if not need_dw:
        return dx, None, None
    if _direct is not None:
        side.wait_event(ready)
        with torch.cuda.stream(side):
            if only_db:
                ops.conv_wgrad(x, dy, g, want_dbias=True, out_db=b.grad, accumulate=False)
            elif defer is not None:
                # split partial sums stay in the convolution's persistent workspace; FusedSGD reduces them
                ops.conv_wgrad(x, dy, g, want_dbias=want_db, out_db=b.grad if want_db else None, accumulate=acc,
                               partials=defer[1])
                w._mcd_deferred = True
            else:
                ops.conv_wgrad(x, dy, g, want_dbias=want_db, out_dw=w.grad, out_db=b.grad if want_db else None,
                               accumulate=acc)
        if not only_db:
            w._mcd_written = True
        _direct.keep.append((x, dy))          # keep the operands alive until the side stream is joined
        for p in ((w, b) if want_db else (w,)):
            sync = getattr(p, "_mcd_sync", None)
            if sync is not None and not (p is w and only_db):
                sync.mark_ready(p, side)
        return dx, None, None
    if need_dx and _overlap_wgrad:
        # dgrad and wgrad both consume dy and are independent: wgrad runs on a side stream so that its CTAs
        # fill the SMs the other kernel's last (partial) wave of tiles leaves idle.  Every use of the side
        # stream is preceded by side.wait_stream(main), which also makes the caching allocator's per-stream
        # reuse of the workspace / gradient blocks safe.
        main = torch.cuda.current_stream(dy.device)
        side = _side_stream(dy.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            dw, db = ops.conv_wgrad(x, dy, g, want_dbias=want_db)
        main.wait_stream(side)
        return dx, (None if only_db else dw), db
    dw, db = ops.conv_wgrad(x, dy, g, want_dbias=want_db)
    return dx, (None if only_db else dw), db


class _ConvFn(torch.autograd.Function):
    """a convolution on its own: the planar fp32 score-map heads (`seg`, decoder outputs) and generic use."""

    @staticmethod
    def forward(ctx, x, weight, bias, mod, planar):
        g = mod.geom(x.shape)
        y, _ = ops.conv_fprop(x, mod.packed(0, g), bias, g, planar=planar)
        ctx.mod, ctx.g, ctx.planar = mod, g, planar
        ctx.has_bias = bias is not None
        ctx.w_tag = _weight_tag(weight)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = ops.to_nhwc(dy, grad=True) if ctx.planar else _as_nhwc_grad(dy)
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        dx, dw, db = _conv_backward(ctx.mod, ctx.g, x, dy, ctx.needs_input_grad[0], ctx.needs_input_grad[1], want_db,
                                    None, ctx.w_tag)
        return dx, dw, db, None, None


class Conv2d(nn.Conv2d):
    """nn.Conv2d parameter container whose forward is the library's implicit-GEMM convolution.

    Input: an nhwc activation (IEEE half or its bf16 twin) or a planar fp32 tensor; output: IEEE-half nhwc, or with
    `planar_out=True` an fp32 NCHW score map.
    """

    def __init__(self, *args, planar_out=False, **kwargs):
        super().__init__(*args, **kwargs)
        assert self.groups == 1 and self.padding_mode == "zeros"
        assert self.kernel_size[0] == self.kernel_size[1] and self.stride[0] == self.stride[1]
        assert self.dilation[0] == self.dilation[1] and self.padding[0] == self.padding[1]
        self.planar_out = planar_out
        self._packs = {}
        self._geoms = {}
        self._folds = {}

    def geom(self, x_shape):
        key = tuple(x_shape)
        g = self._geoms.get(key)
        if g is None:
            g = ops.conv_geom(x_shape, self.in_channels, self.out_channels, self.kernel_size[0],
                              self.kernel_size[1], self.stride[0], self.dilation[0], self.padding[0])
            self._geoms[key] = g
        return g

    def packed(self, mode, g):
        """16-bit packed shadow of the fp32 master weight (layout chosen by the library for geometry `g`; IEEE half
        for the forward operand, bfloat16 for the dgrad operand), refreshed when the parameter changes (optimizer
        step, load_state_dict, .to())."""
        w = self.weight
        tag = (w._version, w.data_ptr())
        key = ops.pack_key(g, mode)
        hit = self._packs.get(key)
        if hit is None or hit[0] != tag:
            hit = (tag, ops.pack_weight_for(w, g, mode))
            self._packs[key] = hit
        return hit[1]

    def folded(self, bn, g):
        """eval-mode BatchNorm folded into this convolution: (packed IEEE-half weights scale[o] * W[o], fp32 bias
        beta - mean * scale (+ scale * conv bias)), scale = gamma / sqrt(running_var + eps); rebuilt when any of the
        tensors involved changes (optimizer step, load_state_dict, a training-mode forward of `bn`)."""
        ts = (self.weight, self.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
        tag = tuple((t._version, t.data_ptr(), getattr(t, "_mcd_step", 0)) for t in ts if t is not None) + \
            (getattr(bn, "_mcd_stat_step", 0), float(bn.eps))
        key = ("fold", id(bn)) + tuple(ops.pack_key(g, 0))
        hit = self._folds.get(key)
        if hit is None or hit[0] != tag:
            with torch.no_grad():
                scale = bn.weight.float() * torch.rsqrt(bn.running_var.float() + bn.eps)
                shift = bn.bias.float() - bn.running_mean.float() * scale
                if self.bias is not None:
                    shift = shift + self.bias.float() * scale
                w = (self.weight.float() * scale[:, None, None, None]).contiguous()
                hit = (tag, ops.pack_weight_for(w, g, 0), shift.contiguous())
            self._folds[key] = hit
        return hit[1], hit[2]

    def prepare(self, x):
        x = to_nhwc_act(x)
        if x.shape[1] < self.in_channels:
            raise ValueError("Conv2d expected >= %d input channels, got %d" % (self.in_channels, x.shape[1]))
        return x

    def forward(self, x):
        return _ConvFn.apply(self.prepare(x), self.weight, self.bias, self, self.planar_out)


# ---------------------------------------------------------------------------------------------
def _bn_backward(ctx_relu, training, res_training, has_res_bn, want_dres, res_ptr, dz, z, y, gamma, aff, res, res_gamma,
                 res_aff):
    """BatchNorm (+ residual) + ReLU backward of one unit; returns dy, dgamma, dbeta, dres, dres_gamma, dres_beta."""
    fused = _direct.bnsums.pop(dz.data_ptr(), None) if _direct is not None else None
    if fused is not None and (fused[1] != y.data_ptr() or has_res_bn or not ctx_relu or not training):
        raise RuntimeError("mcd_b200: fused BatchNorm-backward sums reached the wrong unit")
    dy, dgamma, dbeta, dres, dres_gamma, dres_beta = ops.bn_bwd(
        dz, z, y, gamma, aff, training, ctx_relu, res=res, res_gamma=res_gamma, res_aff=res_aff,
        res_training=res_training, want_dres=want_dres or has_res_bn, raw_sums=fused[0] if fused is not None else None)
    if not (want_dres or has_res_bn):
        dres = None
    elif _direct is not None and not has_res_bn and res_ptr is not None:
        # direct-gradient mode: the identity-shortcut gradient is not returned to autograd (which would add
        # it to conv1's dgrad in a separate pass) but handed to conv1's dgrad kernel, whose epilogue adds it
        _direct.stash[res_ptr] = dres
        dres = None
    return dy, dgamma, dbeta, dres, dres_gamma, dres_beta


class _BNActFn(torch.autograd.Function):
    """BatchNorm (+ residual) (+ ReLU) on its own (BatchNorm2d.forward / .fused); units use _UnitFn."""

    @staticmethod
    def forward(ctx, y, stats, gamma, beta, res, res_stats, res_gamma, res_beta, bn, res_bn, relu, twin):
        training = bn.training
        z, aff, res_aff = ops.bn_forward(y, stats, bn, relu, res=res, res_stats=res_stats, res_bn=res_bn,
                                         repeat=_bn_repeat, twin=twin)
        ctx.relu, ctx.training = relu, training
        ctx.res_training = res_bn.training if res_bn is not None else False
        ctx.has_res, ctx.has_res_bn = res is not None, res_bn is not None
        ctx.res_ptr = res.data_ptr() if (res is not None and res_bn is None and getattr(res, "_mcd_shortcut", False)) else None
        ctx.save_for_backward(y, z, gamma, aff, res if res_bn is not None else None, res_gamma, res_aff)
        ctx.set_materialize_grads(False)       # no zero-filled "gradients" for the non-differentiable outputs
        z16 = getattr(z, "_mcd_h16", None)
        if z16 is None:
            return z, torch.empty(0, dtype=F16, device=z.device)
        ctx.mark_non_differentiable(z16)
        return z, z16

    @staticmethod
    def backward(ctx, dz, _):
        if dz is None:
            return (None,) * 12
        y, z, gamma, aff, res, res_gamma, res_aff = ctx.saved_tensors
        dz = _as_nhwc_grad(dz)
        want_dres = ctx.has_res and ctx.needs_input_grad[4]
        dy, dgamma, dbeta, dres, dres_gamma, dres_beta = _bn_backward(
            ctx.relu, ctx.training, ctx.res_training, ctx.has_res_bn, want_dres, ctx.res_ptr, dz, z, y, gamma, aff, res,
            res_gamma, res_aff)
        if not want_dres:
            dres = None
        return dy, None, dgamma, dbeta, dres, None, dres_gamma, dres_beta, None, None, None, None


def _with_twin(out):
    """(handle, IEEE-half original or empty) as returned by the Functions -> the handle with `_mcd_h16` attached."""
    z, z16 = out
    if z16.numel():
        z._mcd_h16 = z16
    return z


class BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d container (eps/momentum/affine/running stats as in models/drn.py:34,129,202).

    `fused(y, stats, ...)` = normalise (+ residual) (+ ReLU) in one pass; statistics come from the
    producing convolution's epilogue.  `.training = False` (eval() or fix_batchnorm_when_training,
    models/model_util.py:305-310) switches to the running statistics.
    """

    def fused(self, y, stats, relu=True, res=None, res_stats=None, res_bn=None):
        assert self.affine and self.track_running_stats
        if res_bn is not None:
            return _with_twin(_BNActFn.apply(y, stats, self.weight, self.bias, res, res_stats, res_bn.weight,
                                             res_bn.bias, self, res_bn, relu, ops.want_twin()))
        return _with_twin(_BNActFn.apply(y, stats, self.weight, self.bias, res, None, None, None, self, None, relu,
                                         ops.want_twin()))

    def forward(self, x):
        x = ops.to_nhwc(x)
        stats = ops.bn_stats(x, self.num_features) if self.training else None
        return self.fused(x, stats, relu=False)


# ---------------------------------------------------------------------------------------------
class _UnitFn(torch.autograd.Function):
    """z = act(bn(conv(x)) + residual), residual = res (identity) | res_bn(res_conv(res)) - the BasicBlock half /
    conv-BN-ReLU unit of models/drn.py:43-59,126-131,195-205 and CBR of models/dilated_fcn.py:632-644.

    x_bn_y: x = relu(bn(x_bn_y)) was produced by the previous unit and this convolution is x's only consumer (an
    identity shortcut of the same block aside): in direct-gradient mode the dgrad epilogue then applies the ReLU mask
    and accumulates the BatchNorm-backward sums of that unit, which removes its reduction pass."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, res, res_weight, res_gamma, res_beta, conv, bn, res_conv, res_bn,
                relu, x_bn_y, twin):
        g = conv.geom(x.shape)
        training = bn.training
        y, stats = ops.conv_fprop(x, conv.packed(0, g), bias, g, want_stats=training)
        ry = rstats = rg = None
        if res_conv is not None:
            rg = res_conv.geom(res.shape)
            ry, rstats = ops.conv_fprop(res, res_conv.packed(0, rg), None, rg, want_stats=res_bn.training)
        # twin: autograd is recording (decided by the caller - grad mode is always off inside forward)
        z, aff, res_aff = ops.bn_forward(y, stats, bn, relu, res=ry if res_conv is not None else res,
                                         res_stats=rstats, res_bn=res_bn, repeat=_bn_repeat, twin=twin)
        ctx.conv, ctx.g, ctx.res_conv, ctx.rg = conv, g, res_conv, rg
        ctx.relu, ctx.training = relu, training
        ctx.res_training = res_bn.training if res_bn is not None else False
        ctx.has_res, ctx.has_res_bn, ctx.has_bias = res is not None, res_bn is not None, bias is not None
        ctx.res_ptr = res.data_ptr() if (res is not None and res_bn is None and getattr(res, "_mcd_shortcut", False)) else None
        # downsample shortcut: x feeds conv1 (an earlier unit) and this unit's 1x1 convolution; the gradient that
        # reaches x through the 1x1 convolution is handed to conv1's dgrad epilogue instead of an autograd add
        ctx.ds_ptr = res.data_ptr() if (res_conv is not None and getattr(res, "_mcd_shortcut", False)) else None
        ctx.x_bn_y = x_bn_y
        ctx.w_tag = _weight_tag(weight)
        ctx.rw_tag = _weight_tag(res_weight) if res_conv is not None else None
        # backward needs: x (wgrad operand / mask: its bf16 twin), y (BatchNorm input), z (ReLU mask: the bf16 twin)
        ctx.save_for_backward(x, y, z, gamma, aff, ry, res if res_conv is not None else None, res_gamma, res_aff)
        # autograd would otherwise zero-fill a full-size "gradient" for each of the two non-differentiable outputs of
        # every unit (370 fill kernels, 5 % of an iteration)
        ctx.set_materialize_grads(False)
        z16 = getattr(z, "_mcd_h16", None)
        if z16 is None:
            return z, torch.empty(0, dtype=F16, device=z.device), y
        ctx.mark_non_differentiable(z16, y)
        return z, z16, y

    @staticmethod
    def backward(ctx, dz, _z16, _y):
        if dz is None:
            return (None,) * 16
        x, y, z, gamma, aff, ry, res, res_gamma, res_aff = ctx.saved_tensors
        need = ctx.needs_input_grad
        dz = _as_nhwc_grad(dz)
        want_dres = ctx.has_res and need[5]
        dy, dgamma, dbeta, dres, dres_gamma, dres_beta = _bn_backward(
            ctx.relu, ctx.training, ctx.res_training, ctx.has_res_bn, want_dres, ctx.res_ptr, dz, z, y, gamma, aff, ry,
            res_gamma, res_aff)
        drw = None
        if ctx.has_res_bn:
            # downsample branch: dres is the gradient of the 1x1 convolution's output
            dres, drw, _ = _conv_backward(ctx.res_conv, ctx.rg, res, dres, need[5], need[6], False, None, ctx.rw_tag)
            if _direct is not None and ctx.ds_ptr is not None and dres is not None:
                _direct.stash[ctx.ds_ptr] = dres
                dres = None
        elif not want_dres:
            dres = None
        dx, dw, db = _conv_backward(ctx.conv, ctx.g, x, dy, need[0], need[1] , ctx.has_bias and need[2], ctx.x_bn_y,
                                    ctx.w_tag)
        return (dx, dw, db, dgamma, dbeta, dres, drw, dres_gamma, dres_beta, None, None, None, None, None, None, None)


_fold_eval = True


def set_fold_eval(flag):
    """A/B switch of the inference path: False = convolution, then BatchNorm as a separate pass (the training layout)."""
    global _fold_eval
    prev, _fold_eval = _fold_eval, bool(flag)
    return prev


def folded_unit(conv, bn, x, relu, res=None):
    """inference (no autograd, BatchNorm in eval mode): z = act(conv'(x) + bias' + res) in one kernel, the BatchNorm
    folded into conv' / bias' (Conv2d.folded) - no pre-BatchNorm tensor, no normalisation pass."""
    g = conv.geom(x.shape)
    w, b = conv.folded(bn, g)
    return ops.conv_fprop_act(x, w, b, g, res=res, relu=relu)


def conv_bn_act(conv, bn, x, relu=True, res=None, res_conv=None, res_bn=None, sole=False):
    """z = act(bn(conv(x)) + residual); residual = res (identity) or res_bn(res_conv(res)).
    sole: `conv` is the only consumer of x (an identity shortcut of the same block aside), see _UnitFn."""
    bn_y = getattr(x, "_mcd_bn_y", None) if sole else None
    x = conv.prepare(x)
    if (_fold_eval and not torch.is_grad_enabled() and not bn.training and (res_bn is None or not res_bn.training)
            and bn.affine and bn.track_running_stats):
        r = res
        if res is not None:
            r = folded_unit(res_conv, res_bn, res_conv.prepare(res), False) if res_conv is not None else ops.to_nhwc(res)
        z = folded_unit(conv, bn, x, relu, r) if (res is None or r is not None) else None
        if z is not None:
            return z
    if bn_y is not None and (tuple(bn_y.shape) != tuple(x.shape) or x.shape[1] != conv.in_channels):
        bn_y = None
    if res is not None:
        res = res_conv.prepare(res) if res_conv is not None else ops.to_nhwc(res)
    assert bn.affine and bn.track_running_stats
    out = _UnitFn.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, res,
                        res_conv.weight if res_conv is not None else None,
                        res_bn.weight if res_bn is not None else None, res_bn.bias if res_bn is not None else None,
                        conv, bn, res_conv, res_bn, relu, bn_y, ops.want_twin())
    z = _with_twin(out[:2])
    if relu and bn.training and res_bn is None and out[2].shape[1] == bn.num_features:
        z._mcd_bn_y = out[2]     # lets a sole-consumer convolution start this unit's backward in its dgrad epilogue
    return z


class ConvBNReLU(nn.Sequential):
    """`nn.Sequential(conv, bn, relu)` of models/drn.py:126-131,195-205 executed as one fused unit."""

    def forward(self, x):
        mods = list(self.children())
        i = 0
        sole = getattr(x, "_mcd_sole", False)
        while i < len(mods):
            conv, bn = mods[i], mods[i + 1]
            x = conv_bn_act(conv, bn, x, relu=True, sole=sole)
            sole = True          # intermediate activations never leave this container
            i += 3
        return x


class SoleChain(nn.Sequential):
    """nn.Sequential whose intermediate activations are read by the next child only (the DRN trunk stages,
    reference models/dilated_fcn.py `self.base = nn.Sequential(*list(model.children())[:-2])`).  It flags them so
    that the next stage's first convolution may fuse the producing unit's BatchNorm backward (_UnitFn)."""

    def forward(self, x):
        mods = list(self.children())
        i = 0
        while i < len(mods):
            m = mods[i]
            if (isinstance(m, Conv2d) and i + 2 < len(mods) and isinstance(mods[i + 1], BatchNorm2d)
                    and isinstance(mods[i + 2], nn.ReLU)):
                # DRN arch 'C' registers its stem as three separate children (models/drn.py:113-119)
                x = conv_bn_act(m, mods[i + 1], x, relu=True, sole=getattr(x, "_mcd_sole", False))
                i += 3
            else:
                x = m(x)
                i += 1
            if i < len(mods):
                x._mcd_sole = True
        return x


# ---------------------------------------------------------------------------------------------
class _Deconv16s8Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, x2, w2, out_dtype):
        out = ops.deconv16s8_fwd(x, w, x2, w2, out_dtype)
        ctx.save_for_backward(x, w, x2, w2)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, x2, w2 = ctx.saved_tensors
        need = ctx.needs_input_grad
        dx = dw = dx2 = dw2 = None
        if need[0] or need[1]:
            dx, dw = ops.deconv16s8_bwd(dout, x, w, want_dx=need[0], want_dw=need[1])
        if x2 is not None and (need[2] or need[3]):
            dx2, dw2 = ops.deconv16s8_bwd(dout, x2, w2, want_dx=need[2], want_dw=need[3])
        return dx, dw, dx2, dw2, None


def _planar_f32(x):
    if x.dtype != F32:
        x = x.float()
    return x.contiguous()


class DepthwiseDeconv16s8(nn.ConvTranspose2d):
    """ConvTranspose2d(C, C, 16, stride=8, padding=4, groups=C, bias=False): the *learned* upsampling of
    the MCD heads (models/dilated_fcn.py:357-360).  fp32 planar score map in, planar logits out (`logits_dtype`)."""

    def __init__(self, n_class):
        super().__init__(n_class, n_class, 16, stride=8, padding=4, output_padding=0, groups=n_class,
                         bias=False)

    def forward(self, x, x2=None, other=None):
        x = _planar_f32(x)
        if x2 is None:
            return _Deconv16s8Fn.apply(x, self.weight, None, None, _logits_dtype)
        return _Deconv16s8Fn.apply(x, self.weight, _planar_f32(x2), other.weight, _logits_dtype)


class _BilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s, out_f32):
        ctx.s = s
        return ops.bilinear_up_fwd(x, s, out_f32)

    @staticmethod
    def backward(ctx, dout):
        return ops.bilinear_up_bwd(dout.contiguous(), ctx.s), None, None


class BilinearUpsample(nn.Module):
    """nn.Upsample(scale_factor=s, mode='bilinear') (align_corners=False), models/dilated_fcn.py:676,817-819.
    Output dtype: `logits_dtype` (out_f32=True forces fp32)."""

    def __init__(self, scale_factor, out_f32=False):
        super().__init__()
        self.scale_factor, self.out_f32 = int(scale_factor), out_f32

    def forward(self, x):
        return _BilinearFn.apply(_planar_f32(x), self.scale_factor, self.out_f32 or _logits_dtype == F32)


# ---------------------------------------------------------------------------------------------
# Option surface around the hot path (csrc/variants.cu): the pieces the Gate / Concat fusion heads, FuseDRNSegBase,
# the `ver2` / `use_torch_up` heads and the shortcut / seg2bd decoder options are composed of.
class _AddNHWCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, twin):
        z = ops.add_nhwc(a, b, twin=twin)
        ctx.set_materialize_grads(False)
        z16 = getattr(z, "_mcd_h16", None)
        if z16 is None:
            return z, torch.empty(0, dtype=F16, device=z.device)
        ctx.mark_non_differentiable(z16)
        return z, z16

    @staticmethod
    def backward(ctx, dz, _):
        if dz is None:
            return None, None, None
        dz = _as_nhwc_grad(dz)
        return dz, dz, None


def add_nhwc(a, b):
    """torch.add of two nhwc activations (FuseDRNSegBase, models/dilated_fcn.py:308-329; AddFusion on trunk features)."""
    return _with_twin(_AddNHWCFn.apply(ops.to_nhwc(a), ops.to_nhwc(b), ops.want_twin()))


class _ToPlanarFn(torch.autograd.Function):
    """nhwc activation -> planar fp32 NCHW (first c channels); backward: the gradient as a bf16 nhwc tensor."""

    @staticmethod
    def forward(ctx, x, c):
        ctx.cs = x.shape[1]
        return ops.to_nchw_f32(x, c)

    @staticmethod
    def backward(ctx, d):
        if d is None:
            return None, None
        d = ops._pf32(d)
        if d.shape[1] != ctx.cs:
            raise RuntimeError("mcd_b200: to_planar with a channel subset is forward-only")
        return ops.to_nhwc(d, grad=True), None


class _ToNHWCFn(torch.autograd.Function):
    """planar fp32 NCHW (produced by an autograd op) -> nhwc activation; backward: planar fp32 gradient."""

    @staticmethod
    def forward(ctx, x, twin):
        ctx.c = x.shape[1]
        z = ops.to_nhwc(x, twin=twin)
        ctx.set_materialize_grads(False)
        z16 = getattr(z, "_mcd_h16", None)
        if z16 is None:
            return z, torch.empty(0, dtype=F16, device=z.device)
        ctx.mark_non_differentiable(z16)
        return z, z16

    @staticmethod
    def backward(ctx, dz, _):
        if dz is None:
            return None, None
        return ops.to_nchw_f32(_as_nhwc_grad(dz), ctx.c), None


def to_planar(x):
    """planar fp32 view of an activation for the fp32 element-wise fusion kernels (autograd-aware)."""
    if ops.is_nhwc(x):
        return _ToPlanarFn.apply(x, None)
    return _planar_f32(x)


def to_nhwc_act(x):
    """nhwc activation from whatever a module hands to a convolution (autograd-aware for planar fp32 inputs)."""
    if ops.is_nhwc(x):
        return x
    if torch.is_grad_enabled() and x.requires_grad:
        return _with_twin(_ToNHWCFn.apply(_planar_f32(x), True))
    return ops.to_nhwc(x)


class _GateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2, a):
        ctx.save_for_backward(x1, x2, a)
        return ops.gate_fuse_fwd(x1, x2, a)

    @staticmethod
    def backward(ctx, dout):
        x1, x2, a = ctx.saved_tensors
        return tuple(ops.gate_fuse_bwd(x1, x2, a, dout, want=ctx.needs_input_grad[:3]))


def gate_fuse(x1, x2, gate_logits):
    """x1 * sigmoid(a) + x2 * (1 - sigmoid(a)), planar fp32 (GateFusion, models/fusion.py:17-21)."""
    return _GateFn.apply(_planar_f32(x1), _planar_f32(x2), _planar_f32(gate_logits))


class _SoftmaxChFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        p = ops.softmax_ch_fwd(x)
        ctx.save_for_backward(p)
        return p

    @staticmethod
    def backward(ctx, dp):
        (p,) = ctx.saved_tensors
        return ops.softmax_ch_bwd(p, dp)


def softmax_ch(x):
    return _SoftmaxChFn.apply(_planar_f32(x))


class _Cat2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.ca = a.shape[1]
        return ops.cat2_f32(a, b)

    @staticmethod
    def backward(ctx, d):
        return ops.split2_f32(d, ctx.ca, want=ctx.needs_input_grad[:2])


def cat2(a, b):
    """torch.cat([a, b], 1) on planar fp32 tensors."""
    return _Cat2Fn.apply(_planar_f32(a), _planar_f32(b))


class _SigmoidFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = ops.sigmoid_fwd(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return ops.sigmoid_bwd(y, dy)


def sigmoid(x):
    return _SigmoidFn.apply(_planar_f32(x))


class _Add3Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        return ops.add3_f32(a, b, c)

    @staticmethod
    def backward(ctx, d):
        d = ops._pf32(d)
        return d, d, (d if ctx.needs_input_grad[2] else None)


def add3(a, b, c=None):
    """a + b (+ c) on planar fp32 tensors of one shape."""
    return _Add3Fn.apply(_planar_f32(a), _planar_f32(b), None if c is None else _planar_f32(c))


class _BilinearACFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s, out_f32):
        ctx.s = s
        return ops.bilinear_ac_up_fwd(x, s, out_f32)

    @staticmethod
    def backward(ctx, dout):
        return ops.bilinear_ac_up_bwd(dout, ctx.s), None, None


class UpsamplingBilinear2d(nn.Module):
    """nn.UpsamplingBilinear2d(scale_factor=s): bilinear with align_corners=True (`use_torch_up`,
    models/dilated_fcn.py:354-355,443-444).  Output dtype: `logits_dtype`."""

    def __init__(self, scale_factor, out_f32=False):
        super().__init__()
        self.scale_factor, self.out_f32 = int(scale_factor), out_f32

    def forward(self, x):
        return _BilinearACFn.apply(_planar_f32(x), self.scale_factor, self.out_f32 or _logits_dtype == F32)
