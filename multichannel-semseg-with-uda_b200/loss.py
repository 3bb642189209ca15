"""Per-pixel losses of the MCD step on libmcd_sm100 (drop-in for the reference's loss.py).

  CrossEntropyLoss2d(weight, size_average, ignore_index)(inputs[B,C,H,W], targets[B,H,W] int64)   loss.py:7-13
  ProbCrossEntropyLoss2d(weight, size_average)(probabilities, targets)   (Gate fusions)          loss.py:16-30
  Diff2d()(inputs1, inputs2) = mean |softmax - softmax|                                          loss.py:93-100
  JSD, Symkl2d, MySymkl2d, MisSymKLD, SpatialJSD2d                                               loss.py:68-171
  bce2d(input, target)   class-balanced binary cross entropy                                     loss.py:131-138
  get_prob_distance_criterion(name, n_class)   every name the reference knows                    loss.py:192-210

Each criterion is ONE fused forward kernel and ONE fused backward kernel over the full-resolution
logits (softmax is never materialised).
Extra entry points used by the multitask decoders: mse_loss, sigmoid3_mean, sigmoid3_bce2d.
"""
import os

import torch
import torch.nn as nn

from mcd_b200 import ops

BF16, F32 = torch.bfloat16, torch.float32
_CHECK_LABELS = os.environ.get("MCD_CHECK_LABELS", "0") == "1"


def _logits(x):
    """full-resolution predictions are bf16 or fp32 NCHW-contiguous tensors (mcd_b200.nn.logits_dtype); anything
    else is converted to fp32 (differentiably)."""
    if x.dtype not in (BF16, F32):
        x = x.float()
    return x.contiguous()


def _gscale(go):
    return go.reshape(1).to(F32).contiguous()


# data-parallel runs (one process per GPU, mcd_b200/parallel.py): with a process group set, every criterion returns
# its rank's SHARE of the loss nn.DataParallel computes on the gathered outputs, so that the shares sum to the
# global-batch loss and SUM-all-reduced gradients equal the global-batch gradients:
#   CrossEntropyLoss2d   local sum w*nll / GLOBAL sum w (the normaliser is all-reduced)
#   Diff2d, mse, bce2d   local mean / world; the class balance beta of bce2d (loss.py:133 `1 - mean(target)`) is
#                        computed from the all-reduced target sum.
_dp_group = False


def set_process_group(group=None):
    """None = the default group (when torch.distributed is initialised with world > 1), False = single process."""
    global _dp_group
    import torch.distributed as dist
    ok = group is not False and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    _dp_group = group if ok else False


def _target_sum(target):
    """(sum(target) over the global batch, global element count)"""
    tsum = ops.sum_f32(target)
    if _dp_group is False:
        return tsum, target.numel()
    import torch.distributed as dist
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM, group=_dp_group)
    return tsum, target.numel() * dist.get_world_size(_dp_group)


def dp_world():
    """number of ranks the criteria share the global batch with (1 = single process)."""
    if _dp_group is False:
        return 1
    import torch.distributed as dist
    return dist.get_world_size(_dp_group)


def _share(go):
    """mean-type criteria return local_mean / world: their rank's share of the global-batch mean"""
    g = _gscale(go)
    w = dp_world()
    return g / w if w > 1 else g


def _check_target(logits, targets, who):
    """nll_loss's own check (torch raises for a target whose batch / spatial sizes differ from the input's): the kernels
    index the label map with the logits' geometry, so a smaller map would be read out of bounds"""
    want = (logits.shape[0],) + tuple(logits.shape[2:])
    if logits.dim() != 4 or tuple(targets.shape) != want:
        raise ValueError("%s: input and target batch or spatial sizes don't match: target %s, input %s"
                         % (who, list(targets.shape), list(logits.shape)))


class _CE2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, weight, ignore_index, size_average, dist_group=False):
        acc = ops.ce2d_fwd(logits, target, weight, ignore_index)
        if dist_group is False:
            dist_group = _dp_group
        if dist_group is not False and size_average:
            # data parallel: the normaliser sum_i w[y_i] is global (what nn.DataParallel computes on the
            # gathered outputs); the numerator stays local so that SUM-all-reduced gradients are exact.
            import torch.distributed as dist
            dist.all_reduce(acc[1:2], op=dist.ReduceOp.SUM, group=dist_group)
        if _CHECK_LABELS and float(acc[2]) != 0:
            raise IndexError("CrossEntropyLoss2d: %d target labels outside [0, %d) and != ignore_index" %
                             (int(acc[2]), logits.shape[1]))
        ctx.save_for_backward(logits, target, weight, acc)
        ctx.ignore_index, ctx.size_average = ignore_index, size_average
        return acc[0] / acc[1] if size_average else acc[0].clone()

    @staticmethod
    def backward(ctx, go):
        logits, target, weight, acc = ctx.saved_tensors
        if not ctx.size_average:
            acc = torch.ones_like(acc)
        d = ops.ce2d_bwd(logits, target, weight, ctx.ignore_index, acc, _gscale(go))
        return d, None, None, None, None, None


class _NLLState(nn.Module):
    """holds the class-weight buffer under the reference's key `nll_loss.weight` (the multitask decoders
    register their criterion as a sub-module, so that key is part of their checkpoints)."""

    def __init__(self, weight):
        super().__init__()
        self.register_buffer("weight", None if weight is None else weight.detach().to(F32).contiguous())


class CrossEntropyLoss2d(nn.Module):
    def __init__(self, weight=None, size_average=True, ignore_index=-100):
        super().__init__()
        self.nll_loss = _NLLState(weight)
        self.size_average = size_average
        self.ignore_index = ignore_index
        self._dist_group = False

    def set_process_group(self, group=None):
        """data-parallel runs: normalise by the GLOBAL sum of class weights (mcd_b200/parallel.py)."""
        import torch.distributed as dist
        self._dist_group = group if (dist.is_initialized() and dist.get_world_size(group) > 1) else False

    def forward(self, inputs, targets):
        logits = _logits(inputs)
        _check_target(logits, targets, "CrossEntropyLoss2d")
        if targets.dtype != torch.int64:
            targets = targets.long()
        w = self.nll_loss.weight
        if w is not None and w.device != logits.device:
            w = w.to(logits.device)
        return _CE2dFn.apply(logits, targets.contiguous(), w, self.ignore_index, self.size_average,
                             self._dist_group)


class _ProbCE2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, target, weight, ignore_index, size_average):
        acc = ops.prob_ce2d_fwd(p, target, weight, ignore_index)
        if _dp_group is not False and size_average:
            import torch.distributed as dist        # global normaliser, local numerator (as CrossEntropyLoss2d)
            dist.all_reduce(acc[1:2], op=dist.ReduceOp.SUM, group=_dp_group)
        if _CHECK_LABELS and float(acc[2]) != 0:
            raise IndexError("ProbCrossEntropyLoss2d: %d target labels outside [0, %d)" % (int(acc[2]), p.shape[1]))
        ctx.save_for_backward(p, target, weight, acc)
        ctx.ignore_index, ctx.size_average = ignore_index, size_average
        return acc[0] / acc[1] if size_average else acc[0].clone()

    @staticmethod
    def backward(ctx, go):
        p, target, weight, acc = ctx.saved_tensors
        if not ctx.size_average:
            acc = torch.ones_like(acc)
        return ops.prob_ce2d_bwd(p, target, weight, ctx.ignore_index, acc, _gscale(go)), None, None, None, None


class ProbCrossEntropyLoss2d(nn.Module):
    """cross entropy between a probability tensor (0..1) and the labels: NLLLoss2d(weight, size_average)(log(inputs),
    targets) - reference loss.py:16-30, the criterion adapt_mfnet_trainer.py:149 selects for the Gate fusions."""

    def __init__(self, weight=None, size_average=True):
        super().__init__()
        self.nll_loss = _NLLState(weight)
        self.size_average = size_average
        self.ignore_index = -100

    def forward(self, inputs, targets):
        p = inputs if inputs.dtype == F32 else inputs.float()
        _check_target(p, targets, "ProbCrossEntropyLoss2d")
        if targets.dtype != torch.int64:
            targets = targets.long()
        w = self.nll_loss.weight
        if w is not None and w.device != p.device:
            w = w.to(p.device)
        return _ProbCE2dFn.apply(p.contiguous(), targets.contiguous(), w, self.ignore_index, self.size_average)


class _Diff2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        need_bwd = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        if need_bwd:
            acc, stats = ops.diff2d_fwd(a, b, want_stats=True)
        else:
            acc, stats = ops.diff2d_fwd(a, b), None
        ctx.save_for_backward(a, b, stats)
        return acc[0] / float(a.numel() * dp_world())

    @staticmethod
    def backward(ctx, go):
        a, b, stats = ctx.saved_tensors
        da, db = ops.diff2d_bwd(a, b, _share(go), stats)
        return da, db


class Diff2d(nn.Module):
    def __init__(self, weight=None, size_average=True):
        super().__init__()
        self.weight = weight

    def forward(self, inputs1, inputs2):
        return _Diff2dFn.apply(_logits(inputs1), _logits(inputs2))


class _PairDistFn(torch.autograd.Function):
    """mode 1 symmetric KL, 2 Jensen-Shannon, 3 kl_div on probabilities (ops.pairdist); mean or sum over elements"""

    @staticmethod
    def forward(ctx, a, b, mode, mean):
        acc, _, _ = ops.pairdist(mode, a, b)
        ctx.save_for_backward(a, b)
        ctx.mode, ctx.inv = mode, (1.0 / float(a.numel()) if mean else 1.0)
        return acc[0] * (ctx.inv / dp_world())

    @staticmethod
    def backward(ctx, go):
        a, b = ctx.saved_tensors
        _, da, db = ops.pairdist(ctx.mode, a, b, gscale=_share(go), inv_numel=ctx.inv, want_loss=False)
        return da, db, None, None


def _pairdist(mode, inputs1, inputs2, mean=True):
    a, b = _logits(inputs1), _logits(inputs2)
    if a.dtype != b.dtype:
        a, b = a.float(), b.float()
    return _PairDistFn.apply(a.contiguous(), b.contiguous(), mode, mean)


class JSD(nn.Module):
    """loss.py:79-90: 0.5 (kl_div(log_softmax(m), softmax(a)) + kl_div(log_softmax(m), softmax(b))), m = (a + b) / 2"""

    def __init__(self, weight=None, size_average=True):
        super().__init__()
        self.weight, self.size_average = weight, size_average

    def forward(self, inputs1, inputs2):
        return _pairdist(2, inputs1, inputs2, self.size_average)


class Symkl2d(nn.Module):
    """loss.py:103-117: 0.5 (kl_div(log p1, p2) + kl_div(log p2, p1)); the view(-1, n_target_ch) of the reference only
    regroups the elements of an element-wise mean / sum"""

    def __init__(self, weight=None, n_target_ch=None, size_average=True):
        super().__init__()
        self.weight, self.size_average, self.n_target_ch = weight, size_average, n_target_ch

    def forward(self, inputs1, inputs2):
        return _pairdist(1, inputs1, inputs2, self.size_average)


class MySymkl2d(nn.Module):
    """loss.py:141-151: mean 0.5 (p1 log(p1 / p2) + p2 log(p2 / p1))"""

    def __init__(self, weight=None, size_average=True):
        super().__init__()
        self.weight = weight

    def forward(self, inputs1, inputs2):
        return _pairdist(1, inputs1, inputs2)


class MisSymKLD(nn.Module):
    """loss.py:68-76 ("strange but somehow works well"): F.kl_div fed with probabilities where it expects
    log-probabilities: mean 0.5 (p2 (log p2 - p1) + p1 (log p1 - p2))"""

    def __init__(self, weight=None, size_average=True):
        super().__init__()
        self.weight = weight

    def forward(self, inputs1, inputs2):
        return _pairdist(3, inputs1, inputs2)


class SpatialJSD2d(MisSymKLD):
    """loss.py:154-170: the spatial views it builds are unused; what it returns is MisSymKLD's expression"""


class _MSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        acc = ops.mse_fwd(pred, target)
        ctx.save_for_backward(pred, target)
        return acc[0] / float(pred.numel() * dp_world())

    @staticmethod
    def backward(ctx, go):
        pred, target = ctx.saved_tensors
        return ops.mse_bwd(pred, target, _share(go)), None


def mse_loss(pred, target):
    """F.mse_loss(pred, target) for the HHA regression head (reference models/dilated_fcn.py:712,958)."""
    return _MSEFn.apply(_logits(pred), target.to(F32).contiguous())


class _Sigmoid3BCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h1, h2, h3, target):
        tsum, nglobal = _target_sum(target)
        acc, _ = ops.sigmoid3_bce_fwd(h1, h2, h3, target, tsum, numel_global=nglobal)
        ctx.save_for_backward(h1, h2, h3, target, tsum)
        ctx.nglobal = nglobal
        return acc[0] / float(nglobal)

    @staticmethod
    def backward(ctx, go):
        h1, h2, h3, target, tsum = ctx.saved_tensors
        d1, d2, d3 = ops.sigmoid3_bce_bwd(h1, h2, h3, target, tsum, _share(go), numel_global=ctx.nglobal)
        return d1, d2, d3, None


class _BCE2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, target):
        tsum, nglobal = _target_sum(target)
        acc = ops.bce2d_fwd(p, target, tsum, numel_global=nglobal)
        ctx.save_for_backward(p, target, tsum)
        ctx.nglobal = nglobal
        return acc[0] / float(nglobal)

    @staticmethod
    def backward(ctx, go):
        p, target, tsum = ctx.saved_tensors
        return ops.bce2d_bwd(p, target, tsum, _share(go), numel_global=ctx.nglobal), None


def sigmoid3_bce2d(h1, h2, h3, target):
    """bce2d((sigmoid(h1)+sigmoid(h2)+sigmoid(h3))/3, target) fused (reference models/dilated_fcn.py:913-923,
    1002-1004 + loss.py:131-138)."""
    assert not target.requires_grad, "nn criterions don't compute the gradient w.r.t. targets"
    return _Sigmoid3BCEFn.apply(_logits(h1), _logits(h2), _logits(h3), target.to(F32).contiguous())


def sigmoid3_mean(h1, h2, h3):
    """(sigmoid(h1)+sigmoid(h2)+sigmoid(h3))/3 as a bf16 probability map (inference / testers)."""
    with torch.no_grad():
        _, p = ops.sigmoid3_bce_fwd(_logits(h1), _logits(h2), _logits(h3), want_p=True)
    return p


def bce2d(input, target):
    """Class-balanced binary cross entropy on a probability map (reference loss.py:130-138):
    beta = 1 - mean(target); weights = 1 - beta + (2 beta - 1) target; F.binary_cross_entropy(input, target, weights)
    (mean, torch's log clamp at -100).  One forward and one backward kernel over fp32 maps; the MCD decoders call the
    fused `sigmoid3_bce2d`, which never materialises the probability map."""
    assert not target.requires_grad, "nn criterions don't compute the gradient w.r.t. targets"
    assert input.shape == target.shape, "bce2d: input %s vs target %s" % (tuple(input.shape), tuple(target.shape))
    return _BCE2dFn.apply(input.to(F32).contiguous(), target.to(F32).contiguous())


def get_prob_distance_criterion(criterion_name, n_class=None):
    if criterion_name == "jsd":
        return JSD()
    if criterion_name == 'diff':
        return Diff2d()
    if criterion_name == "symkl":
        return Symkl2d(n_target_ch=n_class)
    if criterion_name == "nmlsymkl":
        return Symkl2d(n_target_ch=n_class, size_average=True)
    if criterion_name == "mysymkl":
        return MySymkl2d()
    if criterion_name == "spatial_jsd":
        return SpatialJSD2d()
    if criterion_name == 'mis_symkl':
        return MisSymKLD()
    raise NotImplementedError()
