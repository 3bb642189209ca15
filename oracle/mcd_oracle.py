"""ORACLE - TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

A plain fp32 restatement (stock torch.nn.functional ops, NCHW, no custom kernels) of the reference's MCD hot
path, written functionally over state_dicts that use the REFERENCE's key names.  It is what the CUDA path is
compared against in tests/, in __graft_entry__.smoke() and what bench.py times as `cpu_baseline` /
`--impl reference`.

The arithmetic of the reference lives in a third-party dependency that is not under /root/reference:
PyTorch (pinned `torch-0.4.1` CPU wheel, requirements.txt:7).  The oracle therefore calls the same torch
operators the reference's modules call, from the same call sites:

  conv / BN / ReLU / residual      models/drn.py:21-23,26-59,126-131,171-205
  seg 1x1 conv                     models/dilated_fcn.py:226-232,243
  depthwise deconv heads           models/dilated_fcn.py:357-366,465-470,479-491
  decoders, bilinear upsample      models/dilated_fcn.py:632-658,661-739,790-1019
  CrossEntropyLoss2d / Diff2d / bce2d   loss.py:7-13,93-100,131-138
  class weights, entropy           util.py:99-111,44-48
  MCD phases A / B / C x num_k     adapt_trainer.py:162-212, adapt_mfnet_trainer.py:181-235,
                                   adapt_triple_multitask_trainer.py:202-287
  SGD(momentum, weight decay)      models/model_util.py:289-302 (torch.optim.SGD semantics)

Pinning: tests/golden/make_golden.py imports the real reference modules (py3-patched throw-away copy) in the
build container, runs them on seeded inputs with `fill_state_dict_` weights and stores the outputs under
tests/golden/; tests/test_oracle_golden.py checks this file against those vectors.  The reference's own
tests (test/test_loss.py) do not touch this path, so those golden vectors are the pin.
"""
import math
import zlib

import numpy as np
import torch
import torch.nn.functional as F

LAYERS = {"drn_d_22": [1, 1, 2, 2, 2, 2, 1, 1], "drn_d_38": [1, 1, 3, 4, 6, 3, 1, 1],
          "drn_d_54": [1, 1, 3, 4, 6, 3, 1, 1], "drn_d_105": [1, 1, 3, 4, 23, 3, 1, 1],
          "drn_c_26": [1, 1, 2, 2, 2, 2, 1, 1], "drn_c_42": [1, 1, 3, 4, 6, 3, 1, 1],
          "drn_c_58": [1, 1, 3, 4, 6, 3, 1, 1]}
BOTTLENECK = ("drn_d_54", "drn_d_105", "drn_c_58")  # models/drn.py:337-348,317-320: Bottleneck blocks, expansion 4
ARCH_C = ("drn_c_26", "drn_c_42", "drn_c_58")     # models/drn.py:113-124,147-153: BasicBlock stages 1, 2, 7, 8
CHANNELS = (16, 32, 64, 128, 256, 512, 512, 512)
BN_EPS, BN_MOMENTUM = 1e-5, 0.1


# ------------------------------------------------------------------------------------------------
# architecture description (models/drn.py:103-205): a list of stages; each stage is a list of units
#   ("cbr", key_conv, key_bn, stride, dil)                               conv3x3/7x7 + BN + ReLU
#   ("block", prefix, stride, dil1, dil2, has_downsample)                 BasicBlock
#   ("bneck", prefix, stride, dil1, dil2, has_downsample)                 Bottleneck (1x1, 3x3 dilation dil2, 1x1 x4)
# a 7th element False marks a BasicBlock built with residual=False (arch C stages 7, 8)
def trunk_spec(name="drn_d_38", prefix="base."):
    layers = LAYERS[name]
    kind, exp = ("bneck", 4) if name in BOTTLENECK else ("block", 1)
    if name in ARCH_C:
        return _trunk_spec_c(layers, kind, exp, prefix)
    spec = [[("cbr", "%s0.0" % prefix, "%s0.1" % prefix, 1, 1, 3)]]          # 7x7 pad 3
    inplanes = CHANNELS[0]

    def conv_stack(idx, ch, convs, stride=1, dil=1):
        nonlocal inplanes
        units = []
        for j in range(convs):
            units.append(("cbr", "%s%d.%d" % (prefix, idx, 3 * j), "%s%d.%d" % (prefix, idx, 3 * j + 1),
                          stride if j == 0 else 1, dil, dil))
            inplanes = ch
        return units

    def block_stack(idx, planes, blocks, stride=1, dilation=1, new_level=True):
        nonlocal inplanes
        units = []
        ds = stride != 1 or inplanes != planes * exp
        d0 = (1, 1) if dilation == 1 else ((dilation // 2 if new_level else dilation), dilation)
        units.append((kind, "%s%d.0" % (prefix, idx), stride, d0[0], d0[1], ds))
        inplanes = planes * exp
        for b in range(1, blocks):
            units.append((kind, "%s%d.%d" % (prefix, idx, b), 1, dilation, dilation, False))
        return units

    spec.append(conv_stack(1, CHANNELS[0], layers[0], stride=1))
    spec.append(conv_stack(2, CHANNELS[1], layers[1], stride=2))
    spec.append(block_stack(3, CHANNELS[2], layers[2], stride=2))
    spec.append(block_stack(4, CHANNELS[3], layers[3], stride=2))
    spec.append(block_stack(5, CHANNELS[4], layers[4], dilation=2, new_level=False))
    spec.append(block_stack(6, CHANNELS[5], layers[5], dilation=4, new_level=False))
    spec.append(conv_stack(7, CHANNELS[6], layers[6], dil=2))
    spec.append(conv_stack(8, CHANNELS[7], layers[7], dil=1))
    return spec


def _trunk_spec_c(layers, kind, exp, prefix):
    """arch C (models/drn.py:113-124,147-153): children conv1, bn1, relu, layer1..layer8 -> `prefix`0, 1, 2, 3..10."""
    spec = [[("cbr", "%s0" % prefix, "%s1" % prefix, 1, 1, 3)]]
    state = {"inplanes": CHANNELS[0]}

    def stack(stage, kind_, exp_, planes, blocks, stride=1, dilation=1, new_level=True, residual=True):
        idx = stage + 2
        ds = stride != 1 or state["inplanes"] != planes * exp_
        d0 = (1, 1) if dilation == 1 else ((dilation // 2 if new_level else dilation), dilation)
        units = [(kind_, "%s%d.0" % (prefix, idx), stride, d0[0], d0[1], ds, residual)]
        state["inplanes"] = planes * exp_
        units += [(kind_, "%s%d.%d" % (prefix, idx, b), 1, dilation, dilation, False, residual)
                  for b in range(1, blocks)]
        return units

    spec.append(stack(1, "block", 1, CHANNELS[0], layers[0], stride=1))
    spec.append(stack(2, "block", 1, CHANNELS[1], layers[1], stride=2))
    spec.append(stack(3, kind, exp, CHANNELS[2], layers[2], stride=2))
    spec.append(stack(4, kind, exp, CHANNELS[3], layers[3], stride=2))
    spec.append(stack(5, kind, exp, CHANNELS[4], layers[4], dilation=2, new_level=False))
    spec.append(stack(6, kind, exp, CHANNELS[5], layers[5], dilation=4, new_level=False))
    spec.append(stack(7, "block", 1, CHANNELS[6], layers[6], dilation=2, new_level=False, residual=False))
    spec.append(stack(8, "block", 1, CHANNELS[7], layers[7], dilation=1, new_level=False, residual=False))
    return spec


# ------------------------------------------------------------------------------------------------
# storage emulation: `with storage(torch.bfloat16):` makes the oracle round tensors to the CUDA path's STORAGE
# precision at the CUDA path's storage points (network input, conv weights, conv outputs, unit outputs,
# full-resolution logits; gradients of the same tensors on the way back) while all arithmetic stays fp32.
# DRN-D-38 with train-mode BatchNorm at random weights amplifies any perturbation by ~1.2x per layer, so two
# correct implementations with different storage precision drift apart by ~20 % after 41 layers; this mode
# separates that inherent drift from kernel errors (see DESIGN.md "numerics").
_STORAGE = None


class storage:
    """storage(dtype): every storage point rounds to `dtype`.  Keyword overrides per role (None = no rounding at
    that point, "same" = `dtype`): y = pre-BatchNorm convolution outputs, w = packed weights, grad = every gradient
    tensor on the way back, act = unit inputs / outputs and full-resolution logits."""

    def __init__(self, dtype, y="same", w="same", grad="same", act="same"):
        pick = lambda v: dtype if isinstance(v, str) else v
        self.cfg = None if dtype is None and all(isinstance(v, str) or v is None for v in (y, w, grad, act)) else \
            {"act": pick(act), "y": pick(y), "w": pick(w), "grad": pick(grad)}

    def __enter__(self):
        global _STORAGE
        self.prev, _STORAGE = _STORAGE, self.cfg

    def __exit__(self, *a):
        global _STORAGE
        _STORAGE = self.prev


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fdt, bdt):
        ctx.bdt = bdt
        return x.to(fdt).to(x.dtype) if fdt is not None else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (g.to(ctx.bdt).to(g.dtype) if ctx.bdt is not None else g), None, None


def _q(x, role="act"):      # stored activation: value and its gradient are rounded
    return x if _STORAGE is None else _Round.apply(x, _STORAGE[role], _STORAGE["grad"])


def _qw(w):     # packed weight shadow: rounded copy, fp32 gradient
    return w if _STORAGE is None else _Round.apply(w, _STORAGE["w"], None)


def _qg(x):     # fp32 score map whose incoming gradient is stored rounded
    return x if _STORAGE is None else _Round.apply(x, None, _STORAGE["grad"])


def _bn(sd, key, x, train):
    if train:
        sd[key + ".num_batches_tracked"] += 1
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"],
                        sd[key + ".bias"], train, BN_MOMENTUM, BN_EPS)


def _bn_train_flag(train, fix_bn):
    return train and not fix_bn


def unit_forward(sd, unit, x, bn_train=True, taps=None):
    """one DRN unit of trunk_spec(): conv+BN+ReLU (models/drn.py:195-205) or BasicBlock (models/drn.py:43-59)."""
    if unit[0] == "cbr":
        _, kc, kb, stride, dil, pad = unit
        y = _q(F.conv2d(x, _qw(sd[kc + ".weight"]), None, stride, pad, dil), "y")
        out = _q(F.relu(_bn(sd, kb, y, bn_train)))
        if taps is not None:
            taps[kc + ":conv"] = y
            taps[kc + ":out"] = out
        return out
    _, p, stride, d1, d2, ds = unit[:6]
    residual = unit[6] if len(unit) > 6 else True
    if unit[0] == "bneck":          # Bottleneck.forward, models/drn.py:80-100
        y1 = _q(F.conv2d(x, _qw(sd[p + ".conv1.weight"])), "y")
        o = _q(F.relu(_bn(sd, p + ".bn1", y1, bn_train)))
        y2 = _q(F.conv2d(o, _qw(sd[p + ".conv2.weight"]), None, stride, d2, d2), "y")
        o = _q(F.relu(_bn(sd, p + ".bn2", y2, bn_train)))
        y3 = _q(F.conv2d(o, _qw(sd[p + ".conv3.weight"])), "y")
        o = _bn(sd, p + ".bn3", y3, bn_train)
        res = x
        if ds:
            yd = _q(F.conv2d(x, _qw(sd[p + ".downsample.0.weight"]), None, stride, 0, 1), "y")
            res = _bn(sd, p + ".downsample.1", yd, bn_train)
        out = _q(F.relu(o + res))
        if taps is not None:
            taps[p + ".conv1:conv"], taps[p + ".conv2:conv"], taps[p + ".conv3:conv"] = y1, y2, y3
            taps[p + ":out"] = out
        return out
    y1 = _q(F.conv2d(x, _qw(sd[p + ".conv1.weight"]), None, stride, d1, d1), "y")
    o = _q(F.relu(_bn(sd, p + ".bn1", y1, bn_train)))
    y2 = _q(F.conv2d(o, _qw(sd[p + ".conv2.weight"]), None, 1, d2, d2), "y")
    o = _bn(sd, p + ".bn2", y2, bn_train)
    res = x
    if ds:                          # evaluated even with residual=False (models/drn.py:52-53): BatchNorm bookkeeping
        yd = _q(F.conv2d(x, _qw(sd[p + ".downsample.0.weight"]), None, stride, 0, 1), "y")
        res = _bn(sd, p + ".downsample.1", yd, bn_train)
    out = _q(F.relu(o + res if residual else o))
    if taps is not None:
        taps[p + ".conv1:conv"] = y1
        taps[p + ".conv2:conv"] = y2
        taps[p + ":out"] = out
    return out


def unit_key(unit):
    return (unit[1] if unit[0] == "cbr" else unit[1]) + ":out"


def trunk_forward(sd, x, name="drn_d_38", prefix="base.", train=True, fix_bn=False, taps=None):
    """returns [h0..h8]; `taps` (dict) collects every conv output / unit output by key."""
    bn_train = _bn_train_flag(train, fix_bn)
    outs = []
    x = _q(x)
    for stage in trunk_spec(name, prefix):
        for unit in stage:
            x = unit_forward(sd, unit, x, bn_train, taps)
        outs.append(x)
    return outs


def seg_base_forward(sd, x, name="drn_d_38", train=True, fix_bn=False, taps=None):
    """DRNSegBase.forward: trunk + 1x1 seg conv (models/dilated_fcn.py:238-244)."""
    h = trunk_forward(sd, x, name, "base.", train, fix_bn, taps)[-1]
    return _qg(F.conv2d(h, _qw(sd["seg.weight"]), sd["seg.bias"]))


def up_head(w, x):
    """ConvTranspose2d(C,C,16,stride=8,padding=4,groups=C,bias=False) (models/dilated_fcn.py:357-366)."""
    return _q(F.conv_transpose2d(x, w, None, stride=8, padding=4, groups=x.shape[1]))


def head_forward(sd, feats, kind="single"):
    """kind: 'single' DRNSegPixelClassifier, 'add' FusionDRNSegPixelClassifier(AddFusion),
    'scoreadd' ScoreFusionDRNSegPixelClassifier."""
    if kind == "single":
        return up_head(sd["up.weight"], feats)
    if kind == "add":
        return up_head(sd["up.weight"], feats[0] + feats[1])
    return up_head(sd["up1.weight"], feats[0]) + up_head(sd["up2.weight"], feats[1])


def bilinear_up(x, s):
    return _q(F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=False))


def three_layer_decoder(sd, p, x, train=True, fix_bn=False):
    """ThreeLayerDecoder: CBR 3x3 -> CBR 1x1 -> conv 1x1, all with bias (models/dilated_fcn.py:632-658)."""
    bn_train = _bn_train_flag(train, fix_bn)
    y = _q(F.conv2d(x, _qw(sd[p + ".cbr1.conv.weight"]), sd[p + ".cbr1.conv.bias"], padding=1), "y")
    x = _q(F.relu(_bn(sd, p + ".cbr1.bn", y, bn_train)))
    y = _q(F.conv2d(x, _qw(sd[p + ".cbr2.conv.weight"]), sd[p + ".cbr2.conv.bias"]), "y")
    x = _q(F.relu(_bn(sd, p + ".cbr2.bn", y, bn_train)))
    return _qg(F.conv2d(x, _qw(sd[p + ".conv3.weight"]), sd[p + ".conv3.bias"]))


# ------------------------------------------------------------------------------------------------
# losses
def class_weight(n_class, add_bg_loss=False):
    w = torch.ones(n_class)
    if not add_bg_loss:
        w[n_class - 1] = 0
    return w


def ce2d(logits, target, weight=None, ignore_index=-100):
    return F.nll_loss(F.log_softmax(logits, dim=1), target, weight, ignore_index=ignore_index)


def prob_ce2d(p, target, weight=None, ignore_index=-100):
    """ProbCrossEntropyLoss2d (loss.py:16-30): NLLLoss2d(weight)(log(p), target) on a probability map"""
    return F.nll_loss(torch.log(p), target, weight, ignore_index=ignore_index)


def diff2d(a, b):
    return torch.mean(torch.abs(F.softmax(a, dim=1) - F.softmax(b, dim=1)))


def bce2d(p, t):
    beta = 1 - torch.mean(t)
    w = 1 - beta + (2 * beta - 1) * t
    return F.binary_cross_entropy(p, t, w)


def calc_entropy(logits):
    p = F.softmax(logits, dim=1)
    return -torch.mean(p * torch.log(p + 1e-6))


def predict_labels(logits, n_valid):
    return logits[:, :n_valid].max(1)[1]


# ------------------------------------------------------------------------------------------------
# triple-task decoder (models/dilated_fcn.py:790-1019), default flags
def triple_semseg(sd, hd, train=True, fix_bn=False):
    return (bilinear_up(three_layer_decoder(sd, "semsegcls_dec1", hd["h8"], train, fix_bn), 8),
            bilinear_up(three_layer_decoder(sd, "semsegcls_dec2", hd["h8"], train, fix_bn), 8))


def triple_depth(sd, hd, train=True, fix_bn=False):
    return bilinear_up(three_layer_decoder(sd, "deprgr_dec", hd["h8"], train, fix_bn), 8)


def triple_boundary(sd, hd):
    h1 = bilinear_up(_qg(F.conv2d(hd["h2"], _qw(sd["conv1.weight"]), sd["conv1.bias"])), 2)
    h2 = bilinear_up(_qg(F.conv2d(hd["h3"], _qw(sd["conv2.weight"]), sd["conv2.bias"])), 4)
    h3 = bilinear_up(_qg(F.conv2d(hd["h8"], _qw(sd["conv3.weight"]), sd["conv3.bias"])), 8)
    return (torch.sigmoid(h1) + torch.sigmoid(h2) + torch.sigmoid(h3)) / 3


def _uw(s, value):
    return torch.exp(-s) * value + s


def triple_get_loss(sd, hd, gt_semseg, gt_dep, gt_bd, weight, train=True):
    s1, s2 = triple_semseg(sd, hd, train)
    l1, l2 = ce2d(s1, gt_semseg, weight), ce2d(s2, gt_semseg, weight)
    semseg = (_uw(sd["s_semsegcls"], l1) + _uw(sd["s_semsegcls"], l2)) / 2
    dep = _uw(sd["s_deprgr"], F.mse_loss(triple_depth(sd, hd, train), gt_dep))
    bd = _uw(sd["s_boundary"], bce2d(triple_boundary(sd, hd), gt_bd))
    return semseg, dep, bd


def encoder_dict(sd, x, name="drn_d_38", train=True, fix_bn=False):
    hs = trunk_forward(sd, x, name, "main_layer", train, fix_bn)
    return {"h%d" % i: h for i, h in enumerate(hs)}


# ------------------------------------------------------------------------------------------------
# parameter creation with the reference's initialisation distributions
def _he_normal(shape, gen):
    fan = shape[2] * shape[3] * shape[0]
    return torch.randn(shape, generator=gen) * math.sqrt(2.0 / fan)


def _bn_state(sd, key, c):
    sd[key + ".weight"] = torch.ones(c)
    sd[key + ".bias"] = torch.zeros(c)
    sd[key + ".running_mean"] = torch.zeros(c)
    sd[key + ".running_var"] = torch.ones(c)
    sd[key + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)


def init_trunk(name="drn_d_38", input_ch=3, prefix="base.", gen=None):
    """He-normal convs, BN gamma=1 beta=0 (models/drn.py:163-169); 4-6 channel first conv duplicates the RGB
    filters (models/drn.py:285-288)."""
    gen = gen or torch.Generator().manual_seed(0)
    sd = {}
    cin = 3
    for stage in trunk_spec(name, prefix):
        for unit in stage:
            if unit[0] == "cbr":
                _, kc, kb, stride, dil, pad = unit
                idx = int(kc[len(prefix):].split(".")[0])
                cout = CHANNELS[0] if idx == 0 else (CHANNELS[0] if idx == 1 else
                                                     CHANNELS[1] if idx == 2 else CHANNELS[idx - 1])
                k = 7 if idx == 0 else 3
                sd[kc + ".weight"] = _he_normal((cout, cin, k, k), gen)
                _bn_state(sd, kb, cout)
                cin = cout
            else:
                _, p, stride, d1, d2, ds = unit[:6]
                idx = int(p[len(prefix):].split(".")[0])
                planes = CHANNELS[idx - 1]
                if unit[0] == "bneck":
                    sd[p + ".conv1.weight"] = _he_normal((planes, cin, 1, 1), gen)
                    _bn_state(sd, p + ".bn1", planes)
                    sd[p + ".conv2.weight"] = _he_normal((planes, planes, 3, 3), gen)
                    _bn_state(sd, p + ".bn2", planes)
                    sd[p + ".conv3.weight"] = _he_normal((4 * planes, planes, 1, 1), gen)
                    _bn_state(sd, p + ".bn3", 4 * planes)
                    if ds:
                        sd[p + ".downsample.0.weight"] = _he_normal((4 * planes, cin, 1, 1), gen)
                        _bn_state(sd, p + ".downsample.1", 4 * planes)
                    cin = 4 * planes
                    continue
                sd[p + ".conv1.weight"] = _he_normal((planes, cin, 3, 3), gen)
                _bn_state(sd, p + ".bn1", planes)
                sd[p + ".conv2.weight"] = _he_normal((planes, planes, 3, 3), gen)
                _bn_state(sd, p + ".bn2", planes)
                if ds:
                    sd[p + ".downsample.0.weight"] = _he_normal((planes, cin, 1, 1), gen)
                    _bn_state(sd, p + ".downsample.1", planes)
                cin = planes
    if input_ch != 3:
        w3 = sd[prefix + "0.0.weight"]
        if input_ch == 1:
            sd[prefix + "0.0.weight"] = w3[:, 0:1].clone()
        else:
            sd[prefix + "0.0.weight"] = torch.cat([w3, w3[:, 0:input_ch - 3]], 1)
    return sd


def _default_conv(sd, key, cout, cin, k, gen, bias=True):
    """nn.Conv2d default init: kaiming_uniform(a=sqrt(5)) = U(+-1/sqrt(fan_in)) for weight and bias."""
    bound = 1.0 / math.sqrt(cin * k * k)
    sd[key + ".weight"] = (torch.rand((cout, cin, k, k), generator=gen) * 2 - 1) * bound
    if bias:
        sd[key + ".bias"] = (torch.rand(cout, generator=gen) * 2 - 1) * bound


def init_seg_base(name="drn_d_38", input_ch=6, n_class=41, gen=None):
    gen = gen or torch.Generator().manual_seed(0)
    sd = init_trunk(name, input_ch, "base.", gen)
    sd["seg.weight"] = _he_normal((n_class, 512, 1, 1), gen)
    sd["seg.bias"] = torch.zeros(n_class)
    return sd


def init_head(n_class=41, kind="single", gen=None):
    """nn.ConvTranspose2d default init (weight [C,1,16,16]: fan_in = 256 -> U(+-1/16))."""
    gen = gen or torch.Generator().manual_seed(1)
    keys = ["up.weight"] if kind in ("single", "add") else ["up1.weight", "up2.weight"]
    return {k: (torch.rand((n_class, 1, 16, 16), generator=gen) * 2 - 1) / 16.0 for k in keys}


def init_three_layer_decoder(sd, p, out_ch, gen):
    _default_conv(sd, p + ".cbr1.conv", 512, 512, 3, gen)
    _bn_state(sd, p + ".cbr1.bn", 512)
    _default_conv(sd, p + ".cbr2.conv", 512, 512, 1, gen)
    _bn_state(sd, p + ".cbr2.bn", 512)
    _default_conv(sd, p + ".conv3", out_ch, 512, 1, gen)


def init_triple_decoder(n_class=41, depth_ch=3, gen=None):
    gen = gen or torch.Generator().manual_seed(2)
    sd = {"s_semsegcls": torch.ones(1), "s_deprgr": torch.ones(1), "s_boundary": torch.ones(1)}
    init_three_layer_decoder(sd, "semsegcls_dec1", n_class, gen)
    init_three_layer_decoder(sd, "semsegcls_dec2", n_class, gen)
    init_three_layer_decoder(sd, "deprgr_dec", depth_ch, gen)
    init_three_layer_decoder(sd, "nmlrgr_dec", depth_ch, gen)
    _default_conv(sd, "conv1", 1, 32, 1, gen)
    _default_conv(sd, "conv2", 1, 64, 1, gen)
    _default_conv(sd, "conv3", 1, 512, 1, gen)
    return sd


def fill_state_dict_(sd, seed=0):
    """Deterministic, key-addressed refill used to give the reference modules, the oracle and the CUDA modules
    IDENTICAL non-trivial weights without shipping them: every tensor is regenerated from crc32(key)."""
    for key in sorted(sd):
        t = sd[key]
        if not torch.is_floating_point(t):
            continue
        g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(key.encode()))
        if key.endswith("running_var"):
            v = torch.rand(t.shape, generator=g) + 0.5
        elif key.endswith("running_mean") or key.endswith(".bias"):
            v = torch.randn(t.shape, generator=g) * 0.1
        elif t.dim() == 1 and key.endswith(".weight"):       # BN gamma
            v = torch.rand(t.shape, generator=g) + 0.5
        elif t.dim() == 4:
            fan_in = t.shape[1] * t.shape[2] * t.shape[3]
            v = torch.randn(t.shape, generator=g) * math.sqrt(2.0 / fan_in)
        else:                                                # task-uncertainty scalars
            v = torch.rand(t.shape, generator=g) + 0.5
        t.copy_(v.to(t.dtype))
    return sd


def to_device(sd, dev):
    return {k: v.to(dev) for k, v in sd.items()}


def trainable(sd):
    """names of the tensors torch.optim would update (nn.Parameters of the reference modules)."""
    return [k for k, v in sd.items() if torch.is_floating_point(v) and not k.endswith("running_mean")
            and not k.endswith("running_var")]


# ------------------------------------------------------------------------------------------------
# SGD exactly as torch.optim.SGD(momentum, weight_decay) (dampening 0, no nesterov)
class SGD:
    def __init__(self, lr=1e-3, momentum=0.9, weight_decay=2e-5):
        self.lr, self.momentum, self.wd = lr, momentum, weight_decay
        self.buf = {}

    def step(self, sd, grads):
        with torch.no_grad():
            for k, g in grads.items():
                if g is None:
                    continue
                d = g + self.wd * sd[k] if self.wd else g
                if self.momentum:
                    b = self.buf.get((id(sd), k))
                    b = d.clone() if b is None else b.mul_(self.momentum).add_(d)
                    self.buf[(id(sd), k)] = b
                    d = b
                sd[k].add_(d, alpha=-self.lr)


def _grads(loss, sds):
    """d loss / d params for a list of state dicts; returns one {key: grad} per dict."""
    names, leaves = [], []
    for i, sd in enumerate(sds):
        for k in trainable(sd):
            names.append((i, k))
            leaves.append(sd[k])
    gs = torch.autograd.grad(loss, leaves, allow_unused=True)
    out = [dict() for _ in sds]
    for (i, k), g in zip(names, gs):
        out[i][k] = g
    return out


def _req(sds, flag=True):
    for sd in sds:
        for k in trainable(sd):
            sd[k].requires_grad_(flag)


def _dp_seg_base_forward(G, x, name, world):
    """nn.DataParallel(model_g) (--is_data_parallel, models/model_util.py:283-284): the batch is scattered over
    `world` replicas that share the parameters; every replica normalises with ITS OWN batch statistics and only
    replica 0's running-statistics buffers survive; the outputs are gathered (criteria see the global batch)."""
    if world == 1:
        return seg_base_forward(G, x, name)
    outs = []
    for r, xs in enumerate(x.chunk(world)):
        sd = G if r == 0 else {k: (v.clone() if (k.endswith("running_mean") or k.endswith("running_var") or
                                                  k.endswith("num_batches_tracked")) else v) for k, v in G.items()}
        outs.append(seg_base_forward(sd, xs, name))
    return torch.cat(outs)


def mcd_step_early_dp(G, F1, F2, src, lbl, tgt, weight, opt_g, opt_f, world, **kw):
    """mcd_step_early with nn.DataParallel semantics over `world` replicas (global batch in, see above)."""
    return mcd_step_early(G, F1, F2, src, lbl, tgt, weight, opt_g, opt_f, world=world, **kw)


def mcd_step_early(G, F1, F2, src, lbl, tgt, weight, opt_g, opt_f, num_k=4, num_multiply_d_loss=1.0,
                   name="drn_d_38", record=None, world=1):
    """One iteration of adapt_trainer.py:162-212 (phase A, B, num_k x C).  G/F1/F2 are state dicts and are
    updated in place; returns (c_loss, d_loss) like the trainer's running numbers.  `record` (dict) receives
    intermediate tensors for parity tests."""
    _req([G, F1, F2])

    def seg_base_forward(sd, x, nm):      # noqa: F811  (world > 1: nn.DataParallel semantics)
        return _dp_seg_base_forward(sd, x, nm, world)
    # ---- A: source supervised, updates G, F1, F2 (adapt_trainer.py:163-185)
    feat = seg_base_forward(G, src, name)
    o1, o2 = head_forward(F1, feat), head_forward(F2, feat)
    loss = ce2d(o1, lbl, weight) + ce2d(o2, lbl, weight)
    gg, g1, g2 = _grads(loss, [G, F1, F2])
    if record is not None:
        record.update(A_feat=feat.detach(), A_out1=o1.detach(), A_loss=loss.detach(),
                      A_grad_g=gg, A_grad_f1=g1)
    c_loss = float(loss)
    opt_g.step(G, gg)
    opt_f.step(F1, g1)
    if F2 is not F1:
        opt_f.step(F2, g2)
    # ---- B: classifiers maximise the discrepancy on target (adapt_trainer.py:187-200)
    feat = seg_base_forward(G, src, name)
    o1, o2 = head_forward(F1, feat), head_forward(F2, feat)
    loss = ce2d(o1, lbl, weight) + ce2d(o2, lbl, weight)
    feat_t = seg_base_forward(G, tgt, name)
    t1, t2 = head_forward(F1, feat_t), head_forward(F2, feat_t)
    loss = loss - diff2d(t1, t2)
    _, g1, g2 = _grads(loss, [G, F1, F2])
    if record is not None:
        record.update(B_loss=loss.detach(), B_grad_f1=g1)
    opt_f.step(F1, g1)
    if F2 is not F1:
        opt_f.step(F2, g2)
    # ---- C x num_k: generator minimises the discrepancy (adapt_trainer.py:204-212)
    for i in range(num_k):
        feat_t = seg_base_forward(G, tgt, name)
        t1, t2 = head_forward(F1, feat_t), head_forward(F2, feat_t)
        loss = diff2d(t1, t2) * num_multiply_d_loss
        gg, _, _ = _grads(loss, [G, F1, F2])
        if record is not None:
            record.setdefault("C_losses", []).append(float(loss))
            if i == 0:
                record.update(C0_grad_g=gg)
        opt_g.step(G, gg)
    _req([G, F1, F2], False)
    d_loss = float(loss) / num_k
    return c_loss, d_loss


# ------------------------------------------------------------------------------------------------
# MFNet iteration (adapt_mfnet_trainer.py:181-235): two generators (RGB stream, HHA stream) stepped by ONE optimizer_g,
# heads take both score maps.  Quirks kept: phase C is NOT multiplied by num_multiply_d_loss (:233); the
# `optimizer_f.zero_grad()` after phase B (:222) has no effect on any result (the next use re-zeroes anyway).
def mcd_step_mfnet(G3, G1, F1, F2, src, lbl, tgt, weight, opt_g, opt_f, kind="add", num_k=4, name="drn_d_38",
                   record=None):
    sds = [G3, G1, F1, F2]
    _req(sds)

    def fwd(x):
        f = (seg_base_forward(G3, x[:, :3], name), seg_base_forward(G1, x[:, 3:], name))
        return head_forward(F1, f, kind), head_forward(F2, f, kind)

    o1, o2 = fwd(src)
    loss = ce2d(o1, lbl, weight) + ce2d(o2, lbl, weight)
    g3, g1, gf1, gf2 = _grads(loss, sds)
    c_loss = float(loss)
    if record is not None:
        record.update(A_loss=c_loss)
    opt_g.step(G3, g3), opt_g.step(G1, g1)
    opt_f.step(F1, gf1), opt_f.step(F2, gf2)
    o1, o2 = fwd(src)
    loss = ce2d(o1, lbl, weight) + ce2d(o2, lbl, weight)
    t1, t2 = fwd(tgt)
    loss = loss - diff2d(t1, t2)
    _, _, gf1, gf2 = _grads(loss, sds)
    if record is not None:
        record.update(B_loss=float(loss))
    opt_f.step(F1, gf1), opt_f.step(F2, gf2)
    for i in range(num_k):
        t1, t2 = fwd(tgt)
        loss = diff2d(t1, t2)
        g3, g1, _, _ = _grads(loss, sds)
        if record is not None:
            record.setdefault("C_losses", []).append(float(loss))
        opt_g.step(G3, g3), opt_g.step(G1, g1)
    _req(sds, False)
    return c_loss, float(loss) / num_k


# ------------------------------------------------------------------------------------------------
# seg + HHA decoder (models/dilated_fcn.py:661-739) and the two multitask iterations
def init_multitask_decoder(n_class=41, depth_ch=3, gen=None):
    gen = gen or torch.Generator().manual_seed(2)
    sd = {"s_semsegcls": torch.ones(1), "s_deprgr": torch.ones(1)}
    init_three_layer_decoder(sd, "semsegcls_dec1", n_class, gen)
    init_three_layer_decoder(sd, "semsegcls_dec2", n_class, gen)
    init_three_layer_decoder(sd, "deprgr_dec", depth_ch, gen)
    return sd


def _feat(E, x, triple, name):
    hs = trunk_forward(E, x[:, :3], name, "main_layer" if triple else "base.")
    return {"h%d" % i: h for i, h in enumerate(hs)} if triple else {"h8": hs[-1]}


def _weighted_semseg(D, hd, lbl, weight):
    s1, s2 = triple_semseg(D, hd)
    return (_uw(D["s_semsegcls"], ce2d(s1, lbl, weight)) + _uw(D["s_semsegcls"], ce2d(s2, lbl, weight))) / 2


def mcd_step_multitask(E, D, src, lbl, tgt, weight, opt_e, opt_d, triple=True, num_k=4, num_multiply_d_loss=1.0,
                       name="drn_d_38", record=None):
    """adapt_triple_multitask_trainer.py:202-287 (triple: src [N,7,H,W] = RGB + HHA + boundary) or
    adapt_multitask_trainer.py:194-262 (src [N,6,H,W]); tgt [N,6,H,W].  Parameters without a gradient are skipped by
    the optimizer (torch >= 2.0 zero_grad semantics, see mcd_b200/step.py)."""
    sds = [E, D]
    _req(sds)
    gt_dep = src[:, 3:-1] if triple else src[:, 3:]
    # ---- A
    fs, ft = _feat(E, src, triple, name), _feat(E, tgt, triple, name)
    semseg = _weighted_semseg(D, fs, lbl, weight)
    dep = _uw(D["s_deprgr"], F.mse_loss(triple_depth(D, fs), gt_dep))
    tdep = F.mse_loss(triple_depth(D, ft), tgt[:, 3:])
    loss = semseg + dep + tdep
    if triple:
        loss = loss + _uw(D["s_boundary"], bce2d(triple_boundary(D, fs), src[:, -1:]))
    ge, gd = _grads(loss, sds)
    c_loss = float(loss)
    if record is not None:
        record.update(A_loss=c_loss, A_grad_d=gd)
    opt_e.step(E, ge), opt_d.step(D, gd)
    # ---- B (only optimizer_dec steps)
    fs = _feat(E, src, triple, name)
    loss = _weighted_semseg(D, fs, lbl, weight)
    dep_b = _uw(D["s_deprgr"], F.mse_loss(triple_depth(D, fs), gt_dep))
    if not triple:
        loss = loss + dep_b
    # triple (:256-276): get_loss() evaluates depth and boundary too and the objective drops them - but the depth
    # decoder's train-mode BatchNorm layers have taken their running-statistics update by then
    ft = _feat(E, tgt, triple, name)
    if not triple:
        loss = loss + F.mse_loss(triple_depth(D, ft), tgt[:, 3:])
    loss = loss - diff2d(*triple_semseg(D, ft))
    _, gd = _grads(loss, sds)
    if record is not None:
        record.update(B_loss=float(loss))
    opt_d.step(D, gd)
    # ---- C x num_k (only optimizer_enc steps)
    for i in range(num_k):
        ft = _feat(E, tgt, triple, name)
        loss = diff2d(*triple_semseg(D, ft)) * num_multiply_d_loss
        ge, _ = _grads(loss, sds)
        if record is not None:
            record.setdefault("C_losses", []).append(float(loss))
        opt_e.step(E, ge)
    _req(sds, False)
    return c_loss, float(loss) / num_k


def label_boundary(x):
    """get_boundary of models/dilated_fcn.py:769-773"""
    v = x.float()
    return F.max_pool2d(v, kernel_size=3, stride=1, padding=1) != -F.max_pool2d(-v, kernel_size=3, stride=1, padding=1)


def get_boundary_loss(pred, gt, pred_type="semseg", gt_type="semseg"):
    """models/dilated_fcn.py:743-787"""
    gt_b = label_boundary(gt) if gt_type == "semseg" else gt.detach().clone()
    pred_b = label_boundary(pred) if pred_type == "semseg" else pred
    return bce2d(pred_b.float(), gt_b.float())


# ---- byte-side neighbours of the step: numpy / torch restatements --------------------------
IMAGENET_MEAN = (.485, .456, .406, .485, .485, .485)
IMAGENET_STD = (.229, .224, .225, .229, .229, .229)
CITY_MEAN = (0.290101, 0.328081, 0.286964)
CITY_STD = (0.182954, 0.186566, 0.184475)


def img_transform(img_u8_hwc, normalize_way="imagenet"):
    """transform.py:302-314 on one decoded image: ToTensor (uint8 HWC -> float CHW / 255) and Normalize as the
    torchvision 0.2 loop `for t, m, s in zip(tensor, mean, std): t.sub_(m).div_(s)` (the first c entries are used)."""
    x = torch.from_numpy(np.ascontiguousarray(img_u8_hwc)).permute(2, 0, 1).contiguous().float().div(255)
    c = x.shape[0]
    if normalize_way == "imagenet":
        mean, std = IMAGENET_MEAN[:c], IMAGENET_STD[:c]
    elif normalize_way == "city":
        mean, std = CITY_MEAN[:c], CITY_STD[:c]
    else:
        return x
    m = torch.tensor(mean, dtype=torch.float32)[:, None, None]
    s = torch.tensor(std, dtype=torch.float32)[:, None, None]
    return (x - m) / s


def lbl_transform(lbl_u8, n_class, background_id=255):
    """transform.py:21-48,317-324: ToLabel (long) + ReLabel(background_id, n_class - 1)"""
    t = torch.from_numpy(np.ascontiguousarray(lbl_u8)).long()
    t[t == background_id] = n_class - 1
    return t


def assemble_input(rgb, hha, boundary=None, normalize_way="imagenet"):
    """datasets.py:667-695: cat([img_transform(rgb), img_transform(hha)[, ReLabel(255, 1)(boundary).float()]])"""
    parts = [img_transform(rgb, normalize_way), img_transform(hha, normalize_way)]
    if boundary is not None:
        parts.append(lbl_transform(boundary, 2).unsqueeze(0).float())
    return torch.cat(parts)


def unnormalize(x_hwc):
    """transform.py:285-294 (float64 arithmetic of numpy, uint8 cast)"""
    std, mean = np.array([.229, .224, .225]), np.array([.485, .456, .406])
    return np.uint8((np.asarray(x_hwc) * std + mean) * 255)


def fast_hist(a, b, n):
    """eval.py:21-23"""
    k = (a >= 0) & (a < n)
    return np.bincount(n * a[k].astype(int) + b[k], minlength=n ** 2).reshape(n, n)


def per_class_iu(hist):
    """eval.py:26-27"""
    return np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))


def resize_nearest(lbl_u8, size):
    """adapt_tester.py:124-126: PIL's NEAREST resize, restated: the source index of output coordinate x is
    int(0.5 * s + x * s) with s = n_in / n_out accumulated step by step in double precision (PIL Geometry.c affine
    scale path); size = (width, height)."""
    def table(n_in, n_out):
        s, xin, tab = float(n_in) / float(n_out), 0.0, []
        xin = 0.5 * s
        for _ in range(n_out):
            tab.append(min(max(int(xin), 0), n_in - 1))
            xin += s
        return np.array(tab)
    h, w = lbl_u8.shape
    return lbl_u8[table(h, size[1])][:, table(w, size[0])]


# ---- the other discrepancy criteria (loss.py:68-171; get_prob_distance_criterion names) ---------------------------
def _kl_div(inp, target, mean=True):
    """F.kl_div(input, target) of torch: target * (log(target) - input), 0 where target == 0; element mean or sum"""
    out = torch.xlogy(target, target) - target * inp
    return out.mean() if mean else out.sum()


def pair_distance(name, a, b, size_average=True):
    pa, pb = F.softmax(a, dim=1), F.softmax(b, dim=1)
    la, lb = F.log_softmax(a, dim=1), F.log_softmax(b, dim=1)
    if name == "diff":
        return diff2d(a, b)
    if name in ("symkl", "nmlsymkl"):                    # loss.py:103-117
        return 0.5 * (_kl_div(la, pb, size_average) + _kl_div(lb, pa, size_average))
    if name == "mysymkl":                                # loss.py:141-151
        return torch.mean(0.5 * (pa * torch.log(pa / pb) + pb * torch.log(pb / pa)))
    if name == "jsd":                                    # loss.py:79-90
        lm = F.log_softmax(0.5 * (a + b), dim=1)
        return 0.5 * (_kl_div(lm, pa, size_average) + _kl_div(lm, pb, size_average))
    if name in ("mis_symkl", "spatial_jsd"):             # loss.py:68-76,154-170: kl_div fed with probabilities
        return 0.5 * (_kl_div(pa, pb) + _kl_div(pb, pa))
    raise NotImplementedError(name)


# ------------------------------------------------------------------------------------------------
# option surface around the hot path
def fusion(sd, kind, x1, x2, prefix="fusion."):
    """models/fusion.py:6-50.  kind: gate | scoregate (GateFusion with apply_softmax) | add | concat | concatconv."""
    if kind in ("gate", "scoregate"):
        if kind == "scoregate":
            x1, x2 = F.softmax(x1, 1), F.softmax(x2, 1)
        g = torch.sigmoid(F.conv2d(torch.cat([x1, x2], 1), sd[prefix + "conv.weight"], sd[prefix + "conv.bias"]))
        return x1 * g + x2 * (1 - g)
    if kind == "add":
        return x1 + x2
    if kind == "concat":
        return torch.cat([x1, x2], 1)
    if kind == "concatconv":
        return F.conv2d(torch.cat([x1, x2], 1), sd[prefix + "conv.weight"], sd[prefix + "conv.bias"], padding=1)
    raise ValueError(kind)


def bilinear_up_ac(x, s):
    """nn.UpsamplingBilinear2d(scale_factor=s): align_corners=True (models/dilated_fcn.py:354-355,443-444)."""
    return F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=True)


def fusion_head_forward(sd, kind, x1, x2, ver="ver1", torch_up=False):
    """FusionDRNSegPixelClassifier.forward (models/dilated_fcn.py:465-470)."""
    h = fusion(sd, kind, x1, x2)
    if ver == "ver2":
        h = F.conv2d(h, sd["seg.weight"], sd["seg.bias"])
    if torch_up:
        return bilinear_up_ac(h, 8)
    w = sd["up.weight"]
    groups = w.shape[0] // 2 if kind == "concat" else w.shape[0]
    return F.conv_transpose2d(h, w, None, stride=8, padding=4, groups=groups)


def score_fusion_head_forward(sd, kind, x1, x2):
    """ScoreFusionDRNSegPixelClassifier.forward (models/dilated_fcn.py:487-491)."""
    return fusion(sd, kind, up_head(sd["up1.weight"], x1), up_head(sd["up2.weight"], x2))


def single_head_forward(sd, x, ver="ver1", torch_up=False):
    """DRNSegPixelClassifier.forward (models/dilated_fcn.py:362-366)."""
    if ver == "ver2":
        x = F.conv2d(x, sd["seg.weight"], sd["seg.bias"])
    return bilinear_up_ac(x, 8) if torch_up else up_head(sd["up.weight"], x)


def fuse_seg_base_forward(sd, x, name="drn_d_22", train=True, fix_bn=False):
    """FuseDRNSegBase.forward (models/dilated_fcn.py:294-331): the HHA half, then the RGB half, both through
    `main_layerK`; the HHA activation of every stage is added to the RGB one."""
    bn_train = _bn_train_flag(train, fix_bn)
    spec = trunk_spec(name, "main_layer")
    h, x_d = _q(x[:, 3:]), []
    for stage in spec:
        for unit in stage:
            h = unit_forward(sd, unit, h, bn_train)
        x_d.append(h)
    h = _q(x[:, :3])
    for stage, d in zip(spec, x_d):
        for unit in stage:
            h = unit_forward(sd, unit, h, bn_train)
        h = _q(h + d)
    return _qg(F.conv2d(h, _qw(sd["seg.weight"]), sd["seg.bias"]))


# ---- decoder options (models/dilated_fcn.py:797-1019, 1027-1222, 1258-1398) ---------------------------------------
def _shortcut_sum(sd, hd, prefix):
    """upsample1(c1(h2)) + upsample2(c2(h3)) + upsample3(c3(h8)), 512 channels at full resolution (:866-875)"""
    hs = [bilinear_up(F.conv2d(hd[k], _qw(sd[(prefix % i) + ".weight"]), sd[(prefix % i) + ".bias"]), s)
          for i, k, s in ((1, "h2", 2), (2, "h3", 4), (3, "h8", 8))]
    return hs[0] + hs[1] + hs[2]


def opt_semseg(sd, hd, dec, conv_prefix, shortcut, train=True):
    if shortcut:
        return three_layer_decoder(sd, dec, _shortcut_sum(sd, hd, conv_prefix), train)
    return bilinear_up(three_layer_decoder(sd, dec, hd["h8"], train), 8)


def opt_semseg_losses(sd, hd, gt_semseg, weight, shortcut=False, add_pred_seg_boundary_loss=False, train=True,
                      decs=(("semsegcls_dec1", "seg_conv%d_1"), ("semsegcls_dec2", "seg_conv%d_2"))):
    """get_semseg_loss(separately_returning=True) (:936-954) and the predictions"""
    preds = [opt_semseg(sd, hd, d, c, shortcut, train) for d, c in decs]
    losses = [ce2d(p, gt_semseg, weight) for p in preds]
    if add_pred_seg_boundary_loss:
        losses = [l + get_boundary_loss(p.max(1)[1], gt_semseg) for l, p in zip(losses, preds)]
    return losses, preds


def seg2bd_losses(sd, preds, gt_bd):
    """get_boundary_loss_by_extra_conv (:960-981) given the classifiers' predictions"""
    return [bce2d(torch.sigmoid(F.conv2d(p, _qw(sd["seg2bd_conv.weight"]), sd["seg2bd_conv.bias"], padding=2)),
                  gt_bd.reshape(p.shape[0], 1, *p.shape[2:])) for p in preds]
