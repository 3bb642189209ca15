"""Dilated Residual Network, arch 'D' (BasicBlock: DRN-D-22 / -38; Bottleneck: DRN-D-54 / -105) and arch 'C'
(DRN-C-26 / -42 / -58: BasicBlock stages 1, 2, 7, 8 instead of plain convolution stacks) on libmcd_sm100.

Drop-in for the reference's models/drn.py on the MCD hot path: same factory names, constructor
arguments, module tree and therefore the same state_dict keys (reference models/drn.py:103-253 DRN,
:26-59 BasicBlock, :256-299 replace_first_conv, :323-334 drn_d_22 / drn_d_38).  Convolutions,
BatchNorm, ReLU and the residual add run as fused library kernels on 16-bit NHWC activations.  Bottleneck
(:62-100) and drn_d_54 / drn_d_105 (:337-348) are the SURVEY 8(f) rank-4 widening: same units, 1x1 / 3x3 / 1x1.

Arch 'C' (:113-124,147-153,303-320) registers `conv1`, `bn1`, `relu` as separate children; DRNSegBase's
`nn.Sequential(*children[:-2])` keeps them separate (state_dict keys base.0 / base.1), mcd_b200.nn.SoleChain runs the
triple as one fused unit.  Out of scope: the model-zoo download (`pretrained=True`, no network: see _load_pretrained).
"""
import math

import torch
import torch.nn as nn

from mcd_b200.nn import BatchNorm2d, Conv2d, ConvBNReLU, conv_bn_act

__all__ = ['DRN', 'BasicBlock', 'Bottleneck', 'drn_c_26', 'drn_c_42', 'drn_c_58', 'drn_d_22', 'drn_d_38', 'drn_d_54',
           'drn_d_105', 'replace_first_conv']

# stage -> (channels, dilation) of DRN-D; stages 3..6 are residual, 0..2 and 7..8 plain conv stacks
_CHANNELS = (16, 32, 64, 128, 256, 512, 512, 512)


def conv3x3(in_planes, out_planes, stride=1, padding=1, dilation=1):
    return Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=padding, bias=False,
                  dilation=dilation)


class BasicBlock(nn.Module):
    """conv-bn-relu-conv-bn (+identity | +bn(1x1 conv)) -relu   (reference models/drn.py:26-59)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=(1, 1), residual=True):
        super().__init__()
        self.conv1 = conv3x3(inplanes, planes, stride, padding=dilation[0], dilation=dilation[0])
        self.bn1 = BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)  # parameter-free; kept for module-tree parity
        self.conv2 = conv3x3(planes, planes, padding=dilation[1], dilation=dilation[1])
        self.bn2 = BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride
        self.residual = residual
        self.inner_input = False   # True: the input is the previous block's output and nothing else reads it

    def forward(self, x):
        # `sole`: conv1 is the only convolution reading x (the downsample branch would be a second one); x comes
        # from the previous block of the same stage (DRN._make_layer sets inner_input) or is flagged by DRN.forward
        sole = (self.inner_input or getattr(x, "_mcd_sole", False)) and self.downsample is None
        out = conv_bn_act(self.conv1, self.bn1, x, relu=True, sole=sole)
        if not self.residual:
            if self.downsample is not None and self.downsample[1].training:
                # the reference evaluates the downsample branch even when it does not add it (models/drn.py:52-56):
                # its BatchNorm running statistics move; forward only, the result is dropped
                with torch.no_grad():
                    conv_bn_act(self.downsample[0], self.downsample[1], x, relu=False)
            return conv_bn_act(self.conv2, self.bn2, out, relu=True, sole=True)
        if self.downsample is not None:
            ds_conv, ds_bn = self.downsample[0], self.downsample[1]
            x._mcd_shortcut = True     # (its gradient through the 1x1 convolution joins conv1's dgrad epilogue)
            return conv_bn_act(self.conv2, self.bn2, out, relu=True, res=x, res_conv=ds_conv, res_bn=ds_bn,
                               sole=True)
        # identity shortcut: x feeds conv1 AND the residual add; the flag lets MCDStep fuse the shortcut gradient
        # into conv1's dgrad epilogue (mcd_b200/nn.py, direct-gradient mode)
        x._mcd_shortcut = True
        return conv_bn_act(self.conv2, self.bn2, out, relu=True, res=x, sole=True)


class Bottleneck(nn.Module):
    """1x1 conv-bn-relu, 3x3 (dilated) conv-bn-relu, 1x1 conv-bn (+identity | +bn(1x1 conv)) -relu, output width
    4 x planes (reference models/drn.py:62-100; DRN-D-54 / -105).  `residual` is accepted and ignored like there."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=(1, 1), residual=True):
        super().__init__()
        self.conv1 = Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = BatchNorm2d(planes)
        self.conv2 = Conv2d(planes, planes, kernel_size=3, stride=stride, padding=dilation[1], bias=False,
                            dilation=dilation[1])
        self.bn2 = BatchNorm2d(planes)
        self.conv3 = Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)  # parameter-free; kept for module-tree parity
        self.downsample = downsample
        self.stride = stride
        self.inner_input = False

    def forward(self, x):
        sole = (self.inner_input or getattr(x, "_mcd_sole", False)) and self.downsample is None
        out = conv_bn_act(self.conv1, self.bn1, x, relu=True, sole=sole)
        out = conv_bn_act(self.conv2, self.bn2, out, relu=True, sole=True)
        if self.downsample is not None:
            ds_conv, ds_bn = self.downsample[0], self.downsample[1]
            x._mcd_shortcut = True
            return conv_bn_act(self.conv3, self.bn3, out, relu=True, res=x, res_conv=ds_conv, res_bn=ds_bn, sole=True)
        x._mcd_shortcut = True
        return conv_bn_act(self.conv3, self.bn3, out, relu=True, res=x, sole=True)


class DRN(nn.Module):
    def __init__(self, block, layers, num_classes=1000, channels=_CHANNELS, out_map=False,
                 out_middle=False, pool_size=28, arch='D'):
        super().__init__()
        if arch not in ('C', 'D') or block not in (BasicBlock, Bottleneck):
            raise NotImplementedError("libmcd_sm100 build covers DRN arch 'C' / 'D' (BasicBlock / Bottleneck) only")
        self.inplanes = channels[0]
        self.out_map = out_map
        self.out_dim = channels[-1]
        self.out_middle = out_middle
        self.arch = arch

        if arch == 'C':
            self.conv1 = Conv2d(3, channels[0], kernel_size=7, stride=1, padding=3, bias=False)
            self.bn1 = BatchNorm2d(channels[0])
            self.relu = nn.ReLU(inplace=True)
            self.layer1 = self._make_layer(BasicBlock, channels[0], layers[0], stride=1)
            self.layer2 = self._make_layer(BasicBlock, channels[1], layers[1], stride=2)
        else:
            self.layer0 = ConvBNReLU(
                Conv2d(3, channels[0], kernel_size=7, stride=1, padding=3, bias=False),
                BatchNorm2d(channels[0]), nn.ReLU(inplace=True))
            self.layer1 = self._make_conv_layers(channels[0], layers[0], stride=1)
            self.layer2 = self._make_conv_layers(channels[1], layers[1], stride=2)
        self.layer3 = self._make_layer(block, channels[2], layers[2], stride=2)
        self.layer4 = self._make_layer(block, channels[3], layers[3], stride=2)
        self.layer5 = self._make_layer(block, channels[4], layers[4], dilation=2, new_level=False)
        self.layer6 = None if layers[5] == 0 else \
            self._make_layer(block, channels[5], layers[5], dilation=4, new_level=False)
        if arch == 'C':
            self.layer7 = None if layers[6] == 0 else \
                self._make_layer(BasicBlock, channels[6], layers[6], dilation=2, new_level=False, residual=False)
            self.layer8 = None if layers[7] == 0 else \
                self._make_layer(BasicBlock, channels[7], layers[7], dilation=1, new_level=False, residual=False)
        else:
            self.layer7 = None if layers[6] == 0 else self._make_conv_layers(channels[6], layers[6], dilation=2)
            self.layer8 = None if layers[7] == 0 else self._make_conv_layers(channels[7], layers[7], dilation=1)

        if num_classes > 0:
            # classification head of the ImageNet model: never executed on the MCD path (DRNSegBase drops
            # the last two children) but it is part of the reference's module tree.
            self.avgpool = nn.AvgPool2d(pool_size)
            self.fc = nn.Conv2d(self.out_dim, num_classes, kernel_size=1, stride=1, padding=0, bias=True)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / fan))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1, new_level=True, residual=True):
        assert dilation == 1 or dilation % 2 == 0
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                BatchNorm2d(planes * block.expansion))
        first_dil = (1, 1) if dilation == 1 else (dilation // 2 if new_level else dilation, dilation)
        stack = [block(self.inplanes, planes, stride, downsample, dilation=first_dil, residual=residual)]
        self.inplanes = planes * block.expansion
        stack += [block(self.inplanes, planes, residual=residual, dilation=(dilation, dilation))
                  for _ in range(1, blocks)]
        for b in stack[1:]:
            b.inner_input = True
        return nn.Sequential(*stack)

    def _make_conv_layers(self, channels, convs, stride=1, dilation=1):
        mods = []
        for i in range(convs):
            mods += [Conv2d(self.inplanes, channels, kernel_size=3, stride=stride if i == 0 else 1,
                            padding=dilation, bias=False, dilation=dilation),
                     BatchNorm2d(channels), nn.ReLU(inplace=True)]
            self.inplanes = channels
        return ConvBNReLU(*mods)

    def stages(self):
        if self.arch == 'C':
            stem = ConvBNReLU(self.conv1, self.bn1, self.relu)     # a view: same modules, fused execution
            return [stem] + [s for s in (self.layer1, self.layer2, self.layer3, self.layer4, self.layer5, self.layer6,
                                         self.layer7, self.layer8) if s is not None]
        return [s for s in (self.layer0, self.layer1, self.layer2, self.layer3, self.layer4, self.layer5,
                            self.layer6, self.layer7, self.layer8) if s is not None]

    def forward(self, x):
        # The ImageNet classifier forward (avgpool + fc) is outside the MCD path.
        feats = []
        stages = self.stages()
        for i, stage in enumerate(stages):
            x = stage(x)
            feats.append(x)
            if not self.out_middle and i + 1 < len(stages):
                x._mcd_sole = True     # read by the next stage only (mcd_b200.nn: fused BatchNorm backward)
        if self.out_middle:
            return x, feats[1:]
        return x


def replace_first_conv(res_model, input_ch, arch):
    """1-, 4-, 5- or 6-channel first convolution (reference models/drn.py:256-299): channels 3.. are a copy
    of channels 0.. of the 3-channel filter."""
    if input_ch == 3:
        return res_model
    old_conv = res_model.conv1 if arch == "C" else res_model.layer0[0]
    new_conv = Conv2d(input_ch, 16, kernel_size=7, stride=1, padding=3, bias=False)
    if input_ch == 1:
        new_conv.weight.data = old_conv.weight.data[:, 0:1, :, :].clone()
    elif 3 < input_ch <= 6:
        extra = input_ch - 3
        new_conv.weight.data[:, :3] = old_conv.weight.data
        new_conv.weight.data[:, 3:3 + extra] = old_conv.weight.data[:, 0:extra]
    else:
        raise NotImplementedError()
    if arch == "C":
        res_model.conv1 = new_conv
    else:
        res_model.layer0 = ConvBNReLU(new_conv, res_model.layer0[1], nn.ReLU(inplace=True))
    return res_model


def _load_pretrained(model, name):
    """The reference pulls ImageNet weights with model_zoo.load_url (models/drn.py:8-18,330-334).  There is
    no network here: weights are taken from $MCD_PRETRAINED_DIR/<name>*.pth when present, otherwise the
    He-normal initialisation is kept and a warning is issued."""
    import glob
    import os
    import warnings

    import torch
    root = os.environ.get("MCD_PRETRAINED_DIR", "")
    hits = sorted(glob.glob(os.path.join(root, name.replace("_", "-") + "*.pth")) +
                  glob.glob(os.path.join(root, name + "*.pth"))) if root else []
    if hits:
        model.load_state_dict(torch.load(hits[0], map_location="cpu"))
    else:
        warnings.warn("%s: pretrained=True but no local checkpoint (set MCD_PRETRAINED_DIR); using the "
                      "random He-normal initialisation" % name)


def _build(name, layers, pretrained, input_ch, block=BasicBlock, arch='D', **kwargs):
    model = DRN(block, layers, arch=arch, **kwargs)
    if pretrained:
        _load_pretrained(model, name)
    return replace_first_conv(model, input_ch=input_ch, arch=arch)


def drn_c_26(pretrained=False, input_ch=3, **kwargs):
    return _build("drn_c_26", [1, 1, 2, 2, 2, 2, 1, 1], pretrained, input_ch, arch='C', **kwargs)


def drn_c_42(pretrained=False, input_ch=3, **kwargs):
    return _build("drn_c_42", [1, 1, 3, 4, 6, 3, 1, 1], pretrained, input_ch, arch='C', **kwargs)


def drn_c_58(pretrained=False, input_ch=3, **kwargs):
    return _build("drn_c_58", [1, 1, 3, 4, 6, 3, 1, 1], pretrained, input_ch, block=Bottleneck, arch='C', **kwargs)


def drn_d_22(pretrained=False, input_ch=3, **kwargs):
    return _build("drn_d_22", [1, 1, 2, 2, 2, 2, 1, 1], pretrained, input_ch, **kwargs)


def drn_d_38(pretrained=False, input_ch=3, **kwargs):
    return _build("drn_d_38", [1, 1, 3, 4, 6, 3, 1, 1], pretrained, input_ch, **kwargs)


def drn_d_54(pretrained=False, input_ch=3, **kwargs):
    return _build("drn_d_54", [1, 1, 3, 4, 6, 3, 1, 1], pretrained, input_ch, block=Bottleneck, **kwargs)


def drn_d_105(pretrained=False, input_ch=3, **kwargs):
    return _build("drn_d_105", [1, 1, 3, 4, 23, 3, 1, 1], pretrained, input_ch, block=Bottleneck, **kwargs)
