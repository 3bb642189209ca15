"""Fusion modules of the two-stream (RGB + HHA) MFNet heads (reference models/fusion.py:6-65), same class names,
constructor arguments and parameter names (=> state_dict keys `fusion.conv.weight` / `.bias`).

AddFusion is on the benchmarked hot path (configs "MFNet-AddFusion" / "MFNet-ScoreAddFusion"); the gate / concat
variants are the SURVEY.md section 8(f4) option surface and are composed of the library's element-wise kernels
(csrc/variants.cu) and its convolution.  Inputs are either fp32 planar score maps (ver1: n_class channels at 1/8
resolution; score fusion: full resolution) or nhwc trunk activations (ver2: 512 channels); the gate / concat variants
compute in planar fp32 and return planar fp32, which the `seg` convolution or the upsampling head that follows takes.

`get_fusion_model` keeps the reference's substring dispatch verbatim, including its consequence that
"ScoreAddFusion" is an AddFusion and "ScoreGateFusion" a GateFusion on channel softmaxes.
"""
import torch.nn as nn

from mcd_b200 import nn as mnn
from mcd_b200 import ops
from mcd_b200.nn import Conv2d


class GateFusion(nn.Module):
    """p = x1 * g + x2 * (1 - g), g = sigmoid(conv1x1(cat(x1, x2))); `apply_softmax`: on channel softmaxes."""

    def __init__(self, inplanes, apply_softmax=False):
        super().__init__()
        self.conv = Conv2d(inplanes * 2, inplanes, kernel_size=1, stride=1, planar_out=True)
        self.apply_softmax = apply_softmax

    def forward(self, x1, x2):
        x1, x2 = mnn.to_planar(x1), mnn.to_planar(x2)
        if self.apply_softmax:
            x1, x2 = mnn.softmax_ch(x1), mnn.softmax_ch(x2)     # F.softmax(x) on a 4-D tensor: dim 1
        gate_logits = self.conv(mnn.cat2(x1, x2))
        return mnn.gate_fuse(x1, x2, gate_logits)


class AddFusion(nn.Module):
    """x1 + x2.  On low-resolution fp32 score maps this is a 0.2 M-element add; the full-resolution
    ScoreAdd variant never materialises its operands - the head's dual-input upsampling kernel adds them.
    nhwc trunk activations (ver2) are added by the library's kernel, which also writes the bf16 twin."""

    def forward(self, x1, x2):
        if ops.is_nhwc(x1) and ops.is_nhwc(x2):
            return mnn.add_nhwc(x1, x2)
        return x1 + x2


class ConcatFusion(nn.Module):
    def forward(self, x1, x2):
        return mnn.cat2(mnn.to_planar(x1), mnn.to_planar(x2))


class ConcatConvFusion(nn.Module):
    def __init__(self, inplanes):
        super().__init__()
        self.conv = Conv2d(inplanes * 2, inplanes, kernel_size=3, padding=1, planar_out=True)

    def forward(self, x1, x2):
        return self.conv(mnn.cat2(mnn.to_planar(x1), mnn.to_planar(x2)))


# first substring match wins, in the reference's order (models/fusion.py:53-65)
_BY_SUBSTRING = (
    ("ScoreGateFusion", lambda n_ch: GateFusion(n_ch, apply_softmax=True)),
    ("GateFusion", lambda n_ch: GateFusion(n_ch)),
    ("AddFusion", lambda n_ch: AddFusion()),
    ("ConcatFusion", lambda n_ch: ConcatFusion()),
    ("ConcatConvFusion", lambda n_ch: ConcatConvFusion(n_ch)),
)


def get_fusion_model(fusion_type, n_ch):
    for key, make in _BY_SUBSTRING:
        if key in fusion_type:
            return make(n_ch)
    raise NotImplementedError()
