"""Segmentation wrappers over DRN for the MCD step, on libmcd_sm100 kernels.

Mirrors the hot classes of the reference's models/dilated_fcn.py with identical constructor signatures,
module names (=> state_dict keys) and call signatures:

  DRNSegBase (:217-250)                       generator G: DRN trunk + 1x1 `seg` conv -> [B,n_class,H/8,W/8]
                                              (ver2: the 512-channel trunk feature, `seg` lives in the head)
  FuseDRNSegBase (:253-337)                   `*_fusenet` generator: RGB and HHA through the trunk, added per stage
  DRNSegPixelClassifier (:340-366)            head F: (ver2: 1x1 `seg` ->) learned depthwise 16x16/s8 deconv, or
                                              nn.UpsamplingBilinear2d (`use_torch_up`) -> [B,n_class,H,W]
  FusionDRNSegPixelClassifier (:431-470)      MFNet head  up([seg](fusion(x1, x2))), any fusion of models/fusion.py
  ScoreFusionDRNSegPixelClassifier (:473-491) MFNet score head  fusion(up1(x1), up2(x2))
  MultiTaskEncoder (:554-566), MultiTaskEncoderReturningMultipleFeaturemaps (:569-629)
  CBR (:632-644), ThreeLayerDecoder (:647-658)
  MCDMultiTaskDecoder (:661-739), MCDTripleMultiTaskDecoder (:790-1024), MCDSegBDMultiTaskDecoder (:1027-1222),
  MultiTaskDecoder (:1225-1255), TripleMultiTaskDecoder (:1258-1398)   incl. the shortcut / seg2bd / pseudo-boundary options

  get_boundary_loss (:743-787)                morphological label-map boundary + bce2d

Score maps (n_class / depth / boundary channels at 1/8, 1/4, 1/2 resolution) are fp32 NCHW tensors; the full-resolution
predictions are fp32 NCHW by default (what the reference's testers call `.cpu().numpy()` on; bfloat16 inside MCDStep,
mcd_b200.nn.logits_dtype); trunk activations are IEEE-half channels_last with a bfloat16 twin while autograd records.

Everything else in the reference file (DRNSeg / UncertainDRNSeg and the vendored fyu/drn CLI, DRNSegBase_2, the DANN
domain classifier) is outside SURVEY.md section 8.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.nn import Parameter

import loss as _loss
from mcd_b200 import nn as mnn
from mcd_b200 import ops
from mcd_b200.nn import (BatchNorm2d, BilinearUpsample, Conv2d, DepthwiseDeconv16s8, SoleChain, UpsamplingBilinear2d,
                         conv_bn_act)
from models import drn
from models.fusion import AddFusion, ConcatFusion, get_fusion_model


def util_predict(logits):
    """pred.max(1)[1]: argmax over ALL channels (reference :994-995), int64 [N,H,W], one kernel."""
    lg = logits if logits.dtype in (torch.bfloat16, torch.float32) else logits.float()
    labels, _ = ops.argmax_entropy(lg.contiguous(), want_labels=True, want_entropy=False)
    return labels


def _he_init(conv):
    import math
    fan = conv.kernel_size[0] * conv.kernel_size[1] * conv.out_channels
    conv.weight.data.normal_(0, math.sqrt(2. / fan))
    conv.bias.data.zero_()


def _trunk(model_name, pretrained, input_ch):
    factory = drn.__dict__.get(model_name)
    if factory is None:
        raise NotImplementedError("%s: only drn_c_26 / _42 / _58 and drn_d_22 / _38 / _54 / _105 are built on "
                                  "libmcd_sm100" % model_name)
    return factory(pretrained=pretrained, num_classes=1000, input_ch=input_ch)


class DRNSegBase(nn.Module):
    def __init__(self, model_name, n_class, pretrained=True, input_ch=3, ver="ver1"):
        super().__init__()
        model = _trunk(model_name, pretrained, input_ch)
        self.base = SoleChain(*list(model.children())[:-2])
        self.ver = ver
        if ver == "ver1":
            self.seg = Conv2d(model.out_dim, n_class, kernel_size=1, bias=True, planar_out=True)
            _he_init(self.seg)
        elif ver == "ver2":
            print("ver2 will be used")

    def forward(self, x):
        x = self.base(x)
        if self.ver == "ver2":
            return x
        return self.seg(x)

    def optim_parameters(self, memo=None):
        for param in self.base.parameters():
            yield param
        for param in self.seg.parameters():
            yield param


class FuseDRNSegBase(nn.Module):
    """`drn_d_*_fusenet` generator (reference :253-337): the HHA half and the RGB half of the input both run through
    `main_layerK` (the `sub_layerK` trunk is constructed and never used - kept: it is part of the state_dict), the HHA
    activations are added after every stage.  Each stage therefore runs twice per forward on shared weights: BatchNorm
    running statistics take two updates, weight gradients accumulate over both passes."""
    _mcd_shared_weights = True      # tells mcd_b200.step.MCDStep: no deferred split-K gradients, no folded BatchNorm updates

    def __init__(self, model_name, n_class, pretrained=True, input_ch=3, ver="ver1"):
        super().__init__()
        assert input_ch in [4, 6]
        self.ver = ver
        model = _trunk(model_name, pretrained, 3)
        sub_model = _trunk(model_name, pretrained, input_ch - 3)
        assert model.arch == "D"
        for i in range(9):
            setattr(self, "main_layer%d" % i, getattr(model, "layer%d" % i))
        for i in range(9):
            setattr(self, "sub_layer%d" % i, getattr(sub_model, "layer%d" % i))
        self.seg = Conv2d(model.out_dim, n_class, kernel_size=1, bias=True, planar_out=True)
        _he_init(self.seg)

    def forward(self, x):
        rgb_inputs = x[:, :3, :, :]
        depth_inputs = x[:, 3:, :, :]
        x_d, h = [], depth_inputs
        for i in range(9):
            h = getattr(self, "main_layer%d" % i)(h)
            x_d.append(h)
        x = rgb_inputs
        for i in range(9):
            x = mnn.add_nhwc(getattr(self, "main_layer%d" % i)(x), x_d[i])
        return self.seg(x)

    def optim_parameters(self, memo=None):
        # the reference iterates `self.base`, which this class does not have (AttributeError when called)
        for param in self.base.parameters():
            yield param
        for param in self.seg.parameters():
            yield param


def _score_conv(n_class):
    seg = Conv2d(512, n_class, kernel_size=1, bias=True, planar_out=True)   # in_ch hard-coded as in the reference
    _he_init(seg)
    return seg


class PairedDeconv16s8(nn.ConvTranspose2d):
    """ConvTranspose2d(2C, C, 16, stride=8, padding=4, groups=C, bias=False): the upsampling of the ConcatFusion head
    (reference :446-449).  Group g reads channels 2g and 2g + 1 of the concatenated score maps, so the output is
    up(h[:, 0::2]; w[0::2]) + up(h[:, 1::2]; w[1::2]) - the library's dual-input upsampling kernel."""

    def __init__(self, n_class):
        super().__init__(2 * n_class, n_class, 16, stride=8, padding=4, output_padding=0, groups=n_class, bias=False)

    def forward(self, h):
        h = mnn._planar_f32(h)
        return mnn._Deconv16s8Fn.apply(h[:, 0::2].contiguous(), self.weight[0::2].contiguous(),
                                       h[:, 1::2].contiguous(), self.weight[1::2].contiguous(), mnn.logits_dtype.get())


class DRNSegPixelClassifier(nn.Module):
    def __init__(self, n_class, use_torch_up=False, dropout=False, ver="ver1"):
        super().__init__()
        self.dropout = dropout
        self.ver = ver
        if ver == "ver2":
            self.seg = _score_conv(n_class)
        self.up = UpsamplingBilinear2d(8) if use_torch_up else DepthwiseDeconv16s8(n_class)

    def forward(self, x):
        if self.ver == "ver2":
            x = self.seg(x)
        return self.up(x)


class FusionDRNSegPixelClassifier(nn.Module):
    def __init__(self, fusion_type, n_class, use_torch_up=False, ver="ver1"):
        super().__init__()
        if ver == "ver1":
            self.fusion = get_fusion_model(fusion_type, n_class)
        elif ver == "ver2":
            self.fusion = get_fusion_model(fusion_type, 512)
        self.ver = ver
        if use_torch_up:
            self.up = UpsamplingBilinear2d(8)
        elif type(self.fusion) is ConcatFusion:
            self.up = PairedDeconv16s8(n_class)
        else:
            self.up = DepthwiseDeconv16s8(n_class)
        if ver == "ver2":
            self.seg = _score_conv(n_class)

    def forward(self, x1, x2):
        h = self.fusion(x1, x2)
        if self.ver == "ver2":
            h = self.seg(h)
        return self.up(h)


class ScoreFusionDRNSegPixelClassifier(nn.Module):
    def __init__(self, fusion_type, n_class):
        super().__init__()
        self.fusion = get_fusion_model(fusion_type, n_class)
        self.up1 = DepthwiseDeconv16s8(n_class)
        self.up2 = DepthwiseDeconv16s8(n_class)

    def forward(self, x1, x2):
        if isinstance(self.fusion, AddFusion):
            # up1(x1) + up2(x2) in ONE pass over the full-resolution tensor (no operand materialised)
            return self.up1(x1, x2, self.up2)
        return self.fusion(self.up1(x1), self.up2(x2))


# ---- multitask encoder / decoders -----------------------------------------------------------------
class MultiTaskEncoder(nn.Module):
    def __init__(self, model_name, pretrained=True, input_ch=3):
        super().__init__()
        model = _trunk(model_name, pretrained, input_ch)
        self.base = SoleChain(*list(model.children())[:-2])

    def forward(self, x):
        return self.base(x)


class MultiTaskEncoderReturningMultipleFeaturemaps(nn.Module):
    """returns {'h0'..'h8'}: h2 32@1/2, h3 64@1/4, h8 512@1/8 feed the triple-task decoder."""

    def __init__(self, model_name, pretrained=True, input_ch=3):
        super().__init__()
        model = _trunk(model_name, pretrained, input_ch)
        for i in range(9):
            setattr(self, "main_layer%d" % i, getattr(model, "layer%d" % i))
        self.up = BilinearUpsample(8)  # parameter-free, unused in forward (as in the reference)

    def forward(self, x):
        out = {}
        for i in range(9):
            x = getattr(self, "main_layer%d" % i)(x)
            out["h%d" % i] = x
        return out


class CBR(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True):
        super().__init__()
        self.conv = Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                           dilation=dilation, groups=groups, bias=bias)
        self.bn = BatchNorm2d(out_channels)

    def forward(self, x):
        return conv_bn_act(self.conv, self.bn, x, relu=True)


class ThreeLayerDecoder(nn.Module):
    def __init__(self, output_ch, input_ch=512):
        super().__init__()
        self.cbr1 = CBR(input_ch, 512, kernel_size=3, padding=1)
        self.cbr2 = CBR(512, 512, kernel_size=1)
        self.conv3 = Conv2d(512, output_ch, kernel_size=1, planar_out=True)

    def forward(self, x):
        return self.conv3(self.cbr2(self.cbr1(x)))


def _scalar_param():
    p = Parameter(torch.Tensor(1))
    p.data.fill_(1)
    return p


def _fused_semseg_ce(dec, scores, gt_semseg):
    """CrossEntropyLoss2d(upsample8(score), gt) for each of the two classifiers' 1/8-resolution score maps as ONE kernel
    per classifier (mcd_b200/headloss.py: the full-resolution logits are never written), or None when the criterion is
    not the library's CrossEntropyLoss2d."""
    from mcd_b200 import headloss
    c = dec.semseg_criterion
    if not headloss.enabled() or type(c) is not _loss.CrossEntropyLoss2d or not headloss.fits(1, 1, scores[0].shape[1], False, True):
        return None
    return tuple(headloss.head_ce2d([sc], None, gt_semseg, c.nll_loss.weight, c.ignore_index, c.size_average)
                 for sc in scores)


def _fused_discrepancy(dec, scores):
    from mcd_b200 import headloss
    if (not headloss.enabled() or type(dec.discrepancy_criterion) is not _loss.Diff2d
            or not headloss.fits(2, 1, scores[0].shape[1], False, True)):
        return None
    return headloss.head_diff2d([scores[0]], None, [scores[1]], None)


def _weighted(s, value):
    """learned log-variance task weighting exp(-s) * L + s (reference :1008-1014).  Data parallel: `value` is this
    rank's share of the global loss (loss.set_process_group), so the regulariser s is shared out as well."""
    return torch.exp(-s) * value + s / _loss.dp_world()


def get_boundary_loss(pred, gt, pred_type="semseg", gt_type="semseg"):
    """reference models/dilated_fcn.py:743-787: the morphological boundary (3x3 max-pool of x and of -x differ) of an
    integer label map - for the prediction and / or the ground truth - compared with bce2d.  The boundary maps are
    hard 0/1 (no gradient flows into a "semseg"-type prediction, exactly as in the reference)."""
    assert pred_type in ["semseg", "boundary"]
    assert gt_type in ["semseg", "boundary"]
    gt_boundary = ops.label_boundary(gt.detach()) if gt_type == "semseg" else gt.detach().clone()
    pred_boundary = ops.label_boundary(pred.detach()) if pred_type == "semseg" else pred
    return _loss.bce2d(pred_boundary.float(), gt_boundary.float().reshape(pred_boundary.shape))


class MCDMultiTaskDecoder(nn.Module):
    def __init__(self, n_class, depth_ch, semseg_criterion=None, discrepancy_criterion=None):
        super().__init__()
        self.s_semsegcls = _scalar_param()
        self.s_deprgr = _scalar_param()
        self.semsegcls_dec1 = ThreeLayerDecoder(n_class)
        self.semsegcls_dec2 = ThreeLayerDecoder(n_class)
        self.deprgr_dec = ThreeLayerDecoder(depth_ch)
        self.semseg_criterion = semseg_criterion
        self.discrepancy_criterion = discrepancy_criterion
        self.upsample = BilinearUpsample(8)

    def _semseg_scores(self, x):
        return self.semsegcls_dec1(x), self.semsegcls_dec2(x)

    def semseg_forward(self, x):
        s1, s2 = self._semseg_scores(x)
        return self.upsample(s1), self.upsample(s2)

    def depth_forward(self, x):
        return self.upsample(self.deprgr_dec(x))

    def forward(self, x):
        pred_semseg1, pred_semseg2 = self.semseg_forward(x)
        return pred_semseg1, pred_semseg2, self.depth_forward(x)

    def get_cls_descrepancy(self, x):
        scores = self._semseg_scores(x)
        d = _fused_discrepancy(self, scores)
        if d is not None:
            return d
        return self.discrepancy_criterion(self.upsample(scores[0]), self.upsample(scores[1]))

    def get_semseg_loss(self, x, gt_semseg, separately_returning=False):
        scores = self._semseg_scores(x)
        fused = _fused_semseg_ce(self, scores, gt_semseg)
        if fused is not None:
            loss1, loss2 = fused
        else:
            loss1 = self.semseg_criterion(self.upsample(scores[0]), gt_semseg)
            loss2 = self.semseg_criterion(self.upsample(scores[1]), gt_semseg)
        return (loss1, loss2) if separately_returning else loss1 + loss2

    def get_depth_loss(self, x, gt_dep):
        return _loss.mse_loss(self.depth_forward(x), gt_dep)

    def get_loss(self, x, gt_semseg, gt_dep, separately_returning=False):
        l1, l2 = self.get_semseg_loss(x, gt_semseg, separately_returning=True)
        semseg_loss = (_weighted(self.s_semsegcls, l1) + _weighted(self.s_semsegcls, l2)) / 2
        depreg_loss = _weighted(self.s_deprgr, self.get_depth_loss(x, gt_dep))
        if separately_returning:
            return semseg_loss, depreg_loss
        return semseg_loss + depreg_loss

    def get_task_weights(self):
        std_semseg = np.sqrt(np.exp(2 * self.s_semsegcls.data.cpu().numpy()))
        std_depth = np.sqrt(np.exp(2 * self.s_deprgr.data.cpu().numpy()))
        return std_semseg, std_depth


class _HedBranches(nn.Module):
    """What the three "triple" decoders of the reference share (:790-1024, :1027-1222, :1258-1398): the HED-style
    boundary branch (1x1 convolutions on h2 / h3 / h8, bilinear x2 / x4 / x8, mean of sigmoids), the optional
    shortcut inputs of the segmentation / depth decoders (1x1 convolutions to 512 channels on the same three feature
    maps, upsampled to FULL resolution and summed - the ThreeLayerDecoder then runs at full resolution), the optional
    5x5 seg -> boundary convolution and the pseudo-boundary losses."""

    def _init_branches(self, n_class, semseg_shortcut, depth_shortcut, add_pred_seg_boundary_loss, use_seg2bd_conv,
                       n_semseg_heads, with_depth):
        self.upsample1 = BilinearUpsample(2)
        self.upsample2 = BilinearUpsample(4)
        self.upsample3 = BilinearUpsample(8)
        self.conv1 = Conv2d(32, 1, kernel_size=1, stride=1, padding=0, planar_out=True)
        self.conv2 = Conv2d(64, 1, kernel_size=1, stride=1, padding=0, planar_out=True)
        self.conv3 = Conv2d(512, 1, kernel_size=1, stride=1, padding=0, planar_out=True)
        self.semseg_shortcut = semseg_shortcut
        self.depth_shortcut = depth_shortcut
        self.add_pred_seg_boundary_loss = add_pred_seg_boundary_loss
        if add_pred_seg_boundary_loss:
            self.s_pred_seg_boundary = _scalar_param()
        self.use_seg2bd_conv = use_seg2bd_conv

        def shortcut_convs(prefix):
            for i, cin in ((1, 32), (2, 64), (3, 512)):
                setattr(self, prefix % i, Conv2d(cin, 512, kernel_size=1, stride=1, padding=0, planar_out=True))
        if semseg_shortcut:
            for k in range(1, n_semseg_heads + 1):
                shortcut_convs("seg_conv%d_" + str(k))
        if depth_shortcut and with_depth:
            shortcut_convs("dep_conv%d")
        if use_seg2bd_conv:
            self.seg2bd_conv = Conv2d(n_class, 1, kernel_size=5, padding=2, planar_out=True)

    def _shortcut_sum(self, x_dic, prefix):
        """upsample1(c1(h2)) + upsample2(c2(h3)) + upsample3(c3(h8)): 512 channels at full resolution, planar fp32"""
        hs = [mnn._BilinearFn.apply(getattr(self, prefix % i)(x_dic[k]), sc, True)
              for i, k, sc in ((1, "h2", 2), (2, "h3", 4), (3, "h8", 8))]
        return mnn.add3(*hs)

    def _boundary_maps(self, x_dic):
        return (self.upsample1(self.conv1(x_dic["h2"])), self.upsample2(self.conv2(x_dic["h3"])),
                self.upsample3(self.conv3(x_dic["h8"])))

    def boundary_forward(self, x_dic):
        return _loss.sigmoid3_mean(*self._boundary_maps(x_dic))

    def _extra_pred_seg_boundary(self, preds, gt_semseg):
        """--add_pred_seg_boundary_loss (reference :941-949): bce2d between the morphological boundaries of each
        classifier's argmax label map and of the ground truth; hard 0/1 maps, so it adds to the value only."""
        return [get_boundary_loss(util_predict(p), gt_semseg) for p in preds]

    def get_boundary_loss_by_extra_conv(self, x_dic, gt_bdry=None, separately_returning=False):
        """--use_seg2bd_conv (reference :960-981): sigmoid(5x5 conv) of each classifier's logits against the boundary
        ground truth or, without one, against the (detached) prediction of the boundary branch."""
        assert self.use_seg2bd_conv
        preds = self.semseg_forward(x_dic)
        preds = preds if isinstance(preds, tuple) else (preds,)
        if gt_bdry is None:
            gt_bdry = self.boundary_forward(x_dic).detach().float()
        losses = []
        for p in preds:
            pred_bdry = mnn.sigmoid(self.seg2bd_conv(p))
            losses.append(_loss.bce2d(pred_bdry, gt_bdry.reshape(pred_bdry.shape)))
        if separately_returning:
            return tuple(losses)
        return sum(losses[1:], losses[0])

    def get_psuedo_boundary_loss(self, x_dic, separately_returning=False):
        """--add_pred_seg_boundary_loss (reference :983-1000): boundary of each classifier's argmax label map against
        the (detached) boundary head.  The reference passes a keyword `pred_semseg=` that get_boundary_loss does not
        have (TypeError); this is the evident intent: pred_type "semseg", gt_type "boundary"."""
        assert self.add_pred_seg_boundary_loss
        psuedo_boundary = self.boundary_forward(x_dic).detach().float()
        preds = self.semseg_forward(x_dic)
        preds = preds if isinstance(preds, tuple) else (preds,)
        losses = [get_boundary_loss(util_predict(p), psuedo_boundary[:, 0], gt_type="boundary") for p in preds]
        if separately_returning:
            return tuple(losses)
        return sum(losses[1:], losses[0])


class _MCDSemsegPair(_HedBranches):
    """the two MCD classifiers of MCDTripleMultiTaskDecoder / MCDSegBDMultiTaskDecoder."""

    def _plain(self):
        """1/8-resolution score maps upsampled x8: the configuration the fused head + loss kernel covers"""
        return not self.semseg_shortcut

    def _semseg_scores(self, x_dic):
        h8 = x_dic["h8"]
        return self.semsegcls_dec1(h8), self.semsegcls_dec2(h8)

    def semseg_forward(self, x_dic):
        if self.semseg_shortcut:
            return (self.semsegcls_dec1(self._shortcut_sum(x_dic, "seg_conv%d_1")),
                    self.semsegcls_dec2(self._shortcut_sum(x_dic, "seg_conv%d_2")))
        s1, s2 = self._semseg_scores(x_dic)
        return self.upsample3(s1), self.upsample3(s2)

    def get_cls_descrepancy(self, x_dic):
        if self._plain():
            scores = self._semseg_scores(x_dic)
            d = _fused_discrepancy(self, scores)
            if d is not None:
                return d
            return self.discrepancy_criterion(self.upsample3(scores[0]), self.upsample3(scores[1]))
        pred_semseg1, pred_semseg2 = self.semseg_forward(x_dic)
        return self.discrepancy_criterion(pred_semseg1, pred_semseg2)

    def _wants_extra(self):
        return self.add_pred_seg_boundary_loss

    def get_semseg_loss(self, x_dic, gt_semseg, separately_returning=False):
        fused = None
        if self._plain() and not self._wants_extra():
            scores = self._semseg_scores(x_dic)
            fused = _fused_semseg_ce(self, scores, gt_semseg)
            if fused is None:
                preds = self.upsample3(scores[0]), self.upsample3(scores[1])
        else:
            preds = self.semseg_forward(x_dic)
        if fused is not None:
            loss1, loss2 = fused
        else:
            loss1 = self.semseg_criterion(preds[0], gt_semseg)
            loss2 = self.semseg_criterion(preds[1], gt_semseg)
            if self._wants_extra():
                extra1, extra2 = self._extra_pred_seg_boundary(preds, gt_semseg)
                loss1, loss2 = loss1 + extra1, loss2 + extra2
        return (loss1, loss2) if separately_returning else loss1 + loss2

    def get_weighted_semseg_loss(self, x, gt_semseg):
        """the `semseg_loss` term of get_loss() alone (phase B of adapt_triple_multitask_trainer.py:256-276 uses
        nothing else of get_loss's three results)."""
        l1, l2 = self.get_semseg_loss(x, gt_semseg, separately_returning=True)
        return (_weighted(self.s_semsegcls, l1) + _weighted(self.s_semsegcls, l2)) / 2


class MCDTripleMultiTaskDecoder(_MCDSemsegPair):
    """semantic segmentation (two MCD classifiers) + HHA regression + boundary detection."""

    def __init__(self, n_class, depth_ch, semseg_criterion=None, discrepancy_criterion=None,
                 semseg_shortcut=False, depth_shortcut=False, add_pred_seg_boundary_loss=False,
                 use_seg2bd_conv=False):
        super().__init__()
        self.s_semsegcls = _scalar_param()
        self.s_deprgr = _scalar_param()
        self.s_boundary = _scalar_param()
        self.semsegcls_dec1 = ThreeLayerDecoder(n_class)
        self.semsegcls_dec2 = ThreeLayerDecoder(n_class)
        self.deprgr_dec = ThreeLayerDecoder(depth_ch)
        self.nmlrgr_dec = ThreeLayerDecoder(depth_ch)  # constructed, never used (reference :813)
        self.semseg_criterion = semseg_criterion
        self.discrepancy_criterion = discrepancy_criterion
        self._init_branches(n_class, semseg_shortcut, depth_shortcut, add_pred_seg_boundary_loss, use_seg2bd_conv,
                            n_semseg_heads=2, with_depth=True)

    def depth_forward(self, x_dic):
        if self.depth_shortcut:
            return self.deprgr_dec(self._shortcut_sum(x_dic, "dep_conv%d"))
        return self.upsample3(self.deprgr_dec(x_dic["h8"]))

    def forward(self, x_dic):
        pred_semseg1, pred_semseg2 = self.semseg_forward(x_dic)
        return pred_semseg1, pred_semseg2, self.depth_forward(x_dic), self.boundary_forward(x_dic)

    def get_depth_loss(self, x_dic, gt_dep):
        return _loss.mse_loss(self.depth_forward(x_dic), gt_dep)

    def get_boundary_loss(self, x_dic, gt_boundary):
        # sigmoid-average + bce2d fused: the averaged probability map is never materialised
        return _loss.sigmoid3_bce2d(*self._boundary_maps(x_dic), gt_boundary)

    def get_loss(self, x, gt_semseg, gt_dep, gt_boundary, separately_returning=False):
        l1, l2 = self.get_semseg_loss(x, gt_semseg, separately_returning=True)
        semseg_loss = (_weighted(self.s_semsegcls, l1) + _weighted(self.s_semsegcls, l2)) / 2
        depreg_loss = _weighted(self.s_deprgr, self.get_depth_loss(x, gt_dep))
        boundary_loss = _weighted(self.s_boundary, self.get_boundary_loss(x, gt_boundary))
        if separately_returning:
            return semseg_loss, depreg_loss, boundary_loss
        return semseg_loss + depreg_loss + boundary_loss

    def get_task_weights(self):
        std_semseg = np.sqrt(np.exp(2 * self.s_semsegcls.data.cpu().numpy()))
        std_depth = np.sqrt(np.exp(2 * self.s_deprgr.data.cpu().numpy()))
        return std_semseg, std_depth


class MCDSegBDMultiTaskDecoder(_MCDSemsegPair):
    """semantic segmentation (two MCD classifiers) + boundary detection; the boundary ground truth is the
    morphological boundary of the segmentation labels (reference :1027-1222, adapt_segbd_multitask_trainer.py)."""

    def __init__(self, n_class, depth_ch, semseg_criterion=None, discrepancy_criterion=None,
                 semseg_shortcut=False, depth_shortcut=False, add_pred_seg_boundary_loss=False,
                 use_seg2bd_conv=False):
        super().__init__()
        self.s_semsegcls = _scalar_param()
        self.s_boundary = _scalar_param()
        self.semsegcls_dec1 = ThreeLayerDecoder(n_class)
        self.semsegcls_dec2 = ThreeLayerDecoder(n_class)
        self.semseg_criterion = semseg_criterion
        self.discrepancy_criterion = discrepancy_criterion
        self._init_branches(n_class, semseg_shortcut, depth_shortcut, add_pred_seg_boundary_loss, use_seg2bd_conv,
                            n_semseg_heads=2, with_depth=False)

    def forward(self, x_dic):
        pred_semseg1, pred_semseg2 = self.semseg_forward(x_dic)
        return pred_semseg1, pred_semseg2, self.boundary_forward(x_dic)

    def _wants_extra(self):
        return True          # this decoder adds the predicted-boundary term unconditionally (reference :1145-1153)

    def get_boundary_loss(self, x_dic, gt_semseg):
        # get_boundary_loss(pred=boundary_forward(x), gt=gt_semseg, pred_type="boundary"), sigmoid mean + bce2d fused
        return _loss.sigmoid3_bce2d(*self._boundary_maps(x_dic), ops.label_boundary(gt_semseg.detach()).unsqueeze(1))

    def get_loss(self, x, gt_semseg, separately_returning=False):
        l1, l2 = self.get_semseg_loss(x, gt_semseg, separately_returning=True)
        semseg_loss = (_weighted(self.s_semsegcls, l1) + _weighted(self.s_semsegcls, l2)) / 2
        boundary_loss = _weighted(self.s_boundary, self.get_boundary_loss(x, gt_semseg))
        if separately_returning:
            return semseg_loss, boundary_loss
        return semseg_loss + boundary_loss

    def get_task_weights(self):
        std_semseg = np.sqrt(np.exp(2 * self.s_semsegcls.data.cpu().numpy()))
        std_depth = np.sqrt(np.exp(2 * self.s_deprgr.data.cpu().numpy()))   # AttributeError, as in the reference
        return std_semseg, std_depth


# ---- source-only decoders (`is_src_only`, reference :1225-1398) ---------------------------------------------------
class MultiTaskDecoder(nn.Module):
    """one segmentation decoder + one depth decoder, predictions at 1/8 resolution (no upsampling in the reference)."""

    def __init__(self, n_class, depth_ch, semseg_criterion, discrepancy_criterion=None):
        super().__init__()
        self.s_semsegcls = Parameter(torch.Tensor(1))   # uninitialised in the reference as well
        self.s_deprgr = Parameter(torch.Tensor(1))
        self.semsegcls_dec = ThreeLayerDecoder(n_class)
        self.deprgr_dec = ThreeLayerDecoder(depth_ch)
        self.semseg_criterion = semseg_criterion
        self.discrepancy_criterion = _loss.Diff2d() if discrepancy_criterion is None else discrepancy_criterion

    def forward(self, x):
        return self.semsegcls_dec(x), self.deprgr_dec(x)

    def get_loss(self, x, gt_semseg, gt_dep, separately_returning=False):
        pred_semseg, pred_dep = self.forward(x)
        semseg_loss = _weighted(self.s_semsegcls, self.semseg_criterion(pred_semseg, gt_semseg))
        depreg_loss = _weighted(self.s_deprgr, _loss.mse_loss(pred_dep, gt_dep))
        if separately_returning:
            return semseg_loss, depreg_loss
        return semseg_loss + depreg_loss

    def get_task_weights(self):
        std_semseg = np.sqrt(np.exp(2 * self.s_semsegcls.data.cpu().numpy()))
        std_depth = np.sqrt(np.exp(2 * self.s_deprgr.data.cpu().numpy()))
        return std_semseg, std_depth


class TripleMultiTaskDecoder(_HedBranches):
    """source-only variant of MCDTripleMultiTaskDecoder: ONE segmentation classifier."""

    def __init__(self, n_class, depth_ch=3, semseg_criterion=None, semseg_shortcut=False, depth_shortcut=False,
                 add_pred_seg_boundary_loss=False, conv_seg2bd=False):
        super().__init__()
        self.s_semsegcls = _scalar_param()
        self.s_deprgr = _scalar_param()
        self.s_boundary = _scalar_param()
        self.semsegcls_dec = ThreeLayerDecoder(n_class)
        self.deprgr_dec = ThreeLayerDecoder(depth_ch)
        self.nmlrgr_dec = ThreeLayerDecoder(depth_ch)
        self.semseg_criterion = semseg_criterion
        self._init_branches(n_class, semseg_shortcut, depth_shortcut, add_pred_seg_boundary_loss, False,
                            n_semseg_heads=2, with_depth=True)

    def semseg_forward(self, x_dic):
        if self.semseg_shortcut:
            return self.semsegcls_dec(self._shortcut_sum(x_dic, "seg_conv%d_1"))
        return self.upsample3(self.semsegcls_dec(x_dic["h8"]))

    def depth_forward(self, x_dic):
        if self.depth_shortcut:
            return self.deprgr_dec(self._shortcut_sum(x_dic, "dep_conv%d"))
        return self.upsample3(self.deprgr_dec(x_dic["h8"]))

    def forward(self, x_dic):
        return self.semseg_forward(x_dic), self.depth_forward(x_dic), self.boundary_forward(x_dic)

    def get_semseg_loss(self, x_dic, gt_semseg):
        if not self.semseg_shortcut:
            fused = _fused_semseg_ce(self, (self.semsegcls_dec(x_dic["h8"]),), gt_semseg)
            if fused is not None:
                return fused[0]
        return self.semseg_criterion(self.semseg_forward(x_dic), gt_semseg)

    def get_depth_loss(self, x_dic, gt_dep):
        return _loss.mse_loss(self.depth_forward(x_dic), gt_dep)

    def get_boundary_loss(self, x_dic, gt_boundary):
        return _loss.sigmoid3_bce2d(*self._boundary_maps(x_dic), gt_boundary)

    def get_loss(self, x, gt_semseg, gt_dep, gt_boundary, separately_returning=False):
        semseg_loss = _weighted(self.s_semsegcls, self.get_semseg_loss(x, gt_semseg))
        depreg_loss = _weighted(self.s_deprgr, self.get_depth_loss(x, gt_dep))
        boundary_loss = _weighted(self.s_boundary, self.get_boundary_loss(x, gt_boundary))
        if separately_returning:
            return semseg_loss, depreg_loss, boundary_loss
        return semseg_loss + depreg_loss + boundary_loss

    def get_task_weights(self):
        std_semseg = np.sqrt(np.exp(2 * self.s_semsegcls.data.cpu().numpy()))
        std_depth = np.sqrt(np.exp(2 * self.s_deprgr.data.cpu().numpy()))
        return std_semseg, std_depth
