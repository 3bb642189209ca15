"""Segmentation wrappers over DRN for the MCD step, on libmcd_sm100 kernels.

Mirrors the hot classes of the reference's models/dilated_fcn.py with identical constructor signatures,
module names (=> state_dict keys) and call signatures:

  DRNSegBase (:217-250)                       generator G: DRN trunk + 1x1 `seg` conv -> [B,n_class,H/8,W/8]
  DRNSegPixelClassifier (:340-366)            head F: learned depthwise 16x16/s8 deconv -> [B,n_class,H,W]
  FusionDRNSegPixelClassifier (:431-470)      MFNet AddFusion head     up(x1 + x2)
  ScoreFusionDRNSegPixelClassifier (:473-491) MFNet ScoreAddFusion head up1(x1) + up2(x2)
  MultiTaskEncoder (:554-566), MultiTaskEncoderReturningMultipleFeaturemaps (:569-629)
  CBR (:632-644), ThreeLayerDecoder (:647-658)
  MCDMultiTaskDecoder (:661-739), MCDTripleMultiTaskDecoder (:790-1024)

  get_boundary_loss (:743-787)                morphological label-map boundary + bce2d

Score maps (n_class / depth / boundary channels at 1/8, 1/4, 1/2 resolution) are fp32 NCHW tensors; the full-resolution
predictions are fp32 NCHW by default (what the reference's testers call `.cpu().numpy()` on; bfloat16 inside MCDStep,
mcd_b200.nn.logits_dtype); trunk activations are IEEE-half channels_last with a bfloat16 twin while autograd records.

Everything else in the reference file (DRNSeg, ver2 heads, FuseDRNSegBase, domain classifiers, the
vendored fyu/drn CLI, shortcut / seg2bd options) is outside SURVEY.md section 8 and raises NotImplementedError.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.nn import Parameter

import loss as _loss
from mcd_b200 import ops
from mcd_b200.nn import (BatchNorm2d, BilinearUpsample, Conv2d, DepthwiseDeconv16s8, SoleChain, conv_bn_act)
from models import drn
from models.fusion import AddFusion, get_fusion_model


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
        raise NotImplementedError("%s: only drn_d_22 / _38 / _54 / _105 are built on libmcd_sm100" % model_name)
    return factory(pretrained=pretrained, num_classes=1000, input_ch=input_ch)


class DRNSegBase(nn.Module):
    def __init__(self, model_name, n_class, pretrained=True, input_ch=3, ver="ver1"):
        super().__init__()
        if ver != "ver1":
            raise NotImplementedError("ver2 heads are outside the libmcd_sm100 hot-path scope")
        model = _trunk(model_name, pretrained, input_ch)
        self.base = SoleChain(*list(model.children())[:-2])
        self.ver = ver
        self.seg = Conv2d(model.out_dim, n_class, kernel_size=1, bias=True, planar_out=True)
        _he_init(self.seg)

    def forward(self, x):
        return self.seg(self.base(x))

    def optim_parameters(self, memo=None):
        for param in self.base.parameters():
            yield param
        for param in self.seg.parameters():
            yield param


class DRNSegPixelClassifier(nn.Module):
    def __init__(self, n_class, use_torch_up=False, dropout=False, ver="ver1"):
        super().__init__()
        if ver != "ver1" or use_torch_up:
            raise NotImplementedError("ver2 / UpsamplingBilinear2d heads are outside the hot-path scope")
        self.dropout = dropout
        self.ver = ver
        self.up = DepthwiseDeconv16s8(n_class)

    def forward(self, x):
        return self.up(x)


class FusionDRNSegPixelClassifier(nn.Module):
    def __init__(self, fusion_type, n_class, use_torch_up=False, ver="ver1"):
        super().__init__()
        if ver != "ver1" or use_torch_up:
            raise NotImplementedError("ver2 / UpsamplingBilinear2d heads are outside the hot-path scope")
        self.fusion = get_fusion_model(fusion_type, n_class)
        self.ver = ver
        self.up = DepthwiseDeconv16s8(n_class)

    def forward(self, x1, x2):
        return self.up(self.fusion(x1, x2))


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


class MCDTripleMultiTaskDecoder(nn.Module):
    """semantic segmentation (two MCD classifiers) + HHA regression + boundary detection."""

    def __init__(self, n_class, depth_ch, semseg_criterion=None, discrepancy_criterion=None,
                 semseg_shortcut=False, depth_shortcut=False, add_pred_seg_boundary_loss=False,
                 use_seg2bd_conv=False):
        super().__init__()
        if semseg_shortcut or depth_shortcut or use_seg2bd_conv:
            raise NotImplementedError("shortcut / seg2bd options are default-off in the reference trainer and "
                                      "outside the libmcd_sm100 hot-path scope (SURVEY.md section 8f4)")
        self.s_semsegcls = _scalar_param()
        self.s_deprgr = _scalar_param()
        self.s_boundary = _scalar_param()
        self.semsegcls_dec1 = ThreeLayerDecoder(n_class)
        self.semsegcls_dec2 = ThreeLayerDecoder(n_class)
        self.deprgr_dec = ThreeLayerDecoder(depth_ch)
        self.nmlrgr_dec = ThreeLayerDecoder(depth_ch)  # constructed, never used (reference :813)
        self.semseg_criterion = semseg_criterion
        self.discrepancy_criterion = discrepancy_criterion
        self.upsample1 = BilinearUpsample(2)
        self.upsample2 = BilinearUpsample(4)
        self.upsample3 = BilinearUpsample(8)
        self.conv1 = Conv2d(32, 1, kernel_size=1, stride=1, padding=0, planar_out=True)
        self.conv2 = Conv2d(64, 1, kernel_size=1, stride=1, padding=0, planar_out=True)
        self.conv3 = Conv2d(512, 1, kernel_size=1, stride=1, padding=0, planar_out=True)
        self.semseg_shortcut = semseg_shortcut
        self.depth_shortcut = depth_shortcut
        self.add_pred_seg_boundary_loss = add_pred_seg_boundary_loss
        self.use_seg2bd_conv = use_seg2bd_conv

    def _semseg_scores(self, x_dic):
        h8 = x_dic["h8"]
        return self.semsegcls_dec1(h8), self.semsegcls_dec2(h8)

    def semseg_forward(self, x_dic):
        s1, s2 = self._semseg_scores(x_dic)
        return self.upsample3(s1), self.upsample3(s2)

    def depth_forward(self, x_dic):
        return self.upsample3(self.deprgr_dec(x_dic["h8"]))

    def _boundary_maps(self, x_dic):
        return (self.upsample1(self.conv1(x_dic["h2"])), self.upsample2(self.conv2(x_dic["h3"])),
                self.upsample3(self.conv3(x_dic["h8"])))

    def boundary_forward(self, x_dic):
        return _loss.sigmoid3_mean(*self._boundary_maps(x_dic))

    def forward(self, x_dic):
        pred_semseg1, pred_semseg2 = self.semseg_forward(x_dic)
        return pred_semseg1, pred_semseg2, self.depth_forward(x_dic), self.boundary_forward(x_dic)

    def get_cls_descrepancy(self, x_dic):
        scores = self._semseg_scores(x_dic)
        d = _fused_discrepancy(self, scores)
        if d is not None:
            return d
        return self.discrepancy_criterion(self.upsample3(scores[0]), self.upsample3(scores[1]))

    def get_semseg_loss(self, x_dic, gt_semseg, separately_returning=False):
        scores = self._semseg_scores(x_dic)
        fused = _fused_semseg_ce(self, scores, gt_semseg)
        if fused is not None:
            loss1, loss2 = fused
        else:
            loss1 = self.semseg_criterion(self.upsample3(scores[0]), gt_semseg)
            loss2 = self.semseg_criterion(self.upsample3(scores[1]), gt_semseg)
        return (loss1, loss2) if separately_returning else loss1 + loss2

    def get_depth_loss(self, x_dic, gt_dep):
        return _loss.mse_loss(self.depth_forward(x_dic), gt_dep)

    def get_boundary_loss(self, x_dic, gt_boundary):
        # sigmoid-average + bce2d fused: the averaged probability map is never materialised
        return _loss.sigmoid3_bce2d(*self._boundary_maps(x_dic), gt_boundary)

    def get_psuedo_boundary_loss(self, x_dic, separately_returning=False):
        """--add_pred_seg_boundary_loss (reference :990-1000): boundary of each classifier's argmax label map against
        the (detached) boundary head.  The reference passes a keyword `pred_semseg=` that get_boundary_loss does not
        have (TypeError); this is the evident intent: pred_type "semseg", gt_type "boundary"."""
        assert self.add_pred_seg_boundary_loss
        psuedo_boundary = self.boundary_forward(x_dic).detach().float()
        pred_semseg1, pred_semseg2 = self.semseg_forward(x_dic)
        loss1 = get_boundary_loss(util_predict(pred_semseg1), psuedo_boundary[:, 0], gt_type="boundary")
        loss2 = get_boundary_loss(util_predict(pred_semseg2), psuedo_boundary[:, 0], gt_type="boundary")
        return (loss1, loss2) if separately_returning else loss1 + loss2

    def get_weighted_semseg_loss(self, x, gt_semseg):
        """the `semseg_loss` term of get_loss() alone (phase B of adapt_triple_multitask_trainer.py:256-276 uses
        nothing else of get_loss's three results)."""
        l1, l2 = self.get_semseg_loss(x, gt_semseg, separately_returning=True)
        return (_weighted(self.s_semsegcls, l1) + _weighted(self.s_semsegcls, l2)) / 2

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
