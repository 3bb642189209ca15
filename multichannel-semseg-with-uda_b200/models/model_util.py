"""Model / optimiser factories of the MCD step (drop-in for the reference's models/model_util.py).

Same function names, argument meaning and error behaviour as the reference (file:line below) for the DRN
branches that are on the hot path; every other `net_name` raises NotImplementedError like the reference does
for unknown names.  `is_data_parallel=True` wraps the models in mcd_b200.parallel.DataParallel: the reference's
single-process nn.DataParallel (models/model_util.py:283-284) becomes one process per GPU with NCCL gradient
all-reduce, behind the same `.module` / `module.`-prefixed state_dict interface.
"""
import torch
from torch import nn
from torch.nn.modules.batchnorm import _BatchNorm

_DRN_NAMES = ("drn_c_26", "drn_c_42", "drn_c_58", "drn_d_22", "drn_d_38", "drn_d_54", "drn_d_105")


def _check_drn(net_name):
    if "drn" not in net_name:
        raise NotImplementedError("Only FCN (Including Dilated FCN), SegNet, PSPNet UNet are supported!")
    if net_name not in _DRN_NAMES:
        raise NotImplementedError("libmcd_sm100 builds %s only (got %s)" % (", ".join(_DRN_NAMES), net_name))


def _data_parallel(models, flag):
    """`is_data_parallel=True` (reference models/model_util.py:75-76,96-97,283-284 wraps every model in
    torch.nn.DataParallel): one process per GPU here, see mcd_b200.parallel.DataParallel - same `.module` attribute
    and `module.`-prefixed checkpoints, gradient all-reduce over the ranks of a torchrun job."""
    if not flag:
        return models
    from mcd_b200.parallel import DataParallel
    return [DataParallel(m) for m in models]


def get_models(net_name, input_ch, n_class, res="50", method="MCD", is_data_parallel=False):
    """(model_g, model_f1, model_f2) for method "MCD" (`drn_*`, `drn_*_ver2`, `drn_*_fusenet`);
    (model_g_3ch, model_g_1ch, model_f1, model_f2) for "MCD-MFNet-<Fusion>" / "MCD-MFNet-Score<Fusion>" with any
    fusion of models/fusion.py (reference models/model_util.py:160-286)."""
    from models.dilated_fcn import (DRNSegBase, DRNSegPixelClassifier, FuseDRNSegBase, FusionDRNSegPixelClassifier,
                                    ScoreFusionDRNSegPixelClassifier)
    if method == "MCD":
        if "drn" not in net_name:
            raise NotImplementedError("Only FCN (Including Dilated FCN), SegNet, PSPNet UNet are supported!")
        if "fusenet" in net_name:
            drn_name = net_name.replace("_fusenet", "")
            _check_drn(drn_name)
            model_list = [FuseDRNSegBase(model_name=drn_name, n_class=n_class, input_ch=input_ch),
                          DRNSegPixelClassifier(n_class=n_class), DRNSegPixelClassifier(n_class=n_class)]
        else:
            ver = "ver2" if "ver2" in net_name else "ver1"
            drn_name = net_name.replace("_ver2", "")
            _check_drn(drn_name)
            model_list = [DRNSegBase(model_name=drn_name, n_class=n_class, input_ch=input_ch, ver=ver),
                          DRNSegPixelClassifier(n_class=n_class, ver=ver),
                          DRNSegPixelClassifier(n_class=n_class, ver=ver)]
    elif "MFNet" in method:
        assert input_ch in [4, 6]
        if "drn" not in net_name:
            raise NotImplementedError("Only Dilated FCN is supported!")
        ver = "ver2" if "ver2" in net_name else "ver1"
        drn_name = net_name.replace("_ver2", "")
        _check_drn(drn_name)
        fusion_type = method.split("-")[-1]
        print("fusion type: %s" % fusion_type)
        model_list = [DRNSegBase(model_name=drn_name, n_class=n_class, input_ch=3, ver=ver),
                      DRNSegBase(model_name=drn_name, n_class=n_class, input_ch=input_ch - 3, ver=ver)]
        if "score" in method.lower():
            print("Score Fusion!!!")
            model_list += [ScoreFusionDRNSegPixelClassifier(fusion_type=fusion_type, n_class=n_class)
                           for _ in range(2)]
        else:
            model_list += [FusionDRNSegPixelClassifier(fusion_type=fusion_type, n_class=n_class, ver=ver)
                           for _ in range(2)]
    else:
        # the reference *returns* (does not raise) the exception object here (models/model_util.py:281)
        return NotImplementedError("Sorry... Only MCD is supported!")
    return _data_parallel(model_list, is_data_parallel)


def get_multitask_models(net_name, input_ch, n_class, semseg_criterion=None, discrepancy_criterion=None,
                         is_data_parallel=False, is_src_only=False):
    """(model_enc, model_dec): RGB encoder + seg/HHA decoder - the MCD pair of classifiers, or one classifier with
    `is_src_only` (reference models/model_util.py:81-99)."""
    from models.dilated_fcn import MCDMultiTaskDecoder, MultiTaskDecoder, MultiTaskEncoder
    _check_drn(net_name)
    model_enc = MultiTaskEncoder(model_name=net_name, input_ch=3)  # RGB is 3 channel
    dec = MultiTaskDecoder if is_src_only else MCDMultiTaskDecoder
    model_dec = dec(n_class=n_class, depth_ch=input_ch - 3, semseg_criterion=semseg_criterion,
                    discrepancy_criterion=discrepancy_criterion)
    return tuple(_data_parallel([model_enc, model_dec], is_data_parallel))


def get_triple_multitask_models(net_name, input_ch, n_class, semseg_criterion=None, discrepancy_criterion=None,
                                is_data_parallel=False, semseg_shortcut=False, depth_shortcut=False,
                                add_pred_seg_boundary_loss=False, is_src_only=False, use_seg2bd_conv=False):
    """(model_enc, model_dec): RGB encoder returning h0..h8 + seg/HHA/boundary decoder; `input_ch` is
    ignored exactly as in the reference (models/model_util.py:102-130)."""
    from models.dilated_fcn import (MCDTripleMultiTaskDecoder, MultiTaskEncoderReturningMultipleFeaturemaps,
                                    TripleMultiTaskDecoder)
    _check_drn(net_name)
    model_enc = MultiTaskEncoderReturningMultipleFeaturemaps(model_name=net_name, input_ch=3)
    if is_src_only:
        model_dec = TripleMultiTaskDecoder(n_class=n_class, depth_ch=3, semseg_criterion=semseg_criterion,
                                           semseg_shortcut=semseg_shortcut, depth_shortcut=depth_shortcut,
                                           add_pred_seg_boundary_loss=add_pred_seg_boundary_loss)
    else:
        model_dec = MCDTripleMultiTaskDecoder(n_class=n_class, depth_ch=3, semseg_criterion=semseg_criterion,
                                              discrepancy_criterion=discrepancy_criterion,
                                              semseg_shortcut=semseg_shortcut, depth_shortcut=depth_shortcut,
                                              add_pred_seg_boundary_loss=add_pred_seg_boundary_loss,
                                              use_seg2bd_conv=use_seg2bd_conv)
    return tuple(_data_parallel([model_enc, model_dec], is_data_parallel))


def get_segbd_multitask_models(net_name, input_ch, n_class, semseg_criterion=None, discrepancy_criterion=None,
                               is_data_parallel=False, semseg_shortcut=False, depth_shortcut=False,
                               add_pred_seg_boundary_loss=False, is_src_only=False, use_seg2bd_conv=False):
    """(model_enc, model_dec): RGB encoder returning h0..h8 + seg/boundary MCD decoder (reference
    models/model_util.py:133-157); `depth_shortcut` is accepted and not forwarded, as there."""
    from models.dilated_fcn import MCDSegBDMultiTaskDecoder, MultiTaskEncoderReturningMultipleFeaturemaps
    _check_drn(net_name)
    model_enc = MultiTaskEncoderReturningMultipleFeaturemaps(model_name=net_name, input_ch=3)
    if is_src_only:
        raise NotImplementedError()
    model_dec = MCDSegBDMultiTaskDecoder(n_class=n_class, depth_ch=3, semseg_criterion=semseg_criterion,
                                         discrepancy_criterion=discrepancy_criterion,
                                         semseg_shortcut=semseg_shortcut,
                                         add_pred_seg_boundary_loss=add_pred_seg_boundary_loss,
                                         use_seg2bd_conv=use_seg2bd_conv)
    return tuple(_data_parallel([model_enc, model_dec], is_data_parallel))


def get_optimizer(model_parameters, opt, lr, momentum, weight_decay):
    """torch.optim stays the optimiser (reference models/model_util.py:289-302)."""
    params = filter(lambda p: p.requires_grad, model_parameters)
    if opt == "sgd":
        return torch.optim.SGD(params, lr=lr, momentum=momentum, weight_decay=weight_decay)
    if opt == "adadelta":
        return torch.optim.Adadelta(params, lr=lr, weight_decay=weight_decay)
    if opt == "adam":
        return torch.optim.Adam(params, lr=lr, betas=[0.5, 0.999], weight_decay=weight_decay)
    raise NotImplementedError("Only (Momentum) SGD, Adadelta, Adam are supported!")


def fix_batchnorm_when_training(model):
    """--fix_bn: BatchNorm layers keep using their running statistics (reference :305-310)."""
    if issubclass(type(model), _BatchNorm):
        model.training = False
    for module in model.children():
        fix_batchnorm_when_training(module)


def fix_dropout_when_training(model):
    if type(model) in [nn.Dropout, nn.Dropout2d, nn.Dropout3d, nn.AlphaDropout]:
        model.training = False
    for module in model.children():
        fix_dropout_when_training(module)
