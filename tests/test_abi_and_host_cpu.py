"""CPU-side checks (no GPU, no compute calls): the C-ABI library loads and exports every symbol that
include/mcd_sm100.h declares, the ctypes binding covers them, the host-side mirror of the reference interface
keeps the reference's names / parameter counts / state_dict keys, and the product path refuses to run on CPU."""
import ctypes
import os
import re
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcd_sm100.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mcd_b200 import abi
    names = _declared()
    assert len(names) >= 30
    handle = ctypes.CDLL(abi.LIB_PATH)
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing
    # the binding covers the header and nothing else
    assert sorted(abi.EXPORTS) == names
    lib = abi.lib()
    assert lib.mcd_version() == abi.ABI_VERSION
    m = re.search(r"#define MCD_ABI_VERSION (\d+)", open(HEADER).read())
    assert int(m.group(1)) == abi.ABI_VERSION
    assert lib.mcd_launch_count() == 0


def test_kernel_dispatch_is_host_logic():
    """mcd_conv2d_kernel_id / mcd_conv2d_wgrad_partials are pure host code: which kernel each DRN-D-38 layer gets
    (kind * 1000 + tile width, include/mcd_sm100.h) at 22 images of 480x640."""
    from mcd_b200 import abi, ops
    lib = abi.lib()

    def kid(g, p):
        return lib.mcd_conv2d_kernel_id(ctypes.byref(g), p, abi.OUT_NHWC_BF16, abi.ALGO_AUTO)

    g0 = ops.conv_geom((22, 8, 480, 640), 6, 16, 7, 7, 1, 1, 3)          # layer0
    g1 = ops.conv_geom((22, 16, 480, 640), 16, 16, 3, 3, 1, 1, 1)        # layer1
    g3 = ops.conv_geom((22, 64, 120, 160), 64, 64, 3, 3, 1, 1, 1)        # layer3
    g4 = ops.conv_geom((22, 128, 60, 80), 128, 128, 3, 3, 1, 1, 1)       # layer4
    g5 = ops.conv_geom((22, 256, 60, 80), 256, 256, 3, 3, 1, 2, 2)       # layer5 (dilation 2)
    g6 = ops.conv_geom((22, 512, 60, 80), 512, 512, 3, 3, 1, 4, 4)       # layer6 (dilation 4)
    gd = ops.conv_geom((22, 256, 60, 80), 256, 512, 1, 1, 1, 1, 0)       # 1x1 downsample: one tap, no halo
    assert [kid(g, 0) for g in (g0, g1)] == [3016, 3016]                 # row convolution
    assert [kid(g, 0) for g in (g3, g4)] == [6064, 6128]                 # halo-tile staging, one CTA per tile
    assert [kid(g, 0) for g in (g5, g6)] == [7256, 7256]                 # CTA pairs + halo
    assert [kid(g, 1) for g in (g5, g6)] == [7256, 7256]                 # dgrad is the same kernel
    assert kid(gd, 0) == 1256                                            # CTA pairs, per-tap staging
    assert [kid(g, 2) for g in (g0, g1)] == [8016, 8016]                 # Toeplitz stem wgrad
    assert [kid(g, 2) for g in (g3, g4)] == [4064, 4128]
    assert [kid(g, 2) for g in (g5, g6)] == [10256, 10256]               # CTA-pair wgrad
    assert kid(g6, 0) != lib.mcd_conv2d_kernel_id(ctypes.byref(g6), 0, abi.OUT_NHWC_BF16, abi.ALGO_DIRECT)
    # split partial-sum layout [ksplit][T][CoutP][CinP] consumed by mcd_sgd_pack_multi
    assert ops.wgrad_partial_layout(g6) == (2, 9, 512, 512)
    assert ops.wgrad_partial_layout(g5) == (8, 9, 256, 256)
    assert ops.wgrad_partial_layout(g0) is None and ops.wgrad_partial_layout(g1) is None
    ks, t, coutp, cinp = ops.wgrad_partial_layout(g6)
    assert lib.mcd_conv2d_wgrad_workspace(ctypes.byref(g6), abi.ALGO_AUTO) == 4 * ks * t * coutp * cinp


def test_conv_geom_struct_matches_header():
    from mcd_b200.abi import ConvGeom
    fields = re.search(r"typedef struct mcd_conv_geom \{(.*?)\} mcd_conv_geom;", open(HEADER).read(), re.S).group(1)
    fields = re.sub(r"/\*.*?\*/", "", fields, flags=re.S)
    names = [n.strip() for decl in re.findall(r"int32_t ([^;]+);", fields) for n in decl.split(",")]
    assert names == [n for n, _ in ConvGeom._fields_]
    assert ctypes.sizeof(ConvGeom) == 4 * len(names)


def test_pack_kind_is_host_logic():
    """mcd_conv2d_pack_kind is pure host code: thin stem layers get the row-convolution (2) or row-packed (1) layout."""
    from mcd_b200 import abi, ops
    lib = abi.lib()
    g0 = ops.conv_geom((4, 8, 480, 640), 6, 16, 7, 7, 1, 1, 3)          # layer0
    g1 = ops.conv_geom((4, 16, 480, 640), 16, 16, 3, 3, 1, 1, 1)        # layer1
    g2 = ops.conv_geom((4, 16, 480, 640), 16, 32, 3, 3, 2, 1, 1)        # layer2 (stride 2)
    g6 = ops.conv_geom((4, 512, 60, 80), 512, 512, 3, 3, 1, 4, 4)       # layer6 (dilated)
    assert [lib.mcd_conv2d_pack_kind(ctypes.byref(g), 0, abi.ALGO_AUTO) for g in (g0, g1, g2, g6)] == [2, 2, 1, 0]
    assert [lib.mcd_conv2d_pack_kind(ctypes.byref(g), 1, abi.ALGO_AUTO) for g in (g0, g1, g2, g6)] == [2, 2, 0, 0]
    assert lib.mcd_conv2d_pack_kind(ctypes.byref(g0), 0, abi.ALGO_DIRECT) == 0
    assert (g2.Ho, g2.Wo) == (240, 320) and (g6.Ho, g6.Wo) == (60, 80)
    assert lib.mcd_conv2d_wgrad_workspace(ctypes.byref(g6), abi.ALGO_AUTO) > 0
    assert lib.mcd_conv2d_wgrad_workspace(ctypes.byref(g6), abi.ALGO_DIRECT) == 0


def test_factories_keep_reference_surface():
    from oracle import mcd_oracle as O
    from models.model_util import (fix_batchnorm_when_training, get_models, get_multitask_models,
                                   get_optimizer, get_triple_multitask_models)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g, f1, f2 = get_models("drn_d_38", 6, 41)
        mf = get_models("drn_d_38", 6, 41, method="MCD-MFNet-ScoreAddFusion")
        ma = get_models("drn_d_38", 6, 41, method="MCD-MFNet-AddFusion")
        enc, dec = get_triple_multitask_models("drn_d_38", 6, 41)
        enc2, dec2 = get_multitask_models("drn_d_38", 6, 41)
    count = lambda m: sum(p.numel() for p in m.parameters())
    # parameter counts measured on the reference
    assert count(g) == 26012297 and count(f1) == 10496 and count(mf[2]) == 20992
    assert count(mf[0]) == count(mf[1]) == 26009945
    assert count(enc) == 25988912 and count(dec) == 10543806 and count(dec2) == 7917143
    assert [type(m).__name__ for m in ma] == ["DRNSegBase", "DRNSegBase", "FusionDRNSegPixelClassifier",
                                              "FusionDRNSegPixelClassifier"]
    # state_dict keys identical to the oracle's (= the reference's, tests/test_oracle_golden.py)
    assert sorted(g.state_dict()) == sorted(O.init_seg_base("drn_d_38", 6, 41))
    assert sorted(f1.state_dict()) == ["up.weight"] and sorted(mf[2].state_dict()) == ["up1.weight", "up2.weight"]
    assert sorted(enc.state_dict()) == sorted(O.init_trunk("drn_d_38", 3, "main_layer"))
    dec_keys = set(dec.state_dict())
    assert set(O.init_triple_decoder(41, 3)) <= dec_keys
    # error behaviour of the reference
    assert isinstance(get_models("drn_d_38", 6, 41, method="nope"), NotImplementedError)
    with pytest.raises(NotImplementedError):
        get_models("unet", 6, 41)
    with pytest.raises(NotImplementedError):
        get_optimizer(g.parameters(), "lbfgs", 1e-3, 0.9, 2e-5)
    with pytest.raises(AssertionError):
        get_models("drn_d_38", 5, 41, method="MCD-MFNet-AddFusion")
    opt = get_optimizer(g.parameters(), "sgd", 1e-3, 0.9, 2e-5)
    assert opt.defaults["momentum"] == 0.9 and opt.defaults["weight_decay"] == 2e-5
    g.train()
    fix_batchnorm_when_training(g)
    assert not g.base[3][0].bn1.training and g.base[3][0].conv1.training
    # 6-channel stem: channels 3..5 duplicate the RGB filters (models/drn.py:285-288)
    w0 = g.base[0][0].weight
    assert torch.equal(w0[:, :3], w0[:, 3:])


def test_product_path_has_no_cpu_fallback():
    import loss
    from mcd_b200 import abi
    from models.model_util import get_models
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g, f1, _ = get_models("drn_d_38", 6, 41)
    with pytest.raises(abi.McdError):
        g(torch.randn(1, 6, 32, 32))
    with pytest.raises(abi.McdError):
        f1(torch.randn(1, 41, 4, 4))
    with pytest.raises(abi.McdError):
        loss.Diff2d()(torch.randn(1, 41, 8, 8), torch.randn(1, 41, 8, 8))
    with pytest.raises(abi.McdError):
        loss.CrossEntropyLoss2d()(torch.randn(1, 41, 8, 8), torch.zeros(1, 8, 8, dtype=torch.long))


def test_class_weights_and_lr_schedule():
    import util
    w = util.get_class_weight_from_file(41)
    assert w.shape == (41,) and float(w[40]) == 0.0 and float(w[:40].min()) == 1.0
    assert float(util.get_class_weight_from_file(41, add_bg_loss=True)[40]) == 1.0
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    assert util.adjust_learning_rate(opt, 1e-3, 0.1, 0, 10) == 1e-3
    assert abs(util.adjust_learning_rate(opt, 1e-3, 0.1, 5, 10) - 1e-4) < 1e-12
    assert abs(util.adjust_learning_rate(opt, 1e-3, 0.1, 8, 10) - 1e-5) < 1e-12
    assert abs(opt.param_groups[0]["lr"] - 1e-5) < 1e-12


def test_criteria_reject_mismatched_label_maps():
    """torch's nll_loss raises when the label map does not have the logits' batch / spatial sizes (e.g. the source-only
    MultiTaskDecoder of the reference, whose predictions stay at 1/8 resolution); the kernels index the labels with the
    logits' geometry, so the check has to happen on the host - before anything touches the device."""
    import torch
    from loss import CrossEntropyLoss2d, ProbCrossEntropyLoss2d
    from mcd_b200 import headloss
    logits = torch.zeros(2, 41, 16, 24)
    for crit in (CrossEntropyLoss2d(), ProbCrossEntropyLoss2d()):
        for bad in (torch.zeros(2, 2, 3, dtype=torch.int64), torch.zeros(1, 16, 24, dtype=torch.int64),
                    torch.zeros(2, 128, 192, dtype=torch.int64)):
            with pytest.raises(ValueError, match="sizes don't match"):
                crit(logits, bad)
    with pytest.raises(ValueError, match="sizes don't match"):
        headloss.head_ce2d([torch.zeros(2, 41, 2, 3)], None, torch.zeros(2, 2, 3, dtype=torch.int64))
    with pytest.raises(ValueError, match="sizes don't match"):
        headloss.head_ce2d_pair([torch.zeros(2, 41, 2, 3)], [torch.zeros(41, 1, 16, 16)], [torch.zeros(41, 1, 16, 16)],
                                torch.zeros(2, 16, 25, dtype=torch.int64))
