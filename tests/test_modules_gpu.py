"""Module-level checks on the GPU: the reference-named modules run forward + backward through the C-ABI,
and the tcgen05 path agrees with the CUDA-core cross-check kernels end to end."""
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_early_fusion(dev, algo, seed=0, size=(64, 96)):
    from mcd_b200 import ops
    from loss import CrossEntropyLoss2d, Diff2d
    from models.model_util import get_models
    from util import get_class_weight_from_file
    prev = ops.set_conv_algo(algo)
    try:
        torch.manual_seed(seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g, f1, f2 = get_models("drn_d_38", 6, 41)
        g, f1, f2 = g.to(dev).train(), f1.to(dev).train(), f2.to(dev).train()
        gen = torch.Generator().manual_seed(seed + 1)
        x = torch.randn(2, 6, *size, generator=gen).to(dev)
        lbl = torch.randint(0, 41, (2, *size), generator=gen).to(dev)
        crit = CrossEntropyLoss2d(get_class_weight_from_file(41).to(dev))
        feat = g(x)
        o1, o2 = f1(feat), f2(feat)
        loss = crit(o1, lbl) + crit(o2, lbl) - Diff2d()(o1, o2)
        loss.backward()
        torch.cuda.synchronize()
        grads = {k: p.grad.clone() for k, p in g.named_parameters()}
        grads.update({"f1." + k: p.grad.clone() for k, p in f1.named_parameters()})
        return feat.detach(), o1.detach(), float(loss), grads
    finally:
        ops.set_conv_algo(prev)


def test_early_fusion_step_runs_and_umma_matches_direct(cuda_dev):
    from mcd_b200 import abi
    feat_d, o_d, loss_d, gr_d = _run_early_fusion(cuda_dev, abi.ALGO_DIRECT)
    assert feat_d.shape == (2, 41, 8, 12) and o_d.shape == (2, 41, 64, 96)
    assert o_d.dtype == torch.bfloat16 and torch.isfinite(feat_d).all()
    assert all(torch.isfinite(v).all() for v in gr_d.values())
    feat_u, o_u, loss_u, gr_u = _run_early_fusion(cuda_dev, abi.ALGO_AUTO)
    assert abs(loss_u - loss_d) / abs(loss_d) < 2e-3
    err = float((feat_u - feat_d).abs().max() / feat_d.abs().max())
    assert err < 3e-2, err
    bad = []
    for k in gr_d:
        e = float((gr_u[k] - gr_d[k]).abs().max() / (gr_d[k].abs().max() + 1e-12))
        if e > 6e-2:
            bad.append((k, e))
    assert not bad, bad[:10]
