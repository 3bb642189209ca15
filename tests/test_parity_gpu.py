"""Parity of the CUDA path (reference-named modules over libmcd_sm100, bf16 storage / fp32 accumulate)
against the fp32 oracle on identical weights and synthetic inputs (SURVEY.md section 8c protocol):

  per-layer activations and gradients   max|a-b| / max|b|  <= 2e-2
  losses                                 relative           <= 1e-3
  argmax label maps                      agreement          >= 99.5 %
  labels / ignore_index / argmax handling: integer paths bit-exact (tests/test_kernels_gpu.py)
"""
import os
import warnings

import pytest
import torch

from oracle import mcd_oracle as O

pytestmark = pytest.mark.gpu
N_CLASS = 41
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def nerr(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-20))


def _log(name, lines):
    if os.path.isdir(OUT):
        with open(os.path.join(OUT, name), "w") as f:
            f.write("\n".join(lines) + "\n")


def _models(dev, method="MCD"):
    from models.model_util import get_models
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ms = get_models("drn_d_38", 6, N_CLASS, method=method)
    return [m.to(dev) for m in ms]


def _load(module, sd):
    module.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)


def _inputs(seed, n, size, dev):
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(n, 6, *size, generator=g).to(dev)
    tgt = torch.randn(n, 6, *size, generator=g).to(dev)
    lbl = torch.randint(0, N_CLASS, (n, *size), generator=g).to(dev)
    return src, tgt, lbl


def _hook_units(model_g):
    """unit outputs keyed like the oracle's taps ('base.3.1:out', ...)."""
    from mcd_b200 import ops
    outs, handles = {}, []

    def add(mod, key):
        handles.append(mod.register_forward_hook(
            lambda m, i, o, key=key: outs.__setitem__(key, ops.to_nchw_f32(o.detach()))))

    for i, stage in enumerate(model_g.base):
        if i in (0, 1, 2, 7, 8):
            add(stage, "base.%d.0:out" % i)
        else:
            for b, blk in enumerate(stage):
                add(blk, "base.%d.%d:out" % (i, b))
    return outs, handles


@pytest.mark.parametrize("size,n", [((240, 320), 2), ((480, 640), 1)])
def test_early_fusion_forward_backward_vs_oracle(cuda_dev, size, n):
    from loss import CrossEntropyLoss2d, Diff2d
    dev = cuda_dev
    G = O.to_device(O.fill_state_dict_(O.init_seg_base("drn_d_38", 6, N_CLASS), 1), dev)
    F1 = O.to_device(O.fill_state_dict_(O.init_head(N_CLASS), 2), dev)
    F2 = O.to_device(O.fill_state_dict_(O.init_head(N_CLASS), 3), dev)
    mg, mf1, mf2 = _models(dev)
    _load(mg, G), _load(mf1, F1), _load(mf2, F2)
    mg.train(), mf1.train(), mf2.train()
    src, tgt, lbl = _inputs(5, n, size, dev)
    w = O.class_weight(N_CLASS).to(dev)

    # oracle: phase-B style objective touches every kernel: CE(src) - Diff2d(tgt)
    O._req([G, F1, F2])
    taps = {}
    feat_o = O.seg_base_forward(G, src, taps=taps)
    o1, o2 = O.head_forward(F1, feat_o), O.head_forward(F2, feat_o)
    ce_o = O.ce2d(o1, lbl, w) + O.ce2d(o2, lbl, w)
    ft_o = O.seg_base_forward(G, tgt)
    d_o = O.diff2d(O.head_forward(F1, ft_o), O.head_forward(F2, ft_o))
    gG, gF1, gF2 = O._grads(ce_o - d_o, [G, F1, F2])

    outs, handles = _hook_units(mg)
    feat = mg(src)
    for h in handles:
        h.remove()
    p1, p2 = mf1(feat), mf2(feat)
    crit = CrossEntropyLoss2d(w)
    ce = crit(p1, lbl) + crit(p2, lbl)
    ft = mg(tgt)
    d = Diff2d()(mf1(ft), mf2(ft))
    (ce - d).backward()
    torch.cuda.synchronize()

    lines, worst_act = [], 0.0
    for key in sorted(outs):
        e = nerr(outs[key], taps[key])
        worst_act = max(worst_act, e)
        lines.append("act  %-22s %.3e" % (key, e))
    e_feat, e_out = nerr(feat, feat_o), nerr(p1, o1)
    lines += ["feat %.3e" % e_feat, "out1 %.3e" % e_out,
              "ce %.6f vs %.6f" % (float(ce), float(ce_o)), "diff %.6e vs %.6e" % (float(d), float(d_o))]
    worst_grad, gl = 0.0, []
    for k, p in mg.named_parameters():
        e = nerr(p.grad, gG[k])
        worst_grad = max(worst_grad, e)
        gl.append("grad %-34s %.3e" % (k, e))
    e_up = nerr(mf1.up.weight.grad, gF1["up.weight"])
    lines += gl + ["grad f1.up.weight %.3e" % e_up]
    _log("parity_early_%dx%d.txt" % size, lines)

    assert worst_act <= 2e-2, "activation error %.3e" % worst_act
    assert e_feat <= 2e-2 and e_out <= 2e-2
    assert abs(float(ce) - float(ce_o)) / abs(float(ce_o)) <= 1e-3
    assert abs(float(d) - float(d_o)) / abs(float(d_o)) <= 1e-2   # |p1-p2| of near-identical heads
    assert worst_grad <= 2e-2, "gradient error %.3e" % worst_grad
    assert e_up <= 2e-2
    # running statistics took the same two momentum updates
    assert nerr(mg.base[8][1].running_var, G["base.8.1.running_var"]) < 2e-2
    assert int(mg.base[0][1].num_batches_tracked) == 2


def test_mcd_iteration_and_tester_vs_oracle(cuda_dev):
    """Full A / B / 4xC iteration (adapt_trainer.py:162-212) through the drop-in modules with torch.optim.SGD,
    then the tester path (adapt_tester.py:104-124)."""
    import util
    from loss import CrossEntropyLoss2d, get_prob_distance_criterion
    from models.model_util import get_optimizer
    dev, size, n = cuda_dev, (240, 320), 2
    G = O.to_device(O.fill_state_dict_(O.init_seg_base("drn_d_38", 6, N_CLASS), 1), dev)
    F1 = O.to_device(O.fill_state_dict_(O.init_head(N_CLASS), 2), dev)
    F2 = O.to_device(O.fill_state_dict_(O.init_head(N_CLASS), 3), dev)
    model_g, model_f1, model_f2 = _models(dev)
    _load(model_g, G), _load(model_f1, F1), _load(model_f2, F2)
    src_imgs, tgt_imgs, src_lbls = _inputs(9, n, size, dev)
    w = O.class_weight(N_CLASS).to(dev)
    rec = {}
    c_o, d_o = O.mcd_step_early(G, F1, F2, src_imgs, src_lbls, tgt_imgs, w, O.SGD(), O.SGD(), num_k=4,
                                record=rec)

    optimizer_g = get_optimizer(model_g.parameters(), lr=1e-3, momentum=0.9, opt="sgd", weight_decay=2e-5)
    optimizer_f = get_optimizer(list(model_f1.parameters()) + list(model_f2.parameters()), opt="sgd", lr=1e-3,
                                momentum=0.9, weight_decay=2e-5)
    criterion = CrossEntropyLoss2d(w)
    criterion_d = get_prob_distance_criterion("diff")
    model_g.train(), model_f1.train(), model_f2.train()
    # ---- the reference loop body, verbatim modulo python-3 spellings
    optimizer_g.zero_grad(), optimizer_f.zero_grad()
    outputs = model_g(src_imgs)
    outputs1, outputs2 = model_f1(outputs), model_f2(outputs)
    loss = criterion(outputs1, src_lbls) + criterion(outputs2, src_lbls)
    loss.backward()
    c_loss = float(loss)
    optimizer_g.step(), optimizer_f.step()
    optimizer_g.zero_grad(), optimizer_f.zero_grad()
    outputs = model_g(src_imgs)
    outputs1, outputs2 = model_f1(outputs), model_f2(outputs)
    loss = criterion(outputs1, src_lbls) + criterion(outputs2, src_lbls)
    outputs = model_g(tgt_imgs)
    outputs1, outputs2 = model_f1(outputs), model_f2(outputs)
    loss = loss - criterion_d(outputs1, outputs2)
    loss.backward()
    b_loss = float(loss)
    optimizer_f.step()
    c_losses = []
    for i in range(4):
        optimizer_g.zero_grad()
        outputs = model_g(tgt_imgs)
        outputs1, outputs2 = model_f1(outputs), model_f2(outputs)
        loss = criterion_d(outputs1, outputs2) * 1.0
        loss.backward()
        c_losses.append(float(loss))
        optimizer_g.step()
    torch.cuda.synchronize()

    lines = ["A %.6f vs %.6f" % (c_loss, c_o), "B %.6f vs %.6f" % (b_loss, float(rec["B_loss"]))]
    lines += ["C%d %.6e vs %.6e" % (i, a, b) for i, (a, b) in enumerate(zip(c_losses, rec["C_losses"]))]
    werr = {k: nerr(p.detach() - 0, G[k]) for k, p in model_g.named_parameters()}
    lines += ["w %-34s %.3e" % kv for kv in sorted(werr.items())]
    _log("parity_iteration.txt", lines)
    assert abs(c_loss - c_o) / abs(c_o) <= 1e-3
    assert abs(b_loss - float(rec["B_loss"])) / abs(float(rec["B_loss"])) <= 1e-3
    for a, b in zip(c_losses, rec["C_losses"]):
        assert abs(a - b) / abs(b) <= 2e-2
    assert max(werr.values()) <= 1e-3          # weights after 5 G-steps at lr 1e-3
    assert nerr(model_f1.up.weight.detach(), F1["up.weight"]) <= 1e-3

    # ---- tester: eval-mode forward, argmax over the first n_class-1 channels, entropy
    model_g.eval(), model_f1.eval()
    with torch.no_grad():
        out = model_f1(model_g(tgt_imgs[:1]))
        ref = O.head_forward(F1, O.seg_base_forward(G, tgt_imgs[:1], train=False))
    pred = util.predict_labels(out, N_CLASS - 1)
    agree = float((pred == O.predict_labels(ref, N_CLASS - 1)).float().mean())
    ent, ent_o = float(util.calc_entropy(out)), float(O.calc_entropy(ref))
    _log("parity_tester.txt", ["argmax agreement %.5f" % agree, "entropy %.6e vs %.6e" % (ent, ent_o)])
    assert pred.dtype == torch.int64 and agree >= 0.995
    assert abs(ent - ent_o) / abs(ent_o) <= 1e-2
