"""Parity of the CUDA path (reference-named modules over libmcd_sm100: IEEE-half forward storage, bfloat16 gradient
storage, fp32 accumulate) against the fp32 oracle on identical weights and synthetic inputs.  Protocol (SURVEY.md
section 8c) and what is asserted here; "max-norm" = max|a-b| / max|b|, "rel-L2" = |a-b|_2 / |b|_2:

  per-layer activations                  max-norm <= 2e-3 against fp32 (north_star bound 2e-2; measured 8.6e-4); every
                                         DRN unit of the real network is fed the ORACLE's input / upstream gradient, so
                                         the number measures that layer's kernels only
  per-layer gradients                    rel-L2 <= 2e-2 against the oracle run with the SAME storage formats (kernel
                                         correctness: measured <= 1.4e-2), and against fp32: dx <= 3e-2, parameter
                                         gradients <= 6e-2, with >= 99 % of all gradient elements within 2e-2 of max|ref|
  losses (CE, Diff2d, phases A/B/C, MFNet, multitask)   relative <= 1e-3 against fp32
  updated weights after one iteration    max-norm <= 1e-3
  argmax label maps (testers)            agreement >= 99.5 % against fp32 (measured 99.9 %)
  integer label / ignore_index / argmax handling: bit-exact (tests/test_kernels_gpu.py)

Why gradients are not within 2e-2 of fp32 element for element: the backward pass multiplies by the ReLU mask of the
forward pass.  A pre-activation within rounding distance of zero has a different sign in any two implementations that
round differently - a 100 % error on that element - and the rel-L2 error of a unit's gradients is 1.5 * sqrt(flip
rate) (tests/tools/precision_study.py, tests/tools/debug_unit_real.py).  16-bit tensor-core operands bound the flip
rate from below: 1e-4 with IEEE half (rel-L2 1.5e-2 ... 3e-2; what this library stores), 8e-4 with bfloat16
(4e-2 ... 1e-1; round 1).  tcgen05 has no wider operand format at full rate, so the remaining distance to fp32 is the
hardware's, not the kernels': against the oracle run with the same storage formats every unit agrees to <= 1.4e-2.
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
    a, b = a.detach().float(), b.detach().float()
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
    module.load_state_dict({k: v.detach().clone() for k, v in sd.items()}, strict=True)


def _clone(sd):
    return {k: v.detach().clone() for k, v in sd.items()}


def _inputs(seed, n, size, dev):
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(n, 6, *size, generator=g).to(dev)
    tgt = torch.randn(n, 6, *size, generator=g).to(dev)
    lbl = torch.randint(0, N_CLASS, (n, *size), generator=g).to(dev)
    return src, tgt, lbl


def _state(dev, seed=1):
    G = O.to_device(O.fill_state_dict_(O.init_seg_base("drn_d_38", 6, N_CLASS), seed), dev)
    F1 = O.to_device(O.fill_state_dict_(O.init_head(N_CLASS), seed + 1), dev)
    F2 = O.to_device(O.fill_state_dict_(O.init_head(N_CLASS), seed + 2), dev)
    return G, F1, F2


def _units(model_g):
    """(oracle tap key, module, oracle parameter prefix) for every DRN unit in forward order."""
    units = []
    for i, stage in enumerate(model_g.base):
        if i in (0, 1, 2, 7, 8):
            units.append(("base.%d.0:out" % i, stage, "base.%d." % i))
        else:
            for b, blk in enumerate(stage):
                units.append(("base.%d.%d:out" % (i, b), blk, "base.%d.%d." % (i, b)))
    return units


@pytest.mark.parametrize("size,n", [((240, 320), 2), ((480, 640), 1)])
def test_per_layer_forward_backward_vs_oracle(cuda_dev, size, n):
    from mcd_b200 import ops
    dev = cuda_dev
    G, F1, F2 = _state(dev)
    mg, mf1, mf2 = _models(dev)
    _load(mg, G), _load(mf1, F1), _load(mf2, F2)
    mg.train(), mf1.train(), mf2.train()
    src, tgt, lbl = _inputs(5, n, size, dev)
    w = O.class_weight(N_CLASS).to(dev)

    # fp32 oracle: forward, CE loss, gradients w.r.t. every unit output and every parameter
    Go, F1o, F2o = _clone(G), _clone(F1), _clone(F2)
    O._req([Go, F1o, F2o])
    taps = {}
    feat_o = O.seg_base_forward(Go, src, taps=taps)
    o1, o2 = O.head_forward(F1o, feat_o), O.head_forward(F2o, feat_o)
    loss_o = O.ce2d(o1, lbl, w) + O.ce2d(o2, lbl, w)
    keys = [k for k in taps if k.endswith(":out")]
    pnames = O.trainable(Go)
    grads = torch.autograd.grad(loss_o, [taps[k] for k in keys] + [feat_o, o1] + [Go[k] for k in pnames] +
                                [F1o["up.weight"]])
    d_out = dict(zip(keys, grads[:len(keys)]))
    d_feat, d_o1 = grads[len(keys)], grads[len(keys) + 1]
    gG = dict(zip(pnames, grads[len(keys) + 2:len(keys) + 2 + len(pnames)]))
    g_up = grads[-1]

    def l2err(a, b):
        a, b = a.detach().float(), b.detach().float()
        return float((a - b).norm() / (b.norm() + 1e-30))

    spec_units = [u for stage in O.trunk_spec("drn_d_38", "base.") for u in stage]
    lines = ["# unit | act max-norm vs fp32 | dx, worst param-grad rel-L2 vs the SAME-STORAGE oracle unit | "
             "dx, worst param-grad rel-L2 vs fp32 oracle | fraction of gradient elements off by > 2e-2 of max|ref|"]
    worst = [0.0, 0.0, 0.0, 0.0, 0.0]
    n_far, n_all = 0, 0

    def far(a, b):
        a, b = a.detach().float(), b.detach().float()
        return int(((a - b).abs() > 2e-2 * b.abs().max()).sum()), a.numel()
    x_in, dx_ref_key = src, None
    for (key, mod, prefix), unit in zip(_units(mg), spec_units):
        assert O.unit_key(unit) == key
        first = dx_ref_key is None
        # bf16-storage emulation of this unit on the same input / upstream gradient
        sd_u = {k: v.detach().clone().requires_grad_(k in gG) for k, v in Go.items() if k.startswith(prefix)}
        xe = x_in.detach().clone().requires_grad_(not first)
        with O.storage(torch.float16, grad=torch.bfloat16):
            oe = O.unit_forward(sd_u, unit, O._q(xe), True)
        pk = [k for k in sd_u if sd_u[k].requires_grad]
        ge = torch.autograd.grad(oe, ([] if first else [xe]) + [sd_u[k] for k in pk], d_out[key])
        ge_x = None if first else ge[0]
        ge_p = dict(zip(pk, ge[0 if first else 1:]))
        # CUDA unit
        xin = ops.to_nhwc(x_in.detach()).requires_grad_(not first)
        mod.zero_grad()
        out = mod(xin)
        out.backward(ops.to_nhwc(d_out[key], grad=True))
        e_act = nerr(ops.to_nchw_f32(out), taps[key])
        e_dx = 0.0 if first else l2err(ops.to_nchw_f32(xin.grad), ge_x)
        l_dx = 0.0 if first else l2err(ops.to_nchw_f32(xin.grad), d_out[dx_ref_key])
        e_p = max(l2err(p.grad, ge_p[prefix + name]) for name, p in mod.named_parameters())
        l_p = max(l2err(p.grad, gG[prefix + name]) for name, p in mod.named_parameters())
        fa = [far(p.grad, gG[prefix + name]) for name, p in mod.named_parameters()]
        if not first:
            fa.append(far(ops.to_nchw_f32(xin.grad), d_out[dx_ref_key]))
        n_far, n_all = n_far + sum(f[0] for f in fa), n_all + sum(f[1] for f in fa)
        lines.append("%-16s %.3e | %.3e %.3e | %.3e %.3e | %.2e" % (key, e_act, e_dx, e_p, l_dx, l_p,
                                                                   sum(f[0] for f in fa) / sum(f[1] for f in fa)))
        worst = [max(a, b) for a, b in zip(worst, (e_act, e_dx, e_p, l_dx, l_p))]
        x_in, dx_ref_key = taps[key], key
    # seg conv, head and loss, each on the oracle's input
    h8 = ops.to_nhwc(taps[keys[-1]].detach()).requires_grad_(True)
    mg.seg.zero_grad()
    f = mg.seg(h8)
    f.backward(d_feat)
    e_seg = (nerr(f, feat_o), nerr(ops.to_nchw_f32(h8.grad), d_out[keys[-1]]),
             max(nerr(mg.seg.weight.grad, gG["seg.weight"]), nerr(mg.seg.bias.grad, gG["seg.bias"])))
    fin = feat_o.detach().clone().requires_grad_(True)
    mf1.zero_grad()
    p1 = mf1(fin)
    from loss import CrossEntropyLoss2d
    (CrossEntropyLoss2d(w)(p1, lbl) + 0).backward()
    e_head = (nerr(p1, o1), nerr(mf1.up.weight.grad, g_up))
    lines += ["seg              %.3e %.3e %.3e" % e_seg, "head+ce          %.3e %.3e" % e_head]
    lines.append("worst            %.3e | %.3e %.3e | %.3e %.3e | %.2e" % (tuple(worst) + (n_far / n_all,)))
    _log("parity_per_layer_%dx%d.txt" % size, lines)
    assert worst[0] <= 2e-3, "activation %.3e" % worst[0]
    # gradients: see the module docstring (ReLU-mask flips bound the distance to fp32 for 16-bit operands)
    assert worst[1] <= 2e-2, "input gradient vs same-storage oracle %.3e" % worst[1]
    assert worst[2] <= 2e-2, "parameter gradient vs same-storage oracle %.3e" % worst[2]
    assert worst[3] <= 3e-2, "input gradient vs fp32: rel-L2 %.3e" % worst[3]
    assert worst[4] <= 6e-2, "parameter gradient vs fp32: rel-L2 %.3e" % worst[4]
    assert n_far / n_all <= 1e-2, "gradient elements off by more than 2e-2 of max|ref|: %.3e" % (n_far / n_all)
    assert max(e_seg) <= 5e-3 and max(e_head) <= 2e-3


def test_end_to_end_losses_and_drift_vs_oracle(cuda_dev):
    from loss import CrossEntropyLoss2d, Diff2d
    dev, size, n = cuda_dev, (240, 320), 2
    G, F1, F2 = _state(dev)
    mg, mf1, mf2 = _models(dev)
    _load(mg, G), _load(mf1, F1), _load(mf2, F2)
    mg.train(), mf1.train(), mf2.train()
    src, tgt, lbl = _inputs(5, n, size, dev)
    w = O.class_weight(N_CLASS).to(dev)

    def run_oracle(st):
        g_, f1_, f2_ = _clone(G), _clone(F1), _clone(F2)
        taps_ = {}
        with torch.no_grad(), O.storage(st):
            feat_ = O.seg_base_forward(g_, src, taps=taps_)
            a1, a2 = O.head_forward(f1_, feat_), O.head_forward(f2_, feat_)
            ce_ = O.ce2d(a1, lbl, w) + O.ce2d(a2, lbl, w)
            ft_ = O.seg_base_forward(g_, tgt)
            d_ = O.diff2d(O.head_forward(f1_, ft_), O.head_forward(f2_, ft_))
        return g_, taps_, feat_, float(ce_), float(d_)

    g32, taps32, feat32, ce32, d32 = run_oracle(None)
    _, taps16, feat16, ce16, d16 = run_oracle(torch.float16)
    outs = {}
    handles = [mod.register_forward_hook(lambda m, i, o, key=key: outs.__setitem__(key, o.detach()))
               for key, mod, _ in _units(mg)]
    with torch.no_grad():
        feat = mg(src)
        for h in handles:
            h.remove()
        p1, p2 = mf1(feat), mf2(feat)
        crit = CrossEntropyLoss2d(w)
        ce = float(crit(p1, lbl) + crit(p2, lbl))
        ft = mg(tgt)
        d = float(Diff2d()(mf1(ft), mf2(ft)))
    from mcd_b200 import ops
    lines = ["# unit   cuda-vs-fp32   same-storage-oracle-vs-fp32   cuda-vs-same-storage-oracle   (max-norm)"]
    ratio_ok = True
    for key in outs:
        a = ops.to_nchw_f32(outs[key])
        e_c, e_e = nerr(a, taps32[key]), nerr(taps16[key], taps32[key])
        lines.append("%-16s %.3e %.3e %.3e" % (key, e_c, e_e, nerr(a, taps16[key])))
        ratio_ok &= e_c <= 1.5 * e_e + 5e-3
    lines += ["feat %.3e %.3e" % (nerr(feat, feat32), nerr(feat16, feat32)),
              "ce   cuda %.6f  fp32 %.6f  same-storage %.6f" % (ce, ce32, ce16),
              "diff cuda %.6e  fp32 %.6e  same-storage %.6e" % (d, d32, d16)]
    _log("parity_end_to_end_drift.txt", lines)
    assert abs(ce - ce32) / abs(ce32) <= 1e-3
    assert abs(d - d32) / abs(d32) <= 1e-3
    assert ratio_ok, "drift exceeds 1.5x the same-storage emulation of the oracle"
    assert nerr(feat, feat32) <= 6e-2          # 41 layers end to end (bf16 storage: 0.26)
    # running statistics took the same two momentum updates (src, tgt)
    assert nerr(mg.base[8][1].running_var, g32["base.8.1.running_var"]) < 2e-2
    assert nerr(mg.base[0][1].running_mean, g32["base.0.1.running_mean"]) < 2e-2
    assert int(mg.base[0][1].num_batches_tracked) == 2 == int(g32["base.0.1.num_batches_tracked"])


def test_mcd_iteration_vs_oracle(cuda_dev):
    """Full A / B / 4xC iteration (adapt_trainer.py:162-212) through the drop-in modules with torch.optim.SGD."""
    from loss import CrossEntropyLoss2d, get_prob_distance_criterion
    from models.model_util import get_optimizer
    dev, size, n = cuda_dev, (240, 320), 2
    G, F1, F2 = _state(dev)
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
    c_loss = loss.item()
    optimizer_g.step(), optimizer_f.step()
    optimizer_g.zero_grad(), optimizer_f.zero_grad()
    outputs = model_g(src_imgs)
    outputs1, outputs2 = model_f1(outputs), model_f2(outputs)
    loss = criterion(outputs1, src_lbls) + criterion(outputs2, src_lbls)
    outputs = model_g(tgt_imgs)
    outputs1, outputs2 = model_f1(outputs), model_f2(outputs)
    loss = loss - criterion_d(outputs1, outputs2)
    loss.backward()
    b_loss = loss.item()
    optimizer_f.step()
    c_losses = []
    for i in range(4):
        optimizer_g.zero_grad()
        outputs = model_g(tgt_imgs)
        outputs1, outputs2 = model_f1(outputs), model_f2(outputs)
        loss = criterion_d(outputs1, outputs2) * 1.0
        loss.backward()
        c_losses.append(loss.item())
        optimizer_g.step()
    torch.cuda.synchronize()

    lines = ["A %.6f vs %.6f" % (c_loss, c_o), "B %.6f vs %.6f" % (b_loss, float(rec["B_loss"]))]
    lines += ["C%d %.6e vs %.6e" % (i, a, b) for i, (a, b) in enumerate(zip(c_losses, rec["C_losses"]))]
    werr = {k: nerr(p, G[k]) for k, p in model_g.named_parameters()}
    lines += ["w %-34s %.3e" % kv for kv in sorted(werr.items())]
    _log("parity_iteration.txt", lines)
    assert abs(c_loss - c_o) / abs(c_o) <= 1e-3
    assert abs(b_loss - float(rec["B_loss"])) / abs(float(rec["B_loss"])) <= 1e-3
    for a, b in zip(c_losses, rec["C_losses"]):
        assert abs(a - b) / abs(b) <= 1e-3
    assert max(werr.values()) <= 1e-3          # weights after 5 G-steps at lr 1e-3
    assert nerr(model_f1.up.weight, F1["up.weight"]) <= 1e-3
    assert nerr(model_g.base[4][0].bn1.running_var, G["base.4.0.bn1.running_var"]) <= 2e-2


def test_tester_argmax_entropy_vs_oracle(cuda_dev):
    """adapt_tester.py:104-124 on a 'trained-like' state: BatchNorm shifts that keep most ReLUs active (which
    makes the network non-chaotic, as trained networks are) and running statistics calibrated to the data."""
    import util
    dev = cuda_dev
    G, F1, _ = _state(dev, seed=7)
    for k in list(G):
        if k.endswith(".bias") and not k.startswith("seg") and G[k].dim() == 1:
            G[k] += 2.0
    _, tgt, _ = _inputs(11, 1, (480, 640), dev)
    momentum, O.BN_MOMENTUM = O.BN_MOMENTUM, 1.0      # one calibration pass: running stats := batch stats
    try:
        with torch.no_grad():
            O.seg_base_forward(G, tgt, train=True)
    finally:
        O.BN_MOMENTUM = momentum
    mg, mf1, _ = _models(dev)
    _load(mg, G), _load(mf1, F1)
    mg.eval(), mf1.eval()
    with torch.no_grad():
        out = mf1(mg(tgt))
        ref = O.head_forward(F1, O.seg_base_forward(G, tgt, train=False))
        with O.storage(torch.float16, act=None):     # trunk storage as ours; the predictions stay fp32
            ref16 = O.head_forward(F1, O.seg_base_forward(G, tgt, train=False))
    pred = util.predict_labels(out, N_CLASS - 1)
    ref_pred = O.predict_labels(ref, N_CLASS - 1)
    agree = float((pred == ref_pred).float().mean())
    agree16 = float((pred == O.predict_labels(ref16, N_CLASS - 1)).float().mean())
    agree_oo = float((O.predict_labels(ref16, N_CLASS - 1) == ref_pred).float().mean())
    ent, ent_o = float(util.calc_entropy(out)), float(O.calc_entropy(ref))
    _log("parity_tester.txt", ["argmax agreement cuda vs fp32 oracle %.5f" % agree,
                               "argmax agreement cuda vs same-storage oracle %.5f" % agree16,
                               "argmax agreement same-storage oracle vs fp32 oracle %.5f" % agree_oo,
                               "entropy %.6e vs %.6e" % (ent, ent_o),
                               "logit err vs fp32 %.3e vs same-storage %.3e" % (nerr(out, ref), nerr(out, ref16)),
                               "distinct labels %d" % int(ref_pred.unique().numel())])
    assert pred.dtype == torch.int64 and pred.shape == (1, 480, 640)
    assert out.dtype == torch.float32          # the drop-in default: `.cpu().numpy()` of adapt_tester.py:114-118 works
    assert out[0].data.cpu().numpy().shape == (N_CLASS, 480, 640)
    # north_star: >= 99.5 % of the pixels (synthetic random weights give 40-way near-ties: bf16 storage moved 0.8 %)
    assert agree >= 0.995 and agree16 >= 0.995, (agree, agree16)
    assert abs(ent - ent_o) / abs(ent_o) <= 1e-3


@pytest.mark.parametrize("graph", [False, True, "prefetch", "fused-bn-bwd", "fused-bn-bwd-512"])
def test_mcdstep_runner_vs_oracle(cuda_dev, graph):
    """mcd_b200.step.MCDStep (dead phase-B backward skipped, phase-B target forward re-used for C[0] with folded
    BatchNorm updates, optional CUDA-graph replay) produces the reference iteration's results.
    "fused-bn-bwd[-512]": the same with the opt-in dgrad epilogue that applies the ReLU mask and accumulates the
    BatchNorm-backward sums of the producing unit (everywhere / for >= 512-channel inputs only) - measured slower than
    the separate reduction pass and off by default, kept and tested as an A/B switch."""
    from loss import CrossEntropyLoss2d, get_prob_distance_criterion
    from mcd_b200 import nn as mnn
    from mcd_b200.step import MCDStep
    if isinstance(graph, str) and graph.startswith("fused-bn-bwd"):
        prev = (mnn._fuse_bn_bwd, mnn._fuse_bn_bwd_min_c)
        mnn.set_fuse_bn_bwd(True, 512 if graph.endswith("512") else 0)
        try:
            return _mcdstep_runner_case(cuda_dev, False)
        finally:
            mnn.set_fuse_bn_bwd(*prev)
    return _mcdstep_runner_case(cuda_dev, graph)


def _mcdstep_runner_case(cuda_dev, graph):
    from loss import CrossEntropyLoss2d, get_prob_distance_criterion
    from mcd_b200.step import MCDStep
    dev, size, n = cuda_dev, (240, 320), 2
    G, F1, F2 = _state(dev)
    models = _models(dev)
    for m, sd in zip(models, (G, F1, F2)):
        _load(m, sd)
        m.train()
    src, tgt, lbl = _inputs(9, n, size, dev)
    w = O.class_weight(N_CLASS).to(dev)
    step = MCDStep(models, CrossEntropyLoss2d(w), get_prob_distance_criterion("diff"), num_k=4)
    iters = 2
    og, of = O.SGD(), O.SGD()
    for _ in range(iters):
        c_o, d_o = O.mcd_step_early(G, F1, F2, src, lbl, tgt, w, og, of, num_k=4)
    if graph:
        # 1 eager iteration + 1 graph replay = 2 iterations (capturing does not execute anything)
        step(src, lbl, tgt)
        step.capture(src, lbl, tgt, warmup=0)
        if graph == "prefetch":
            # the e2e path of bench.py: pinned host batch -> staging (copy stream) -> static inputs -> graph
            host = [t.cpu().pin_memory() for t in (src, lbl, tgt)]
            step._static[0].zero_()                       # the graph must see the prefetched data, not the captured
            step.prefetch(*host)
            c, d = step.replay_prefetched()
        else:
            c, d = step.replay(src, lbl, tgt)
    else:
        for _ in range(iters):
            c, d = step(src, lbl, tgt)
    torch.cuda.synchronize()
    assert abs(float(c) - c_o) / abs(c_o) <= 1e-3
    assert abs(float(d) - d_o) / abs(d_o) <= 1e-3
    werr = max(nerr(p, G[k]) for k, p in models[0].named_parameters())
    assert werr <= 2e-3, werr
    assert nerr(models[1].up.weight, F1["up.weight"]) <= 2e-3
    # BatchNorm bookkeeping: 7 forward passes per iteration (A, B-src, B-tgt == C0 folded, C1..C3)
    assert int(models[0].base[5][2].bn2.num_batches_tracked) == 7 * iters == int(G["base.5.2.bn2.num_batches_tracked"])
    assert nerr(models[0].base[5][2].bn2.running_var, G["base.5.2.bn2.running_var"]) <= 2e-2
    assert nerr(models[0].base[0][1].running_mean, G["base.0.1.running_mean"]) <= 2e-2


# ---- config 3: MFNet two-stream heads (adapt_mfnet_trainer.py:181-235) -------------------------------------
@pytest.mark.parametrize("method,kind", [("MCD-MFNet-AddFusion", "add"), ("MCD-MFNet-ScoreAddFusion", "scoreadd")])
def test_mfnet_step_vs_oracle(cuda_dev, method, kind):
    from loss import CrossEntropyLoss2d, Diff2d
    dev, size, n = cuda_dev, (240, 320), 2
    G3 = O.to_device(O.fill_state_dict_(O.init_seg_base("drn_d_38", 3, N_CLASS), 21), dev)
    G1 = O.to_device(O.fill_state_dict_(O.init_seg_base("drn_d_38", 3, N_CLASS), 22), dev)
    F1 = O.to_device(O.fill_state_dict_(O.init_head(N_CLASS, kind), 23), dev)
    F2 = O.to_device(O.fill_state_dict_(O.init_head(N_CLASS, kind), 24), dev)
    g3, g1, f1, f2 = _models(dev, method)
    for m, sd in ((g3, G3), (g1, G1), (f1, F1), (f2, F2)):
        _load(m, sd)
        m.train()
    src, tgt, lbl = _inputs(31, n, size, dev)
    w = O.class_weight(N_CLASS).to(dev)
    # oracle: B-phase objective CE(src) - Diff2d(tgt) exercises both streams, both heads, both criteria
    O._req([G3, G1, F1, F2])
    o3, o1 = O.seg_base_forward(G3, src[:, :3]), O.seg_base_forward(G1, src[:, 3:])
    p1, p2 = O.head_forward(F1, (o3, o1), kind), O.head_forward(F2, (o3, o1), kind)
    ce_o = O.ce2d(p1, lbl, w) + O.ce2d(p2, lbl, w)
    t3, t1 = O.seg_base_forward(G3, tgt[:, :3]), O.seg_base_forward(G1, tgt[:, 3:])
    d_o = O.diff2d(O.head_forward(F1, (t3, t1), kind), O.head_forward(F2, (t3, t1), kind))
    _, _, gF1, _ = O._grads(ce_o - d_o, [G3, G1, F1, F2])
    # ours, written like the trainer
    outputs_3ch, outputs_1ch = g3(src[:, :3, :, :]), g1(src[:, 3:, :, :])
    outputs1, outputs2 = f1(outputs_3ch, outputs_1ch), f2(outputs_3ch, outputs_1ch)
    crit = CrossEntropyLoss2d(w)
    ce = crit(outputs1, lbl) + crit(outputs2, lbl)
    t3_, t1_ = g3(tgt[:, :3, :, :]), g1(tgt[:, 3:, :, :])
    d = Diff2d()(f1(t3_, t1_), f2(t3_, t1_))
    (ce - d).backward()
    torch.cuda.synchronize()
    assert outputs1.shape == (n, N_CLASS, *size)
    assert abs(float(ce) - float(ce_o)) / abs(float(ce_o)) <= 1e-3
    assert abs(float(d) - float(d_o)) / abs(float(d_o)) <= 1e-3
    # head parameters: dW = sum x (x) dout inherits the end-to-end drift of the 41-layer features x (3 % at layer 8)
    for k, p in f1.named_parameters():
        e = float((p.grad - gF1[k]).norm() / gF1[k].norm())
        assert e <= 0.05, (k, e)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for m in (g3, g1) for p in m.parameters())
    # the fused ScoreAdd head equals up1(x1) + up2(x2) on identical inputs
    with torch.no_grad():
        ref = O.head_forward(F1, (o3.detach(), o1.detach()), kind)
        got = f1(o3.detach(), o1.detach())
    assert nerr(got, ref) <= 8e-3


# ---- config 4 / 5: triple multitask (adapt_triple_multitask_trainer.py:194-287, tester :117-142) ------------
def test_triple_multitask_vs_oracle(cuda_dev):
    import util
    from loss import CrossEntropyLoss2d, Diff2d
    from models.model_util import get_triple_multitask_models
    dev, size, n = cuda_dev, (240, 320), 2
    E = O.to_device(O.fill_state_dict_(O.init_trunk("drn_d_38", 3, "main_layer"), 31), dev)
    D = O.to_device(O.fill_state_dict_(O.init_triple_decoder(N_CLASS, 3), 32), dev)
    w = O.class_weight(N_CLASS).to(dev)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model_enc, model_dec = get_triple_multitask_models("drn_d_38", 6, N_CLASS,
                                                           semseg_criterion=CrossEntropyLoss2d(w),
                                                           discrepancy_criterion=Diff2d())
    model_enc, model_dec = model_enc.to(dev).train(), model_dec.to(dev).train()
    _load(model_enc, E)
    model_dec.load_state_dict({k: v.clone() for k, v in D.items()}, strict=False)   # criterion buffer stays
    g = torch.Generator().manual_seed(303)
    src = torch.randn(n, 7, *size, generator=g)
    src[:, 6] = (torch.rand(n, *size, generator=g) < 0.1).float()
    tgt = torch.randn(n, 6, *size, generator=g)
    lbl = torch.randint(0, N_CLASS, (n, *size), generator=g)
    src, tgt, lbl = src.to(dev), tgt.to(dev), lbl.to(dev)
    # oracle (phase A objective + discrepancy)
    O._req([E, D])
    src_f, tgt_f = O.encoder_dict(E, src[:, :3]), O.encoder_dict(E, tgt[:, :3])
    semseg_o, dep_o, bd_o = O.triple_get_loss(D, src_f, lbl, src[:, 3:-1], src[:, -1:], w)
    tdep_o = torch.nn.functional.mse_loss(O.triple_depth(D, tgt_f), tgt[:, 3:])
    disc_o = O.diff2d(*O.triple_semseg(D, tgt_f))
    gE, gD = O._grads(semseg_o + dep_o + bd_o + tdep_o - disc_o, [E, D])
    # ours, written like the trainer
    src_rgbs, src_depths, src_boundary = src[:, :3, :, :], src[:, 3:-1, :, :], src[:, -1:, :, :]
    tgt_rgbs, tgt_depths = tgt[:, :3, :, :], tgt[:, 3:, :, :]
    src_fet, tgt_fet = model_enc(src_rgbs), model_enc(tgt_rgbs)
    assert sorted(src_fet) == ["h%d" % i for i in range(9)] and src_fet["h2"].shape[1:] == (32, 120, 160)
    semseg, dep, bd = model_dec.get_loss(src_fet, lbl, src_depths, src_boundary, separately_returning=True)
    tdep = model_dec.get_depth_loss(tgt_fet, tgt_depths)
    disc = model_dec.get_cls_descrepancy(tgt_fet)
    (semseg + dep + bd + tdep - disc).backward()
    torch.cuda.synchronize()
    for name, a, b in (("semseg", semseg, semseg_o), ("depth", dep, dep_o), ("boundary", bd, bd_o),
                       ("tgt_depth", tdep, tdep_o)):
        assert abs(float(a) - float(b)) / abs(float(b)) <= 1e-3, (name, float(a), float(b))
    assert abs(float(disc) - float(disc_o)) / abs(float(disc_o)) <= 1e-3
    # decoder-side gradients (short path): uncertainty scalars, boundary convs, last decoder layers
    for k in ("s_semsegcls", "s_deprgr", "s_boundary", "conv3.bias", "conv1.weight", "deprgr_dec.conv3.weight",
              "semsegcls_dec1.conv3.bias"):
        p = dict(model_dec.named_parameters())[k]
        e = float((p.grad - gD[k]).norm() / (gD[k].norm() + 1e-20))
        assert e <= 8e-2, (k, e)
    assert model_dec.nmlrgr_dec.conv3.weight.grad is None        # never used (reference :813)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model_enc.parameters())
    # tester: forward, argmax without background, entropy, depth and boundary maps
    model_enc.eval(), model_dec.eval()
    O._req([E, D], False)
    with torch.no_grad():
        semseg1, semseg2, depth, boundary = model_dec(model_enc(tgt_rgbs[:1]))
        f = O.encoder_dict(E, tgt[:1, :3], train=False)
        s1_o, _ = O.triple_semseg(D, f, train=False)
        depth_o, bd_map_o = O.triple_depth(D, f, train=False), O.triple_boundary(D, f)
    assert semseg1.shape == (1, N_CLASS, *size) and depth.shape == (1, 3, *size) and boundary.shape == (1, 1, *size)
    agree = float((util.predict_labels(semseg1, N_CLASS - 1) == O.predict_labels(s1_o, N_CLASS - 1)).float().mean())
    assert agree >= 0.995, agree
    # eval mode on un-calibrated random running statistics is ill-conditioned for the deep h8 branch: compare the
    # probability maps on average, the shallow-branch-dominated structure must agree
    assert float((boundary.float() - bd_map_o).abs().mean()) <= 2e-2
    assert float((depth.float() - depth_o).pow(2).mean().sqrt() / depth_o.pow(2).mean().sqrt()) <= 0.3
