#!/usr/bin/env python
"""Generate the golden vectors that pin oracle/mcd_oracle.py to the REAL reference implementation.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It stages a throw-away py3-patched copy of the reference under /tmp (four one-line Python-2-isms, SURVEY.md
section 8c; nothing of the reference is written into this repository), imports the reference's own modules
(models/model_util.py factories, loss.py criteria), gives them deterministic weights with
oracle.fill_state_dict_ (regenerable from the key names, so no weights are shipped), replays the reference's
trainer / tester loop bodies on seeded synthetic inputs with torch.optim.SGD, and stores the resulting losses,
activations, gradient norms and updated-weight checksums in tests/golden/*.npz.

Re-running it reproduces every array bit for bit except the multi-step quantities of iterations.npz (phase-C losses and
weight checksums after 3-5 optimizer steps), which move by 1e-8 ... 3e-7 relative from run to run with the summation
order of torch's multi-threaded CPU kernels; the tests compare those at 2e-4 or looser.
"""
import os
import shutil
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import mcd_oracle as O  # noqa: E402

REF = "/root/reference"
STAGE = "/tmp/mcd_ref_py3"
SIZE = (96, 128)
N_CLASS = 41


def stage_reference():
    if os.path.exists(STAGE):
        shutil.rmtree(STAGE)
    shutil.copytree(REF, STAGE, ignore=shutil.ignore_patterns("*.png", "*.jpg", "docs", "sample_img", "_static"))

    def patch(rel, pairs):
        p = os.path.join(STAGE, rel)
        s = open(p).read()
        for a, b in pairs:
            assert a in s, (rel, a)
            s = s.replace(a, b)
        open(p, "w").write(s)

    patch("loss.py", [("print prob1", "print(prob1)")])
    patch("models/dilated_fcn.py", [("cuda(async=True)", "cuda(non_blocking=True)"),
                                    ("\nimport drn\n", "\nfrom models import drn\n"),
                                    # CPU-only container: the option branches call .cuda() unconditionally
                                    ("loss1 += extra_loss1.cuda()", "loss1 += extra_loss1"),
                                    ("loss2 += extra_loss2.cuda()", "loss2 += extra_loss2")])
    patch("models/drn.py", [("gen.next()", "next(gen)")])
    sys.path.insert(0, STAGE)
    import models.drn as drn
    for name in ("drn_d_22", "drn_d_38", "drn_c_26"):
        orig = getattr(drn, name)
        setattr(drn, name, (lambda f: lambda pretrained=False, **kw: f(pretrained=False, **kw))(orig))


def filled(module, seed):
    # the multitask decoders register their criterion as a sub-module, so its class-weight buffer
    # (`semseg_criterion.nll_loss.weight`) shows up in state_dict(): keep that one as configured.
    sd = {k: v for k, v in module.state_dict().items() if "criterion" not in k}
    O.fill_state_dict_(sd, seed)
    module.load_state_dict(sd, strict=False)
    return module


def inputs(seed, n=2, ch=6, size=SIZE):
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(n, ch, *size, generator=g)
    tgt = torch.randn(n, 6, *size, generator=g)
    lbl = torch.randint(0, N_CLASS, (n, *size), generator=g)
    return src, tgt, lbl


def norms(named_grads):
    keys = sorted(named_grads)
    return keys, np.array([float(named_grads[k].norm()) if named_grads[k] is not None else -1.0 for k in keys],
                          dtype=np.float64)


def summarize(sd):
    keys = sorted(k for k in sd if torch.is_floating_point(sd[k]))
    return keys, np.array([[float(sd[k].double().sum()), float(sd[k].double().norm())] for k in keys])


def golden_early_fusion():
    """adapt_trainer.py:151-215 replayed verbatim (python-3 spellings) on synthetic tensors."""
    from loss import CrossEntropyLoss2d, get_prob_distance_criterion
    from models.model_util import get_models, get_optimizer
    from util import get_class_weight_from_file
    model_g, model_f1, model_f2 = get_models(net_name="drn_d_38", res="50", input_ch=6, n_class=N_CLASS,
                                             method="MCD", is_data_parallel=False)
    filled(model_g, 11), filled(model_f1, 12), filled(model_f2, 13)
    optimizer_g = get_optimizer(model_g.parameters(), lr=1e-3, momentum=0.9, opt="sgd", weight_decay=2e-5)
    optimizer_f = get_optimizer(list(model_f1.parameters()) + list(model_f2.parameters()), opt="sgd", lr=1e-3,
                                momentum=0.9, weight_decay=2e-5)
    weight = get_class_weight_from_file(n_class=N_CLASS, weight_filename=None, add_bg_loss=False)
    criterion = CrossEntropyLoss2d(weight)
    criterion_d = get_prob_distance_criterion("diff")
    model_g.train(), model_f1.train(), model_f2.train()
    src_imgs, tgt_imgs, src_lbls = inputs(101)
    num_k, out = 4, {}

    # phase A
    optimizer_g.zero_grad(), optimizer_f.zero_grad()
    outputs = model_g(src_imgs)
    outputs1, outputs2 = model_f1(outputs), model_f2(outputs)
    loss = criterion(outputs1, src_lbls) + criterion(outputs2, src_lbls)
    loss.backward()
    out["A_loss"] = float(loss)
    out["A_feat"] = outputs.detach().numpy()
    out["A_out1_sub"] = outputs1.detach()[:, :, ::8, ::8].numpy()
    gk, gv = norms({k: p.grad for k, p in model_g.named_parameters()})
    out["A_grad_g_keys"], out["A_grad_g_norms"] = np.array(gk), gv
    out["A_grad_up1"] = model_f1.up.weight.grad.numpy().copy()
    out["A_grad_seg_bias"] = model_g.seg.bias.grad.numpy().copy()
    optimizer_g.step(), optimizer_f.step()
    # phase B
    optimizer_g.zero_grad(), optimizer_f.zero_grad()
    outputs = model_g(src_imgs)
    outputs1, outputs2 = model_f1(outputs), model_f2(outputs)
    loss = criterion(outputs1, src_lbls) + criterion(outputs2, src_lbls)
    outputs = model_g(tgt_imgs)
    outputs1, outputs2 = model_f1(outputs), model_f2(outputs)
    loss = loss - criterion_d(outputs1, outputs2)
    loss.backward()
    out["B_loss"] = float(loss)
    out["B_grad_up1"] = model_f1.up.weight.grad.numpy().copy()
    optimizer_f.step()
    # phase C
    c_losses = []
    for i in range(num_k):
        optimizer_g.zero_grad()
        outputs = model_g(tgt_imgs)
        outputs1, outputs2 = model_f1(outputs), model_f2(outputs)
        loss = criterion_d(outputs1, outputs2) * 1.0
        loss.backward()
        if i == 0:
            gk, gv = norms({k: p.grad for k, p in model_g.named_parameters()})
            out["C0_grad_g_norms"] = gv
        c_losses.append(float(loss))
        optimizer_g.step()
    out["C_losses"] = np.array(c_losses)
    sk, sv = summarize(model_g.state_dict())
    out["final_g_keys"], out["final_g_sums"] = np.array(sk), sv
    out["final_up1"] = model_f1.up.weight.detach().numpy().copy()
    # tester (adapt_tester.py:104-124): eval forward, argmax without the background channel, entropy
    from util import calc_entropy
    model_g.eval(), model_f1.eval()
    with torch.no_grad():
        o = model_f1(model_g(tgt_imgs[:1]))
    out["test_labels"] = o[0, :N_CLASS - 1].max(0)[1].numpy()
    out["test_entropy"] = float(calc_entropy(o))
    np.savez_compressed(os.path.join(HERE, "early_fusion.npz"), **out)
    print("early_fusion: A %.6f B %.6f C %s" % (out["A_loss"], out["B_loss"], c_losses))


def golden_mfnet():
    from loss import CrossEntropyLoss2d, Diff2d
    from models.model_util import get_models
    from util import get_class_weight_from_file
    out = {}
    for tag, method in (("add", "MCD-MFNet-AddFusion"), ("scoreadd", "MCD-MFNet-ScoreAddFusion")):
        g3, g1, f1, f2 = get_models(net_name="drn_d_38", res="50", input_ch=6, n_class=N_CLASS, method=method)
        filled(g3, 21), filled(g1, 22), filled(f1, 23), filled(f2, 24)
        for m in (g3, g1, f1, f2):
            m.train()
        src, tgt, lbl = inputs(202, size=(64, 96))
        crit = CrossEntropyLoss2d(get_class_weight_from_file(n_class=N_CLASS))
        # adapt_mfnet_trainer.py:186-192 forward pattern + the B-phase loss (CE on src - Diff2d on tgt)
        o3, o1 = g3(src[:, :3, :, :]), g1(src[:, 3:, :, :])
        p1, p2 = f1(o3, o1), f2(o3, o1)
        loss = crit(p1, lbl) + crit(p2, lbl)
        t3, t1 = g3(tgt[:, :3, :, :]), g1(tgt[:, 3:, :, :])
        q1, q2 = f1(t3, t1), f2(t3, t1)
        d = Diff2d()(q1, q2)
        (loss - d).backward()
        out[tag + "_ce"], out[tag + "_diff"] = float(loss), float(d)
        out[tag + "_feat3"] = o3.detach().numpy()
        out[tag + "_p1_sub"] = p1.detach()[:, :, ::8, ::8].numpy()
        gk, gv = norms({k: p.grad for k, p in g1.named_parameters()})
        out[tag + "_grad_g1_keys"], out[tag + "_grad_g1_norms"] = np.array(gk), gv
        for k, p in f1.named_parameters():
            out[tag + "_grad_f1_" + k] = p.grad.numpy().copy()
        print("mfnet %s: ce %.6f diff %.6f" % (tag, float(loss), float(d)))
    np.savez_compressed(os.path.join(HERE, "mfnet.npz"), **out)


def golden_triple():
    from loss import CrossEntropyLoss2d, Diff2d
    from models.model_util import get_triple_multitask_models
    from util import calc_entropy, get_class_weight_from_file
    crit = CrossEntropyLoss2d(get_class_weight_from_file(n_class=N_CLASS))
    enc, dec = get_triple_multitask_models(net_name="drn_d_38", input_ch=6, n_class=N_CLASS,
                                           semseg_criterion=crit, discrepancy_criterion=Diff2d())
    filled(enc, 31), filled(dec, 32)
    enc.train(), dec.train()
    g = torch.Generator().manual_seed(303)
    size = (64, 96)
    src = torch.randn(2, 7, *size, generator=g)
    src[:, 6] = (torch.rand(2, *size, generator=g) < 0.1).float()
    tgt = torch.randn(2, 6, *size, generator=g)
    lbl = torch.randint(0, N_CLASS, (2, *size), generator=g)
    # adapt_triple_multitask_trainer.py:194-248 (phase A)
    src_rgbs, src_depths, src_boundary = src[:, :3], src[:, 3:-1], src[:, -1:]
    tgt_rgbs, tgt_depths = tgt[:, :3], tgt[:, 3:]
    src_fet, tgt_fet = enc(src_rgbs), enc(tgt_rgbs)
    semseg, dep, bd = dec.get_loss(src_fet, lbl, src_depths, src_boundary, separately_returning=True)
    tgt_dep = dec.get_depth_loss(tgt_fet, tgt_depths)
    disc = dec.get_cls_descrepancy(tgt_fet)
    total = semseg + dep + bd + tgt_dep - disc
    total.backward()
    out = dict(semseg=float(semseg), dep=float(dep), bd=float(bd), tgt_dep=float(tgt_dep), disc=float(disc))
    out["h8_sub"] = src_fet["h8"].detach()[:, ::16].numpy()
    gk, gv = norms({k: p.grad for k, p in enc.named_parameters()})
    out["grad_enc_keys"], out["grad_enc_norms"] = np.array(gk), gv
    gk, gv = norms({k: p.grad for k, p in dec.named_parameters()})
    out["grad_dec_keys"], out["grad_dec_norms"] = np.array(gk), gv
    # tester (adapt_triple_multitask_tester.py:117-142)
    enc.eval(), dec.eval()
    with torch.no_grad():
        s1, s2, depth, boundary = dec(enc(tgt_rgbs[:1]))
    out["test_labels"] = s1[0, :N_CLASS - 1].max(0)[1].numpy()
    out["test_entropy"] = float(calc_entropy(s1))
    out["test_depth_sub"] = depth[:, :, ::8, ::8].numpy()
    out["test_boundary_sub"] = boundary[:, :, ::8, ::8].numpy()
    np.savez_compressed(os.path.join(HERE, "triple.npz"), **out)
    print("triple:", {k: v for k, v in out.items() if isinstance(v, float)})


def golden_losses():
    from loss import CrossEntropyLoss2d, Diff2d, bce2d
    g = torch.Generator().manual_seed(404)
    a = (torch.randn(2, N_CLASS, 12, 16, generator=g) * 3).requires_grad_(True)
    b = (torch.randn(2, N_CLASS, 12, 16, generator=g) * 3).requires_grad_(True)
    t = torch.randint(0, N_CLASS, (2, 12, 16), generator=g)
    t[0, 0, :4] = -100
    w = torch.ones(N_CLASS)
    w[N_CLASS - 1] = 0
    w[5] = 2.0
    ce = CrossEntropyLoss2d(w)(a, t)
    df = Diff2d()(a, b)
    (ce + df).backward()
    p = torch.rand(2, 1, 12, 16, generator=g).clamp(0.01, 0.99).requires_grad_(True)
    tb = (torch.rand(2, 1, 12, 16, generator=g) < 0.2).float()
    bc = bce2d(p, tb)
    bc.backward()
    np.savez_compressed(os.path.join(HERE, "losses.npz"), a=a.detach().numpy(), b=b.detach().numpy(), t=t.numpy(),
                        w=w.numpy(), ce=float(ce), diff=float(df), da=a.grad.numpy(), db=b.grad.numpy(),
                        p=p.detach().numpy(), tb=tb.numpy(), bce=float(bc), dp=p.grad.numpy())
    print("losses: ce %.6f diff %.6f bce %.6f" % (float(ce), float(df), float(bc)))


def golden_iterations():
    """Full A / B / num_k x C iterations of the MFNet and the two multitask trainers replayed verbatim (python-3
    spellings) with torch.optim.SGD: adapt_mfnet_trainer.py:181-235, adapt_triple_multitask_trainer.py:202-287,
    adapt_multitask_trainer.py:194-262; plus get_boundary_loss (models/dilated_fcn.py:743-787)."""
    from loss import CrossEntropyLoss2d, Diff2d, get_prob_distance_criterion
    from models.dilated_fcn import get_boundary_loss
    from models.model_util import get_models, get_multitask_models, get_optimizer, get_triple_multitask_models
    from util import get_class_weight_from_file
    out, size, num_k = {}, (48, 64), 2
    weight = get_class_weight_from_file(n_class=N_CLASS)
    kw = dict(lr=1e-3, momentum=0.9, opt="sgd", weight_decay=2e-5)

    # ---- MFNet ScoreAddFusion
    g3, g1, f1, f2 = get_models(net_name="drn_d_38", res="50", input_ch=6, n_class=N_CLASS,
                                method="MCD-MFNet-ScoreAddFusion")
    filled(g3, 41), filled(g1, 42), filled(f1, 43), filled(f2, 44)
    optimizer_g = get_optimizer(list(g3.parameters()) + list(g1.parameters()), **kw)
    optimizer_f = get_optimizer(list(f1.parameters()) + list(f2.parameters()), **kw)
    criterion, criterion_d = CrossEntropyLoss2d(weight), get_prob_distance_criterion("diff")
    for m in (g3, g1, f1, f2):
        m.train()
    src_imgs, tgt_imgs, src_lbls = inputs(505, size=size)
    optimizer_g.zero_grad(), optimizer_f.zero_grad()
    a, b = g3(src_imgs[:, :3, :, :]), g1(src_imgs[:, 3:, :, :])
    loss = criterion(f1(a, b), src_lbls) + criterion(f2(a, b), src_lbls)
    loss.backward()
    out["mf_A"] = float(loss)
    optimizer_g.step(), optimizer_f.step()
    optimizer_g.zero_grad(), optimizer_f.zero_grad()
    a, b = g3(src_imgs[:, :3, :, :]), g1(src_imgs[:, 3:, :, :])
    loss = criterion(f1(a, b), src_lbls) + criterion(f2(a, b), src_lbls)
    a, b = g3(tgt_imgs[:, :3, :, :]), g1(tgt_imgs[:, 3:, :, :])
    loss = loss - criterion_d(f1(a, b), f2(a, b))
    loss.backward()
    out["mf_B"] = float(loss)
    optimizer_f.step()
    optimizer_f.zero_grad()
    cl = []
    for i in range(num_k):
        optimizer_g.zero_grad()
        a, b = g3(tgt_imgs[:, :3, :, :]), g1(tgt_imgs[:, 3:, :, :])
        loss = criterion_d(f1(a, b), f2(a, b))
        loss.backward()
        optimizer_g.step()
        cl.append(float(loss))
    out["mf_C"] = np.array(cl)
    sk, sv = summarize(g1.state_dict())
    out["mf_g1_keys"], out["mf_g1_sums"] = np.array(sk), sv
    out["mf_up1"] = f1.up1.weight.detach().numpy().copy()
    print("mfnet iteration: A %.6f B %.6f C %s" % (out["mf_A"], out["mf_B"], cl))

    # ---- multitask trainers
    for tag, triple in (("tri", True), ("mt", False)):
        crit = CrossEntropyLoss2d(weight)
        if triple:
            enc, dec = get_triple_multitask_models(net_name="drn_d_38", input_ch=6, n_class=N_CLASS,
                                                   semseg_criterion=crit, discrepancy_criterion=Diff2d())
        else:
            enc, dec = get_multitask_models(net_name="drn_d_38", input_ch=6, n_class=N_CLASS,
                                            semseg_criterion=crit, discrepancy_criterion=Diff2d())
        filled(enc, 51), filled(dec, 52)
        enc.train(), dec.train()
        optimizer_enc = get_optimizer(enc.parameters(), **kw)
        optimizer_dec = get_optimizer(dec.parameters(), **kw)
        g = torch.Generator().manual_seed(606)
        src_imgs = torch.randn(2, 7 if triple else 6, *size, generator=g)
        if triple:
            src_imgs[:, 6] = (torch.rand(2, *size, generator=g) < 0.1).float()
        tgt_imgs = torch.randn(2, 6, *size, generator=g)
        src_gt_semseg = torch.randint(0, N_CLASS, (2, *size), generator=g)
        src_rgbs = src_imgs[:, :3, :, :]
        src_depths = src_imgs[:, 3:-1, :, :] if triple else src_imgs[:, 3:, :, :]
        src_boundary = src_imgs[:, -1:, :, :]
        tgt_rgbs, tgt_depths = tgt_imgs[:, :3, :, :], tgt_imgs[:, 3:, :, :]
        # A
        optimizer_enc.zero_grad(), optimizer_dec.zero_grad()
        src_fet, tgt_fet = enc(src_rgbs), enc(tgt_rgbs)
        if triple:
            terms = dec.get_loss(src_fet, src_gt_semseg, src_depths, src_boundary, separately_returning=True)
        else:
            terms = dec.get_loss(src_fet, src_gt_semseg, src_depths, separately_returning=True)
        tgt_depth_loss = dec.get_depth_loss(tgt_fet, tgt_depths)
        loss = sum(terms) + tgt_depth_loss
        loss.backward()
        out[tag + "_A"] = float(loss)
        out[tag + "_A_terms"] = np.array([float(t) for t in terms] + [float(tgt_depth_loss)])
        optimizer_enc.step(), optimizer_dec.step()
        # B
        optimizer_enc.zero_grad(), optimizer_dec.zero_grad()
        src_fet = enc(src_rgbs)
        if triple:
            src_semseg_loss, _, _ = dec.get_loss(src_fet, src_gt_semseg, src_depths, src_boundary, separately_returning=True)
            tgt_fet = enc(tgt_rgbs)
            loss = src_semseg_loss - dec.get_cls_descrepancy(tgt_fet)
        else:
            src_semseg_loss, src_depth_loss = dec.get_loss(src_fet, src_gt_semseg, src_depths, separately_returning=True)
            tgt_fet = enc(tgt_rgbs)
            tgt_depth_loss = dec.get_depth_loss(tgt_fet, tgt_depths)
            loss = src_semseg_loss + src_depth_loss + tgt_depth_loss - dec.get_cls_descrepancy(tgt_fet)
        loss.backward()
        out[tag + "_B"] = float(loss)
        optimizer_dec.step()
        # C
        cl = []
        for i in range(num_k):
            optimizer_enc.zero_grad()
            tgt_fet = enc(tgt_rgbs)
            loss = dec.get_cls_descrepancy(tgt_fet) * 1.0
            loss.backward()
            optimizer_enc.step()
            cl.append(float(loss))
        out[tag + "_C"] = np.array(cl)
        sk, sv = summarize(enc.state_dict())
        out[tag + "_enc_keys"], out[tag + "_enc_sums"] = np.array(sk), sv
        sk, sv = summarize({k: v for k, v in dec.state_dict().items() if "criterion" not in k})
        out[tag + "_dec_keys"], out[tag + "_dec_sums"] = np.array(sk), sv
        print("%s iteration: A %.6f B %.6f C %s" % (tag, out[tag + "_A"], out[tag + "_B"], cl))

    # ---- get_boundary_loss
    g = torch.Generator().manual_seed(707)
    lab_p = torch.randint(0, 5, (2, 20, 24), generator=g)
    lab_g = torch.randint(0, 5, (2, 20, 24), generator=g)
    lab_p[:, 5:15, 6:18] = 2          # uniform regions: boundary only at their rims
    lab_g[:, 2:12, 3:20] = 1
    lab_g[:, 14:, :] = 3
    bmap = (torch.rand(2, 20, 24, generator=g) < 0.3).float()
    out["bd_lab_p"], out["bd_lab_g"], out["bd_map"] = lab_p.numpy(), lab_g.numpy(), bmap.numpy()
    out["bd_ss"] = float(get_boundary_loss(lab_p, lab_g))
    out["bd_sb"] = float(get_boundary_loss(lab_p, bmap, gt_type="boundary"))
    v = lab_p.float()
    out["bd_boundary_of_p"] = (torch.nn.functional.max_pool2d(v, 3, 1, 1) != -torch.nn.functional.max_pool2d(-v, 3, 1, 1)).numpy()
    print("boundary losses:", out["bd_ss"], out["bd_sb"])
    np.savez_compressed(os.path.join(HERE, "iterations.npz"), **out)


def golden_pipeline():
    """The byte-side neighbours of the step, from the reference's own code: transform.py ToLabel / ReLabel /
    get_img_transform / get_lbl_transform / unnormalize (imported from the staged copy; `Scale` is spelled `Resize` in
    today's torchvision, and Normalize is the torchvision-0.2 loop `for t, m, s in zip(tensor, mean, std)` the reference
    was written against - with 6 means on a 3-channel image it uses the first three), eval.py fast_hist and scores
    (function sources executed from the file: the module itself imports matplotlib), PIL's NEAREST resize."""
    import ast
    import torchvision.transforms as T
    from PIL import Image
    T.Scale = T.Resize

    class ZipNormalize:                                  # torchvision 0.2.x transforms.Normalize.__call__
        def __init__(self, mean, std):
            self.mean, self.std = mean, std

        def __call__(self, tensor):
            for t, m, s in zip(tensor, self.mean, self.std):
                t.sub_(m).div_(s)
            return tensor

    T.Normalize = ZipNormalize
    import transform as RT
    rng = np.random.RandomState(7)
    H, W = 10, 14
    rgb = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    hha = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    lbl = rng.randint(0, 41, (H, W)).astype(np.uint8)
    lbl[rng.rand(H, W) < 0.2] = 255
    bd = np.where(rng.rand(H, W) < 0.3, 255, 0).astype(np.uint8)
    out = {"rgb": rgb, "hha": hha, "lbl": lbl, "bd": bd}
    img_t = RT.get_img_transform((W, H), "imagenet", use_crop=True)
    lbl_t = RT.get_lbl_transform((W, H), N_CLASS, use_crop=True)
    x_rgb, x_hha = img_t(Image.fromarray(rgb)), img_t(Image.fromarray(hha))
    out["img6"] = torch.cat([x_rgb, x_hha]).numpy()                                  # datasets.py:667-680
    out["lbl_out"] = lbl_t(Image.fromarray(lbl)).numpy()
    to_bd = T.Compose(lbl_t.transforms[:-1] + [RT.ReLabel(255, 1)])                   # datasets.py:688-690
    out["img7"] = torch.cat([x_rgb, x_hha, to_bd(Image.fromarray(bd)).unsqueeze(0).float()]).numpy()
    city = RT.get_img_transform((W, H), "city", use_crop=True)
    out["img3_city"] = city(Image.fromarray(rgb)).numpy()
    un = RT.unnormalize(np.transpose(x_rgb.numpy(), (1, 2, 0)))
    out["unnorm"] = np.array(un)
    # eval.py functions
    src = open(os.path.join(REF, "eval.py")).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("fast_hist", "per_class_iu", "calc_fw_iu",
                                                                 "calc_pixel_accuracy", "calc_mean_accuracy"):
            exec(compile(ast.Module([node], []), "eval.py", "exec"), ns)
    gt = rng.randint(0, 41, 4000).astype(np.int64)
    gt[rng.rand(4000) < 0.1] = 255
    pred = np.where(rng.rand(4000) < 0.6, np.minimum(gt, 39), rng.randint(0, 40, 4000)).astype(np.int64)
    hist = ns["fast_hist"](gt, pred, 40)
    out.update(h_gt=gt, h_pred=pred, hist=hist, iu=ns["per_class_iu"](hist), fw_iu=ns["calc_fw_iu"](hist),
               pix_acc=ns["calc_pixel_accuracy"](hist), mean_acc=ns["calc_mean_accuracy"](hist))
    # PIL NEAREST resize of a label map: (in, out) size pairs incl. non-integer ratios
    for i, ((ih, iw), (oh, ow)) in enumerate([((10, 14), (20, 28)), ((10, 14), (7, 9)), ((48, 64), (53, 71)),
                                              ((30, 40), (1, 3)), ((9, 7), (64, 50))]):
        m = rng.randint(0, 41, (ih, iw)).astype(np.uint8)
        out["rs%d_in" % i] = m
        out["rs%d_out" % i] = np.array(Image.fromarray(m).resize((ow, oh), Image.NEAREST))
    np.savez_compressed(os.path.join(HERE, "pipeline.npz"), **out)


def golden_discrepancies():
    """every name get_prob_distance_criterion (loss.py:192-210) knows, from the reference's own classes: value and the
    gradients w.r.t. both logit tensors (torch 2.x differentiates F.kl_div through its target as well)."""
    from loss import get_prob_distance_criterion
    g = torch.Generator().manual_seed(505)
    a0 = torch.randn(2, N_CLASS, 6, 8, generator=g) * 2.5
    b0 = a0 + torch.randn(2, N_CLASS, 6, 8, generator=g) * 1.5
    out = {"a": a0.numpy(), "b": b0.numpy()}
    for name in ("diff", "jsd", "symkl", "nmlsymkl", "mysymkl", "spatial_jsd", "mis_symkl"):
        a, b = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
        v = get_prob_distance_criterion(name, N_CLASS)(a, b)
        v.backward()
        out[name] = float(v)
        out[name + "_da"], out[name + "_db"] = a.grad.numpy(), b.grad.numpy()
        print("discrepancy %-12s %.8e" % (name, float(v)))
    np.savez_compressed(os.path.join(HERE, "discrepancies.npz"), **out)


def golden_bottleneck():
    """DRN-D-54 (Bottleneck blocks, models/drn.py:62-100,337-341) through the reference's DRNSegBase: train-mode forward
    and the gradient norms of a quadratic objective, plus the BatchNorm bookkeeping of one forward."""
    from models.dilated_fcn import DRNSegBase
    torch.manual_seed(0)
    ref = DRNSegBase("drn_d_54", N_CLASS, pretrained=False, input_ch=6)
    G = O.fill_state_dict_(O.init_seg_base("drn_d_54", 6, N_CLASS), 54)
    assert set(G) == set(ref.state_dict())
    ref.load_state_dict({k: v.clone() for k, v in G.items()})
    ref.train()
    x = torch.randn(2, 6, 64, 96, generator=torch.Generator().manual_seed(540))
    feat = ref(x)
    feat.square().mean().backward()
    gk, gn = norms({k: p.grad for k, p in ref.named_parameters()})
    sk, sv = summarize(ref.state_dict())
    np.savez_compressed(os.path.join(HERE, "drn_d_54.npz"), feat=feat.detach().numpy(), grad_keys=np.array(gk),
                        grad_norms=gn, state_keys=np.array(sk), state_sums=sv)
    print("drn_d_54: %d state entries, feat norm %.6f" % (len(G), float(feat.norm())))


def _grads_of(obj, named):
    """obj.backward() and the gradients of the named leaves as numpy arrays"""
    for t in named.values():
        t.grad = None
    obj.backward()
    return {k: (t.grad.numpy().copy() if t.grad is not None else None) for k, t in named.items()}


def sample(a, n=512):
    """a strided sample of at most ~n elements of an array (the fixtures stay small)"""
    f = np.asarray(a).reshape(-1)
    return f[::max(1, f.size // n)].copy()


def variant_head_inputs(i, cin, h, w):
    """seeded inputs of head case i (regenerated, not stored, by tests/test_variants_gpu.py): two score maps
    (41 channels) or two post-ReLU trunk features (512 channels) and the weights r of the objective mean(out * r)"""
    g = torch.Generator().manual_seed(7100 + i)
    scale = 2.0 if cin != 512 else 0.5
    x1 = torch.randn(2, cin, h, w, generator=g) * scale
    x2 = torch.randn(2, cin, h, w, generator=g) * scale
    if cin == 512:
        x1, x2 = x1.clamp_(min=0), x2.clamp_(min=0)
    r = torch.randn(2, N_CLASS, 8 * h, 8 * w, generator=g)
    return x1.requires_grad_(True), x2.requires_grad_(True), r


def golden_variants():
    """The option surface around the hot path from the reference's own classes:
    fusion heads (models/fusion.py:6-65 through models/dilated_fcn.py:340-366,431-491 incl. ver2 and use_torch_up),
    FuseDRNSegBase (:253-337), DRN arch C (models/drn.py:113-153,303-306), the triple decoder's shortcut / seg2bd /
    add_pred_seg_boundary_loss options (:797-1019), MCDSegBDMultiTaskDecoder (:1027-1222) and the source-only decoders
    (:1225-1398)."""
    from loss import CrossEntropyLoss2d, Diff2d
    from models import dilated_fcn as D
    from util import get_class_weight_from_file
    out = {}
    h, w = 6, 8

    # ---- heads: objective = mean(out * r), r seeded; gradients w.r.t. the inputs and every parameter
    cases = [
        ("gate_v1", lambda: D.FusionDRNSegPixelClassifier("GateFusion", N_CLASS), N_CLASS),
        ("concat_v1", lambda: D.FusionDRNSegPixelClassifier("ConcatFusion", N_CLASS), N_CLASS),
        ("concatconv_v1", lambda: D.FusionDRNSegPixelClassifier("ConcatConvFusion", N_CLASS), N_CLASS),
        ("add_v2", lambda: D.FusionDRNSegPixelClassifier("AddFusion", N_CLASS, ver="ver2"), 512),
        ("gate_v2", lambda: D.FusionDRNSegPixelClassifier("GateFusion", N_CLASS, ver="ver2"), 512),
        ("add_torchup", lambda: D.FusionDRNSegPixelClassifier("AddFusion", N_CLASS, use_torch_up=True), N_CLASS),
        ("scoregate", lambda: D.ScoreFusionDRNSegPixelClassifier("ScoreGateFusion", N_CLASS), N_CLASS),
        ("scoregate_nosm", lambda: D.ScoreFusionDRNSegPixelClassifier("GateFusion", N_CLASS), N_CLASS),
        ("single_v2", lambda: D.DRNSegPixelClassifier(N_CLASS, ver="ver2"), 512),
        ("single_torchup", lambda: D.DRNSegPixelClassifier(N_CLASS, use_torch_up=True), N_CLASS),
    ]
    for i, (tag, make, cin) in enumerate(cases):
        head = filled(make(), 70 + i)
        head.train()
        two = not tag.startswith("single")
        x1, x2, r = variant_head_inputs(i, cin, h, w)
        o = head(x1, x2) if two else head(x1)
        named = {"x1": x1, **({"x2": x2} if two else {}), **{"p:" + k: p for k, p in head.named_parameters()}}
        gr = _grads_of((o * r).mean(), named)
        out[tag + ":out_sub"] = o.detach()[:, :, ::4, ::4].numpy()
        out[tag + ":out_norm"] = float(o.norm())
        for k, v in gr.items():
            out[tag + ":gn:" + k] = float(np.linalg.norm(v))
            out[tag + ":gs:" + k] = sample(v)
        print("variant head %-16s out %s |out| %.5f" % (tag, tuple(o.shape), float(o.norm())))

    # ---- FuseDRNSegBase (drn_d_22_fusenet) and DRN arch C (drn_c_26): eval forward, then one train-mode
    #      forward + backward of a quadratic objective (gradient norms, BatchNorm bookkeeping)
    for tag, make in (("fusenet", lambda: D.FuseDRNSegBase("drn_d_22", N_CLASS, pretrained=False, input_ch=6)),
                      ("drn_c_26", lambda: D.DRNSegBase("drn_c_26", N_CLASS, pretrained=False, input_ch=6))):
        torch.manual_seed(0)
        net = filled(make(), 81)
        x = torch.randn(2, 6, 64, 96, generator=torch.Generator().manual_seed(808))
        net.eval()
        with torch.no_grad():
            out[tag + ":eval"] = net(x).numpy()
        net.train()
        feat = net(x)
        feat.square().mean().backward()
        gk, gn = norms({k: p.grad for k, p in net.named_parameters()})
        sk, sv = summarize(net.state_dict())
        out[tag + ":train"] = feat.detach().numpy()
        out[tag + ":grad_keys"], out[tag + ":grad_norms"] = np.array(gk), gn
        out[tag + ":state_keys"], out[tag + ":state_sums"] = np.array(sk), sv
        out[tag + ":nbt"] = np.array([int(v) for k, v in net.state_dict().items() if k.endswith("num_batches_tracked")])
        print("variant trunk %-10s %d state entries, eval |feat| %.5f" % (tag, len(net.state_dict()),
                                                                           float(np.linalg.norm(out[tag + ":eval"]))))

    # ---- decoders with options, on seeded feature maps (full resolution 32 x 48)
    weight = get_class_weight_from_file(n_class=N_CLASS)
    H, W = 32, 48

    def feats(seed):
        gg = torch.Generator().manual_seed(seed)
        return {k: (torch.randn(2, c, H // d, W // d, generator=gg).clamp_(min=0) * 0.7).requires_grad_(True)
                for k, c, d in (("h2", 32, 2), ("h3", 64, 4), ("h8", 512, 8))}

    gg = torch.Generator().manual_seed(909)
    gt_semseg = torch.randint(0, N_CLASS, (2, H, W), generator=gg)
    gt_semseg[:, 5:20, 10:30] = 7                     # some structure, so that the label boundaries are not everywhere
    gt_semseg[:, 22:, :15] = 3
    gt_dep = torch.randn(2, 3, H, W, generator=gg)
    gt_bd = (torch.rand(2, 1, H, W, generator=gg) < 0.15).float()
    out.update({"dec:gt_semseg": gt_semseg.numpy(), "dec:gt_dep": gt_dep.numpy(), "dec:gt_bd": gt_bd.numpy()})

    def record(tag, dec, fd, total, scalars):
        named = {**{"x:" + k: v for k, v in fd.items()}, **{"p:" + k: p for k, p in dec.named_parameters()}}
        gr = _grads_of(total, named)
        for k, v in scalars.items():
            out[tag + ":" + k] = float(v)
        keys = sorted(k for k, v in gr.items() if v is not None)
        out[tag + ":grad_keys"] = np.array(keys)
        out[tag + ":grad_norms"] = np.array([float(np.linalg.norm(gr[k])) for k in keys])
        out[tag + ":g:x:h8"], out[tag + ":g:x:h2"] = gr["x:h8"], gr["x:h2"]
        print("variant decoder %-14s %s" % (tag, {k: round(float(v), 6) for k, v in scalars.items()}))

    # (a) MCDTripleMultiTaskDecoder, every option on
    dec = filled(D.MCDTripleMultiTaskDecoder(N_CLASS, 3, semseg_criterion=CrossEntropyLoss2d(weight),
                                             discrepancy_criterion=Diff2d(), semseg_shortcut=True,
                                             depth_shortcut=True, add_pred_seg_boundary_loss=True,
                                             use_seg2bd_conv=True), 91)
    dec.train()
    fd = feats(1001)
    l_seg, l_dep, l_bd = dec.get_loss(fd, gt_semseg, gt_dep, gt_bd, separately_returning=True)
    l_x_src = dec.get_boundary_loss_by_extra_conv(fd, gt_bd)
    l_x_tgt = dec.get_boundary_loss_by_extra_conv(fd)
    l_disc = dec.get_cls_descrepancy(fd)
    with torch.no_grad():
        p1, p2, pdep, pbd = dec(fd)
    out["tri_opt:pred1_sub"], out["tri_opt:pdep_sub"] = p1[:, :, ::4, ::4].numpy(), pdep[:, :, ::4, ::4].numpy()
    out["tri_opt:pbd"] = pbd.numpy()
    record("tri_opt", dec, fd, l_seg + l_dep + l_bd + l_x_src + 0.5 * l_x_tgt - l_disc,
           dict(seg=l_seg, dep=l_dep, bd=l_bd, x_src=l_x_src, x_tgt=l_x_tgt, disc=l_disc))

    # (b) MCDSegBDMultiTaskDecoder (default options).  torch >= 1.x refuses F.binary_cross_entropy on [N,1,H,W] vs
    #     [N,H,W]; torch 0.4.1 (the reference's pin) only warned and paired the elements in flat order, which is what
    #     squeezing the channel axis computes.
    dec = filled(D.MCDSegBDMultiTaskDecoder(N_CLASS, 3, semseg_criterion=CrossEntropyLoss2d(weight),
                                            discrepancy_criterion=Diff2d()), 92)
    dec.get_boundary_loss = lambda x_dic, gt: D.get_boundary_loss(pred=dec.boundary_forward(x_dic)[:, 0], gt=gt,
                                                                   pred_type="boundary")
    dec.train()
    fd = feats(1002)
    l_seg, l_bd = dec.get_loss(fd, gt_semseg, separately_returning=True)
    l_disc = dec.get_cls_descrepancy(fd)
    record("segbd", dec, fd, l_seg + l_bd - l_disc, dict(seg=l_seg, bd=l_bd, disc=l_disc))

    # (c) source-only TripleMultiTaskDecoder and MultiTaskDecoder
    dec = filled(D.TripleMultiTaskDecoder(N_CLASS, 3, semseg_criterion=CrossEntropyLoss2d(weight)), 93)
    dec.train()
    fd = feats(1003)
    l_seg, l_dep, l_bd = dec.get_loss(fd, gt_semseg, gt_dep, gt_bd, separately_returning=True)
    record("tri_src", dec, fd, l_seg + l_dep + l_bd, dict(seg=l_seg, dep=l_dep, bd=l_bd))
    dec = D.MultiTaskDecoder(N_CLASS, 3, semseg_criterion=CrossEntropyLoss2d(weight))
    dec.s_semsegcls.data.fill_(1), dec.s_deprgr.data.fill_(1)        # uninitialised memory in the reference
    filled(dec, 94)
    dec.eval()
    with torch.no_grad():
        ps, pd = dec(feats(1004)["h8"])
    out["mt_src:semseg"], out["mt_src:dep"] = ps.numpy(), pd.numpy()
    # ---- ProbCrossEntropyLoss2d (loss.py:16-30): the criterion of the Gate fusions (adapt_mfnet_trainer.py:149)
    from loss import ProbCrossEntropyLoss2d
    gg = torch.Generator().manual_seed(1111)
    p = torch.softmax(torch.randn(2, N_CLASS, 6, 8, generator=gg) * 2, 1).requires_grad_(True)
    t = torch.randint(0, N_CLASS, (2, 6, 8), generator=gg)
    t[0, 0, :3] = -100                                  # NLLLoss2d's default ignore_index
    v = ProbCrossEntropyLoss2d(weight)(p, t)
    v.backward()
    out.update({"pce:p": p.detach().numpy(), "pce:t": t.numpy(), "pce:loss": float(v), "pce:dp": p.grad.numpy()})
    np.savez_compressed(os.path.join(HERE, "variants.npz"), **{k: v for k, v in out.items() if v is not None})


if __name__ == "__main__":
    warnings.simplefilter("ignore")
    torch.set_num_threads(8)
    stage_reference()
    golden_losses()
    golden_early_fusion()
    golden_mfnet()
    golden_triple()
    golden_iterations()
    golden_pipeline()
    golden_discrepancies()
    golden_bottleneck()
    golden_variants()
