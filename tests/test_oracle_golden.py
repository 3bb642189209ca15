"""Pin the oracle (oracle/mcd_oracle.py) to the real reference: every golden vector under tests/golden/ was
produced by tests/golden/make_golden.py running the reference's own modules.  Both sides are fp32 torch CPU
ops on identical weights (fill_state_dict_), so agreement is expected to ~1e-5."""
import os

import numpy as np
import pytest
import torch

from oracle import mcd_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_CLASS = 41


def close(a, b, rtol=2e-4, atol=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b).max() if a.size else 0.0
    assert err <= atol + rtol * np.abs(b).max(), (err, np.abs(b).max())


def _inputs(seed, n=2, ch=6, size=(96, 128)):
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(n, ch, *size, generator=g)
    tgt = torch.randn(n, 6, *size, generator=g)
    lbl = torch.randint(0, N_CLASS, (n, *size), generator=g)
    return src, tgt, lbl


def _norms(keys, grads):
    return np.array([float(grads[k].norm()) if grads.get(k) is not None else -1.0 for k in keys])


def test_losses_match_reference():
    z = np.load(os.path.join(GOLD, "losses.npz"))
    a = torch.tensor(z["a"], requires_grad=True)
    b = torch.tensor(z["b"], requires_grad=True)
    t, w = torch.tensor(z["t"]), torch.tensor(z["w"])
    ce, df = O.ce2d(a, t, w), O.diff2d(a, b)
    (ce + df).backward()
    close(float(ce), z["ce"]), close(float(df), z["diff"])
    close(a.grad.numpy(), z["da"]), close(b.grad.numpy(), z["db"])
    p = torch.tensor(z["p"], requires_grad=True)
    bc = O.bce2d(p, torch.tensor(z["tb"]))
    bc.backward()
    close(float(bc), z["bce"]), close(p.grad.numpy(), z["dp"])


def test_state_dict_keys_match_reference():
    z = np.load(os.path.join(GOLD, "early_fusion.npz"))
    G = O.init_seg_base("drn_d_38", 6, N_CLASS)
    assert sorted(k for k in G if torch.is_floating_point(G[k])) == list(z["final_g_keys"])
    assert sorted(O.trainable(G)) == list(z["A_grad_g_keys"])


@pytest.mark.timeout(900)
def test_early_fusion_iteration_matches_reference():
    z = np.load(os.path.join(GOLD, "early_fusion.npz"))
    G = O.fill_state_dict_(O.init_seg_base("drn_d_38", 6, N_CLASS), 11)
    F1 = O.fill_state_dict_(O.init_head(N_CLASS), 12)
    F2 = O.fill_state_dict_(O.init_head(N_CLASS), 13)
    src, tgt, lbl = _inputs(101)
    rec = {}
    c_loss, d_loss = O.mcd_step_early(G, F1, F2, src, lbl, tgt, O.class_weight(N_CLASS), O.SGD(), O.SGD(),
                                      num_k=4, record=rec)
    close(float(rec["A_loss"]), z["A_loss"])
    close(rec["A_feat"].numpy(), z["A_feat"])
    close(rec["A_out1"][:, :, ::8, ::8].numpy(), z["A_out1_sub"])
    close(_norms(list(z["A_grad_g_keys"]), rec["A_grad_g"]), z["A_grad_g_norms"], rtol=1e-3)
    close(rec["A_grad_f1"]["up.weight"].numpy(), z["A_grad_up1"])
    close(rec["A_grad_g"]["seg.bias"].numpy(), z["A_grad_seg_bias"])
    close(float(rec["B_loss"]), z["B_loss"])
    close(rec["B_grad_f1"]["up.weight"].numpy(), z["B_grad_up1"])
    close(np.array(rec["C_losses"]), z["C_losses"], rtol=1e-3)
    close(_norms(list(z["A_grad_g_keys"]), rec["C0_grad_g"]), z["C0_grad_g_norms"], rtol=2e-3)
    sums = np.array([[float(G[k].double().sum()), float(G[k].double().norm())] for k in z["final_g_keys"]])
    close(sums[:, 1], z["final_g_sums"][:, 1], rtol=1e-5)
    close(sums[:, 0], z["final_g_sums"][:, 0], rtol=1e-4, atol=1e-3)
    close(F1["up.weight"].numpy(), z["final_up1"])
    assert abs(c_loss - float(z["A_loss"])) < 1e-4 and abs(d_loss - z["C_losses"][-1] / 4) < 1e-6
    # tester path: eval-mode forward, argmax without background, entropy (adapt_tester.py:104-124)
    with torch.no_grad():
        o = O.head_forward(F1, O.seg_base_forward(G, tgt[:1], train=False))
    labels = O.predict_labels(o, N_CLASS - 1)[0].numpy()
    assert (labels == z["test_labels"]).mean() > 0.999
    close(float(O.calc_entropy(o)), z["test_entropy"], rtol=1e-3)


@pytest.mark.parametrize("tag,kind", [("add", "add"), ("scoreadd", "scoreadd")])
def test_mfnet_heads_match_reference(tag, kind):
    z = np.load(os.path.join(GOLD, "mfnet.npz"))
    G3 = O.fill_state_dict_(O.init_seg_base("drn_d_38", 3, N_CLASS), 21)
    G1 = O.fill_state_dict_(O.init_seg_base("drn_d_38", 3, N_CLASS), 22)
    F1 = O.fill_state_dict_(O.init_head(N_CLASS, kind), 23)
    F2 = O.fill_state_dict_(O.init_head(N_CLASS, kind), 24)
    src, tgt, lbl = _inputs(202, size=(64, 96))
    w = O.class_weight(N_CLASS)
    O._req([G3, G1, F1, F2])
    o3, o1 = O.seg_base_forward(G3, src[:, :3]), O.seg_base_forward(G1, src[:, 3:])
    p1, p2 = O.head_forward(F1, (o3, o1), kind), O.head_forward(F2, (o3, o1), kind)
    ce = O.ce2d(p1, lbl, w) + O.ce2d(p2, lbl, w)
    t3, t1 = O.seg_base_forward(G3, tgt[:, :3]), O.seg_base_forward(G1, tgt[:, 3:])
    d = O.diff2d(O.head_forward(F1, (t3, t1), kind), O.head_forward(F2, (t3, t1), kind))
    _, gg1, gf1, _ = O._grads(ce - d, [G3, G1, F1, F2])
    close(float(ce), z[tag + "_ce"]), close(float(d), z[tag + "_diff"], rtol=1e-3)
    close(o3.detach().numpy(), z[tag + "_feat3"])
    close(p1.detach()[:, :, ::8, ::8].numpy(), z[tag + "_p1_sub"])
    close(_norms(list(z[tag + "_grad_g1_keys"]), gg1), z[tag + "_grad_g1_norms"], rtol=2e-3)
    for k, g in gf1.items():
        close(g.numpy(), z[tag + "_grad_f1_" + k], rtol=1e-3)


def test_triple_multitask_matches_reference():
    z = np.load(os.path.join(GOLD, "triple.npz"))
    E = O.fill_state_dict_(O.init_trunk("drn_d_38", 3, "main_layer"), 31)
    D = O.fill_state_dict_(O.init_triple_decoder(N_CLASS, 3), 32)
    g = torch.Generator().manual_seed(303)
    size = (64, 96)
    src = torch.randn(2, 7, *size, generator=g)
    src[:, 6] = (torch.rand(2, *size, generator=g) < 0.1).float()
    tgt = torch.randn(2, 6, *size, generator=g)
    lbl = torch.randint(0, N_CLASS, (2, *size), generator=g)
    w = O.class_weight(N_CLASS)
    O._req([E, D])
    src_f, tgt_f = O.encoder_dict(E, src[:, :3]), O.encoder_dict(E, tgt[:, :3])
    semseg, dep, bd = O.triple_get_loss(D, src_f, lbl, src[:, 3:-1], src[:, -1:], w)
    tgt_dep = torch.nn.functional.mse_loss(O.triple_depth(D, tgt_f), tgt[:, 3:])
    disc = O.diff2d(*O.triple_semseg(D, tgt_f))
    ge, gd = O._grads(semseg + dep + bd + tgt_dep - disc, [E, D])
    for name, v in (("semseg", semseg), ("dep", dep), ("bd", bd), ("tgt_dep", tgt_dep), ("disc", disc)):
        close(float(v), z[name], rtol=5e-4)
    close(src_f["h8"].detach()[:, ::16].numpy(), z["h8_sub"])
    close(_norms(list(z["grad_enc_keys"]), ge), z["grad_enc_norms"], rtol=2e-3)
    gd_n = _norms(list(z["grad_dec_keys"]), gd)
    close(gd_n, z["grad_dec_norms"], rtol=2e-3)
    assert (z["grad_dec_norms"] < 0).sum() == 10  # nmlrgr_dec never receives a gradient (reference :813)
    O._req([E, D], False)
    with torch.no_grad():
        f = O.encoder_dict(E, tgt[:1, :3], train=False)
        s1, _ = O.triple_semseg(D, f, train=False)
        depth, boundary = O.triple_depth(D, f, train=False), O.triple_boundary(D, f)
    assert (O.predict_labels(s1, N_CLASS - 1)[0].numpy() == z["test_labels"]).mean() > 0.999
    close(float(O.calc_entropy(s1)), z["test_entropy"], rtol=1e-3)
    close(depth[:, :, ::8, ::8].numpy(), z["test_depth_sub"])
    close(boundary[:, :, ::8, ::8].numpy(), z["test_boundary_sub"])


# ---- full iterations of the other three trainers + get_boundary_loss (tests/golden/iterations.npz) ---------------
def _check_sums(keys, sums, sd):
    for k, (s_ref, n_ref) in zip(keys, sums):
        t = sd[str(k)].double()
        assert abs(float(t.norm()) - n_ref) <= 1e-6 + 2e-5 * abs(n_ref), (k, float(t.norm()), n_ref)
        assert abs(float(t.sum()) - s_ref) <= 1e-4 + 2e-4 * abs(n_ref), (k, float(t.sum()), s_ref)


@pytest.mark.timeout(900)
def test_mfnet_iteration_matches_reference():
    """adapt_mfnet_trainer.py:181-235 (ScoreAddFusion), A + B + 2 x C with torch.optim.SGD."""
    z = np.load(os.path.join(GOLD, "iterations.npz"))
    G3 = O.fill_state_dict_(O.init_seg_base("drn_d_38", 3, N_CLASS), 41)
    G1 = O.fill_state_dict_(O.init_seg_base("drn_d_38", 3, N_CLASS), 42)
    F1 = O.fill_state_dict_(O.init_head(N_CLASS, "scoreadd"), 43)
    F2 = O.fill_state_dict_(O.init_head(N_CLASS, "scoreadd"), 44)
    src, tgt, lbl = _inputs(505, size=(48, 64))
    rec = {}
    og, of = O.SGD(), O.SGD()
    c, d = O.mcd_step_mfnet(G3, G1, F1, F2, src, lbl, tgt, O.class_weight(N_CLASS), og, of, kind="scoreadd", num_k=2,
                            record=rec)
    close(rec["A_loss"], z["mf_A"]), close(rec["B_loss"], z["mf_B"]), close(rec["C_losses"], z["mf_C"])
    _check_sums(z["mf_g1_keys"], z["mf_g1_sums"], G1)
    close(F1["up1.weight"].numpy(), z["mf_up1"])


@pytest.mark.timeout(900)
@pytest.mark.parametrize("tag,triple", [("tri", True), ("mt", False)])
def test_multitask_iteration_matches_reference(tag, triple):
    """adapt_triple_multitask_trainer.py:202-287 / adapt_multitask_trainer.py:194-262, A + B + 2 x C."""
    z = np.load(os.path.join(GOLD, "iterations.npz"))
    E = O.fill_state_dict_(O.init_trunk("drn_d_38", 3, "main_layer" if triple else "base."), 51)
    D = O.fill_state_dict_(O.init_triple_decoder(N_CLASS, 3) if triple else O.init_multitask_decoder(N_CLASS, 3), 52)
    g = torch.Generator().manual_seed(606)
    size = (48, 64)
    src = torch.randn(2, 7 if triple else 6, *size, generator=g)
    if triple:
        src[:, 6] = (torch.rand(2, *size, generator=g) < 0.1).float()
    tgt = torch.randn(2, 6, *size, generator=g)
    lbl = torch.randint(0, N_CLASS, (2, *size), generator=g)
    rec = {}
    c, d = O.mcd_step_multitask(E, D, src, lbl, tgt, O.class_weight(N_CLASS), O.SGD(), O.SGD(), triple=triple, num_k=2,
                                record=rec)
    close(rec["A_loss"], z[tag + "_A"]), close(rec["B_loss"], z[tag + "_B"]), close(rec["C_losses"], z[tag + "_C"])
    _check_sums(z[tag + "_enc_keys"], z[tag + "_enc_sums"], E)
    _check_sums(z[tag + "_dec_keys"], z[tag + "_dec_sums"], D)
    if triple:        # never used, never updated (reference :813)
        assert rec["A_grad_d"]["nmlrgr_dec.conv3.weight"] is None


def test_boundary_loss_matches_reference():
    z = np.load(os.path.join(GOLD, "iterations.npz"))
    lab_p, lab_g, bmap = torch.tensor(z["bd_lab_p"]), torch.tensor(z["bd_lab_g"]), torch.tensor(z["bd_map"])
    assert np.array_equal(O.label_boundary(lab_p).numpy(), z["bd_boundary_of_p"])     # integer part: bit-exact
    close(float(O.get_boundary_loss(lab_p, lab_g)), z["bd_ss"])
    close(float(O.get_boundary_loss(lab_p, bmap, gt_type="boundary")), z["bd_sb"])


# ---- byte-side neighbours of the step: input transform, label transform, evaluation counts ----
def _pipe():
    return np.load(os.path.join(GOLD, "pipeline.npz"))


def test_oracle_input_transform_matches_reference_bitwise():
    d = _pipe()
    img6 = O.assemble_input(d["rgb"], d["hha"]).numpy()
    assert np.array_equal(img6, d["img6"])
    img7 = O.assemble_input(d["rgb"], d["hha"], d["bd"]).numpy()
    assert np.array_equal(img7, d["img7"])
    assert np.array_equal(O.img_transform(d["rgb"], "city").numpy(), d["img3_city"])
    assert np.array_equal(O.lbl_transform(d["lbl"], 41).numpy(), d["lbl_out"])
    assert np.array_equal(O.unnormalize(np.transpose(img6[:3], (1, 2, 0))), d["unnorm"])


def test_oracle_eval_counts_match_reference():
    d = _pipe()
    hist = O.fast_hist(d["h_gt"], d["h_pred"], 40)
    assert np.array_equal(hist, d["hist"])
    assert np.array_equal(O.per_class_iu(hist), d["iu"])


def test_oracle_nearest_resize_matches_pil():
    from PIL import Image
    d = _pipe()
    for i in range(5):
        m, ref = d["rs%d_in" % i], d["rs%d_out" % i]
        assert np.array_equal(O.resize_nearest(m, (ref.shape[1], ref.shape[0])), ref), i
    rng = np.random.RandomState(0)          # and against PIL itself (same image here and on the GPU box)
    for (ih, iw), (oh, ow) in [((480, 640), (530, 730)), ((480, 640), (425, 560)), ((60, 80), (480, 640)), ((33, 47), (100, 31))]:
        m = rng.randint(0, 41, (ih, iw)).astype(np.uint8)
        ref = np.array(Image.fromarray(m).resize((ow, oh), Image.NEAREST))
        assert np.array_equal(O.resize_nearest(m, (ow, oh)), ref), ((ih, iw), (oh, ow))


DISCREPANCIES = ("diff", "jsd", "symkl", "nmlsymkl", "mysymkl", "spatial_jsd", "mis_symkl")


@pytest.mark.parametrize("name", DISCREPANCIES)
def test_oracle_discrepancy_criteria_match_reference(name):
    """loss.py:68-171 through get_prob_distance_criterion of the real reference (tests/golden/discrepancies.npz)"""
    d = np.load(os.path.join(GOLD, "discrepancies.npz"))
    a = torch.from_numpy(d["a"]).requires_grad_(True)
    b = torch.from_numpy(d["b"]).requires_grad_(True)
    v = O.pair_distance(name, a, b)
    v.backward()
    assert abs(float(v) - float(d[name])) <= 1e-6 * abs(float(d[name]))
    close(a.grad.numpy(), d[name + "_da"], rtol=1e-5, atol=1e-12)
    close(b.grad.numpy(), d[name + "_db"], rtol=1e-5, atol=1e-12)


def test_oracle_bottleneck_trunk_matches_reference():
    """DRN-D-54 (Bottleneck, models/drn.py:62-100) restated in the oracle vs the reference's DRNSegBase("drn_d_54")"""
    d = np.load(os.path.join(GOLD, "drn_d_54.npz"))
    G = O.fill_state_dict_(O.init_seg_base("drn_d_54", 6, N_CLASS), 54)
    assert len(G) == 344
    O._req([G], True)
    x = torch.randn(2, 6, 64, 96, generator=torch.Generator().manual_seed(540))
    feat = O.seg_base_forward(G, x, name="drn_d_54")
    close(feat.detach().numpy(), d["feat"], rtol=1e-5)
    grads = O._grads(feat.square().mean(), [G])[0]
    keys = [str(k) for k in d["grad_keys"]]
    got = np.array([float(grads[k].norm()) for k in keys])
    close(got, d["grad_norms"], rtol=2e-4)
    sk = [str(k) for k in d["state_keys"]]
    got_s = np.array([[float(G[k].detach().double().sum()), float(G[k].detach().double().norm())] for k in sk])
    close(got_s, d["state_sums"], rtol=1e-5)          # incl. the running statistics after one train-mode forward


# ---- option surface around the hot path: oracle restatements vs tests/golden/variants.npz ---------
def _variants():
    return np.load(os.path.join(GOLD, "variants.npz"), allow_pickle=False)


_HEAD_CASES = [("gate_v1", "fusion", "gate", "ver1", False, 41), ("concat_v1", "fusion", "concat", "ver1", False, 41),
               ("concatconv_v1", "fusion", "concatconv", "ver1", False, 41), ("add_v2", "fusion", "add", "ver2", False, 512),
               ("gate_v2", "fusion", "gate", "ver2", False, 512), ("add_torchup", "fusion", "add", "ver1", True, 41),
               ("scoregate", "score", "scoregate", "ver1", False, 41), ("scoregate_nosm", "score", "gate", "ver1", False, 41),
               ("single_v2", "single", None, "ver2", False, 512), ("single_torchup", "single", None, "ver1", True, 41)]


def head_state(tag, cls, kind, ver, torch_up, cin, n_class=41):
    """state_dict of one head case with the reference's key names and shapes, filled like make_golden.filled()."""
    sd = {}
    if kind in ("gate", "scoregate"):
        sd["fusion.conv.weight"], sd["fusion.conv.bias"] = torch.empty(cin, 2 * cin, 1, 1), torch.empty(cin)
    if kind == "concatconv":
        sd["fusion.conv.weight"], sd["fusion.conv.bias"] = torch.empty(cin, 2 * cin, 3, 3), torch.empty(cin)
    if ver == "ver2":
        sd["seg.weight"], sd["seg.bias"] = torch.empty(n_class, 512, 1, 1), torch.empty(n_class)
    if cls == "score":
        sd["up1.weight"], sd["up2.weight"] = torch.empty(n_class, 1, 16, 16), torch.empty(n_class, 1, 16, 16)
    elif not torch_up:
        sd["up.weight"] = torch.empty(2 * n_class if kind == "concat" else n_class, 1, 16, 16)
    return sd


def head_case_inputs(i, cin, h=6, w=8, n_class=41):
    """tests/golden/make_golden.py::variant_head_inputs"""
    g = torch.Generator().manual_seed(7100 + i)
    scale = 2.0 if cin != 512 else 0.5
    x1 = torch.randn(2, cin, h, w, generator=g) * scale
    x2 = torch.randn(2, cin, h, w, generator=g) * scale
    if cin == 512:
        x1, x2 = x1.clamp_(min=0), x2.clamp_(min=0)
    r = torch.randn(2, n_class, 8 * h, 8 * w, generator=g)
    return x1.requires_grad_(True), x2.requires_grad_(True), r


def _sample(a, n=512):
    f = np.asarray(a).reshape(-1)
    return f[::max(1, f.size // n)]


@pytest.mark.parametrize("i", range(len(_HEAD_CASES)))
def test_oracle_fusion_heads_match_reference(i):
    tag, cls, kind, ver, torch_up, cin = _HEAD_CASES[i]
    z = _variants()
    sd = O.fill_state_dict_(head_state(tag, cls, kind, ver, torch_up, cin), 70 + i)
    for v in sd.values():
        v.requires_grad_(True)
    x1, x2, r = head_case_inputs(i, cin)
    if cls == "fusion":
        out = O.fusion_head_forward(sd, kind, x1, x2, ver, torch_up)
    elif cls == "score":
        out = O.score_fusion_head_forward(sd, kind, x1, x2)
    else:
        out = O.single_head_forward(sd, x1, ver, torch_up)
    (out * r).mean().backward()
    assert np.allclose(out.detach()[:, :, ::4, ::4].numpy(), z[tag + ":out_sub"], rtol=1e-4, atol=1e-5)
    named = {"x1": x1, **({"x2": x2} if cls != "single" else {}), **{"p:" + k: v for k, v in sd.items()}}
    for k, t in named.items():
        assert np.allclose(_sample(t.grad.numpy()), z[tag + ":gs:" + k], rtol=1e-3, atol=1e-7), (tag, k)


@pytest.mark.parametrize("tag,name", [("fusenet", "drn_d_22"), ("drn_c_26", "drn_c_26")])
def test_oracle_fusenet_and_arch_c_match_reference(tag, name):
    """FuseDRNSegBase (models/dilated_fcn.py:253-337) and DRN arch C through DRNSegBase: eval and train forward,
    BatchNorm bookkeeping (two updates per stage and forward for the fusenet generator)."""
    z = _variants()
    keys = [str(k) for k in z[tag + ":state_keys"]]
    x = torch.randn(2, 6, 64, 96, generator=torch.Generator().manual_seed(808))
    fwd = O.fuse_seg_base_forward if tag == "fusenet" else O.seg_base_forward

    def state():
        sd = {}
        if tag == "fusenet":
            for pre, ch in (("main_layer", 3), ("sub_layer", 3)):
                sd.update(O.init_trunk(name, ch, pre))
            sd["seg.weight"], sd["seg.bias"] = torch.empty(41, 512, 1, 1), torch.empty(41)
        else:
            sd = _arch_c_state(name)
        assert set(k for k in sd if torch.is_floating_point(sd[k])) == set(keys)
        return O.fill_state_dict_(sd, 81)

    with torch.no_grad():
        ev = fwd(state(), x, name=name, train=False)
    close(ev.numpy(), z[tag + ":eval"], rtol=2e-4)
    sd = state()
    with torch.no_grad():
        tr = fwd(sd, x, name=name, train=True)
    ref = z[tag + ":train"]
    assert float(np.abs(tr.numpy() - ref).max()) <= 2e-3 * float(np.abs(ref).max())
    nbt = sorted(int(v) for k, v in sd.items() if k.endswith("num_batches_tracked"))
    assert nbt == sorted(int(v) for v in z[tag + ":nbt"])
    sums = {k: v for k, v in zip(keys, z[tag + ":state_sums"])}
    for k in keys:
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert abs(float(sd[k].double().norm()) - sums[k][1]) <= 2e-3 * sums[k][1] + 1e-6, k


def _arch_c_state(name):
    """state_dict shapes of DRNSegBase(drn_c_*, input_ch=6) from the oracle's architecture description"""
    sd, cin = {}, 6
    for stage in O.trunk_spec(name, "base."):
        for unit in stage:
            if unit[0] == "cbr":
                sd[unit[1] + ".weight"] = torch.empty(16, cin, 7, 7)
                O._bn_state(sd, unit[2], 16)
                cin = 16
                continue
            p, ds = unit[1], unit[5]
            planes = O.CHANNELS[int(p.split(".")[1]) - 3]
            assert unit[0] == "block"
            sd[p + ".conv1.weight"] = torch.empty(planes, cin, 3, 3)
            O._bn_state(sd, p + ".bn1", planes)
            sd[p + ".conv2.weight"] = torch.empty(planes, planes, 3, 3)
            O._bn_state(sd, p + ".bn2", planes)
            if ds:
                sd[p + ".downsample.0.weight"] = torch.empty(planes, cin, 1, 1)
                O._bn_state(sd, p + ".downsample.1", planes)
            cin = planes
    sd["seg.weight"], sd["seg.bias"] = torch.empty(41, 512, 1, 1), torch.empty(41)
    return sd


def decoder_option_feats(seed, H=32, W=48):
    """tests/golden/make_golden.py::golden_variants.feats"""
    gg = torch.Generator().manual_seed(seed)
    return {k: (torch.randn(2, c, H // d, W // d, generator=gg).clamp_(min=0) * 0.7).requires_grad_(True)
            for k, c, d in (("h2", 32, 2), ("h3", 64, 4), ("h8", 512, 8))}


def decoder_option_state(kind, seed):
    """state_dict (reference key names / shapes) of MCDTripleMultiTaskDecoder with every option on ("tri_opt"),
    MCDSegBDMultiTaskDecoder ("segbd") or the source-only TripleMultiTaskDecoder ("tri_src")."""
    g = torch.Generator().manual_seed(0)
    sd = {"s_semsegcls": torch.ones(1), "s_boundary": torch.ones(1)}
    decs = {"tri_opt": ("semsegcls_dec1", "semsegcls_dec2", "deprgr_dec", "nmlrgr_dec"),
            "segbd": ("semsegcls_dec1", "semsegcls_dec2"), "tri_src": ("semsegcls_dec", "deprgr_dec", "nmlrgr_dec")}[kind]
    for d in decs:
        O.init_three_layer_decoder(sd, d, N_CLASS if d.startswith("semseg") else 3, g)
    if kind != "segbd":
        sd["s_deprgr"] = torch.ones(1)
    for i, c in ((1, 32), (2, 64), (3, 512)):
        O._default_conv(sd, "conv%d" % i, 1, c, 1, g)
        if kind == "tri_opt":
            for nm in ("seg_conv%d_1", "seg_conv%d_2", "dep_conv%d"):
                O._default_conv(sd, nm % i, 512, c, 1, g)
    if kind == "tri_opt":
        sd["s_pred_seg_boundary"] = torch.ones(1)
        O._default_conv(sd, "seg2bd_conv", 1, N_CLASS, 5, g)
    return O.fill_state_dict_(sd, seed)


def decoder_option_losses(kind, sd, hd, z):
    """the scalars tests/golden/make_golden.py::golden_variants records for decoder case `kind`, from the oracle"""
    import torch.nn.functional as F
    gt_semseg, gt_dep, gt_bd = (torch.tensor(z["dec:gt_semseg"]).to(hd["h8"].device),
                                torch.tensor(z["dec:gt_dep"]).to(hd["h8"].device),
                                torch.tensor(z["dec:gt_bd"]).to(hd["h8"].device))
    weight = O.class_weight(N_CLASS).to(hd["h8"].device)
    opt = kind == "tri_opt"
    decs = (("semsegcls_dec", "seg_conv%d_1"),) if kind == "tri_src" else \
        (("semsegcls_dec1", "seg_conv%d_1"), ("semsegcls_dec2", "seg_conv%d_2"))
    ls, preds = O.opt_semseg_losses(sd, hd, gt_semseg, weight, shortcut=opt,
                                       add_pred_seg_boundary_loss=opt or kind == "segbd", decs=decs)
    out = {"seg": sum(O._uw(sd["s_semsegcls"], l) for l in ls) / len(ls)}
    if kind != "segbd":
        dep_in = O._shortcut_sum(sd, hd, "dep_conv%d") if opt else hd["h8"]
        dep = O.three_layer_decoder(sd, "deprgr_dec", dep_in)
        out["dep"] = O._uw(sd["s_deprgr"], F.mse_loss(dep if opt else O.bilinear_up(dep, 8), gt_dep))
        out["bd"] = O._uw(sd["s_boundary"], O.bce2d(O.triple_boundary(sd, hd), gt_bd))
    else:
        out["bd"] = O._uw(sd["s_boundary"], O.get_boundary_loss(O.triple_boundary(sd, hd)[:, 0], gt_semseg,
                                                                pred_type="boundary"))
    if opt:
        out["x_src"] = sum(O.seg2bd_losses(sd, preds, gt_bd))
        out["x_tgt"] = sum(O.seg2bd_losses(sd, preds, O.triple_boundary(sd, hd).detach()))
    if kind != "tri_src":
        out["disc"] = O.diff2d(preds[0], preds[1])
    total = out["seg"] + out["bd"] + out.get("dep", 0) + out.get("x_src", 0) + 0.5 * out.get("x_tgt", 0) - out.get("disc", 0)
    return out, total, preds


@pytest.mark.parametrize("kind,seed,fseed", [("tri_opt", 91, 1001), ("segbd", 92, 1002), ("tri_src", 93, 1003)])
def test_oracle_decoder_options_match_reference(kind, seed, fseed):
    z = _variants()
    sd = decoder_option_state(kind, seed)
    for k in O.trainable(sd):
        sd[k].requires_grad_(True)
    hd = decoder_option_feats(fseed)
    out, total, preds = decoder_option_losses(kind, sd, hd, z)
    for k, v in out.items():
        assert abs(float(v) - float(z["%s:%s" % (kind, k)])) <= 2e-4 * abs(float(z["%s:%s" % (kind, k)])) + 1e-6, k
    total.backward()
    named = {**{"x:" + k: v for k, v in hd.items()}, **{"p:" + k: sd[k] for k in O.trainable(sd)}}
    keys = [str(k) for k in z[kind + ":grad_keys"]]
    assert sorted(k for k, v in named.items() if v.grad is not None) == keys
    got = np.array([float(named[k].grad.norm()) for k in keys])
    close(got, z[kind + ":grad_norms"], rtol=2e-3)
    close(hd["h8"].grad.numpy(), z[kind + ":g:x:h8"], rtol=2e-3)
    if kind == "tri_opt":
        close(preds[0].detach()[:, :, ::4, ::4].numpy(), z["tri_opt:pred1_sub"], rtol=1e-3)


def test_oracle_prob_cross_entropy_matches_reference():
    z = _variants()
    p = torch.tensor(z["pce:p"], requires_grad=True)
    v = O.prob_ce2d(p, torch.tensor(z["pce:t"]), O.class_weight(N_CLASS))
    v.backward()
    assert abs(float(v) - float(z["pce:loss"])) <= 1e-6 * abs(float(z["pce:loss"]))
    close(p.grad.numpy(), z["pce:dp"], rtol=1e-6)


def test_variant_modules_keep_the_reference_state_dict_layout():
    """drop-in contract of the option surface, checkable without a GPU: the modules of multichannel-semseg-with-uda_b200
    have exactly the state_dict keys and shapes of the reference's classes (recorded in variants.npz / restated by
    head_state), so the reference's checkpoints load."""
    import warnings
    from models import dilated_fcn as D
    from models.model_util import get_models
    z = _variants()
    fusion_type = {"gate": "GateFusion", "scoregate": "ScoreGateFusion", "add": "AddFusion", "concat": "ConcatFusion",
                   "concatconv": "ConcatConvFusion", None: None}
    for tag, cls, kind, ver, torch_up, cin in _HEAD_CASES:
        if cls == "fusion":
            m = D.FusionDRNSegPixelClassifier(fusion_type[kind], N_CLASS, use_torch_up=torch_up, ver=ver)
        elif cls == "score":
            m = D.ScoreFusionDRNSegPixelClassifier(fusion_type[kind], N_CLASS)
        else:
            m = D.DRNSegPixelClassifier(N_CLASS, use_torch_up=torch_up, ver=ver)
        want = head_state(tag, cls, kind, ver, torch_up, cin)
        got = m.state_dict()
        assert set(got) == set(want), (tag, set(got) ^ set(want))
        assert all(tuple(got[k].shape) == tuple(want[k].shape) for k in want), tag
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nets = {"fusenet": get_models("drn_d_22_fusenet", 6, N_CLASS)[0], "drn_c_26": get_models("drn_c_26", 6, N_CLASS)[0]}
    for tag, net in nets.items():
        keys = sorted(k for k, v in net.state_dict().items() if torch.is_floating_point(v))
        assert keys == [str(k) for k in z[tag + ":state_keys"]], tag
    decs = {"tri_opt": D.MCDTripleMultiTaskDecoder(N_CLASS, 3, semseg_shortcut=True, depth_shortcut=True,
                                                   add_pred_seg_boundary_loss=True, use_seg2bd_conv=True),
            "segbd": D.MCDSegBDMultiTaskDecoder(N_CLASS, 3), "tri_src": D.TripleMultiTaskDecoder(N_CLASS, 3)}
    for tag, dec in decs.items():
        want = decoder_option_state(tag, 0)
        got = {k: v for k, v in dec.state_dict().items() if "criterion" not in k}
        assert set(got) == set(want), (tag, set(got) ^ set(want))
        assert all(tuple(got[k].shape) == tuple(want[k].shape) for k in want), tag
        trained = set(k[2:] for k in (str(s_) for s_ in z[tag + ":grad_keys"]) if k.startswith("p:"))
        assert trained <= set(k for k, _ in dec.named_parameters())
