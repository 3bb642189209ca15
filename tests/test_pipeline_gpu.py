"""Input pipeline and evaluation counts on the GPU (csrc/pipeline.cu through mcd_b200.pipeline / transform.py / eval.py)
against the oracle restatements, which tests/test_oracle_golden.py pins bit for bit to the reference's transform.py,
eval.py and PIL (tests/golden/pipeline.npz).  Everything is byte / integer / IEEE-exact fp32 work: all comparisons are
array_equal.

  transform.py:302-325, datasets.py:667-695   uint8 HWC planes -> normalised input (+ boundary channel), labels
  adapt_tester.py:124-126                     NEAREST resize of the predicted label map
  eval.py:21-47                               confusion matrix counts and the scores derived from them
  transform.py:285-294                        unnormalize
  MCDStep(input_pipeline=...)                 the same iteration from uint8 inputs as from the loader's fp32 tensors
"""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import mcd_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_CLASS = 41


def _planes(seed, n, h, w, dev):
    rng = np.random.RandomState(seed)
    rgb = rng.randint(0, 256, (n, h, w, 3)).astype(np.uint8)
    hha = rng.randint(0, 256, (n, h, w, 3)).astype(np.uint8)
    lbl = rng.randint(0, N_CLASS, (n, h, w)).astype(np.uint8)
    lbl[rng.rand(n, h, w) < 0.15] = 255
    bd = np.where(rng.rand(n, h, w) < 0.1, 255, 0).astype(np.uint8)
    host = dict(rgb=rgb, hha=hha, lbl=lbl, bd=bd)
    return host, {k: torch.from_numpy(v).to(dev) for k, v in host.items()}


def test_golden_fixture_through_the_kernels(cuda_dev):
    """the committed reference outputs (tests/golden/pipeline.npz) reproduced by the CUDA path"""
    import eval as E
    from mcd_b200 import pipeline
    d = np.load(os.path.join(GOLD, "pipeline.npz"))
    t = {k: torch.from_numpy(d[k]).to(cuda_dev) for k in ("rgb", "hha", "lbl", "bd", "h_gt", "h_pred")}
    img6 = pipeline.transform_images([t["rgb"][None], t["hha"][None]], out="nchw")[0]
    assert np.array_equal(img6.cpu().numpy(), d["img6"])
    img7 = pipeline.transform_images([t["rgb"][None], t["hha"][None]], label_planes=[t["bd"][None]], out="nchw")[0]
    assert np.array_equal(img7.cpu().numpy(), d["img7"])
    city = pipeline.transform_images([t["rgb"][None]], "city", out="nchw")[0]
    assert np.array_equal(city.cpu().numpy(), d["img3_city"])
    assert np.array_equal(pipeline.relabel(t["lbl"], N_CLASS).cpu().numpy(), d["lbl_out"])
    assert np.array_equal(pipeline.unnormalize(img6[None, :3].contiguous())[0].cpu().numpy(), d["unnorm"])
    hist = E.fast_hist(t["h_gt"], t["h_pred"], 40)
    assert np.array_equal(hist, d["hist"])
    assert np.array_equal(E.per_class_iu(hist), d["iu"], equal_nan=True)
    assert E.calc_fw_iu(hist) == d["fw_iu"] and E.calc_pixel_accuracy(hist) == d["pix_acc"]
    assert E.calc_mean_accuracy(hist) == d["mean_acc"]
    for i in range(5):
        m, ref = torch.from_numpy(d["rs%d_in" % i]).to(cuda_dev), d["rs%d_out" % i]
        out = pipeline.resize_nearest(m[None], (ref.shape[1], ref.shape[0]))[0]
        assert np.array_equal(out.cpu().numpy(), ref), i


@pytest.mark.parametrize("shape", [(2, 96, 128), (3, 37, 53), (1, 480, 640)])
def test_input_transform_vs_oracle(cuda_dev, shape):
    """ragged sizes (tail block, odd widths) and the full 480 x 640 frame; NCHW fp32 bit-exact, the NHWC 16-bit operand
    forms equal the correctly rounded fp32 values (IEEE half and bfloat16, round to nearest even), padding is zero"""
    import transform as T
    from mcd_b200 import pipeline
    n, h, w = shape
    host, dev = _planes(5, n, h, w, cuda_dev)
    ref6 = torch.stack([O.assemble_input(host["rgb"][i], host["hha"][i]) for i in range(n)])
    ref7 = torch.stack([O.assemble_input(host["rgb"][i], host["hha"][i], host["bd"][i]) for i in range(n)])
    out6 = pipeline.transform_images([dev["rgb"], dev["hha"]], out="nchw")
    out7 = pipeline.transform_images([dev["rgb"], dev["hha"]], label_planes=[dev["bd"]], out="nchw")
    assert torch.equal(out6.cpu(), ref6) and torch.equal(out7.cpu(), ref7)
    # the reference-shaped callables: one image plane at a time, [N,H,W,c] or [H,W,c]
    img_t, lbl_t = T.get_img_transform((w, h), "imagenet"), T.get_lbl_transform((w, h), N_CLASS)
    assert torch.equal(img_t(dev["hha"]).cpu(), ref6[:, 3:])
    assert torch.equal(img_t(dev["rgb"][0]).cpu(), ref6[0, :3])
    ref_l = torch.stack([O.lbl_transform(host["lbl"][i], N_CLASS) for i in range(n)])
    got_l = lbl_t(dev["lbl"])
    assert got_l.dtype == torch.int64 and torch.equal(got_l.cpu(), ref_l)
    # operand form
    with torch.enable_grad():
        xb = pipeline.transform_images([dev["rgb"], dev["hha"]])
    assert xb.dtype == torch.bfloat16 and tuple(xb.shape) == (n, 8, h, w) and xb._mcd_h16.dtype == torch.float16
    assert torch.equal(xb[:, :6].float().cpu(), ref6.bfloat16().float())
    assert torch.equal(xb._mcd_h16[:, :6].float().cpu(), ref6.half().float())
    assert float(xb[:, 6:].abs().max()) == 0.0 and float(xb._mcd_h16[:, 6:].abs().max()) == 0.0
    with torch.no_grad():
        x16 = pipeline.transform_images([dev["rgb"], dev["hha"]])
    assert x16.dtype == torch.float16 and torch.equal(x16, xb._mcd_h16)
    # MFNet streams: channel sub-ranges of a 6-channel plane
    six = torch.cat([dev["rgb"], dev["hha"]], dim=3).contiguous()
    hh = pipeline.transform_images([six], channels=[(3, 3)], out="nchw")
    ref_six = torch.stack([O.img_transform(np.concatenate([host["rgb"][i], host["hha"][i]], 2)) for i in range(n)])
    assert torch.equal(hh.cpu(), ref_six[:, 3:])           # a 6-channel IMAGE uses mean[3:6] = .485 (transform.py:307)


def test_unnormalize_and_round_trip(cuda_dev):
    from mcd_b200 import pipeline
    host, dev = _planes(9, 2, 40, 56, cuda_dev)
    x = pipeline.transform_images([dev["rgb"]], out="nchw")
    back = pipeline.unnormalize(x)
    ref = np.stack([O.unnormalize(np.transpose(x[i].cpu().numpy(), (1, 2, 0))) for i in range(2)])
    assert np.array_equal(back.cpu().numpy(), ref)
    # encode -> decode: the float64 round trip lands within one grey level below (truncation) of the source bytes
    diff = host["rgb"].astype(np.int32) - back.cpu().numpy().astype(np.int32)
    assert diff.min() >= 0 and diff.max() <= 1


@pytest.mark.parametrize("dtypes", [(torch.uint8, torch.int64), (torch.int64, torch.int64), (torch.uint8, torch.uint8),
                                    (torch.int64, torch.uint8)])
def test_fast_hist_vs_oracle(cuda_dev, dtypes):
    """full-size frames, void (255) ground truth skipped, accumulation over images; a checksum property on top:
    the counts sum to the number of valid ground-truth pixels"""
    import eval as E
    n = 40
    rng = np.random.RandomState(3)
    cm = E.ConfusionMatrix(n)
    total = np.zeros((n, n), dtype=np.int64)
    valid = 0
    for _ in range(3):
        gt = rng.randint(0, n, (480, 640))
        gt[rng.rand(480, 640) < 0.07] = 255
        pred = np.where(rng.rand(480, 640) < 0.5, np.minimum(gt, n - 1), rng.randint(0, n, (480, 640)))
        total += O.fast_hist(gt.flatten(), pred.flatten(), n)
        valid += int((gt < n).sum())
        cm.update(torch.from_numpy(gt).to(cuda_dev).to(dtypes[0]), torch.from_numpy(pred).to(cuda_dev).to(dtypes[1]))
    hist = cm.hist()
    assert hist.dtype == np.int64 and np.array_equal(hist, total) and int(hist.sum()) == valid
    assert np.array_equal(E.per_class_iu(hist), O.per_class_iu(total))


def test_fast_hist_edge_cases(cuda_dev):
    import eval as E
    from mcd_b200 import pipeline
    # nothing valid: all-void ground truth -> zero matrix
    gt = torch.full((1000,), 255, dtype=torch.uint8, device=cuda_dev)
    pred = torch.zeros(1000, dtype=torch.uint8, device=cuda_dev)
    assert int(E.fast_hist(gt, pred, 40).sum()) == 0
    # negative int64 ground truth is skipped like the reference's (a >= 0) mask
    gt = torch.tensor([-1, 0, 1, 39, 40], dtype=torch.int64, device=cuda_dev)
    pred = torch.tensor([0, 0, 1, 39, 39], dtype=torch.int64, device=cuda_dev)
    h = E.fast_hist(gt, pred, 40)
    assert int(h.sum()) == 3 and h[0, 0] == 1 and h[1, 1] == 1 and h[39, 39] == 1
    # a prediction outside [0, n): the reference's reshape raises
    bad = pipeline.fast_hist(torch.tensor([3], device=cuda_dev), torch.tensor([77], device=cuda_dev), 40)
    with pytest.raises(ValueError):
        pipeline.hist_matrix(bad, 40)
    # largest supported table and a single-class one
    g = torch.randint(0, 104, (5000,), device=cuda_dev)
    assert np.array_equal(E.fast_hist(g, g, 104), np.diag(np.bincount(g.cpu().numpy(), minlength=104)))
    assert E.fast_hist(torch.zeros(7, dtype=torch.uint8, device=cuda_dev), torch.zeros(7, dtype=torch.uint8, device=cuda_dev), 1)[0, 0] == 7


@pytest.mark.parametrize("sizes", [((480, 640), (530, 730)), ((480, 640), (425, 560)), ((480, 640), (480, 640)),
                                   ((60, 80), (1, 1)), ((33, 47), (100, 31))])
def test_resize_nearest_vs_pil(cuda_dev, sizes):
    from PIL import Image
    from mcd_b200 import pipeline
    (ih, iw), (oh, ow) = sizes
    rng = np.random.RandomState(11)
    m = rng.randint(0, N_CLASS, (2, ih, iw)).astype(np.uint8)
    out = pipeline.resize_nearest(torch.from_numpy(m).to(cuda_dev), (ow, oh)).cpu().numpy()
    for i in range(2):
        assert np.array_equal(out[i], np.array(Image.fromarray(m[i]).resize((ow, oh), Image.NEAREST)))


def test_relabel_ragged_and_idempotent(cuda_dev):
    from mcd_b200 import pipeline
    for numel in (1, 3, 4, 5, 1023, 480 * 640 + 2):
        src = torch.randint(0, 256, (numel,), dtype=torch.uint8, device=cuda_dev)
        out = pipeline.relabel(src, N_CLASS)
        ref = src.long()
        ref[ref == 255] = N_CLASS - 1
        assert torch.equal(out, ref)
        again = pipeline.relabel(out.to(torch.uint8), N_CLASS)
        assert torch.equal(again, out)


def _early_models(dev):
    from models.model_util import get_models
    G = O.to_device(O.fill_state_dict_(O.init_seg_base("drn_d_38", 6, N_CLASS), 1), dev)
    F1 = O.to_device(O.fill_state_dict_(O.init_head(N_CLASS), 2), dev)
    F2 = O.to_device(O.fill_state_dict_(O.init_head(N_CLASS), 3), dev)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        models = [m.to(dev).train() for m in get_models("drn_d_38", 6, N_CLASS)]
    for m, sd in zip(models, (G, F1, F2)):
        m.load_state_dict({k: v.clone() for k, v in sd.items()})
    return models


@pytest.mark.parametrize("graph", [False, True])
def test_mcdstep_from_uint8_equals_fp32_inputs(cuda_dev, graph):
    """adapt_trainer.py:156-215 fed with the decoded uint8 planes (transform on the GPU inside the iteration) and with
    the fp32 tensors the reference's loader would have produced from the same bytes: same losses, same weights"""
    from loss import CrossEntropyLoss2d, Diff2d
    from mcd_b200.pipeline import InputPipeline
    from mcd_b200.step import MCDStep
    dev, (n, h, w) = cuda_dev, (2, 128, 160)
    host, d = _planes(21, n, h, w, dev)
    _, dt = _planes(22, n, h, w, dev)
    wt = O.class_weight(N_CLASS).to(dev)
    src32 = torch.stack([O.assemble_input(host["rgb"][i], host["hha"][i]) for i in range(n)]).to(dev)
    tgt32 = torch.stack([O.assemble_input(dt["rgb"][i].cpu().numpy(), dt["hha"][i].cpu().numpy()) for i in range(n)]).to(dev)
    lbl64 = torch.stack([O.lbl_transform(host["lbl"][i], N_CLASS) for i in range(n)]).to(dev)
    res = []
    for pipe in (None, InputPipeline("early", N_CLASS)):
        models = _early_models(dev)
        step = MCDStep(models, CrossEntropyLoss2d(wt), Diff2d(), num_k=2, input_pipeline=pipe)
        batch = (src32, lbl64, tgt32) if pipe is None else ((d["rgb"], d["hha"]), d["lbl"], (dt["rgb"], dt["hha"]))
        if graph:
            step(*batch)
            step.capture(*batch, warmup=0)
            c, dl = step.replay(*batch)
        else:
            c, dl = step(*batch)
        torch.cuda.synchronize()
        res.append((float(c), float(dl), [p.detach().clone() for m in models for p in m.parameters()]))
    (c0, d0, p0), (c1, d1, p1) = res
    # (the graph variant compares the SECOND iteration: weights already differ by the order of the fp32 atomics)
    assert abs(c0 - c1) <= 1e-4 * abs(c0) and abs(d0 - d1) <= 1e-3 * abs(d0), (c0, c1, d0, d1)
    for a, b in zip(p0, p1):            # identical operands; the order of the fp32 atomics differs between two runs
        assert float((a - b).abs().max()) <= 1e-3 * float(a.abs().max()) + 1e-7


def test_pipelines_for_mfnet_and_multitask(cuda_dev):
    """stream / target layout of the other trainers: MFNet = two 3-channel streams (adapt_mfnet_trainer.py:186-187),
    triple multitask = RGB stream + fp32 HHA and boundary targets (adapt_triple_multitask_trainer.py:194-196)"""
    from mcd_b200.pipeline import InputPipeline
    n, h, w = 2, 64, 96
    host, d = _planes(31, n, h, w, cuda_dev)
    ref7 = torch.stack([O.assemble_input(host["rgb"][i], host["hha"][i], host["bd"][i]) for i in range(n)])
    with torch.no_grad():
        b = InputPipeline("mfnet").images((d["rgb"], d["hha"]))
        assert len(b.streams) == 2
        assert torch.equal(b.streams[0][:, :3].float().cpu(), ref7[:, :3].half().float())
        assert torch.equal(b.streams[1][:, :3].float().cpu(), ref7[:, 3:6].half().float())
        b = InputPipeline("multitask").images((d["rgb"], d["hha"]), d["bd"])
        assert torch.equal(b.streams[0][:, :3].float().cpu(), ref7[:, :3].half().float())
        assert torch.equal(b.aux.cpu(), ref7[:, 3:])
