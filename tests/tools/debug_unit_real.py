"""Like debug_unit.py but on a REAL unit of DRN-D-38 (oracle weights, the oracle's input and upstream gradient).
    python tests/tools/debug_unit_real.py base.8.0 [H W N]"""
import os
import sys
import warnings

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
sys.path.insert(0, ROOT)
from mcd_b200 import abi, ops  # noqa: E402
from oracle import mcd_oracle as O  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
unit_name = sys.argv[1] if len(sys.argv) > 1 else "base.8.0"
h, w, n = (int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (240, 320, 2)
HF, BF = torch.float16, torch.bfloat16


def l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-300))


def rd(t, dt):
    return t.to(dt).float()


G = O.to_device(O.fill_state_dict_(O.init_seg_base("drn_d_38", 6, 41), 1), dev)
F1 = O.to_device(O.fill_state_dict_(O.init_head(41), 2), dev)
g0 = torch.Generator().manual_seed(5)
src = torch.randn(n, 6, h, w, generator=g0).to(dev)
_ = torch.randn(n, 6, h, w, generator=g0)
lbl = torch.randint(0, 41, (n, h, w), generator=g0).to(dev)
wgt = O.class_weight(41).to(dev)
O._req([G, F1])
taps = {}
feat = O.seg_base_forward(G, src, taps=taps)
loss = O.ce2d(O.head_forward(F1, feat), lbl, wgt)
keys = [k for k in taps if k.endswith(":out")]
grads = torch.autograd.grad(loss, [taps[k] for k in keys])
d_out = dict(zip(keys, grads))
units = [u for st in O.trunk_spec("drn_d_38", "base.") for u in st]
idx = [O.unit_key(u) for u in units].index(unit_name + ":out")
unit = units[idx]
assert unit[0] == "cbr", "conv-BN-ReLU units only"
_, kc, kb, stride, dil, pad = unit
x = taps[O.unit_key(units[idx - 1])].detach()
W, gamma, beta = G[kc + ".weight"].detach(), G[kb + ".weight"].detach(), G[kb + ".bias"].detach()
dz_up = d_out[unit_name + ":out"]
C = W.shape[0]
print("unit %s: x %s  |dz| max %.3e  x mean %.3f sparsity %.3f" % (unit_name, tuple(x.shape), float(dz_up.abs().max()),
                                                                  float(x.mean()), float((x == 0).float().mean())))
for algo_name, algo in (("umma", abi.ALGO_UMMA), ("direct", abi.ALGO_DIRECT)):
    with torch.no_grad():
        x16, w16 = rd(x, HF), rd(W, HF)
        y32 = F.conv2d(x16.double(), w16.double(), None, stride, pad, dil)
        y_e = rd(y32.float(), HF)
        mean = y_e.double().mean((0, 2, 3)); var = y_e.double().var((0, 2, 3), unbiased=False)
        rstd = (var + 1e-5).rsqrt()
        xhat = ((y_e.double() - mean[None, :, None, None]) * rstd[None, :, None, None])
        v_e = xhat * gamma.double()[None, :, None, None] + beta.double()[None, :, None, None]
        z_e = torch.relu(v_e).float()
        g_e = rd(dz_up, BF).double() * (v_e > 0)
        m1 = g_e.mean((0, 2, 3)); m2 = (g_e * xhat).mean((0, 2, 3))
        dy_e = rd(((gamma.double() * rstd)[None, :, None, None] * (g_e - m1[None, :, None, None] - xhat * m2[None, :, None, None])).float(), BF)
        dx_e = torch.nn.grad.conv2d_input(x.shape, rd(W, BF).double(), dy_e.double(), stride, pad, dil)
        dW_e = torch.nn.grad.conv2d_weight(rd(x, BF).double(), W.shape, dy_e.double(), stride, pad, dil)
    xn = ops.to_nhwc(x)
    gm = ops.conv_geom(xn.shape, W.shape[1], C, W.shape[2], W.shape[3], stride, dil, pad)
    y, stats = ops.conv_fprop(xn, ops.pack_weight_for(W, gm, 0, algo), None, gm, want_stats=True, algo=algo)
    bn = torch.nn.BatchNorm2d(C).to(dev)
    bn.weight.data.copy_(gamma); bn.bias.data.copy_(beta)
    z, save, _ = ops.bn_forward(y, stats, bn, True)
    dzn = ops.to_nhwc(dz_up, grad=True)
    dy, dgam, dbet, _, _, _ = ops.bn_bwd(dzn, z, y, bn.weight, save, True, True)
    dx = ops.conv_dgrad(dy, ops.pack_weight_for(W, gm, 1, algo), gm, algo=algo)
    dW, _ = ops.conv_wgrad(xn, dy, gm, algo=algo)
    torch.cuda.synchronize()
    yg = ops.to_nchw_f32(y); zg = ops.to_nchw_f32(z); dyg = ops.to_nchw_f32(dy)
    flips = float(((zg > 0) != (z_e > 0)).float().mean())
    flips32 = float(((zg > 0) != (taps[unit_name + ":out"] > 0)).float().mean())
    d = (yg.double() - y32).abs() / y_e.abs().clamp_min(1e-3).double()
    print("[%s] y %.2e (exact-equal %.4f; |y_gpu - y_fp64|/|y| mean %.2e) mean %.2e rstd %.2e z %.2e flips vs emu %.2e vs fp32 %.2e"
          " | dy %.2e | dx %.2e dW %.2e | dgamma %.2e dbeta %.2e" % (
              algo_name, l2(yg, y_e), float((yg == y_e).float().mean()), float(d.mean()), l2(save[0], mean),
              l2(save[1], rstd), l2(zg, z_e), flips, flips32, l2(dyg, dy_e), l2(ops.to_nchw_f32(dx), dx_e), l2(dW, dW_e),
              l2(dgam, (g_e * xhat).sum((0, 2, 3))), l2(dbet, g_e.sum((0, 2, 3)))))
    print("    dx vs fp32 oracle %.3e   emulation vs fp32 oracle %.3e" % (
        l2(ops.to_nchw_f32(dx), d_out[O.unit_key(units[idx - 1])]), l2(dx_e, d_out[O.unit_key(units[idx - 1])])))

# ---- the module path (what tests/test_parity_gpu.py runs)
from models.model_util import get_models  # noqa: E402
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    mg = get_models("drn_d_38", 6, 41)[0].to(dev)
mg.load_state_dict({k: v.detach().clone() for k, v in G.items()})
mg.train()
stage = mg.base[int(unit_name.split(".")[1])]
xin = ops.to_nhwc(x).requires_grad_(True)
out = stage(xin)
out.backward(ops.to_nhwc(dz_up, grad=True))
torch.cuda.synchronize()
print("module path: act %.3e dx vs fp32 %.3e  vs ops-path dx %.3e ; dW vs ops-path %.3e" % (
    l2(ops.to_nchw_f32(out), taps[unit_name + ":out"]), l2(ops.to_nchw_f32(xin.grad), d_out[O.unit_key(units[idx - 1])]),
    l2(ops.to_nchw_f32(xin.grad), ops.to_nchw_f32(dx)), l2(stage[0].weight.grad, dW)))
from mcd_b200 import nn as mnn  # noqa: E402
y_mod = out._mcd_bn_y
print("module y vs ops y: equal %.6f ; module z vs ops z equal %.6f" % (
    float((y_mod == y).float().mean()), float((ops.b16(out) == ops.b16(z)).float().mean())))
for overlap in (False, True):
    mnn.set_overlap_wgrad(overlap)
    stage.zero_grad()
    xin = ops.to_nhwc(x).requires_grad_(True)
    out = stage(xin)
    out.backward(ops.to_nhwc(dz_up, grad=True))
    torch.cuda.synchronize()
    print("overlap=%d: dx vs ops-path %.3e dW vs ops-path %.3e dgamma vs ops %.3e" % (
        overlap, l2(ops.to_nchw_f32(xin.grad), ops.to_nchw_f32(dx)), l2(stage[0].weight.grad, dW),
        l2(stage[1].weight.grad, dgam)))
# manual replay of the module's backward pieces
aff_mod = None
dy2, dgam2, dbet2, _, _, _ = ops.bn_bwd(ops.to_nhwc(dz_up, grad=True), out, y_mod, stage[1].weight, save, True, True)
print("bn_bwd on module tensors vs ops-path dy: %.3e" % l2(ops.to_nchw_f32(dy2), ops.to_nchw_f32(dy)))
# capture what the module's backward hands to the kernels
cap = {}
orig_bn_bwd, orig_dgrad = ops.bn_bwd, ops.conv_dgrad
def bn_bwd_spy(dz_, z_, y_, gamma_, aff_, *a, **k):
    cap["bn"] = (dz_, z_, y_, gamma_, aff_, a, k)
    r = orig_bn_bwd(dz_, z_, y_, gamma_, aff_, *a, **k)
    cap["dy"] = r[0]
    return r
def dgrad_spy(dy_, wp_, g_, *a, **k):
    cap["dgrad_dy"] = dy_
    cap["wp"] = wp_
    return orig_dgrad(dy_, wp_, g_, *a, **k)
ops.bn_bwd, ops.conv_dgrad = bn_bwd_spy, dgrad_spy
stage.zero_grad()
xin = ops.to_nhwc(x).requires_grad_(True)
out = stage(xin)
out.backward(ops.to_nhwc(dz_up, grad=True))
torch.cuda.synchronize()
dz_m, z_m, y_m, gamma_m, aff_m, a_m, k_m = cap["bn"]
print("module bn_bwd args: dz vs given %.3e | z dtype %s y dtype %s | aff mean %.3e rstd %.3e | gamma %.3e | extra %s %s" % (
    l2(ops.to_nchw_f32(dz_m), rd(dz_up, BF)), z_m.dtype, y_m.dtype, l2(aff_m[0], save[0]), l2(aff_m[1], save[1]),
    l2(gamma_m, gamma), a_m, {kk: (None if vv is None else (vv if not torch.is_tensor(vv) else tuple(vv.shape))) for kk, vv in k_m.items()}))
print("module dy vs ops dy %.3e ; dgrad got same dy: %s ; dgrad pack vs fresh %.3e" % (
    l2(ops.to_nchw_f32(cap["dy"]), ops.to_nchw_f32(dy)), cap["dgrad_dy"].data_ptr() == cap["dy"].data_ptr(),
    l2(cap["wp"].float(), ops.pack_weight_for(W, gm, 1, abi.ALGO_UMMA).float())))
